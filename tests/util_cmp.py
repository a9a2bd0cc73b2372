"""Triangle-soup comparison used by the parity tests and smoke()."""
import numpy as np


def _sorted_rows(a):
    a = np.ascontiguousarray(a)
    order = np.lexsort(a.T[::-1])
    return a[order]


def canonical_triangles(points):
    """(3T,3) soup -> (T,9) float32 rows; vertices of each triangle rotated so the
    lexicographically smallest comes first (cyclic rotation keeps the winding)."""
    t = np.asarray(points, dtype=np.float32).reshape(-1, 3, 3)
    if len(t) == 0:
        return t.reshape(0, 9)
    keys = np.stack([np.lexsort((t[:, k, 2], t[:, k, 1], t[:, k, 0])) for k in range(3)], 0)
    # rank of each vertex inside its triangle: do it directly
    first = np.zeros(len(t), np.int64)
    for k in (1, 2):
        a, b = t[np.arange(len(t)), first], t[:, k]
        less = (b[:, 0] < a[:, 0]) | ((b[:, 0] == a[:, 0]) & ((b[:, 1] < a[:, 1]) | ((b[:, 1] == a[:, 1]) & (b[:, 2] < a[:, 2]))))
        first = np.where(less, k, first)
    idx = (first[:, None] + np.arange(3)[None, :]) % 3
    t = t[np.arange(len(t))[:, None], idx]
    del keys
    return _sorted_rows(t.reshape(-1, 9))


def compare_soups(got, ref):
    """got / ref: (3T,3) vertex soups (any float dtype; compared as float32 like the STL sink).

    Returns dict(count_equal, exact_triangles, max_vertex_dist, max_sorted_diff)."""
    a = canonical_triangles(got)
    b = canonical_triangles(ref)
    rep = dict(n_got=len(a), n_ref=len(b), count_equal=len(a) == len(b))
    if len(a) == 0 or len(b) == 0:
        rep.update(exact_triangles=0, max_vertex_dist=0.0 if len(a) == len(b) else float('inf'),
                   max_sorted_diff=0.0)
        return rep
    if len(a) == len(b):
        eq = np.all(a == b, axis=1)
        rep['exact_triangles'] = int(eq.sum())
        rep['max_sorted_diff'] = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))
    else:
        rep['exact_triangles'] = -1
        rep['max_sorted_diff'] = float('inf')
    if rep['exact_triangles'] == len(a):           # bit-identical multisets: nothing left to measure
        rep['max_vertex_dist'] = 0.0
        return rep
    # vertex multiset distance (robust to sort ties): nearest-neighbour both ways
    from scipy.spatial import cKDTree
    va, vb = a.reshape(-1, 3).astype(np.float64), b.reshape(-1, 3).astype(np.float64)
    d1 = cKDTree(vb).query(va)[0].max()
    d2 = cKDTree(va).query(vb)[0].max()
    rep['max_vertex_dist'] = float(max(d1, d2))
    return rep


def mesh_topology(points, decimals=None):
    """Weld bit-identical vertices and return (V, E, F, boundary_edges, nonmanifold_edges)."""
    t = np.asarray(points, dtype=np.float32).reshape(-1, 3)
    uniq, inv = np.unique(t, axis=0, return_inverse=True)
    f = inv.reshape(-1, 3)
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]   # drop degenerate
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und = np.sort(e, axis=1)
    ue, counts = np.unique(und, axis=0, return_counts=True)
    # directed consistency: every undirected edge once in each direction
    base = len(uniq) + 1
    dire = e[:, 0].astype(np.int64) * base + e[:, 1]
    rev = e[:, 1].astype(np.int64) * base + e[:, 0]
    dk, dcounts = np.unique(dire, return_counts=True)
    rk, rcounts = np.unique(rev, return_counts=True)
    # orientation: every directed edge is matched by as many opposite ones
    fwd = dict(zip(dk.tolist(), dcounts.tolist()))
    unbalanced = sum(1 for k, c in zip(rk.tolist(), rcounts.tolist()) if fwd.get(k, 0) != c)
    return dict(V=len(uniq), E=len(ue), F=len(f), boundary=int((counts == 1).sum()),
                nonmanifold=int((counts > 2).sum()), repeated_directed=int((dcounts > 1).sum()),
                unbalanced=int(unbalanced))


def soup_multiset_bits(points):
    """The soup as a multiset of triangles compared by BIT PATTERN (NaN-aware, -0.0 != 0.0): Counter of 9-tuples of
    uint32, each triangle rotated to its smallest representation, every NaN mapped to one pattern.  For volumes with
    NaN / infinite samples, where the float comparisons of compare_soups are meaningless."""
    from collections import Counter
    t = np.array(points, dtype=np.float32).reshape(-1, 3, 3)
    t[np.isnan(t)] = np.float32(np.nan)
    u = t.view(np.uint32)
    return Counter(min(tuple(int(x) for x in np.roll(tri, -k, axis=0).reshape(-1)) for k in range(3)) for tri in u)


def special_value_volumes(count=6, n=12, seed=5, denormals=True):
    """Noise volumes with 15 % of the samples replaced by NaN, +-inf, +-0, +-FLT_EPSILON, +-3e38 (and denormals)."""
    rng = np.random.default_rng(seed)
    specials = [np.nan, np.inf, -np.inf, 0.0, -0.0, 3e38, -3e38, 1.1920929e-07, -1.1920929e-07]
    if denormals:
        specials += [1e-45, -1e-45, 1e-38]
    specials = np.array(specials, np.float32)
    out = []
    for _ in range(count):
        vol = rng.normal(size=(n, n, n)).astype(np.float32)
        m = rng.random(vol.shape) < 0.15
        vol[m] = rng.choice(specials, int(m.sum()))
        out.append(vol)
    return out
