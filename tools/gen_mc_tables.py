#!/usr/bin/env python
"""Generator + checker for the marching-cubes lookup tables of sdf_b200.

What the reference does (sdf/core.py:16-18) is call
``skimage.measure.marching_cubes(volume, 0)`` -- Lewiner et al.'s 33-case,
topologically consistent marching cubes.  scikit-image (a third-party, unpinned
dependency: setup.py:14 ``scikit-image>=0.17``) is absent from this image and
from the wheelhouse, so its ~13k-entry tiling tables cannot be read.  This tool
REGENERATES a table set with the same decision structure:

 * cube index bit k  <=> corner k value > level (strict), Lewiner numbering
 * cases 1,2,5,8,9,11,14 (unambiguous): the classic triangle table, which is
   what Lewiner's TILING1/2/5/8/9/11/14 carry
 * cases 3,4,6,7,10,12,13: face tests (asymptotic decider with the FLT_EPSILON
   dead band) and the interior test choose a sub-tiling; sub-tilings here are
   produced from the iso-contour loops on the cube faces:
     - a loop of n <= 7 vertices is a disc of n-2 triangles,
     - a loop of n >= 8 vertices is a fan of n triangles around the centre
       vertex (edge id 12), as in Lewiner's 7.3 / 10.2 / 12.2 / 13.3 / 13.4,
     - an interior "tunnel" joins two loops of n and m vertices with n+m
       triangles (4.1.2, 7.4.2, 10.1.2, 12.1.2, 13.5.2), except 6.1.2 which
       has one interior edge split at the centre vertex (9 triangles), so the
       per-subcase triangle COUNTS equal Lewiner's.
   Which diagonal a polygon is split along inside an ambiguous sub-tiling may
   differ from Lewiner's table (parity unpinned: no source available).

Every generated tiling is machine-checked: directed boundary == iso-contour
segments of the cube faces for that sign pattern / face resolution, interior
edges paired with opposite direction, expected component count.

Outputs (identical content) : sdf_b200/csrc/mc_tables.inc, oracle/mc_tables.inc
and a python mirror tools/mc_tables_data.py for the tests.
"""
import itertools
import os
import sys
from functools import lru_cache

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mc_classic_table import CLASSIC  # noqa: E402

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0),
           (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4),
         (0, 4), (1, 5), (2, 6), (3, 7)]
# face id -> corners (A,B,C,D) exactly as skimage's test_face reads them
FACES = {1: (0, 4, 5, 1), 2: (1, 5, 6, 2), 3: (2, 6, 7, 3),
         4: (3, 7, 4, 0), 5: (0, 3, 2, 1), 6: (4, 7, 6, 5)}
FACE_NORMAL = {1: (0, -1, 0), 2: (1, 0, 0), 3: (0, 1, 0),
               4: (-1, 0, 0), 5: (0, 0, -1), 6: (0, 0, 1)}
OPPOSITE_FACE = {1: 3, 3: 1, 2: 4, 4: 2, 5: 6, 6: 5}
EDGE_ID = {frozenset(e): i for i, e in enumerate(EDGES)}


def bit(index, c):
    return (index >> c) & 1


def sub(a, b):
    return tuple(x - y for x, y in zip(a, b))


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def edge_mid(e):
    a, b = EDGES[e]
    return tuple((CORNERS[a][k] + CORNERS[b][k]) / 2 for k in range(3))


def vpos(e):
    return (0.5, 0.5, 0.5) if e == 12 else edge_mid(e)


def centroid(cs):
    return tuple(sum(CORNERS[c][k] for c in cs) / len(cs) for k in range(3))


def ambiguous_faces(index):
    out = []
    for f, cs in FACES.items():
        p = [bit(index, c) for c in cs]
        if sum(p) == 2 and p[0] == p[2]:
            out.append(f)
    return out


def segments(index, J):
    """Directed iso-contour segments (edge_from, edge_to) on the six faces.

    Direction convention (that of the classic table / Lewiner's tilings): seen
    from outside the cube, the > level region is on the LEFT of the segment.
    J[f] (ambiguous faces only) == True  <=> the two positive corners of face f
    are joined across the face.
    """
    segs = []
    for f, cs in FACES.items():
        p = [bit(index, c) for c in cs]
        n = sum(p)
        cuts = []  # (corner list cut off, sign of cut off region)
        if n == 1:
            cuts.append(([p.index(1)], 1))
        elif n == 3:
            cuts.append(([p.index(0)], 0))
        elif n == 2:
            if p[0] == p[2]:  # diagonal
                pos = [i for i in range(4) if p[i]]
                neg = [i for i in range(4) if not p[i]]
                if J[f]:
                    cuts += [([i], 0) for i in neg]
                else:
                    cuts += [([i], 1) for i in pos]
            else:
                i = [k for k in range(4) if p[k] and p[(k + 1) % 4]][0]
                cuts.append(([i, (i + 1) % 4], 1))
        for idxs, sign in cuts:
            first, last = idxs[0], idxs[-1]
            P = EDGE_ID[frozenset((cs[(first - 1) % 4], cs[first]))]
            Q = EDGE_ID[frozenset((cs[last], cs[(last + 1) % 4]))]
            mp, mq = edge_mid(P), edge_mid(Q)
            d = sub(mq, mp)
            mid = tuple((a + b) / 2 for a, b in zip(mp, mq))
            side = dot(cross(FACE_NORMAL[f], d), sub(centroid([cs[i] for i in idxs]), mid))
            assert side != 0
            if (side > 0) != (sign == 1):
                P, Q = Q, P
            segs.append((P, Q))
    return segs


def loops_of(segs):
    nxt = {}
    for a, b in segs:
        assert a not in nxt
        nxt[a] = b
    seen, loops = set(), []
    for s in sorted(nxt):
        if s in seen:
            continue
        loop, v = [], s
        while v not in seen:
            seen.add(v)
            loop.append(v)
            v = nxt[v]
        assert v == s
        loops.append(loop)
    return loops


def tri_list(flat):
    return [tuple(flat[i:i + 3]) for i in range(0, len(flat), 3)]


def count_in_face_chords(tris):
    return sum(1 for t in tris for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])) if a < b and in_face(a, b)
               and (b, a) in {(u, v) for tt in tris for u, v in ((tt[0], tt[1]), (tt[1], tt[2]), (tt[2], tt[0]))})


def check_tiling(index, J, tris, n_components=None, strict_chords=False):
    """True iff tris is an oriented manifold patch set bounded by segments().

    strict_chords additionally rejects interior chords that lie inside a cube face (the
    generator avoids them wherever a triangulation without them exists; the tunnels of 7.4.2 /
    13.5.2 cannot avoid them without an extra vertex)."""
    directed = {}
    for t in tris:
        if len(set(t)) != 3:
            return False
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            if (a, b) in directed:
                return False
            directed[(a, b)] = 1
    boundary = {e for e in directed if (e[1], e[0]) not in directed}
    if boundary != set(segments(index, J)):
        return False
    if strict_chords:
        for (a, b) in directed:
            if (a, b) not in boundary and in_face(a, b):
                return False
    if n_components is not None:
        parent = list(range(len(tris)))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        owner = {}
        for ti, t in enumerate(tris):
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                key = frozenset((a, b))
                if key in owner:
                    parent[find(ti)] = find(owner[key])
                else:
                    owner[key] = ti
        if len({find(i) for i in range(len(tris))}) != n_components:
            return False
    return True


def d2(a, b):
    pa, pb = vpos(a), vpos(b)
    return sum((x - y) ** 2 for x, y in zip(pa, pb))


def faces_of_edge(e):
    a, b = EDGES[e]
    return {f for f, cs in FACES.items() if a in cs and b in cs}


def in_face(a, b):
    """True when the chord between edge-vertices a and b lies inside a cube face.  Such a chord
    would make the patch touch the face (and the neighbouring cube's patch) along a line."""
    if a == 12 or b == 12:
        return False
    return bool(faces_of_edge(a) & faces_of_edge(b))


FORBID = 1e6   # cost of a chord lying in a cube face


def disc(loop):
    """Minimum sum-of-squared-diagonals triangulation of a directed loop."""
    n = len(loop)
    if n == 3:
        return [tuple(loop)]

    @lru_cache(None)
    def best(i, j):  # polygon loop[i..j], chord (i,j) already paid
        if j - i < 2:
            return (0.0, ())
        cands = []
        for k in range(i + 1, j):
            c1, t1 = best(i, k)
            c2, t2 = best(k, j)
            cost = c1 + c2
            if k - i > 1:
                cost += d2(loop[i], loop[k]) + (FORBID if in_face(loop[i], loop[k]) else 0)
            if j - k > 1:
                cost += d2(loop[k], loop[j]) + (FORBID if in_face(loop[k], loop[j]) else 0)
            cands.append((round(cost, 9), t1 + ((loop[i], loop[k], loop[j]),) + t2))
        return min(cands)
    return list(best(0, n - 1)[1])


def fan(loop, centre=12):
    n = len(loop)
    return [(loop[i], loop[(i + 1) % n], centre) for i in range(n)]


def tube(A, B):
    """Annulus between directed loops A and B (surface on the left of both).

    State (i, j) is the open directed edge B[j0-j] -> A[i]; A is walked forward
    and B backward.  All monotone lattice paths (0,0)->(n,m) are enumerated
    (<= C(12,6) of them); a path is valid when no interior edge repeats, and the
    cheapest (sum of squared interior edge lengths, ties by triangle list) wins.
    """
    n, m = len(A), len(B)
    best = None
    for j0 in range(m):
        for steps in itertools.combinations(range(n + m), n):
            steps = set(steps)
            i = j = 0
            used = {(0, 0)}
            cost = d2(A[0], B[j0]) + (FORBID if in_face(A[0], B[j0]) else 0)
            tris, ok = [], True
            for s in range(n + m):
                ai, bj = A[i % n], B[(j0 - j) % m]
                if s in steps:
                    a2 = A[(i + 1) % n]
                    tris.append((ai, a2, bj))
                    i += 1
                else:
                    b2 = B[(j0 - j - 1) % m]
                    tris.append((b2, bj, ai))
                    j += 1
                key = (i % n, j % m)
                if s < n + m - 1:
                    if key in used:
                        ok = False
                        break
                    used.add(key)
                    cost += d2(A[i % n], B[(j0 - j) % m]) + (FORBID if in_face(A[i % n], B[(j0 - j) % m]) else 0)
            if not ok:
                continue
            cand = (round(cost, 9), tuple(tris))
            if best is None or cand < best:
                best = cand
    return list(best[1])


def split_edge_at_centre(tris, loops):
    """Split the longest loop-to-loop interior edge at vertex 12 (6.1.2)."""
    where = {}
    for li, lp in enumerate(loops):
        for v in lp:
            where[v] = li
    cands = set()
    for t in tris:
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            if where[a] != where[b]:
                cands.add((round(d2(a, b), 9), min(a, b), max(a, b)))
    _, a, b = max(cands)
    out = []
    for t in tris:
        if a in t and b in t:
            k = [i for i in range(3) if t[i] not in (a, b)][0]
            t = t[k:] + t[:k]                      # (c, x, y) with {x,y}={a,b}
            out.append((t[0], t[1], 12))
            out.append((t[0], 12, t[2]))
        else:
            out.append(t)
    return out


# ---------------------------------------------------------------------------
# case classification (Lewiner / Chernyaev numbering)

def rotations():
    """The 24 proper rotations as corner permutations."""
    perms = set()
    for axes in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            # determinant
            m = [[0] * 3 for _ in range(3)]
            for r in range(3):
                m[r][axes[r]] = signs[r]
            det = (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1])
                   - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
                   + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))
            if det != 1:
                continue
            perm = []
            for c in CORNERS:
                v = [2 * x - 1 for x in c]
                w = [sum(m[r][k] * v[k] for k in range(3)) for r in range(3)]
                perm.append(CORNERS.index(tuple((x + 1) // 2 for x in w)))
            perms.add(tuple(perm))
    assert len(perms) == 24
    return sorted(perms)


ROT = rotations()


def apply_perm(index, perm):
    out = 0
    for c in range(8):
        if bit(index, c):
            out |= 1 << perm[c]
    return out


# representative index -> case, for npos <= 4 (complements share the case)
_REPS = {0: 0, 1: 1, 3: 2, 5: 3, 65: 4, 7: 5, 67: 6, 37: 7, 15: 8, 27: 9,
         195: 10, 23: 11, 135: 12, 165: 13, 29: 14}
# 65 = {0,6} body diagonal; 67 = {0,1,6}; 37 = {0,2,5}; 27 = {0,1,3,4} tripod;
# 195 = {0,1,6,7} two opposite edges; 23 = {0,1,2,4}; 135 = {0,1,2,7};
# 165 = {0,2,5,7}; 29 = {0,2,3,4} (mirror image of 23)


def case_of(index):
    npos = bin(index).count('1')
    idx = index if npos <= 4 else index ^ 255
    for perm in ROT:
        r = apply_perm(idx, perm)
        if r in _REPS:
            return _REPS[r]
    raise AssertionError(index)


# Reference edges of the interior test.  case 6: Lewiner test6[.][2], keyed by
# the positive (minority) corner set; case 7: edge id == isolated corner id
# (test7[.][4]); case 12: test12[.][3]; case 13.5: ring edge starting at the
# isolated positive corner (tiling13_5_1[.][.][0]).  Written from the published
# tables; each entry is checked below to be a sign-changing edge incident to
# the isolated corner.
_TEST6_EDGE = {
    (0, 1, 6): 10, (0, 1, 7): 11, (0, 2, 4): 1, (0, 2, 6): 3, (0, 3, 5): 9,
    (0, 3, 6): 10, (0, 4, 6): 5, (0, 5, 6): 8, (0, 6, 7): 8, (1, 2, 4): 8,
    (1, 2, 7): 11, (1, 3, 5): 2, (1, 3, 7): 0, (1, 4, 7): 9, (1, 5, 7): 6,
    (1, 6, 7): 9, (2, 3, 4): 8, (2, 3, 5): 9, (2, 4, 5): 10, (2, 4, 6): 7,
    (2, 4, 7): 10, (3, 4, 5): 11, (3, 5, 6): 11, (3, 5, 7): 4,
}
# case 12 in lexicographic order of the 4-corner positive set
_TEST12_EDGE_LIST = [11, 10, 5, 7, 9, 1, 2, 0, 3, 6, 4, 8, 8, 4, 6, 3, 0, 2, 1, 9, 7, 5, 10, 11]


def corner_sets(case, npos):
    return [tuple(c for c in range(8) if bit(i, c)) for i in range(256)
            if bin(i).count('1') == npos and case_of(i) == case]


def minority(index):
    npos = bin(index).count('1')
    idx = index if npos <= 4 else index ^ 255
    return tuple(c for c in range(8) if bit(idx, c))


def neighbours(c):
    return [b if a == c else a for a, b in EDGES if c in (a, b)]


def isolated_corner(cs):
    """Corner of cs that has no edge-neighbour inside cs."""
    iso = [c for c in cs if not any(n in cs for n in neighbours(c))]
    return iso


def interior_edge(index):
    """Reference edge for test_interior, or 12 for the z-sweep of cases 4/10."""
    case = case_of(index)
    mn = minority(index)
    if case in (4, 10):
        return 12
    if case == 6:
        e = _TEST6_EDGE[mn]
        iso = isolated_corner(mn)
        assert len(iso) == 1 and iso[0] in EDGES[e]
        return e
    if case == 7:
        common = [c for c in range(8) if all(n in mn for n in neighbours(c))]
        assert len(common) == 1
        e = common[0]
        assert common[0] in EDGES[e]
        return e
    if case == 12:
        sets = sorted(corner_sets(12, 4))
        e = _TEST12_EDGE_LIST[sets.index(mn)]
        iso = isolated_corner(mn)
        if not (len(iso) == 1 and iso[0] in EDGES[e]):
            # recalled entry inconsistent -> fall back to a vertical/ring edge rule
            raise AssertionError(('test12', mn, e, iso))
        return e
    return None


# ---------------------------------------------------------------------------
# tilings

def disc_tiling(index, J):
    """Triangles when every loop is filled as a disc (or centre fan)."""
    loops = loops_of(segments(index, J))
    flat = CLASSIC[index]
    if flat and all(len(lp) <= 7 for lp in loops):
        tris = tri_list(flat)
        if check_tiling(index, J, tris, len(loops)):
            return tris, 'classic'
    tris = []
    for lp in loops:
        tris += fan(lp) if len(lp) >= 8 else disc(lp)
    assert check_tiling(index, J, tris, len(loops)), (index, J)
    return tris, 'generated'


def tunnel_tiling(index, J, loopA, loopB, via_centre=False):
    loops = loops_of(segments(index, J))
    tris = tube(loopA, loopB)
    if via_centre:
        tris = split_edge_at_centre(tris, [loopA, loopB])
    for lp in loops:
        if lp != loopA and lp != loopB:
            tris += fan(lp) if len(lp) >= 8 else disc(lp)
    assert check_tiling(index, J, tris, len(loops) - 1), (index, J)
    return tris


def loop_around(loops, index, corner):
    """The loop that is the 3-cycle around `corner` (its three edges)."""
    want = {EDGE_ID[frozenset((corner, n))] for n in neighbours(corner)}
    for lp in loops:
        if set(lp) == want:
            return lp
    raise AssertionError((index, corner, loops))


def build():
    """Return the full decision structure.

    per index: case, faces (list of face ids, Lewiner sign applied at run time:
    + when npos<=4, - otherwise), and for each of the 2**nfaces test-result
    masks a record (tile_true, tile_false, interior) where interior is -1 (no
    interior test; tile_true is the answer), 0..11 (reference edge) or 12
    (z-sweep).  tile ids index the deduplicated tiling list; -1 = emit nothing.
    """
    tilings, tiling_id = [], {}
    stats = {'classic': 0, 'generated': 0}

    def intern(tris):
        key = tuple(tris)
        if key not in tiling_id:
            tiling_id[key] = len(tilings)
            tilings.append(list(tris))
        return tiling_id[key]

    per_index = []
    for index in range(256):
        case = case_of(index)
        npos = bin(index).count('1')
        sign_pos = npos <= 4
        faces = ambiguous_faces(index)
        subs = []
        for mask in range(1 << len(faces)):
            # bit j = literal result of test_face(sign * faces[j]):
            #   sign +: True <=> positives joined ; sign -: True <=> negatives joined
            res = {f: bool((mask >> j) & 1) for j, f in enumerate(faces)}
            J = {f: (res[f] if sign_pos else not res[f]) for f in faces}
            M = res  # "minority joined" in Lewiner's wording
            nM = sum(M.values())
            rec = None
            if case in (0,):
                rec = (-1, -1, -1)
            elif case in (1, 2, 5, 8, 9, 11, 14):
                tris, how = disc_tiling(index, J)
                assert how == 'classic', ('classic row wrong for unambiguous index', index)
                stats[how] += 1
                rec = (intern(tris), -1, -1)
            else:
                loops = loops_of(segments(index, J))
                mn = minority(index)
                use_interior = ((case == 4) or (case == 6 and nM == 0) or
                                (case == 7 and nM == 3) or
                                (case in (10, 12) and nM == 0))
                if case == 13:
                    joined = [f for f in faces if J[f]]
                    k = len(joined)
                    pos = [c for c in range(8) if bit(index, c)]
                    possible = True
                    if k == 2:
                        possible = OPPOSITE_FACE[joined[0]] != joined[1]
                    elif k == 4:
                        nj = [f for f in faces if not J[f]]
                        possible = OPPOSITE_FACE[nj[0]] != nj[1]
                    elif k == 3:
                        possible = all(OPPOSITE_FACE[a] != b for a in joined for b in joined)
                    if not possible:
                        subs.append((-1, -1, -1))
                        continue
                    if k == 3:
                        # the three faces share a corner; 13.4 if it is positive,
                        # 13.5 (interior test) if it is negative
                        common = set(FACES[joined[0]]) & set(FACES[joined[1]]) & set(FACES[joined[2]])
                        (cc,) = common
                        if not bit(index, cc):
                            # isolated positive corner = the one opposite to cc
                            far = [c for c in pos if c not in neighbours(cc)]
                            assert len(far) == 1
                            ip = far[0]
                            discs, how = disc_tiling(index, J)
                            stats[how] += 1
                            la = loop_around(loops, index, ip)
                            big = [lp for lp in loops if len(lp) == 6][0]
                            tun = tunnel_tiling(index, J, la, big)
                            assert len(discs) == 6 and len(tun) == 10
                            rec = (intern(discs), intern(tun), ip)  # ring edge id == corner id
                            assert ip in EDGES[ip]
                if rec is None and use_interior:
                    discs, how = disc_tiling(index, J)
                    stats[how] += 1
                    e = interior_edge(index)
                    if case == 7:
                        # 7.4: isolated corner of the MAJORITY sign; tunnel joins its
                        # triangle with the hexagon.  test true -> tunnel (7.4.2)
                        la = loop_around(loops, index, e)
                        big = [lp for lp in loops if lp is not la][0]
                        tun = tunnel_tiling(index, J, la, big)
                        assert len(discs) == 5 and len(tun) == 9
                        rec = (intern(tun), intern(discs), e)
                    else:
                        assert len(loops) == 2
                        tun = tunnel_tiling(index, J, loops[0], loops[1], via_centre=(case == 6))
                        want = {4: (2, 6), 6: (3, 9), 10: (4, 8), 12: (4, 8)}[case]
                        assert (len(discs), len(tun)) == want, (case, len(discs), len(tun))
                        rec = (intern(discs), intern(tun), e)
                elif rec is None:
                    tris, how = disc_tiling(index, J)
                    stats[how] += 1
                    rec = (intern(tris), -1, -1)
                    expect = {3: {0: 2, 1: 4}, 6: {1: 5}, 7: {0: 3, 1: 5, 2: 9},
                              10: {1: 8, 2: 4}, 12: {1: 8, 2: 4}}
                    if case in expect:
                        assert len(tris) == expect[case][nM], (case, nM, len(tris))
                    if case == 13:
                        k = sum(J.values())
                        assert len(tris) == {0: 4, 1: 6, 2: 10, 3: 12, 4: 10, 5: 6, 6: 4}[k], (k, len(tris))
            subs.append(rec)
        per_index.append({'case': case, 'faces': faces, 'subs': subs, 'npos': npos})
    return per_index, tilings, stats


# ---------------------------------------------------------------------------
# emit

def flatten(per_index, tilings):
    case = [p['case'] for p in per_index]
    nface = [len(p['faces']) for p in per_index]
    faces = []
    for p in per_index:
        faces += (p['faces'] + [0] * 6)[:6]
    sub_base, sub_true, sub_false, sub_int = [], [], [], []
    for p in per_index:
        sub_base.append(len(sub_true))
        for t, f, e in p['subs']:
            sub_true.append(t)
            sub_false.append(f)
            sub_int.append(e)
    tile_off, tile_ntri, tile_data = [], [], []
    for t in tilings:
        tile_off.append(len(tile_data))
        tile_ntri.append(len(t))
        for tri in t:
            tile_data += list(tri)
    return dict(MC_CASE=case, MC_NFACE=nface, MC_FACES=faces, MC_SUB_BASE=sub_base,
                MC_SUB_TRUE=sub_true, MC_SUB_FALSE=sub_false, MC_SUB_INTERIOR=sub_int,
                MC_TILE_OFF=tile_off, MC_TILE_NTRI=tile_ntri, MC_TILE_DATA=tile_data)


CTYPE = dict(MC_CASE='unsigned char', MC_NFACE='unsigned char', MC_FACES='unsigned char',
             MC_SUB_BASE='unsigned short', MC_SUB_TRUE='short', MC_SUB_FALSE='short',
             MC_SUB_INTERIOR='signed char', MC_TILE_OFF='unsigned short',
             MC_TILE_NTRI='unsigned char', MC_TILE_DATA='unsigned char')


def emit_c(flat, path):
    lines = ['/* GENERATED by tools/gen_mc_tables.py -- do not edit.',
             ' * Marching-cubes decision tables (Lewiner case structure; see the generator',
             ' * docstring for what is restated and what is regenerated).',
             ' * Define MC_TABLE_QUAL before including (e.g. "static const" or',
             ' * "__device__ const"). */',
             '#ifndef MC_TABLE_QUAL', '#define MC_TABLE_QUAL static const', '#endif']
    for name, arr in flat.items():
        lines.append('#define %s_LEN %d' % (name, len(arr)))
        lines.append('MC_TABLE_QUAL %s %s[%d] = {' % (CTYPE[name], name, len(arr)))
        for i in range(0, len(arr), 24):
            lines.append('  ' + ', '.join(str(v) for v in arr[i:i + 24]) + ',')
        lines.append('};')
    with open(path, 'w') as fp:
        fp.write('\n'.join(lines) + '\n')


def emit_py(flat, path):
    with open(path, 'w') as fp:
        fp.write('# GENERATED by tools/gen_mc_tables.py -- do not edit.\n')
        for name, arr in flat.items():
            fp.write('%s = %r\n' % (name, arr))


def main():
    per_index, tilings, stats = build()
    flat = flatten(per_index, tilings)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit_c(flat, os.path.join(root, 'sdf_b200', 'csrc', 'mc_tables.inc'))
    emit_c(flat, os.path.join(root, 'oracle', 'mc_tables.inc'))
    emit_py(flat, os.path.join(root, 'tools', 'mc_tables_data.py'))
    print('tilings: %d, tile data bytes: %d, sub records: %d, classic rows reused: %d, generated: %d'
          % (len(tilings), len(flat['MC_TILE_DATA']), len(flat['MC_SUB_TRUE']),
             stats['classic'], stats['generated']))
    print('max triangles per cube:', max(flat['MC_TILE_NTRI']))


if __name__ == '__main__':
    main()
