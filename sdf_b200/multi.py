"""Multi-GPU mesh generation: Z-slab split of the batch lattice, one process per GPU.

The reference has no distributed path; its only parallelism is a ThreadPool over independent
batches (core.py:131-133).  Batches stay independent here: the field is analytic, so a slab
boundary plane is simply re-evaluated by both neighbours (exactly what adjacent reference
batches do through the ``s + 1`` slice overlap, core.py:115-117) -- no halo exchange, no
data-path collective.  The only collective is one variable-length all-gather of the triangle
buffers at the end (NCCL over NVLink/NVSwitch via torch.distributed; gloo on CPU in the tests).

Slabs are whole Z batch layers, cut so that every rank gets about the same number of
NON-SKIPPED batches (the skip pre-pass over the full lattice costs <= 9 samples per batch).
"""
import numpy as np

from . import core


def partition_layers(active_per_layer, world):
    """Cut range(len(active_per_layer)) into ``world`` contiguous [begin, end) ranges whose
    active-batch sums are as equal as a prefix-sum cut allows.  Ranges cover every layer,
    are disjoint and ordered; a range may be empty when there are fewer layers than ranks."""
    a = np.asarray(active_per_layer, dtype=np.float64)
    n = len(a)
    # a small per-layer floor keeps empty layers (cheap but not free) from piling on one rank
    w = a + max(a.sum(), 1.0) * 1e-3 / max(n, 1)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target, side='left'))
        # choose the nearer of the two neighbouring cut positions
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def plan_slabs(sdf, X, Y, Z, batch_size, world, device=None):
    """Run the skip pre-pass for the whole lattice and return the per-rank Z-layer ranges."""
    h, _ = core.program_handle(sdf)
    dev = core.default_device() if device is None else device
    flags = h.skip_flags(np.ascontiguousarray(X, np.float64), np.ascontiguousarray(Y, np.float64),
                         np.ascontiguousarray(Z, np.float64), batch_size, device=dev)
    active = (flags == 0).sum(axis=(0, 1))
    return partition_layers(active, world)


class _PinnedTorchPool:
    """Page-locked host buffers for the gathered result, recycled when the arrays handed out die
    (same contract as sdf_b200._abi.PinnedPool, but torch-owned so that ``copy_(non_blocking=True)``
    runs at PCIe rate)."""

    def __init__(self, max_free=2):
        self.free, self.max_free = [], max_free

    def take(self, n_floats):
        import weakref
        import torch
        n = max(int(n_floats), 1024)
        best = None
        for i, t in enumerate(self.free):
            if t.numel() >= n and (best is None or t.numel() < self.free[best].numel()):
                best = i
        if best is not None:
            t = self.free.pop(best)
        else:
            t = torch.empty(1 << max(10, (n - 1).bit_length()), dtype=torch.float32, pin_memory=True)
        arr = t.numpy()
        weakref.finalize(arr, self._release, t)
        return t, arr

    def _release(self, t):
        if len(self.free) < self.max_free:
            self.free.append(t)


_pinned = _PinnedTorchPool()


def allgather_triangles(local, group=None, compact=True):
    """Concatenate every rank's (n_r, 9) float32 triangle tensor on every rank.

    Counts are gathered first, buffers are padded to the largest count for one
    ``all_gather_into_tensor``; ``compact=True`` returns the dense (T, 9) tensor in rank (= Z-slab)
    order, ``compact=False`` the padded buffer with ``(counts, stride)`` so that the caller can copy
    the valid segments itself (the host gather does, straight into pinned memory)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.cpu().tolist()
    mx = max(max(counts), 1)
    padded = local
    if local.shape[0] != mx:
        padded = torch.empty((mx, 9), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    gathered = torch.empty((world * mx, 9), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if not compact:
        return gathered, counts, mx
    if all(c == mx for c in counts):
        return gathered
    return torch.cat([gathered[r * mx: r * mx + c] for r, c in enumerate(counts)])


def generate_distributed(sdf, step=None, bounds=None, samples=core.SAMPLES, batch_size=core.BATCH_SIZE,
                         sparse=True, precision='exact', plan=None, gather='host', group=None,
                         compute=None, device=None):
    """SPMD entry: every rank calls this with the same arguments.

    Returns (points, info): ``points`` is the concatenated soup of ALL ranks -- a (3T,3) float32
    ndarray for gather='host', a CUDA tensor (T,9) for gather='device', or only this rank's slab
    for gather='none'.  ``compute(zb) -> (tris ndarray/tensor (n,9), info)`` can be injected (the
    CPU gloo tests use the oracle there)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds, _, (X, Y, Z) = core.make_lattice(sdf, step, bounds, samples)
    if plan is None:
        if compute is not None:
            nbz = (len(Z) + batch_size - 1) // batch_size
            plan = partition_layers(np.ones(nbz), world)
        else:
            plan = plan_slabs(sdf, X, Y, Z, batch_size, world, device=device)
    zb = plan[rank]
    if compute is not None:
        tris, info = compute(zb)
        local = torch.as_tensor(np.ascontiguousarray(tris, np.float32).reshape(-1, 9))
    else:
        dev = core.default_device() if device is None else device
        need, info = _count(sdf, X, Y, Z, batch_size, sparse, precision, zb, dev)
        out = torch.empty((max(need, 1), 9), dtype=torch.float32, device=torch.device('cuda', dev))
        tri, info = core.run_batches(sdf, X, Y, Z, batch_size, sparse, precision, zb=zb, device=dev, out=out,
                                     stream=torch.cuda.current_stream(dev).cuda_stream)
        local = tri
    info = dict(info, zb=zb, rank=rank, world=world)
    if gather == 'host' and world > 1 and local.is_cuda:
        # padded all-gather, then the valid segment of every rank straight into one pinned array
        gathered, counts, mx = allgather_triangles(local, group, compact=False)
        total = int(sum(counts))
        t, arr = _pinned.take(total * 9)
        off = 0
        for r, c in enumerate(counts):
            if c:
                t[off * 9:(off + c) * 9].copy_(gathered[r * mx:r * mx + c].reshape(-1), non_blocking=True)
            off += c
        torch.cuda.current_stream(local.device).synchronize()
        info['triangles_total'] = total
        return arr[:total * 9].reshape(-1, 3), info
    if gather == 'none' or world == 1:
        allt = local
    else:
        allt = allgather_triangles(local, group)
    info['triangles_total'] = int(allt.shape[0])
    if gather == 'device':
        return allt, info
    return allt.cpu().numpy().reshape(-1, 3), info


_count_cache = {}


def _count(sdf, X, Y, Z, batch_size, sparse, precision, zb, dev):
    """Triangle count of a slab (count-only pass), cached per (program, lattice, slab)."""
    h, key = core.program_handle(sdf)
    ck = (key, len(X), len(Y), len(Z), float(X[0]), float(X[-1]), batch_size, bool(sparse), precision, tuple(zb))
    if ck not in _count_cache:
        tau = core.mixed_tau(h.program, X, Y, Z)
        n, _, _ = h.generate(np.ascontiguousarray(X), np.ascontiguousarray(Y), np.ascontiguousarray(Z), batch_size,
                             sparse, precision, tau, None, 0, zb=zb, device=dev)
        _count_cache[ck] = int(n * 1.02) + 1024
    return _count_cache[ck], {}
