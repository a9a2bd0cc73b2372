"""Multi-GPU mesh generation: Z-slab split of the batch lattice, one process per GPU.

The reference has no distributed path; its only parallelism is a ThreadPool over independent
batches (core.py:131-133).  Batches stay independent here: the field is analytic, so a slab
boundary plane is simply re-evaluated by both neighbours (exactly what adjacent reference
batches do through the ``s + 1`` slice overlap, core.py:115-117) -- no halo exchange, no
data-path collective.  The only collective is one variable-length all-gather of the triangle
buffers at the end (NCCL over NVLink/NVSwitch via torch.distributed; gloo on CPU in the tests).

A rank's share is a Z-slab: by default a WINDOW of the Z-major batch order (whole batch layers plus a
ragged first / last layer, ``plan_windows``), cut so that every rank gets the same number of NON-SKIPPED
batches even when the surface lives in fewer batch layers than there are ranks (the skip pre-pass over
the full lattice costs <= 9 samples per batch); ``plan_slabs`` cuts at whole layers only.
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


def partition_windows(weights, world):
    """Cut range(len(weights)) -- the batches in Z-MAJOR order, u = bz * (nbx * nby) + bx * nby + by -- into
    ``world`` contiguous windows [u0, u1) of (nearly) equal weight.  Windows cover every batch, are disjoint
    and ordered; a window may be empty."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    w = w + max(w.sum(), 1.0) * 1e-4 / max(n, 1)          # skipped batches are cheap, not free
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cum, cum[-1] * r / world, side='left'))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def zmajor_active(flags):
    """skip flags [nbx, nby, nbz] (core.py:119 order) -> activity per batch in Z-major order."""
    return (np.asarray(flags) == 0).transpose(2, 0, 1).reshape(-1)


def batch_weights(sdf, X, Y, Z, batch_size, device=None):
    """Per batch, in Z-major order: 1 for a non-skipped batch, 0 for a skipped one (the full-lattice pre-pass)."""
    h, _ = core.program_handle(sdf)
    dev = core.default_device() if device is None else device
    flags = h.skip_flags(np.ascontiguousarray(X, np.float64), np.ascontiguousarray(Y, np.float64),
                         np.ascontiguousarray(Z, np.float64), batch_size, device=dev)
    return zmajor_active(flags).astype(np.float64)


def reweigh(weights, windows, seconds):
    """Measured re-plan: the batches of a window that took longer than the mean weigh more, the others less (the
    cost of a batch depends on how much surface it holds, which the pre-pass cannot see).  Applied repeatedly the
    windows converge to equal measured times."""
    w = np.array(weights, dtype=np.float64)
    t = np.asarray(seconds, dtype=np.float64)
    mean = t.mean() if t.mean() > 0 else 1.0
    for (u0, u1), tr in zip(windows, t):
        if u1 > u0 and tr > 0:
            w[u0:u1] *= tr / mean
    return w


def plan_windows(sdf, X, Y, Z, batch_size, world, device=None, weights=None):
    """The per-rank windows of the Z-major batch order: balanced by non-skipped batches (a skip pre-pass over the
    whole lattice), or by the given per-batch ``weights`` (``batch_weights`` refined by ``reweigh``)."""
    if weights is None:
        weights = batch_weights(sdf, X, Y, Z, batch_size, device)
    return partition_windows(weights, world)


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
    """Concatenate every rank's (n_r, 9) float32 triangle tensor on every rank, in rank (= Z-slab) order.

    Counts are exchanged first (one tiny all-gather); on CUDA every rank then sends its slab to every
    peer and receives each peer's slab directly at its final offset of the dense output (one grouped
    NCCL send/recv batch: no padding to the largest slab, no compaction copy afterwards).  On CPU
    (gloo, the tests) the padded ``all_gather_into_tensor`` + compaction form is used."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.cpu().tolist()
    if local.is_cuda:
        total = int(sum(counts))
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        out = torch.empty((total, 9), dtype=local.dtype, device=local.device)
        local = local.contiguous()
        ops = []
        for r in range(world):
            if r == rank:
                continue
            if counts[rank]:
                ops.append(dist.P2POp(dist.isend, local, r, group))
            if counts[r]:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if counts[rank]:
            out[offs[rank]:offs[rank + 1]].copy_(local)
        for q in reqs:
            q.wait()
        return out if compact else (out, counts, None)
    mx = max(max(counts), 1)
    padded = local
    if local.shape[0] != mx:
        padded = torch.empty((mx, 9), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    gathered = torch.empty((world * mx, 9), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(c == mx for c in counts):
        return gathered
    return torch.cat([gathered[r * mx: r * mx + c] for r, c in enumerate(counts)])


class RootGather:
    """Fused gather to rank 0 (the B200 path): the extraction kernels of all ranks write into a buffer in rank 0's
    HBM over NVLink (CUDA-IPC mapped, system-scope atomic cursor); no collective moves triangles and NO collective or
    host barrier synchronises a step -- the ranks meet through flag words in rank 0's memory (include/sdfb200.h):
    a rank's extraction waits on the device until rank 0 has released the buffer of the step's parity, and signals
    when its triangles are in; rank 0 waits on the device for all signals, reads the count, consumes ``view(total)``
    and releases the buffer.  Two buffers ping-pong, so ranks run at most two steps ahead of rank 0's consumption."""

    def __init__(self, capacity, device, group=None):
        import torch.distributed as dist
        from . import _abi
        self.group, self.device = group, device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.capacity = int(capacity)
        box = [None]
        if self.rank == 0:
            self.shared = _abi.SharedOutput(self.capacity, device)
            box[0] = self.shared.handle
        dist.broadcast_object_list(box, src=0, group=group)
        if self.rank != 0:
            self.shared = _abi.SharedOutput(self.capacity, device, handle=box[0])
        self.step = 0
        self.last = None
        dist.barrier(group)                               # (set-up only) every rank has the buffer mapped

    def run(self, handle, X, Y, Z, batch_size, sparse, precision, tau, zb, stream=None, window=None, release=True):
        """One step on this rank (its Z-slab ``zb`` of whole layers, or its ``window`` of the Z-major batch order);
        returns (n_local, stats, timings, total) -- total only on rank 0.  ``release=False``: rank 0 keeps the
        step's buffer (``view(total)``) until it calls ``release()`` itself."""
        which = self.step & 1
        if window is not None:
            n_local, stats, timings = handle.generate_window(X, Y, Z, batch_size, sparse, precision, tau, window,
                                                             shared=self.shared, which=which, device=self.device, stream=stream)
        else:
            n_local, stats, timings = handle.generate_shared(X, Y, Z, batch_size, sparse, precision, tau, self.shared,
                                                             which, zb=zb, device=self.device, stream=stream)
        total = None
        if self.rank == 0:
            total = self.shared.wait(which, self.world, stream)           # device-side wait for every rank's signal
            if total > self.capacity:
                raise RuntimeError('shared triangle buffer too small: %d needed, %d given' % (total, self.capacity))
            self.last = which
            if release:
                self.release(stream)
        self.step += 1
        return n_local, stats, timings, total

    def release(self, stream=None):
        """Rank 0: done with the buffer of the last step (its cursor is reset, the step after next may use it)."""
        if self.rank == 0 and self.last is not None:
            self.shared.release(self.last, stream)
            self.last = None

    def view(self, total):
        """(total, 9) float32 CUDA tensor over the buffer of the last step (rank 0 only, before ``release()``)."""
        import torch
        ptr = self.shared.buffer_ptr(self.last)
        iface = {'shape': (int(total), 9), 'typestr': '<f4', 'data': (ptr, False), 'version': 2}
        holder = type('CudaView', (), {'__cuda_array_interface__': iface})()
        return torch.as_tensor(holder, device=torch.device('cuda', self.device))


class HostGather:
    """Host output of a multi-GPU call over ALL PCIe links: one page-locked POSIX shared-memory segment that every
    rank process maps; after its slab is generated a rank publishes its triangle count in the segment's header,
    reads the others', and copies its triangles to its offset with its own GPU's DMA engine -- N ranks, N links
    (8 x ~50 GB/s) instead of funnelling the soup through rank 0's HBM and its single link.  Counts and completion are
    exchanged through the header words (every process polls shared memory: no collective, no CUDA call).  Two data
    regions alternate, so the array handed out by step s stays valid until step s + 2 overwrites it.  Single node."""

    HEADER = 4096
    MAX_RANKS = 64

    def __init__(self, capacity, device, group=None):
        import atexit
        import torch.distributed as dist
        from multiprocessing import shared_memory
        from . import _abi
        self.group, self.device = group, int(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= self.MAX_RANKS
        self.capacity = int(capacity)
        size = self.HEADER + 2 * self.capacity * 36
        box = [None]
        if self.rank == 0:
            try:        # (a container with a small /dev/shm cannot hold the segment: every rank then falls back, see _host_gather)
                import os
                self.shm = shared_memory.SharedMemory(create=True, size=size)
                os.posix_fallocate(self.shm._fd, 0, size)          # ENOSPC now, as an exception -- not SIGBUS at the first write
                np.ndarray((self.HEADER // 8,), np.uint64, buffer=self.shm.buf)[:] = 0
                box[0] = self.shm.name
            except (OSError, ValueError) as e:
                box[0] = 'ERROR: %s' % (e,)
                try:
                    self.shm.close(); self.shm.unlink()
                except Exception:
                    pass
                self.shm = None
        dist.broadcast_object_list(box, src=0, group=group)
        if isinstance(box[0], str) and box[0].startswith('ERROR'):
            raise OSError('HostGather: no shared host segment of %d MB (%s)' % (size >> 20, box[0]))
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=box[0])
            try:        # (Python < 3.13 registers attached segments with the resource tracker, which would unlink them)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, 'shared_memory')
            except Exception:
                pass
        self.hdr = np.ndarray((self.HEADER // 8,), np.uint64, buffer=self.shm.buf)
        self.data = np.ndarray((2, self.capacity * 9), np.float32, buffer=self.shm.buf, offset=self.HEADER)
        self.addr = self.data.ctypes.data
        _abi.host_register(self.hdr.ctypes.data, size)
        self.seq = 0
        dist.barrier(group)
        atexit.register(self.close)

    def close(self):
        from . import _abi
        shm, self.shm = getattr(self, 'shm', None), None
        if shm is None:
            return
        try:
            _abi.host_unregister(self.hdr.ctypes.data)
            self.hdr = self.data = None
            shm.close()
            if self.rank == 0:
                shm.unlink()
        except Exception:
            pass

    def _all(self, base, seq):
        """Spin until every rank's word at ``base`` carries ``seq`` in its upper half; returns the lower halves."""
        import time
        words = self.hdr[base:base + self.world]
        deadline = time.time() + 30.0
        while True:
            v = words.copy()
            if np.all((v >> np.uint64(32)) == np.uint64(seq)):
                return (v & np.uint64(0xffffffff)).astype(np.int64)
            if time.time() > deadline:
                raise RuntimeError('HostGather: a rank did not arrive within 30 s')

    def run(self, local, stream=None):
        """``local``: this rank's (n, 9) float32 CUDA tensor.  Returns the (total, 9) ndarray view on rank 0 (None elsewhere)
        -- or the needed capacity as a negative int on every rank when the segment is too small."""
        from . import _abi
        self.seq += 1
        s, p = self.seq, self.seq & 1
        n = int(local.shape[0])
        self.hdr[p * 64 + self.rank] = np.uint64((s << 32) | n)
        counts = self._all(p * 64, s)
        total, off = int(counts.sum()), int(counts[:self.rank].sum())
        fits = total <= self.capacity
        if fits and n:
            _abi.copy_to_host(self.addr + (p * self.capacity * 9 + off * 9) * 4, local.data_ptr(), n * 36,
                              device=self.device, stream=stream)
        self.hdr[128 + p * 64 + self.rank] = np.uint64((s << 32) | 1)
        if not fits:
            return -total
        if self.rank != 0:
            return None
        self._all(128 + p * 64, s)
        return self.data[p, :total * 9].reshape(total, 9)


_host_gathers = {}


def _host_gather(local_need, dev, group, total=None):
    """The cached HostGather of this process; created (collectively: every rank calls this at the same point) for the
    sum of the ranks' needs, or for ``total`` triangles when given."""
    import torch
    import torch.distributed as dist
    key = (dev, id(group))
    hg = _host_gathers.get(key)
    if hg is not None:
        return hg                    # (False: this box cannot hold the segment)
    if total is None:
        t = torch.tensor([int(local_need)], dtype=torch.int64, device=torch.device('cuda', dev))
        dist.all_reduce(t, group=group)
        total = int(t.item())
    cap = 1 << max(16, (int(total * 1.05) - 1).bit_length())
    try:
        hg = _host_gathers[key] = HostGather(cap, dev, group)
    except OSError:              # raised on every rank together (the creation outcome is broadcast)
        hg = _host_gathers[key] = False
    return hg


def generate_distributed(sdf, step=None, bounds=None, samples=core.SAMPLES, batch_size=core.BATCH_SIZE,
                         sparse=True, precision='exact', plan=None, gather='host', group=None,
                         compute=None, device=None, partition=None):
    """SPMD entry: every rank calls this with the same arguments.

    Returns (points, info): ``points`` is the concatenated soup of ALL ranks -- a (3T,3) float32
    ndarray for gather='shm' (rank 0 only: every rank copies its slab to one shared page-locked host segment
    over its own PCIe link -- the fastest host output), 'host' (NCCL all-gather, every rank gets it) / 'root'
    (fused NVLink gather into rank 0's HBM, then one copy; rank 0 only), a CUDA tensor (T,9) for
    gather='device', or only this rank's share for gather='none'.  ``partition``: 'window' (default on the GPU path: windows of
    the Z-major batch order, ``plan_windows``) or 'slab' (whole Z batch layers, ``plan_slabs``); ``plan``
    is the matching list of per-rank ranges.  ``compute(range) -> (tris ndarray/tensor (n,9), info)`` can
    be injected (the CPU gloo tests use the oracle there; default partition 'slab')."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds, _, (X, Y, Z) = core.make_lattice(sdf, step, bounds, samples)
    if partition is None:
        partition = 'slab' if compute is not None else 'window'
    windowed = partition == 'window'
    if plan is None:
        if compute is not None:
            nb = [(len(a) + batch_size - 1) // batch_size for a in (X, Y, Z)]
            plan = (partition_windows(np.ones(nb[0] * nb[1] * nb[2]), world) if windowed
                    else partition_layers(np.ones(nb[2]), world))
        elif windowed:
            plan = plan_windows(sdf, X, Y, Z, batch_size, world, device=device)
        else:
            plan = plan_slabs(sdf, X, Y, Z, batch_size, world, device=device)
    mine = tuple(plan[rank])
    zb, window = ((0, -1), mine) if windowed else (mine, None)
    if gather == 'root' and compute is None and world > 1:
        # fused gather: every rank's extraction kernel writes into rank 0's buffer over NVLink
        dev = core.default_device() if device is None else device
        h, key = core.program_handle(sdf)
        Xc, Yc, Zc = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
        tau = core.mixed_tau(h.program, Xc, Yc, Zc)
        need, _ = _count(sdf, Xc, Yc, Zc, batch_size, sparse, precision, zb, dev, window)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rg = _root_gather(need, dev, group)          # sized from the all-reduced per-rank counts of THIS lattice and plan
        n_local, stats, timings, total = rg.run(h, Xc, Yc, Zc, batch_size, sparse, precision, tau, zb,
                                                stream=stream, window=window, release=False)
        info = dict(triangles=n_local, skipped=int(stats[0]), empty=int(stats[1]), nonempty=int(stats[2]),
                    evaluated=int(stats[3]), refined=int(stats[4]), active_cubes=int(stats[5]), zb=mine, rank=rank,
                    world=world, triangles_total=total, partition=partition)
        if rank != 0:
            return np.zeros((0, 3), np.float32), info
        t, arr = _pinned.take(total * 9)
        t[:total * 9].copy_(rg.view(total).reshape(-1), non_blocking=True)
        rg.release(stream)                                # (stream order: after the copy out of the shared buffer)
        torch.cuda.current_stream(dev).synchronize()
        return arr[:total * 9].reshape(-1, 3), info
    if compute is not None:
        tris, info = compute(mine)
        local = torch.as_tensor(np.ascontiguousarray(tris, np.float32).reshape(-1, 9))
    else:
        dev = core.default_device() if device is None else device
        stream = torch.cuda.current_stream(dev).cuda_stream
        need, info = _count(sdf, X, Y, Z, batch_size, sparse, precision, zb, dev, window)
        for attempt in range(2):
            out = torch.empty((max(need, 1), 9), dtype=torch.float32, device=torch.device('cuda', dev))
            try:
                if windowed:
                    h, _ = core.program_handle(sdf)
                    Xc, Yc, Zc = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
                    n, st, _ = h.generate_window(Xc, Yc, Zc, batch_size, sparse, precision, core.mixed_tau(h.program, Xc, Yc, Zc),
                                                 window, tri_out=out, capacity=out.shape[0], device=dev, stream=stream)
                    local = out[:n]
                    info = dict(triangles=n, skipped=int(st[0]), empty=int(st[1]), nonempty=int(st[2]),
                                evaluated=int(st[3]), refined=int(st[4]), active_cubes=int(st[5]))
                else:
                    local, info = core.run_batches(sdf, X, Y, Z, batch_size, sparse, precision, zb=zb, device=dev, out=out,
                                                   stream=stream)
                break
            except _abi_capacity_error() as e:           # a stale cached count: size from what the call reports
                if attempt:
                    raise
                need = int(e.needed * 1.02) + 1024
    info = dict(info, zb=mine, rank=rank, world=world, partition=partition)
    if gather == 'shm' and world > 1 and local.is_cuda and _host_gather(need, local.device.index, group) is False:
        gather = 'host'              # no room in /dev/shm on this box: NCCL all-gather + one copy instead
    if gather == 'shm' and world > 1 and local.is_cuda:
        # every rank moves its slab to the host over its own PCIe link, into one shared page-locked segment
        hg = _host_gather(need, local.device.index, group)
        res = hg.run(local, torch.cuda.current_stream(local.device).cuda_stream)
        if isinstance(res, int):                        # too small (a stale count): every rank sees the same total
            _host_gathers.pop((local.device.index, id(group))).close()
            hg = _host_gather(0, local.device.index, group, total=-res)
            if hg is False:
                raise RuntimeError('HostGather: /dev/shm cannot hold %d triangles' % -res)
            res = hg.run(local, torch.cuda.current_stream(local.device).cuda_stream)
        if rank != 0:
            return np.zeros((0, 3), np.float32), info
        info['triangles_total'] = int(res.shape[0])
        return res.reshape(-1, 3), info
    if gather == 'host' and world > 1 and local.is_cuda:
        # dense gather on the device, then one copy into a pinned (pooled) host array
        allt = allgather_triangles(local, group)
        total = int(allt.shape[0])
        t, arr = _pinned.take(total * 9)
        t[:total * 9].copy_(allt.reshape(-1), non_blocking=True)
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


def _abi_capacity_error():
    from . import _abi
    return _abi.CapacityError


_root_gathers = {}


def _root_gather(local_need, dev, group):
    """RootGather sized for the sum of the ranks' needs (cached: creating one maps IPC memory)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(local_need)], dtype=torch.int64, device=torch.device('cuda', dev))
    dist.all_reduce(t, group=group)
    cap = 1 << max(16, (int(t.item()) - 1).bit_length())
    key = (cap, dev, id(group))
    if key not in _root_gathers:
        _root_gathers[key] = RootGather(cap, dev, group)
    return _root_gathers[key]


_count_cache = {}


def _count(sdf, X, Y, Z, batch_size, sparse, precision, zb, dev, window=None):
    """Triangle count of a rank's share (count-only pass), cached per (program, lattice axes, share)."""
    import hashlib
    h, key = core.program_handle(sdf)
    X, Y, Z = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
    axes = hashlib.blake2b(X.tobytes() + Y.tobytes() + Z.tobytes(), digest_size=16).digest()      # the WHOLE lattice, not its ends
    ck = (key, axes, batch_size, bool(sparse), precision, tuple(zb), tuple(window) if window is not None else None)
    if ck not in _count_cache:
        if len(_count_cache) > 256:
            _count_cache.clear()
        tau = core.mixed_tau(h.program, X, Y, Z)
        if window is not None:
            n, _, _ = h.generate_window(X, Y, Z, batch_size, sparse, precision, tau, window, device=dev)
        else:
            n, _, _ = h.generate(X, Y, Z, batch_size, sparse, precision, tau, None, 0, zb=zb, device=dev)
        _count_cache[ck] = int(n * 1.02) + 1024
    return _count_cache[ck], {}
