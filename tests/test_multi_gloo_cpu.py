"""CPU suite, part 4: the N > 1 path on world_size-2 gloo (slab plan + variable-length all-gather).
The per-slab compute is injected (the CPU oracle restricted to the slab's batches), so the test
covers exactly the distributed host logic that runs on the GPUs with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import models
import sdf_b200
from sdf_b200 import core, multi
from oracle import sdf_ref
from util_cmp import compare_soups


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _slab_batches(X, Y, Z, bs, zb):
    nb = [(len(a) + bs - 1) // bs for a in (X, Y, Z)]
    return [(bx * nb[1] + by) * nb[2] + bz for bx in range(nb[0]) for by in range(nb[1]) for bz in range(zb[0], zb[1])]


def _worker(rank, world, port, step, L, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    f = models.cfg_csg(sdf_b200)
    bounds = ((-L, -L, -L), (L, L, L))
    _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)

    def compute(zb):
        pts, st = sdf_ref.generate(f, step=step, bounds=bounds, batch_subset=_slab_batches(X, Y, Z, 32, zb))
        return pts.astype(np.float32).reshape(-1, 9), st
    pts, info = multi.generate_distributed(f, step=step, bounds=bounds, gather='host', compute=compute)
    q.put((rank, pts, info['zb'], info['triangles_total']))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_slab_split_and_allgather_equal_single_process(world):
    step, L = 2.0 / 96, 1.0
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, step, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    f = models.cfg_csg(sdf_b200)
    ref_pts, _ = sdf_ref.generate(f, step=step, bounds=((-L,) * 3, (L,) * 3))
    ranges = sorted(g[2] for g in got)
    assert ranges[0][0] == 0 and ranges[-1][1] == 3 and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    for rank, pts, zb, total in got:
        assert total * 3 == len(pts)
        rep = compare_soups(pts, ref_pts)
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (rank, rep)


def _window_batches(X, Y, Z, bs, window):
    """Batch ids (core.py:119 order) of the window [u0, u1) of the Z-major order u = bz * (nbx * nby) + bx * nby + by."""
    nb = [(len(a) + bs - 1) // bs for a in (X, Y, Z)]
    out = []
    for u in range(window[0], window[1]):
        bz, r = divmod(u, nb[0] * nb[1])
        bx, by = divmod(r, nb[1])
        out.append((bx * nb[1] + by) * nb[2] + bz)
    return out


def _worker_windows(rank, world, port, step, L, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    f = models.cfg_csg(sdf_b200)
    bounds = ((-L, -L, -L), (L, L, L))
    _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)

    def compute(window):
        pts, st = sdf_ref.generate(f, step=step, bounds=bounds, batch_subset=_window_batches(X, Y, Z, 32, window))
        return pts.astype(np.float32).reshape(-1, 9), st
    pts, info = multi.generate_distributed(f, step=step, bounds=bounds, gather='host', compute=compute, partition='window')
    q.put((rank, pts, info['zb'], info['triangles_total']))
    dist.barrier()
    dist.destroy_process_group()


def test_window_split_cuts_inside_a_layer_and_equals_single_process():
    """Windows of the Z-major batch order (slabs with ragged first / last layers): 27 batches over 2 ranks cut
    INSIDE the middle layer; the gathered soup equals the single-process one."""
    step, L, world = 2.0 / 96, 1.0, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_windows, args=(r, world, port, step, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    f = models.cfg_csg(sdf_b200)
    ref_pts, _ = sdf_ref.generate(f, step=step, bounds=((-L,) * 3, (L,) * 3))
    ranges = sorted(g[2] for g in got)
    assert ranges[0][0] == 0 and ranges[-1][1] == 27 and ranges[0][1] == ranges[1][0] and ranges[0][1] % 9 != 0
    for rank, pts, window, total in got:
        rep = compare_soups(pts, ref_pts)
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (rank, rep)


def test_partition_windows_balances_active_batches():
    w = np.zeros(32 * 32 * 32)
    w[5000:9000] = 1                                  # the surface lives in 4 of 32 layers
    wins = multi.partition_windows(w, 8)
    assert wins[0][0] == 0 and wins[-1][1] == len(w) and all(wins[i][1] == wins[i + 1][0] for i in range(7))
    act = [w[a:b].sum() for a, b in wins]
    assert max(act) - min(act) <= 2, act
    flags = np.zeros((4, 3, 2), np.uint8)
    flags[1, 2, 0] = 1
    assert list(np.nonzero(~multi.zmajor_active(flags))[0]) == [1 * 3 + 2]


def test_measured_replans_converge_to_equal_times():
    """multi.reweigh: windows re-cut from measured per-rank times converge although the true cost of a batch (how
    much surface it holds) is invisible to the pre-pass."""
    rng = np.random.default_rng(3)
    n, world = 32 * 32 * 32, 8
    active = np.zeros(n)
    active[6000:26000] = (rng.random(20000) < 0.3)
    cost = active * (1.0 + 4.0 * np.exp(-((np.arange(n) - 15000) / 3000.0) ** 2))          # hidden: surface-heavy in the middle
    weights = active.copy()
    wins = multi.partition_windows(weights, world)
    spread = []
    for _ in range(4):
        t = [cost[a:b].sum() + 50.0 for a, b in wins]                                      # + a fixed per-rank cost
        spread.append(max(t) / min(t))
        weights = multi.reweigh(weights, wins, t)
        wins = multi.partition_windows(weights, world)
    assert spread[0] > 1.5 and spread[-1] < 1.06, spread


def _worker_hostgather(rank, world, port, q):
    """HostGather's protocol (counts / completion through header words of one shared segment, offsets, two data regions)
    with the two CUDA calls it makes replaced by host stand-ins: page-locking is a no-op, the DMA copy a memmove."""
    import ctypes
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sdf_b200 import _abi
    _abi.host_register = lambda addr, n: None
    _abi.host_unregister = lambda addr: None
    _abi.copy_to_host = lambda dst, src, n, device=0, stream=None: ctypes.memmove(dst, src, n)
    hg = multi.HostGather(1000, 0)
    copies, views = [], []
    for step in range(3):
        n = 10 * (rank + 1) + step                       # different counts per rank and step
        local = torch.full((n, 9), float(100 * step + rank), dtype=torch.float32)
        res = hg.run(local)
        if rank == 0:
            views.append(res)                            # a VIEW of the shared segment
            copies.append(np.array(res))
    too_big = hg.run(torch.zeros((800, 9), dtype=torch.float32))          # 2 x 800 > capacity: every rank learns the total
    if rank == 0:
        # the view handed out by step 1 is intact after step 2 wrote the OTHER region (two regions alternate), and after the
        # refused step; step 0's region was re-used by step 2, as documented
        assert np.array_equal(views[1], copies[1]) and np.array_equal(views[2], copies[2])
        assert not np.array_equal(np.array(views[0]), copies[0])
        q.put((copies, too_big))
    else:
        q.put((None, too_big))
    dist.barrier()
    hg.close()
    dist.destroy_process_group()


def test_hostgather_protocol_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_hostgather, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    outs = [g for g in got if g[0] is not None][0][0]
    for step, arr in enumerate(outs):
        n0, n1 = 10 + step, 20 + step
        assert arr.shape == (n0 + n1, 9)
        assert np.all(arr[:n0] == 100 * step) and np.all(arr[n0:] == 100 * step + 1)          # rank order, right offsets
    assert all(g[1] == -1600 for g in got)               # the segment was too small: the needed total, on every rank
