#!/usr/bin/env python
"""torchrun entry: N-GPU Z-slab generate == 1-GPU generate (triangle multiset), NCCL all-gather."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist
import models
import sdf_b200
from sdf_b200 import multi
from util_cmp import compare_soups


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok = True
    for name, n in (('csg', 128), ('gearlike', 160), ('knurling', 192)):
        f = models.CONFIGS[name](sdf_b200)
        _, L = models.CONFIG_GRIDS[name]
        bounds, step = ((-L, -L, -L), (L, L, L)), 2.0 * L / n
        one, info1 = sdf_b200.generate(f, step=step, bounds=bounds, verbose=False, device=local, return_info=True)
        for partition in ('window', 'slab'):
            pts, info = multi.generate_distributed(f, step=step, bounds=bounds, gather='host', device=local, partition=partition)
            rep = compare_soups(pts, one)
            good = rep['count_equal'] and rep['exact_triangles'] == rep['n_ref']
            ok = ok and good
            print('rank %d %s: %s %s local %d total %d single %d identical=%s' % (
                rank, name, partition, info['zb'], info['triangles'], info['triangles_total'], info1['triangles'], good), flush=True)
        # host output over every rank's own PCIe link (shared page-locked segment)
        spts, sinfo = multi.generate_distributed(f, step=step, bounds=bounds, gather='shm', device=local)
        if rank == 0:
            srep = compare_soups(spts, one)
            sgood = srep['count_equal'] and srep['exact_triangles'] == srep['n_ref']
            ok = ok and sgood
            print('rank 0 %s: SHM host gather total %d identical=%s' % (name, sinfo['triangles_total'], sgood), flush=True)
        # fused gather to rank 0 (NVLink peer writes from the extraction kernels)
        rpts, rinfo = multi.generate_distributed(f, step=step, bounds=bounds, gather='root', device=local)
        if rank == 0:
            rrep = compare_soups(rpts, one)
            rgood = rrep['count_equal'] and rrep['exact_triangles'] == rrep['n_ref']
            ok = ok and rgood
            print('rank 0 %s: ROOT gather total %d identical=%s' % (name, rinfo['triangles_total'], rgood), flush=True)
    t = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('MULTI-GPU CHECK', 'PASSED' if int(t.item()) else 'FAILED', 'world', world)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) else 1)


if __name__ == '__main__':
    main()
