"""Fill the on-disk cache of run-time specialised kernels (sdf_b200/.jitcache, or $SDFB200_JIT_CACHE)
for the models of the test corpus and the benchmark configurations.  NVRTC needs no GPU, so this runs on
the build machine; a GPU box that receives the directory starts without compiling.

  python tools/jit_warm.py [--jobs 8] [--kernels 0,1,2,3,4,5,6,7,8]
"""
import argparse
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

FORCED = [(None, None), ('-1', '4'), ('0', '11'), ('1', '6'), ('2', '11'), ('2', '3'), ('1', '16'), ('0', '1')]


def _work(item):
    name, kernels, forced = item
    import models
    import sdf_b200
    from sdf_b200 import _abi
    from sdf_b200.ir import lower
    allm = dict(models.build(sdf_b200))
    S = sdf_b200
    allm['grp_cyl_x'] = S.cylinder(0.7).orient(S.X) & S.sphere(1.5)
    allm['grp_cyl_y'] = S.rounded_box((2, 1, 1.5), 0.2) - S.cylinder(0.4).orient(S.Y).k(0.1)
    for k, v in models.CONFIGS.items():
        allm.setdefault('cfg_' + k, v(S))
    done = 0
    for share, group in (FORCED if forced else FORCED[:1]):
        for k, v in (('SDFB200_JIT_SHARE', share), ('SDFB200_JIT_N', group)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        h = _abi.ProgramHandle(lower(allm[name]))
        for kern in (kernels if share is None else [kk for kk in kernels if 1 <= kk <= 4]):
            try:
                h.jit_probe(kern)
                done += 1
            except Exception as e:          # e.g. a program too large for one kernel flavour
                print(name, kern, str(e)[:100])
    return name, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--jobs', type=int, default=os.cpu_count() or 4)
    ap.add_argument('--kernels', default='0,1,2,3,4,5,6,7,8')
    ap.add_argument('--only', default='', help='only models whose name starts with this prefix (e.g. cfg_: the five BASELINE configurations)')
    args = ap.parse_args()
    kernels = [int(k) for k in args.kernels.split(',')]
    import models
    import sdf_b200
    names = list(dict(models.build(sdf_b200)).keys()) + ['grp_cyl_x', 'grp_cyl_y'] + ['cfg_' + k for k in models.CONFIGS]
    names = list(dict.fromkeys(names))
    if args.only:
        names = [n for n in names if n.startswith(args.only)]
    forced = {'grp_cyl_x', 'grp_cyl_y', 'cfg_gearlike', 'cfg_knurling', 'gearlike', 'knurling'}
    items = [(n, kernels, n in forced) for n in names]
    with mp.Pool(args.jobs) as pool:
        total = 0
        for name, done in pool.imap_unordered(_work, items):
            total += done
    print('compiled or found', total, 'kernels for', len(items), 'models')


if __name__ == '__main__':
    main()
