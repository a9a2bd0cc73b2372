"""How much evaluation work per-tile operand pruning could save ("next lever") -- measured on
the BASELINE trees with the CPU oracle, no GPU involved.

For every non-skipped batch (= tile) of a configuration the operands of every CSG fold are evaluated at the tile
centre; an operand is prunable in a tile when it provably cannot influence the fold anywhere inside it, assuming
the operands are Lipschitz-L (true distance fields and rigid transforms: L = 1):

    union          min(d1, d2), smooth K:   d2 is dead if (d2 - L r) - (d1 + L r) >= K   (and symmetrically d1)
    intersection   max(d1, d2)              d2 is dead if (d1 - L r) - (d2 + L r) >= K
    difference     max(d1, -d2)             d2 is dead if (d1 - L r) - (-d2 + L r) >= K

with r the tile's half diagonal (the smooth folds are EXACTLY the surviving operand beyond K: dn.py:16-18,31-33,46-48).
Work is counted in the rough float32 instruction counts per opcode that sdfb200.cu's share_axis() uses (COST below).
Folds below a non-rigid node are not looked into.

    python tools/pruning_potential.py [config ...]      default: csg gearlike blobby knurling
"""
import os
import sys
import time

import numpy as np

# rough instruction counts of the float32 throughput path per program word (sdfb200.cu share_axis)
COST = dict(SPHERE=8, PLANE=3, BOX=18, ROUNDED_BOX=18, WIREFRAME_BOX=55, TORUS=10, CAPSULE=22, CYLINDER=5, CAPPED_CYLINDER=35,
            ROUNDED_CYLINDER=19, CAPPED_CONE=45, ROUNDED_CONE=25, ELLIPSOID=29, PYRAMID=43, TETRAHEDRON=9, OCTAHEDRON=6,
            DODECAHEDRON=23, ICOSAHEDRON=23, TRANSLATE=3, SCALE_P=6, ROTATE=9, TWIST=14, BEND=14, ELONGATE=12, CIRCULAR_PRE=55,
            REPEAT_CELL=15, MULC=1, NEGATE=1, ADDC=1, SHELL=2, ADD_POP=1, UNION=1, DIFFERENCE=1, INTERSECTION=1, BLEND=3,
            UNION_K=12, DIFFERENCE_K=12, INTERSECTION_K=12, CIRCLE=6, LINE2=2, RECTANGLE=12, ROUNDED_RECTANGLE=16,
            EQUILATERAL_TRIANGLE=20, HEXAGON=20, ROUNDED_X=10, VESICA=14, REVOLVE=5, EXTRUDE_PRE=2, EXTRUDE_POST=10,
            SLICE_PRE=0, SLICE_SELECT=2, TPARAM_LINEAR=8, TPARAM_RADIAL=8, TPARAM_Z=4, EASE=10, DISPLACE=3, LERP_POP2=3,
            MIX_POP2=3, WRAP_PRE=35, WRAP_POST=1, PUSH=0, SAVEP=0, LOADP=0)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]


def analyse(name, lipschitz=(1.0, 2.0)):
    import models
    import sdf_b200 as S
    from oracle import sdf_ref
    from sdf_b200.ir import OPNAME, lower, node_of
    f = models.CONFIGS[name](S)
    n, L = models.CONFIG_GRIDS[name]
    X, Y, Z = sdf_ref.lattice(((-L, -L, -L), (L, L, L)), 2.0 * L / n)
    t0 = time.time()
    active = [job for job in sdf_ref.batches_of(X, Y, Z, 32) if not sdf_ref.skip(f, job)]
    C = np.array([[(a[0] + a[-1]) / 2 for a in job] for job in active])
    R = np.array([np.linalg.norm([(a[-1] - a[0]) / 2 for a in job]) for job in active])
    words = {}

    def cost_of(obj):
        if id(obj) not in words:
            words[id(obj)] = sum(COST.get(OPNAME[int(w) & 0xff], 10) for w in lower(obj).code)
        return words[id(obj)]

    def fold_value(op, d1, d2, K):
        if op == 'union':
            if not K:
                return np.minimum(d1, d2)
            h = np.clip(0.5 + 0.5 * (d2 - d1) / K, 0, 1)
            return d2 + (d1 - d2) * h - K * h * (1 - h)
        if op == 'intersection':
            if not K:
                return np.maximum(d1, d2)
            h = np.clip(0.5 - 0.5 * (d2 - d1) / K, 0, 1)
            return d2 + (d1 - d2) * h + K * h * (1 - h)
        if not K:
            return np.maximum(d1, -d2)
        h = np.clip(0.5 - 0.5 * (d2 + d1) / K, 0, 1)
        return d1 + (-d2 - d1) * h + K * h * (1 - h)

    def walk(obj, pts, lip):
        """(value at the tile centres, words needed per tile with pruning)"""
        nd = node_of(obj)
        if nd.kind == 'transform' and nd.op in ('TRANSLATE', 'ROTATE'):
            q = pts - np.array(nd.consts[:3]) if nd.op == 'TRANSLATE' else pts @ np.array(nd.consts).reshape(3, 3)
            v, c = walk(nd.children[0], q, lip)
            return v, c + COST[nd.op]
        if nd.kind != 'fold':
            return sdf_ref.evaluate(obj, pts).reshape(-1), np.full(len(pts), float(cost_of(obj)))
        acc, acc_cost = walk(nd.children[0], pts, lip)
        for child in nd.children[1:]:
            K = nd.extra or getattr(child, '_k', None) or 0.0          # dn.py:12
            v, c = walk(child, pts, lip)
            w = -v if nd.op == 'difference' else v
            if nd.op == 'union':
                rhs_dead = (w - lip * R) - (acc + lip * R) >= K
                lhs_dead = (acc - lip * R) - (w + lip * R) >= K
            else:
                rhs_dead = (acc - lip * R) - (w + lip * R) >= K
                lhs_dead = (w - lip * R) - (acc + lip * R) >= K
            acc_cost = np.where(rhs_dead, acc_cost, np.where(lhs_dead, c, acc_cost + c + (12.0 if K else 1.0)))
            acc = fold_value(nd.op, acc, v, K)
        return acc, acc_cost

    full = float(cost_of(f))
    out = dict(config=name, lattice=n, active_tiles=len(active), instructions_per_sample=int(full), seconds_for_the_skip_test=round(time.time() - t0, 1))
    for lip in lipschitz:
        _, pruned = walk(f, C, lip)
        frac = pruned / full
        out['lipschitz_%g' % lip] = dict(mean_fraction_of_work_left=round(float(frac.mean()), 3),
                                         tiles_with_nothing_pruned=round(float((frac > 0.999).mean()), 3),
                                         quartiles=[round(float(q), 3) for q in np.quantile(frac, [0.25, 0.5, 0.75])])
    return out


if __name__ == '__main__':
    import json
    for name in (sys.argv[1:] or ['csg', 'gearlike', 'blobby', 'knurling']):
        print(json.dumps(analyse(name)), flush=True)
