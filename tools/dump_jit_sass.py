"""SASS of a run-time specialised kernel, for profiles/: compiles the kernel of one BASELINE configuration with
NVRTC into an empty cache directory (no GPU needed) and prints resource usage, an instruction histogram and the
listing (`cuobjdump -sass`).

  python tools/dump_jit_sass.py gearlike 4 > profiles/r2_sass_tile_sign_gearlike.txt
     (kernel numbers: csrc/sdf_jit_entry.cuh -- 4 = tile_sign_body, the dominant kernel of precision `exact`)
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    config, kernel = sys.argv[1], int(sys.argv[2])
    cache = tempfile.mkdtemp(prefix='jitsass')
    os.environ['SDFB200_JIT_CACHE'] = cache
    import models
    import sdf_b200
    from sdf_b200 import _abi
    from sdf_b200.ir import lower
    h = _abi.ProgramHandle(lower(models.CONFIGS[config](sdf_b200)))
    share, n, saved = h.jit_plan()
    h.jit_probe(kernel)
    cubins = [os.path.join(cache, f) for f in os.listdir(cache) if f.endswith('.cubin')]
    cubin = max(cubins, key=os.path.getmtime)
    res = subprocess.run(['cuobjdump', '-res-usage', cubin], capture_output=True, text=True).stdout
    sass = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True).stdout
    hist = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m:
            hist[m.group(1)] += 1
    print('# %s, kernel %d of csrc/sdf_jit_entry.cuh, specialised with NVRTC for sm_100a (tools/dump_jit_sass.py)' % (config, kernel))
    print('# sample grouping chosen by the host: share axis %d, %d samples per thread-run (%.0f %% of the program shared)'
          % (share, n, 100 * saved))
    print('#', ' '.join(res.split()))
    print('# instruction histogram (static):', ', '.join('%s %d' % kv for kv in hist.most_common()))
    print(sass)


if __name__ == '__main__':
    main()
