"""`bench.py --impl reference` (the driver's reference arm) runs on host cores only, so the CPU suite can hold it to
the contract: one JSON line with the keys the driver reads, the UNMODIFIED reference package when baseline/_ref is
installed (`__graft_entry__.build()` installs it where /root/reference exists), a bounded step."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['steps'] == 1 and d['higher_is_better'] is True
    assert d['unit'] == 'Mpoints/s' and d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None
    assert 'gearlike' in d['config']['workload'] and '512^3' in d['config']['workload']
    cb = d['cpu_baseline']
    assert set(('value', 'unit', 'cores', 'kind', 'sample')) <= set(cb) and cb['value'] == d['value'] and cb['cores'] >= 1
    assert cb['kind'] == ('reference' if os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'sdf')) else 'port')
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert d['triangles'] == 895760                       # the oracle's count for the bench workload (tests/models.py lattice)
    assert time.time() - t0 < 600
