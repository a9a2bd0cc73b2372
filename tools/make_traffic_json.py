#!/usr/bin/env python
"""profiles/r2_traffic.json from committed ncu captures: DRAM bytes per launch + kernel time of the dominant kernel.

    python tools/make_traffic_json.py gearlike:exact=gpurun_out/r2d_tile_gearlike.ncu-rep [model:precision=report ...]

bench.py prints `roofline.traffic` from this file only when the capture's kernel time agrees with the time it
measures itself (a stale capture prints null)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda k: (float(vals[hdr.index(k)].replace(',', '')), units[hdr.index(k)])
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    rd, ru = get('dram__bytes_read.sum')
    wr, wu = get('dram__bytes_write.sum')
    t, tu = get('gpu__time_duration.sum')
    t_us = t * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(tu, 1.0)
    return dict(kernel=vals[hdr.index('Kernel Name')], dram_bytes_per_launch=int(rd * scale[ru] + wr * scale[wu]),
                dram_read_bytes=int(rd * scale[ru]), dram_write_bytes=int(wr * scale[wu]), kernel_us=round(t_us, 3),
                grid=vals[hdr.index('launch__grid_size')], block=vals[hdr.index('launch__block_size')],
                capture=os.path.basename(rep))


def main():
    out = {}
    for arg in sys.argv[1:]:
        key, rep = arg.split('=', 1)
        out[key] = read(rep)
    path = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    with open(path, 'w') as fp:
        json.dump(out, fp, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
