#!/usr/bin/env python
"""Phase times of tg_tel.get_tel_repeat_comp_parallel on the bench's scan workload (run under gpurun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from telogator2_b200 import tg_kmer, tg_tel  # noqa: E402

mbases = float(sys.argv[1]) if len(sys.argv) > 1 else 400
_cv, reads, tables = bench.load_fixtures()
sr = bench.make_scan_workload(reads, int(mbases * 1e6))
gp = bench.scan_gtrc_params(tables['human_hifi'], tg_kmer.RC)
ard = [(f'read{i}', r, None) for i, r in enumerate(sr)]
tg_tel.get_tel_repeat_comp_parallel(ard[:64], gp)
res = None
for it in range(3):
    res = None
    t0 = time.perf_counter()
    res = tg_tel.get_tel_repeat_comp_parallel(ard, gp)
    dt = time.perf_counter() - t0
    S = tg_tel.LAST_STATS
    print(f'{it}: {dt * 1e3:.0f} ms  {sum(map(len, sr)) / dt / 1e6:.1f} Mbases/s  kept {len(res[0])}')
    print('   phases', {k: round(v * 1e3, 1) for k, v in S['phases'].items()})
    print('   hits  ', {k: round(v * 1e3, 1) for k, v in S['hits_phases'].items()})
