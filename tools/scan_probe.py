#!/usr/bin/env python
"""Times the scan phases (device resident, CUDA events) on the bench's synthetic read set: python tools/scan_probe.py [Mbases]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from telogator2_b200 import tg_kmer  # noqa: E402

mbases = float(sys.argv[1]) if len(sys.argv) > 1 else 400
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cv, reads, tables = bench.load_fixtures()
T = tables['human_hifi']
KL, CANON, ISSUB = T['KMER_LIST'], T['CANONICAL_STRINGS'], T['KMER_ISSUBSTRING']
KLR, CR = [tg_kmer.RC(k) for k in KL], [tg_kmer.RC(k) for k in CANON]
sr = bench.make_scan_workload(reads, int(mbases * 1e6))
pr = tg_kmer.PackedReads(sr)
nb = int(pr.lens.sum())
print('bases', nb, 'padded', pr.total, 'reads', pr.n)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


t_bc = timed(lambda: pr.basecount(CANON, CR))
t_de = timed(lambda: pr.density_counts(KL, KLR, 100))
sl = (np.arange(pr.n), np.maximum(pr.lens - 6000, 0), pr.lens)
t_hi = timed(lambda: pr.hits_device(*sl, KLR, ISSUB))
for nm, t, bpb in (('basecount', t_bc, 0.375), ('density', t_de, 2.375), ('hits', t_hi, 0)):
    print(f'{nm:10s} {t:8.3f} ms  {nb / t / 1e6:9.1f} Gbases/s  {nb * bpb / t / 1e6:8.1f} GB/s')
