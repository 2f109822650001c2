#!/usr/bin/env python
"""Where one small get_dist_matrix_parallel call (D2 shape: N=48, L=3000, R=3) spends its time (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from telogator2_b200 import tg_align  # noqa: E402
from telogator2_b200.dist_engine import DistEngine, Scoring  # noqa: E402

cv, _reads, _tables = bench.load_fixtures()
rs2 = np.random.default_rng(21)
long_cv = [c for c in (cv['human_ont'] + cv['human_hifi']) if len(c) >= 3000]
groups = []
for gi in range(8):
    grp = []
    for k in range(48):
        sq = list(long_cv[(gi * 48 + k) % len(long_cv)][:3000])
        for ppos in rs2.choice(3000, 60, replace=False):
            sq[ppos] = 'CVHTFDA'[rs2.integers(7)]
        grp.append(''.join(sq))
    groups.append(grp)
aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
tg_align.get_dist_matrix_parallel(groups[0], aligner, True, True, 3, rng_seed=12345)
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for it in range(n_rep):
    g = groups[it % len(groups)]
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    eng = DistEngine(Scoring.from_aligner(aligner)); t.append(time.perf_counter())
    st = eng.prepare(g); torch.cuda.synchronize(); t.append(time.perf_counter())
    D = torch.zeros(st['out_elems'], dtype=torch.float32, device='cuda')
    dev = eng._launch_fast(st, True, True, 3, 12345, None, True, d_matrix=D); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    out = D.cpu().numpy(); t.append(time.perf_counter())
    names = ['engine', 'prepare(encode+h2d)', 'enqueue', 'gpu wait', 'd2h']
    print(it, ' '.join(f'{n}={1e3 * (b - a):.2f}ms' for n, a, b in zip(names, t[:-1], t[1:])), f'total={1e3 * (t[-1] - t[0]):.2f}ms')
t0 = time.perf_counter()
for it in range(n_rep):
    tg_align.get_dist_matrix_parallel(groups[it % len(groups)], aligner, True, True, 3, rng_seed=12345)
print('public call: %.2f ms each' % ((time.perf_counter() - t0) / n_rep * 1e3))
# batched: all groups in one pass
groups64 = (groups * 8)[:64]
tg_align.get_dist_matrices_parallel(groups64[:2], aligner, True, True, 3, rng_seed=12345)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tg_align.get_dist_matrices_parallel(groups64, aligner, True, True, 3, rng_seed=12345)
    print('batched 64 matrices: %.1f ms' % ((time.perf_counter() - t0) * 1e3))
