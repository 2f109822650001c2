"""Host-side phase timing of get_dist_matrix_parallel's body (one GPU): where the e2e - device gap goes."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from telogator2_b200 import tg_align  # noqa: E402
from telogator2_b200.dist_engine import DistEngine, Scoring  # noqa: E402

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
shard = (0, int(sys.argv[2])) if len(sys.argv) > 2 else None
cv, _reads, _tables = bench.load_fixtures()
seqs = bench.make_dp_workload(cv, n_seq)
aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
for it in range(3):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    eng = DistEngine(Scoring.from_aligner(aligner)); t.append(time.perf_counter())
    st = eng.prepare(seqs); torch.cuda.synchronize(); t.append(time.perf_counter())
    D = torch.zeros((n_seq, n_seq), dtype=torch.float32, device='cuda')
    dev = eng._launch_fast(st, True, True, 2, 12345, shard, True, d_matrix=D); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    out = D.cpu().numpy(); t.append(time.perf_counter())
    names = ['engine', 'prepare(encode+h2d)', 'enqueue', 'gpu wait', 'd2h']
    print(it, ' '.join(f'{n}={1e3 * (b - a):.1f}ms' for n, a, b in zip(names, t[:-1], t[1:])), f'total={1e3 * (t[-1] - t[0]):.1f}ms')
