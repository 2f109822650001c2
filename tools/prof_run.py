#!/usr/bin/env python
"""Small, fixed workload for ncu captures: one pass of every hot kernel (run under gpurun)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from telogator2_b200 import tg_align, tg_kmer  # noqa: E402
from telogator2_b200.dist_engine import DistEngine, Scoring  # noqa: E402

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mbases = float(sys.argv[2]) if len(sys.argv) > 2 else 20
cv, reads, tables = bench.load_fixtures()
seqs = bench.make_dp_workload(cv, n_seq)
aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
eng = DistEngine(Scoring.from_aligner(aligner))
st = eng.prepare(seqs)
for _ in range(2):
    D = eng.matrix_device(st, True, True, 2, 12345)
torch.cuda.synchronize()
print('dp cells', eng.stats['cells'], 'early', eng.stats['early_out'], 'pairs', eng.stats['pairs'])
from telogator2_b200 import tg_linkage  # noqa: E402
for _ in range(2):
    Z = tg_linkage.linkage_from_matrix(D, 'ward', divide_by=10.0)
print('linkage n', D.shape[0], 'merges', len(Z))
T = tables['human_hifi']
KL, CANON, ISSUB = T['KMER_LIST'], T['CANONICAL_STRINGS'], T['KMER_ISSUBSTRING']
KLR, CR = [tg_kmer.RC(k) for k in KL], [tg_kmer.RC(k) for k in CANON]
sr = bench.make_scan_workload(reads, int(mbases * 1e6))
pr = tg_kmer.PackedReads(sr)
for _ in range(2):
    bc = pr.basecount(CANON, CR)
    ca, cb = pr.density_counts(KL, KLR, 100)
    sp = pr.hits_device(np.arange(pr.n), np.maximum(pr.lens - 6000, 0), pr.lens, KLR, ISSUB)
torch.cuda.synchronize()
print('scan bases', int(pr.lens.sum()), 'reads', pr.n, 'spans', sp.shape[1])
from telogator2_b200 import tg_boundary  # noqa: E402
for _ in range(2):
    term = tg_boundary.terminating_tl_from_counts(ca, cb, pr.d_base_off, pr.lens, 100, 0.5, 100, 'q')
torch.cuda.synchronize()
print('boundary reads', len(term), 'with telomere', int((term[:, 0] > 0).sum()))
from telogator2_b200 import tg_tvr  # noqa: E402
kd, meta, _flat = bench.make_prep_workload(512)
for _ in range(2):
    out = tg_tvr.tvr_prep(kd, meta, 1000)
torch.cuda.synchronize()
print('prep reads', len(kd), 'letters', sum(k[1] for k in kd))
from telogator2_b200 import tg_prefilter  # noqa: E402
wg = [''.join(np.random.default_rng(5).choice(list('ACGT'), 16384)) for _ in range(64)] * 256
ar = tg_prefilter.AsciiReads(wg + sr[:64])
for _ in range(2):
    ca, cb = ar.count_pair_device('CCCTAACCCTAA', 'TTAGGGTTAGGG')
torch.cuda.synchronize()
print('prefilter reads', ar.n, 'bases', int(ar.lens.astype(np.int64).sum()))
