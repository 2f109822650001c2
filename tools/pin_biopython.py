#!/usr/bin/env python
"""Pin Bio.Align.PairwiseAligner.score (the one third-party arithmetic of the DP half, tg_align.py:125,142) to REAL Biopython.

    python tools/pin_biopython.py [--install] [--out DIR]

Runs anywhere Biopython >= 1.82 is importable (with --install it first tries `pip install "biopython>=1.86" pywavelets`,
which needs an index; the build image and the GPU boxes have none -- the attempt and its output are logged to
DIR/pin_biopython.log so the absence is on record).  It needs neither /root/reference nor a GPU:

  * the aligner is configured through telogator2_b200.tg_align.get_scoring_matrix / get_aligner_object, which build real
    Bio.Align objects when Biopython is importable (the same attribute assignments as tg_align.py:32-106);
  * scores come from aligner.score() -- Biopython's C code -- on the committed colour vectors (tests/golden/colorvecs.json.gz):
    the FULL iteration-0 matrices D0 (N = 83 / 101 / 13, truncate 1000, R = 2, rng_seed 12345) through a literal
    restatement of tg_align.py:109-189 around aligner.score(), plus raw-score cases: ragged shapes, free right end gaps,
    match/mismatch mode, sequences beyond the 16-bit range;
  * output: DIR/dp_biopython_golden.npz (move it to tests/golden/; tests/test_gpu_dp.py and tests/test_oracle_golden.py
    pick it up when present and compare the CUDA matrices / the oracle with it).
"""
import argparse
import gzip
import json
import os
import random
import subprocess
import sys
import time
from itertools import combinations

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def try_install(log):
    cmd = [sys.executable, '-m', 'pip', 'install', '--disable-pip-version-check', '--retries', '1', '--timeout', '10',
           'biopython>=1.86', 'pywavelets']
    log.write('$ ' + ' '.join(cmd) + '\n')
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
        log.write(r.stdout[-4000:] + r.stderr[-4000:] + f'\nrc={r.returncode}\n')
        return r.returncode == 0
    except subprocess.TimeoutExpired:
        log.write('pip install timed out after 180 s\n')
        return False


def tvr_distance(tvr_i, tvr_j, aligner, adjust_lens, min_viable, randshuffle, rng_seed, S_diag):
    """tg_align.py:109-151 around the real aligner.score()."""
    random.seed(rng_seed)
    if adjust_lens:
        ml = min(len(tvr_i), len(tvr_j))
        if min_viable:
            ml = max(ml, 1000)
        si, sj = tvr_i[:ml], tvr_j[:ml]
    else:
        si, sj = tvr_i, tvr_j
    aln = aligner.score(si, sj)
    if S_diag is None:
        iden = 5. * ((len(si) + len(sj)) / 2.)
    else:
        iden = (sum(S_diag[c] for c in si) + sum(S_diag[c] for c in sj)) / 2.
    if aln < -(0.900 * iden):
        return 10.0, aln, []
    rs = []
    for _ in range(randshuffle):
        a = ''.join(random.sample(si, len(si)))
        b = ''.join(random.sample(sj, len(sj)))
        rs.append(aligner.score(a, b))
    r = np.mean(rs)
    d = 10.0 if r >= aln else min(-np.log((aln - r) / (iden - r)), 10.0)
    return max(d, 0.0001), aln, rs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--install', action='store_true')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out'))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    log = open(os.path.join(a.out, 'pin_biopython.log'), 'w')
    log.write(time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime()) + ' host ' + os.uname().nodename + '\n')
    try:
        import Bio
    except ImportError:
        Bio = None
        log.write('import Bio: ModuleNotFoundError\n')
        if a.install and try_install(log):
            import Bio
    if Bio is None:
        log.write('RESULT: Biopython unavailable; aligner.score() stays PARITY UNPINNED\n')
        log.close()
        print(open(os.path.join(a.out, 'pin_biopython.log')).read())
        return 3
    log.write(f'Biopython {Bio.__version__}\n')
    from telogator2_b200 import tg_align
    assert tg_align._BioAligner is not None
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'colorvecs.json.gz'), 'rt') as f:
        cv = json.load(f)['colorvecs']
    out = {'biopython_version': np.array(Bio.__version__)}
    canon = {'human_ont': 'C', 'human_hifi': 'C', 'maize_hifi': 'C'}
    for name, vecs in cv.items():
        S = tg_align.get_scoring_matrix(canon[name], 'tvr')
        al = tg_align.get_aligner_object(S, (True, True), 'tvr')
        diag = {ch: float(S[ch, ch]) for ch in tg_align.AMINO}
        seqs = [c[:1000] for c in vecs]
        n = len(seqs)
        D = np.zeros((n, n), dtype='<f4')
        alns, rnds = [], []
        t0 = time.time()
        for p, (i, j) in enumerate(combinations(range(n), 2)):
            d, aln, rs = tvr_distance(seqs[i], seqs[j], al, True, True, 2, 12345 + p, diag)
            D[i, j] = D[j, i] = d
            alns.append(aln)
            rnds.append(rs + [np.nan] * (2 - len(rs)))
        out[f'D0_{name}|D'] = D
        out[f'D0_{name}|aln'] = np.array(alns)
        out[f'D0_{name}|rnd'] = np.array(rnds)
        log.write(f'D0 {name}: n={n} {time.time() - t0:.1f} s\n')
    # raw scores: every gap_bool, both scoring modes, ragged shapes, > 16-bit range
    rng = np.random.default_rng(2026)
    pool = cv['human_ont'] + cv['human_hifi']
    raw_a, raw_b, raw_cfg, raw_s = [], [], [], []
    for k in range(160):
        x, y = pool[rng.integers(len(pool))], pool[rng.integers(len(pool))]
        lx, ly = int(rng.integers(1, 1 + min(len(x), 2500))), int(rng.integers(1, 1 + min(len(y), 2500)))
        if k >= 150:                                  # long: beyond the 16-bit cell range
            x, y = (x * 4)[:9000 + k], (y * 4)[:8000 + 3 * k]
            lx, ly = len(x), len(y)
        x, y = x[:lx], y[:ly]
        for gb in ((True, True), (True, False), (False, True), (False, False)):
            for mode in ('matrix', 'mm'):
                S = tg_align.get_scoring_matrix('C', 'tvr') if mode == 'matrix' else None
                al = tg_align.get_aligner_object(S, gb, 'tvr')
                raw_a.append(x)
                raw_b.append(y)
                raw_cfg.append([int(gb[0]), int(gb[1]), int(mode == 'matrix')])
                raw_s.append(al.score(x, y))
    out['raw|a'] = np.array(raw_a)
    out['raw|b'] = np.array(raw_b)
    out['raw|cfg'] = np.array(raw_cfg, dtype=np.int8)
    out['raw|score'] = np.array(raw_s)
    np.savez_compressed(os.path.join(a.out, 'dp_biopython_golden.npz'), **out)
    log.write('RESULT: wrote dp_biopython_golden.npz\n')
    log.close()
    print(open(os.path.join(a.out, 'pin_biopython.log')).read())
    return 0


if __name__ == '__main__':
    sys.exit(main())
