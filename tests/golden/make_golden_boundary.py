#!/usr/bin/env python
"""Telomere-boundary calling of ALL bundled reads through the REFERENCE's own code (build container only):

    python tests/golden/make_golden_boundary.py      ->  tests/golden/boundary_golden.json.gz

For every bundled read (tests/golden/reads_all.json.gz), oriented as tg_tel.py:262-266 does, this runs the reference's
unmodified source/tg_kmer.get_telomere_regions and source/tg_tel.get_terminating_tl (imported from /root/reference).
PyWavelets is absent from the image, so `pywt` is `oracle.boundary.fake_pywt()`: a restatement of wavedec / waverec / threshold
for db4 + periodisation (PARITY UNPINNED against the real library, see oracle/boundary.py).  Everything else -- the power
signal, the region walk, the scoring and the interstitial clustering -- is the reference's own Python and is what these
vectors pin.  Per read:
  n_power, power_sha1   length and SHA-1 of the float64 p_vs_q_power the reference returns
  regions_sha1, n_reg   SHA-1 of the int32 (start, end, class) triples (class 0 None / 1 'p' / 2 'q')
  gtt_q, gtt_p          get_terminating_tl(read, 'q' | 'p', gtt_params) -> [tel_len, edge_nontel, interstitial or None]
plus a few synthetic signals (short, odd/even lengths around every early-out of wavelet_smooth) with their full outputs.
"""
import gzip
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor
from unittest import mock

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import boundary  # noqa: E402

sys.path.insert(0, REF)
for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()
sys.modules['pywt'] = boundary.fake_pywt()

from source import tg_kmer, tg_tel, tg_util  # noqa: E402  (the reference)

W, THR, MIN_SCORE = 100, 0.5, 100            # telogator2.py defaults: TEL_WINDOW_SIZE, P_VS_Q_AMP_THRESH, MIN_TEL_SCORE
CLS = {None: 0, 'p': 1, 'q': 2}


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def regions_digest(regs):
    return sha(np.array([(r[0], r[1], CLS[r[2]]) for r in regs], dtype=np.int32).reshape(-1, 3))


def one(job):
    (rdat, KL, CANON) = job
    KLR, CR = [tg_util.RC(k) for k in KL], [tg_util.RC(k) for k in CANON]
    flip = tg_kmer.get_telomere_base_count(rdat, CANON) > tg_kmer.get_telomere_base_count(rdat, CR)
    ori = tg_util.RC(rdat) if flip else rdat
    (p0, p1) = tg_kmer.get_telomere_kmer_density(ori, KL, W)
    (q0, q1) = tg_kmer.get_telomere_kmer_density(ori, KLR, W)
    (power, regs) = tg_kmer.get_telomere_regions(p0, p1, q0, q1, W, THR)
    out = {'n_power': int(len(power)), 'power_sha1': sha(np.asarray(power, dtype=np.float64)), 'n_reg': len(regs),
           'regions_sha1': regions_digest(regs)}
    for pq in 'qp':
        (tl, edge, inter) = tg_tel.get_terminating_tl(ori, pq, [KL, KLR, W, THR, MIN_SCORE])
        out['gtt_' + pq] = [int(tl), int(edge), None if inter is None else [int(inter[0]), int(inter[1])]]
    return out


def synthetic_cases():
    """Signals around the early-outs (tg_kmer.py:158-165) and the length quirks, with full outputs."""
    rng = np.random.default_rng(20260)
    cases = []
    for n in [100, 101, 111, 112, 113, 127, 128, 223, 224, 225, 257, 500, 501, 1023, 1024, 1025, 3001]:
        for kind in ['noise', 'step', 'flat']:
            if kind == 'noise':
                cp, cq = rng.integers(0, 101, n), rng.integers(0, 101, n)
            elif kind == 'step':
                cp = np.where(np.arange(n) < n // 3, 90, 3) + rng.integers(0, 8, n)
                cq = np.where(np.arange(n) > 2 * n // 3, 95, 1) + rng.integers(0, 5, n)
            else:
                cp, cq = np.full(n, 37), np.full(n, 37)
                cp[n // 2] = 38
            dp, dq = (cp / W).tolist(), (cq / W).tolist()
            try:
                (power, regs) = tg_kmer.get_telomere_regions(dp, [0.0] * n, dq, [0.0] * n, W, THR)
                res = {'power': np.asarray(power, dtype=np.float64).tolist(), 'regions': [[r[0], r[1], CLS[r[2]]] for r in regs]}
            except IndexError:
                res = {'error': 'IndexError'}
            cases.append({'n': n, 'kind': kind, 'cnt_p': cp.tolist(), 'cnt_q': cq.tolist(), **res})
    return cases


def main():
    tables = json.load(open(os.path.join(HERE, 'kmer_tables.json')))
    with gzip.open(os.path.join(HERE, 'reads_all.json.gz'), 'rt') as f:
        reads = json.load(f)
    out = {'params': {'window': W, 'pthresh': THR, 'min_tel_score': MIN_SCORE}, 'sets': {}}
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        for key, lst in reads.items():
            T = tables[key]
            jobs = [(r[1], T['KMER_LIST'], T['CANONICAL_STRINGS']) for r in lst]
            out['sets'][key] = list(ex.map(one, jobs, chunksize=2))
            n_tel = sum(1 for e in out['sets'][key] if e['gtt_q'][0] > 0)
            print(key, len(jobs), 'reads;', n_tel, 'with a terminal telomere (q);',
                  sum(1 for e in out['sets'][key] if e['gtt_q'][2] is not None), 'with an interstitial region')
    out['synthetic'] = synthetic_cases()
    with gzip.open(os.path.join(HERE, 'boundary_golden.json.gz'), 'wt') as f:
        json.dump(out, f)
    print('wrote boundary_golden.json.gz', os.path.getsize(os.path.join(HERE, 'boundary_golden.json.gz')), 'bytes')


if __name__ == '__main__':
    main()
