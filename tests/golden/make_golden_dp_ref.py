#!/usr/bin/env python
"""Pin the distance half against the REFERENCE's own code (build container only; needs /root/reference):

    python tests/golden/make_golden_dp_ref.py        ->  tests/golden/dp_ref_golden.npz

source/tg_align.py is imported unmodified (Bio stubbed: Biopython is not installable here) and its tvr_distance /
get_dist_matrix_parallel / quick_compare_tvrs-style calls are run with oracle.ScoreStandInAligner, whose score() is the
oracle's NW restatement.  Everything else -- length adjustment, identity score, early-out, random.seed / shuffle_seq
order, mean, -log ratio, clamps, per-pair seeds, process pool, float32 fill -- is the reference's code.  The matrices must
equal the oracle's own (dp_golden.npz), which the script asserts; so only score() itself remains "parity unpinned".
"""
import gzip
import json
import os
import random
import sys
from unittest import mock

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
for name in ['pywt', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()

from source import tg_align as ref_align   # noqa: E402  (the reference)
import oracle                               # noqa: E402


def main():
    with gzip.open(os.path.join(HERE, 'colorvecs.json.gz'), 'rt') as f:
        cv = json.load(f)['colorvecs']
    old = np.load(os.path.join(HERE, 'dp_golden.npz'))
    S = oracle.scoring_matrix('C', 'tvr')
    out = {}
    cases = (('ont24_tvr_R2', [c[:1000] for c in cv['human_ont']][:24], S, (True, True), True, True, 2),
             ('hifi24_tvr_R2', [c[:1000] for c in cv['human_hifi']][:24], S, (True, True), True, True, 2),
             ('ont12_tvr2000_R3', [c[:2000] for c in cv['human_ont']][:12], S, (True, True), True, True, 3),
             ('ont12_ms_rightfree_R3', [c[:700 + 97 * i] for i, c in enumerate(cv['human_ont'][:12])], None, (True, False), False, False, 3))
    for name, seqs, Sm, gap_bool, adj, mv, R in cases:
        aligner = oracle.ScoreStandInAligner(Sm, gap_bool)
        D = ref_align.get_dist_matrix_parallel(seqs, aligner, adj, mv, R, max_workers=4, rng_seed=12345)
        assert D.dtype == np.dtype('<f4') and D.shape == (len(seqs), len(seqs))
        assert np.array_equal(D, old[name + '|D']), f'{name}: reference code + oracle score != oracle matrix'
        out[name + '|D'] = D
        print(name, 'reference get_dist_matrix_parallel == oracle dist_matrix_c', D.shape)
    # single-pair calls: reseeded, and the quick_compare_tvrs configuration (tg_align.py:379-381: no reseed, the caller's
    # `random` state is consumed)
    ont = [c[:1500] for c in cv['human_ont']][:6]
    al = oracle.ScoreStandInAligner(S, (True, True))
    d1 = [ref_align.tvr_distance(ont[0], ont[k], al, True, True, 3, rng_seed=777 + k) for k in range(1, 6)]
    out['single_reseeded'] = np.array(d1, dtype=np.float64)
    al2 = oracle.ScoreStandInAligner(None, (True, False))
    random.seed(4242)
    d2 = [ref_align.tvr_distance(ont[0], ont[k], al2, adjust_lens=False, min_viable=False) / ref_align.MAX_SEQ_DIST for k in range(1, 6)]
    out['quick_compare_style_seed4242'] = np.array(d2, dtype=np.float64)
    P, P2 = al.params(), al2.params()
    assert d1 == [oracle.tvr_distance_py(ont[0], ont[k], P, True, True, 3, rng_seed=777 + k) for k in range(1, 6)]
    random.seed(4242)
    assert d2 == [oracle.tvr_distance_py(ont[0], ont[k], P2, False, False, 3, rng_seed=None) / 10.0 for k in range(1, 6)]
    np.savez_compressed(os.path.join(HERE, 'dp_ref_golden.npz'), **out)
    print('wrote dp_ref_golden.npz')


if __name__ == '__main__':
    main()
