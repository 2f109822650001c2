"""GPU parity of the colour-vector preparation (reference tg_tvr.py:96-142, 449-498, 524-556, 558-592)
against golden vectors produced by the reference's own code (tests/golden/make_golden_tvr.py) and the oracle."""
import gzip
import json
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def G():
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'tvr_prep_golden.json.gz'), 'rt') as f:
        return json.load(f)


@pytest.fixture(scope='module')
def tgt():
    from telogator2_b200 import tg_tvr
    return tg_tvr


def _kmer_dat(S):
    K = len(S['meta_letters'])
    out = []
    for kd in S['kmer_dat']:
        hits = [[] for _ in range(K)]
        for k, s, e in kd['hits']:
            hits[k].append([s, e])
        out.append([hits, kd['tlen'], 0, 'FWD', kd['name'], 60, ''])
    return out


def _meta(S):
    K = len(S['meta_letters'])
    return [[f'k{i}' for i in range(K)], [None] * K, S['meta_letters'], S['meta_flags']]


def test_find_density_boundary_golden(tgt, G):
    groups = {}
    for c in G['fdb']:
        groups.setdefault((tuple(c['letters']), c['win'], c['thresh'], c['dir'], c['start'], c['lowest']), []).append(c)
    for (letters, win, th, d, st, low), cases in groups.items():
        got = tgt.find_density_boundary_batch([c['seq'] for c in cases], list(letters), win, th, d, st, low)
        for c, (l, r) in zip(cases, got):
            assert (l, r) == (c['left'], c['right']), (letters, win, th, d, st, low, len(c['seq']))
    c = G['fdb'][0]
    one = tgt.find_density_boundary(c['seq'], c['letters'], c['win'], c['thresh'], thresh_dir=c['dir'], starting_coord=c['start'],
                                    use_lowest_dens=c['lowest'])
    assert isinstance(one, tuple) and one == (c['left'], c['right'])
    with pytest.raises(SystemExit):
        tgt.find_density_boundary('CCCC', ['C'], 2, 0.5, thresh_dir='sideways')


def test_denoise_colorvec_golden(tgt, G):
    groups = {}
    for c in G['denoise']:
        groups.setdefault((c['replace'], c['min_size'], c['max_gap_fill'], tuple(c['delete']), tuple(c['merge'])), []).append(c)
    for (rep, ms, gf, dele, mrg), cases in groups.items():
        got = tgt.denoise_colorvec_batch([c['v'] for c in cases], rep, ms, gf, list(dele), list(mrg))
        for c, o in zip(cases, got):
            assert o == c['out'], (rep, ms, gf, dele, mrg, c['v'][:60])
    assert tgt.denoise_colorvec('', replace_char='C') == ''
    assert tgt.denoise_colorvec('DDDCCCCCCCCCCCCDDD', replace_char='C', chars_to_delete=['D']) == 'C' * 19
    from telogator2_b200._lib import TgError
    with pytest.raises(TgError):
        tgt.denoise_colorvec('CCDCC', replace_char='C', min_size=65)


@pytest.mark.parametrize('key', ['human_ont', 'human_hifi', 'maize_hifi'])
def test_cluster_tvrs_preparation_matches_reference(tgt, G, key):
    S = G['sets'][key]
    kd, meta = _kmer_dat(S), _meta(S)
    for trunc in (1000, 3000):
        got = tgt.tvr_prep(kd, meta, trunc)
        want = S[f'prep_trunc{trunc}']
        for k in want:
            assert got[k] == want[k], (key, trunc, k)
    assert tgt.quick_get_tvrtel_lens(kd, meta) == S['quick_tvrtel_lens']
    assert tgt.quick_get_tvrtel_lens([], meta) == []


def test_random_vectors_vs_oracle(tgt):
    """Seeded random colour vectors over a small alphabet, full-size lengths: kernels vs the numpy/Python oracle."""
    rng = np.random.default_rng(3)
    seqs = []
    for t in range(40):
        n = int(rng.integers(1, 30000 if t < 4 else 3000))
        runs, tot = [], 0
        while tot < n:
            ln = int(rng.integers(1, 120 if t % 3 else 15))
            runs.append(str(rng.choice(list('ACDEF'), p=[.25, .45, .1, .1, .1])) * ln)
            tot += ln
        seqs.append(''.join(runs)[:n])
    for (letters, win, th, d, st, low) in ((['A', 'F'], 100, 0.125, 'below', 0, False), (['C'], 100, 0.7, 'above', 0, False),
                                           (['A'], 50, 0.5, 'above', 3, True), (['D', 'E'], 255, 0.2, 'below', 0, True)):
        got = tgt.find_density_boundary_batch(seqs, letters, win, th, d, st, low)
        for s, g in zip(seqs, got):
            assert g == oracle.find_density_boundary_py(s, letters, win, th, d, st, low)
        rgot = tgt.find_density_boundary_batch([s[::-1] for s in seqs], letters, win, th, d, st, low)
        for s, g in zip(seqs, rgot):
            assert g == oracle.find_density_boundary_py(s[::-1], letters, win, th, d, st, low)
    for (rep, ms, gf, dele, mrg) in (('C', 10, 50, ['D', 'E'], ['C', 'D', 'F']), ('A', 3, 63, ['C', 'D', 'E', 'F'], ['C']),
                                     ('C', 64, 1, ['A', 'C', 'F'], ['A', 'C', 'D', 'E', 'F'])):
        got = tgt.denoise_colorvec_batch(seqs, rep, ms, gf, dele, mrg)
        for s, g in zip(seqs, got):
            assert g == oracle.denoise_colorvec_py(s, rep, ms, gf, dele, mrg), (rep, ms, gf, len(s))
