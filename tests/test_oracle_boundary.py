"""The boundary-calling oracle (oracle/boundary.py) against (a) the mathematics that defines db4 and an orthogonal periodised
transform, and (b) the reference's own get_telomere_regions / get_terminating_tl outputs on all bundled reads
(tests/golden/boundary_golden.json.gz).  The three PyWavelets functions themselves are PARITY UNPINNED (library absent)."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import boundary as B
from conftest import GOLDEN, regions_digest, sha1_of


def test_db4_taps_satisfy_the_defining_equations():
    # (the literals are the library's 17-digit table values; they satisfy db4's equations to ~1e-12, not to the last bit)
    h = B.DB4_DEC_LO[::-1]                                   # scaling filter (rec_lo)
    assert abs(h.sum() - np.sqrt(2)) < 5e-12
    assert abs((h * h).sum() - 1) < 5e-12
    for m in (1, 2, 3):                                      # orthogonality to even shifts
        assert abs((h[2 * m:] * h[:-2 * m]).sum()) < 5e-12
    k = np.arange(8)
    for p in range(4):                                       # four vanishing moments of the wavelet
        assert abs((((-1.0) ** k) * (k ** p) * h).sum()) < 5e-10
    g = B.DB4_DEC_HI
    assert np.array_equal(g, [(-1) ** (j + 1) * B.DB4_DEC_LO[7 - j] for j in range(8)])     # quadrature mirror
    for m in range(-3, 4):                                   # low- and high-pass are orthogonal at every even shift
        a = np.zeros(24); b = np.zeros(24)
        a[8:16] = B.DB4_DEC_LO; b[8 + 2 * m:16 + 2 * m] = g
        assert abs((a * b).sum()) < 5e-12


@pytest.mark.parametrize('n', [14, 16, 100, 1000, 4096, 18422])
def test_periodised_transform_is_orthogonal(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    c = B.wavedec_per(x)
    assert len(c) == B.dwt_max_level(n) + 1
    y = B.waverec_per(c)
    assert len(y) == n + (n & 1)
    assert np.abs(y[:n] - x).max() < 1e-10                   # perfect reconstruction
    if n % 2 == 0 and all(len(a) % 2 == 0 for a in c[2:]):
        assert abs(sum((a * a).sum() for a in c) - (x * x).sum()) < 1e-9 * (x * x).sum()     # Parseval


def test_max_level_and_lengths():
    assert [B.dwt_max_level(n) for n in (0, 6, 7, 13, 14, 111, 112, 223, 224, 18422)] == [0, 0, 0, 0, 1, 3, 4, 4, 5, 11]
    c = B.wavedec_per(np.arange(225.0))
    assert [len(a) for a in c] == [8, 8, 15, 29, 57, 113]


def test_soft_threshold():
    d = np.array([-3.0, -1.0, 0.0, 0.5, 1.0, 2.0])
    assert np.allclose(B.soft_threshold(d, 1.0), [-2.0, 0.0, 0.0, 0.0, 0.0, 1.0])


def test_wavelet_smooth_early_outs():
    x = np.random.default_rng(1).standard_normal(111)
    assert B.wavelet_smooth(x) is x                          # 4 coefficient arrays < smoothing_level
    x = np.zeros(5000); x[100] = 1
    assert B.wavelet_smooth(x) is x                          # sigma below fail_sigma


def test_regions_against_reference_on_all_bundled_reads(reads_all, kmer_tables, boundary_golden):
    P = boundary_golden['params']
    W, THR, MINS = P['window'], P['pthresh'], P['min_tel_score']
    n = 0
    for key, lst in reads_all.items():
        T = kmer_tables[key]
        KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
        KLR, CR = [oracle.rc(k) for k in KL], [oracle.rc(k) for k in CANON]
        for (name, rdat), G in zip(lst, boundary_golden['sets'][key]):
            _f, ori, c0, c1, _bc = oracle.scan_read(rdat, KL, KLR, CANON, CR, W)
            dp, dq = c0.astype(np.float64) / W, c1.astype(np.float64) / W
            (power, regs) = B.get_telomere_regions(dp, dp, dq, dq, W, THR)
            assert len(power) == G['n_power'], name
            assert sha1_of(np.asarray(power, dtype=np.float64)) == G['power_sha1'], name
            assert len(regs) == G['n_reg'] and regions_digest(regs) == G['regions_sha1'], name
            for pq in 'qp':
                (tl, edge, inter) = B.terminating_tl_of_regions(regs, pq, W, MINS)
                assert [tl, edge, None if inter is None else list(inter)] == G['gtt_' + pq], (name, pq)
            n += 1
    assert n == 296


def test_synthetic_signals_against_reference(boundary_golden):
    P = boundary_golden['params']
    for c in boundary_golden['synthetic']:
        dp, dq = np.array(c['cnt_p']) / P['window'], np.array(c['cnt_q']) / P['window']
        if 'error' in c:
            with pytest.raises(IndexError):
                B.get_telomere_regions(dp, dp, dq, dq, P['window'], P['pthresh'])
            continue
        (power, regs) = B.get_telomere_regions(dp, dp, dq, dq, P['window'], P['pthresh'])
        assert np.array_equal(np.asarray(power), np.array(c['power'])), (c['n'], c['kind'])
        cls = {None: 0, 'p': 1, 'q': 2}
        assert [[r[0], r[1], cls[r[2]]] for r in regs] == c['regions'], (c['n'], c['kind'])
