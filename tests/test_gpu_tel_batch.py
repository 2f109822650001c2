"""GPU: the batch boundary tg_tel.get_tel_repeat_comp_parallel vs a per-read restatement of
gtrc_parallel_job (tg_tel.py:238-328) built on the CPU oracle.  get_terminating_tl (pywt code, outside the
hot path) is replaced on BOTH sides by the same simple stand-in driven by the density arrays."""
import numpy as np
import pytest

import oracle
from conftest import rc_py

pytestmark = pytest.mark.gpu


def make_gtt(density_fn):
    def gtt(rdat, pq, gtt_params, telplot_dat=None):
        [KL, KLR, W, thr, min_score] = gtt_params
        (p0, p1) = density_fn(rdat, KL, W)
        (q0, q1) = density_fn(rdat, KLR, W)
        assert len(p0) == len(p1) == len(q0) == len(q1) == max(len(rdat) - W, 0)
        power = np.array(p0) - np.array(q0)
        tl = 0
        for v in power[::-1]:                      # trailing run of q-type windows
            if v <= -thr:
                tl += 1
            else:
                break
        tl = tl + W // 2 if tl >= min_score else 0
        nontel = 0 if tl else int(np.argmax(power[::-1] <= -thr)) if (power <= -thr).any() else len(power)
        inter = None
        idx = np.nonzero(np.abs(power) >= thr)[0]
        if tl == 0 and len(idx):
            inter = (int(idx[0]) + W // 2, int(idx[-1]) + W // 2)
        return (tl, nontel, inter)
    return gtt


def oracle_flow(all_read_dat, gtrc_params):
    [READ_TYPE, MAPQ, KL, KLR, ISSUB, CANON, CANON_REV, W, THR, MIN_TEL_SCORE, F_TEL, F_NONTEL, F_SUB, PLOT, SIGDIR,
     MIN_ITL, MIN_ANCHOR] = gtrc_params

    def dens(r, kl, w):
        c = oracle.density_counts(r, kl, w)
        return ((c.astype(np.float64) / w).tolist(), [0.0] * len(c))
    gtt = make_gtt(dens)
    out = [[], [], [], 0, 0, 0, [], {}, {}]
    for (rnm, rdat, qdat) in all_read_dat:
        f, r = oracle.base_count(rdat, CANON), oracle.base_count(rdat, CANON_REV)
        if f > r:
            rdat = rc_py(rdat)
        gmin = min(F_TEL, MIN_TEL_SCORE) if F_TEL > 0 else MIN_TEL_SCORE
        (tl, nontel, inter) = gtt(rdat, 'q', [KL, KLR, W, THR, gmin])
        tl, nontel = min(tl, len(rdat)), min(nontel, len(rdat))
        if tl == 0 and inter is not None and inter[1] - inter[0] >= MIN_ITL:
            left, right = rdat[:inter[0]], rdat[inter[1]:]
            if len(left) >= MIN_ANCHOR and len(right) >= MIN_ANCHOR:
                out[6].append([(f'interstitial-left_{len(rdat)}_{rnm}', left), (f'interstitial-right_{len(rdat)}_{rnm}', right)])
                out[7][rnm] = [oracle.nonoverlapping_hits(rdat, KL, ISSUB), len(rdat), 0, 'q', rnm.split(' ')[0], MAPQ, rdat]
                out[8][rnm] = [oracle.nonoverlapping_hits(rdat, KLR, ISSUB), len(rdat), 0, 'q', rnm.split(' ')[0], MAPQ, rdat]
        s_tel = F_TEL > 0 and tl < F_TEL
        s_unk = F_NONTEL > 0 and nontel > F_NONTEL
        s_sub = F_SUB > 0 and len(rdat) < tl + F_SUB
        if not (s_tel or s_unk or s_sub):
            seq = rdat[max(len(rdat) - tl - F_SUB, 0):]
            if len(seq) == 0:
                seq = rdat
            out[0].append([oracle.nonoverlapping_hits(seq, KLR, ISSUB), len(seq), 0, 'q', rnm.split(' ')[0], MAPQ, rdat])
            out[1].append(tl)
            out[2].append(nontel)
        out[3] += s_tel * 1
        out[4] += s_unk * 1
        out[5] += s_sub * 1
    return out


def tel_cases():
    return [('human_hifi', (400, 100, 1000)), ('human_ont', (0, 0, 0)), ('maize_hifi', (400, 100, 1000))]


def digest_tel_result(res):
    """JSON-able form of the 9-element result with read sequences replaced by SHA-1 digests (golden fixture format)."""
    import hashlib

    def h(s):
        return hashlib.sha1(s.encode('latin-1')).hexdigest()

    def entry(e):
        return [[[list(map(int, sp)) for sp in k] for k in e[0]], int(e[1]), int(e[2]), e[3], e[4], int(e[5]), h(e[6])]
    return [[entry(e) for e in res[0]], [int(x) for x in res[1]], [int(x) for x in res[2]], int(res[3]), int(res[4]), int(res[5]),
            [[[nm, h(seq)] for (nm, seq) in pair] for pair in res[6]],
            {rn: entry(e) for rn, e in res[7].items()}, {rn: entry(e) for rn, e in res[8].items()}]


def _tel_inputs(kmer_tables, golden_reads, tkey, filters):
    T = kmer_tables[tkey]
    KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
    KLR, CANON_REV = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
    reads = [(name + ' extra', seq, None) for name, seq in golden_reads[tkey]]
    reads += [(n, s, None) for n, s in golden_reads['synthetic'] if len(s) > 0 and set(s) <= set('ACGTN')]
    params = [T['read_type'], 60, KL, KLR, ISSUB, CANON, CANON_REV, 100, 0.5, 100, filters[0], filters[1], filters[2],
              False, '', 100, 1000]
    return reads, params


def reference_golden():
    import gzip
    import json
    import os
    from conftest import ROOT
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'tel_batch_golden.json.gz'), 'rt') as f:
        return json.load(f)


@pytest.mark.parametrize('tkey,filters', tel_cases())
def test_batch_boundary_matches_reference_driver(kmer_tables, golden_reads, tkey, filters):
    """GPU batch boundary vs the output of the reference's own get_tel_repeat_comp_parallel (tests/golden/make_golden_tel.py)."""
    from telogator2_b200 import tg_kmer, tg_tel
    reads, params = _tel_inputs(kmer_tables, golden_reads, tkey, filters)
    got = tg_tel.get_tel_repeat_comp_parallel(reads, params, max_workers=7, get_terminating_tl=make_gtt(tg_kmer.get_telomere_kmer_density))
    assert digest_tel_result(got) == reference_golden()[tkey]


@pytest.mark.parametrize('tkey,filters', tel_cases())
def test_batch_boundary_matches_per_read_flow(kmer_tables, golden_reads, tkey, filters):
    from telogator2_b200 import tg_kmer, tg_tel
    T = kmer_tables[tkey]
    KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
    KLR, CANON_REV = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
    reads = [(name + ' extra', seq, None) for name, seq in golden_reads[tkey]]
    reads += [(n, s, None) for n, s in golden_reads['synthetic'] if len(s) > 0 and set(s) <= set('ACGTN')]
    params = [T['read_type'], 60, KL, KLR, ISSUB, CANON, CANON_REV, 100, 0.5, 100, filters[0], filters[1], filters[2],
              False, '', 100, 1000]
    got = tg_tel.get_tel_repeat_comp_parallel(reads, params, max_workers=7, get_terminating_tl=make_gtt(tg_kmer.get_telomere_kmer_density))
    want = oracle_flow(reads, params)
    assert len(got) == 9
    assert got[1] == want[1] and got[2] == want[2] and got[3:6] == want[3:6]
    assert len(got[0]) == len(want[0]) > 0
    for g, w in zip(got[0], want[0]):
        assert g[1:] == w[1:]
        assert g[0] == w[0]
    assert got[6] == want[6]
    assert got[7] == want[7] and got[8] == want[8]


def test_all_bundled_reads_through_the_batch_boundary(kmer_tables, reads_all):
    """get_tel_repeat_comp_parallel on ALL bundled reads of every set (83 / 200 / 13) against the per-read oracle flow."""
    from telogator2_b200 import tg_kmer, tg_tel
    for key, filt in tel_cases():
        T = kmer_tables[key]
        KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
        params = [T['read_type'], 60, KL, [rc_py(k) for k in KL], T['KMER_ISSUBSTRING'], CANON, [rc_py(k) for k in CANON], 100, 0.5, 100,
                  filt[0], filt[1], filt[2], False, '', 100, 1000]
        ard = [(name, r, None) for name, r in reads_all[key]]
        got = tg_tel.get_tel_repeat_comp_parallel(ard, params, get_terminating_tl=make_gtt(tg_kmer.get_telomere_kmer_density))
        want = oracle_flow(ard, params)
        assert got[1:6] == want[1:6], key
        assert len(got[0]) == len(want[0]) and all(a == b for a, b in zip(got[0], want[0])), key
        assert got[6] == want[6] and got[7] == want[7] and got[8] == want[8], key
