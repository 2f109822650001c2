"""CPU: pin the oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and against CPython's `random` (the reference RNG)."""
import random
from itertools import combinations

import numpy as np
import pytest

import oracle
from conftest import hits_digest, hits_to_array, rc_py, scan_cases, sha1_of


def test_scan_oracle_matches_reference_golden(kmer_tables, golden_reads, scan_golden):
    n = 0
    for tag, T, rdat in scan_cases(kmer_tables, golden_reads):
        KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
        KL_REV, CANON_REV = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
        bc = [oracle.base_count(rdat, CANON), oracle.base_count(rdat, CANON_REV)]
        assert bc == scan_golden[tag + '|bc'].tolist(), tag
        ori = oracle.rc(rdat) if bc[0] > bc[1] else rdat
        assert int(bc[0] > bc[1]) == int(scan_golden[tag + '|flip'][0])
        if bc[0] > bc[1]:
            assert ori == rc_py(rdat)
        for sname, kl in (('p', KL), ('q', KL_REV)):
            cnt = oracle.density_counts(ori, kl, 100)
            assert np.array_equal(cnt, scan_golden[tag + '|dens_' + sname].astype(np.int32)), tag
        for sname, s, kl in (('suffix_q', ori[-6000:], KL_REV), ('full_p', ori[:20000], KL), ('full_q', ori[:20000], KL_REV)):
            got = hits_to_array(oracle.nonoverlapping_hits(s, kl, ISSUB))
            assert np.array_equal(got, scan_golden[tag + '|hits_' + sname]), (tag, sname)
        n += 1
    assert n > 60


def test_issubstring_matches_reference(kmer_tables):
    for T in kmer_tables.values():
        assert oracle.kmer_issubstring(T['KMER_LIST']) == T['KMER_ISSUBSTRING']


def test_mt19937_matches_cpython_golden(rng_golden):
    for key, perms in rng_golden.items():
        seed, n = (int(x) for x in key.split('|'))
        base = bytes(range(1, 251)) * (n // 250 + 1)
        s1 = ''.join(chr(33 + (i % 90)) for i in range(n))
        s2 = ''.join(chr(40 + (i % 80)) for i in range(max(n - 1, 1)))
        out = oracle.shuffle_stream(seed, [s1, s2, s1])
        p1, p2, p3 = perms[:n], perms[n:n + max(n - 1, 1)], perms[n + max(n - 1, 1):]
        assert out[0] == ''.join(s1[i] for i in p1), key
        assert out[1] == ''.join(s2[i] for i in p2), key
        assert out[2] == ''.join(s1[i] for i in p3), key
        del base


@pytest.mark.parametrize('seed', [0, 7, 12345, 2 ** 32 + 5, 2 ** 63 + 11])
def test_mt19937_matches_live_cpython(seed):
    rng = np.random.default_rng(seed % 1000)
    seqs = [''.join(rng.choice(list('ACDEFGH'), n)) for n in (1, 17, 1000, 999, 2048, 3)]
    random.seed(seed)
    want = [''.join(random.sample(s, len(s))) for s in seqs]
    assert oracle.shuffle_stream(seed, seqs) == want


def test_mt_base_table():
    t = oracle.mt_base_table()
    assert t[0] == 19650218 and t.dtype == np.uint32
    assert int(t[1]) == (1812433253 * (19650218 ^ (19650218 >> 30)) + 1) & 0xffffffff


def test_scoring_matrix_quirks():
    m = oracle.scoring_matrix('C', 'tvr')
    ix = oracle.ALPHABET.index
    assert m[ix('A'), ix('A')] == -2          # the '= 2.' is overwritten
    assert m[ix('C'), ix('C')] == 0 and m[ix('D'), ix('D')] == 5
    assert m[ix('C'), ix('D')] == -4 and m[ix('A'), ix('D')] == -2 and m[ix('A'), ix('C')] == -2
    assert oracle.scoring_matrix('C', 'consensus')[ix('C'), ix('C')] == 1
    assert np.array_equal(m, m.T)


def test_nw_c_matches_python_full_matrix():
    rng = np.random.default_rng(5)
    letters = list('ACDEFY')
    for trial in range(60):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        a, b = ''.join(rng.choice(letters, n)), ''.join(rng.choice(letters, m))
        for P in (oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True)),
                  oracle.AlignerParams(oracle.scoring_matrix('C', 'consensus'), (True, False)),
                  oracle.AlignerParams(None, (True, False)),
                  oracle.AlignerParams(None, (False, True)),
                  oracle.AlignerParams(None, (False, False))):
            assert oracle.nw_score(a, b, P) == oracle.nw_score_py(a, b, P)


def test_nw_known_answers():
    P = oracle.AlignerParams(None, (True, True))
    assert oracle.nw_score('ACGT', 'ACGT', P) == 20
    assert oracle.nw_score('ACGT', 'AGT', P) == 11          # 3 matches, one internal gap
    assert oracle.nw_score('AAAA', 'TTTT', P) == -16
    Pr = oracle.AlignerParams(None, (True, False))
    assert oracle.nw_score('ACGTTTTT', 'ACG', Pr) == 15      # free right end gap
    assert oracle.nw_score('TTTTTACG', 'ACG', Pr) == -5      # left end gap still costs
    with pytest.raises(ValueError):
        oracle.nw_score('', 'ACG', P)


def test_dist_matrix_c_matches_python_and_golden(colorvecs, dp_golden):
    seqs = [c[:1000] for c in colorvecs['colorvecs']['human_ont']][:24]
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    D, aln, rnd, cells = oracle.dist_matrix_c(seqs, P, True, True, 2, 12345, want_scores=True)
    assert np.array_equal(D, dp_golden['ont24_tvr_R2|D'])
    assert np.array_equal(aln, dp_golden['ont24_tvr_R2|aln'])
    assert np.array_equal(rnd, dp_golden['ont24_tvr_R2|rnd'])
    assert np.array_equal(D, D.T) and (np.diag(D) == 0).all()
    sub = seqs[:5]
    Dpy = oracle.dist_matrix_py(sub, P, True, True, 2, 777)
    Dc = oracle.dist_matrix_c(sub, P, True, True, 2, 777)
    assert np.array_equal(Dpy, Dc)
    # a pair range only fills its own entries
    Dpart = oracle.dist_matrix_c(sub, P, True, True, 2, 777, pair_range=(2, 5))
    pairs = list(combinations(range(5), 2))
    for p, (i, j) in enumerate(pairs):
        assert Dpart[i, j] == (Dc[i, j] if 2 <= p < 5 else 0)


def test_dist_matrix_c_negative_seed_uses_abs_like_cpython(colorvecs):
    seqs = [c[:300] for c in colorvecs['colorvecs']['human_ont']][:5]
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    for seed in (-12345, -3, 2 ** 40 + 7):
        assert np.array_equal(oracle.dist_matrix_py(seqs, P, True, False, 2, seed), oracle.dist_matrix_c(seqs, P, True, False, 2, seed))


def test_tvr_prep_oracle_matches_reference_golden():
    """Colour-vector preparation: the numpy/Python restatement against the reference's own outputs."""
    import gzip
    import json
    import os
    from conftest import ROOT
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'tvr_prep_golden.json.gz'), 'rt') as f:
        G = json.load(f)
    for c in G['fdb']:
        assert oracle.find_density_boundary_py(c['seq'], c['letters'], c['win'], c['thresh'], c['dir'], c['start'], c['lowest']) == \
            (c['left'], c['right'])
    for c in G['denoise']:
        assert oracle.denoise_colorvec_py(c['v'], c['replace'], c['min_size'], c['max_gap_fill'], c['delete'], c['merge']) == c['out']
    for key, S in G['sets'].items():
        meta = [list(range(len(S['meta_letters']))), None, S['meta_letters'], S['meta_flags']]
        for tr in (1000, 3000):
            got = oracle.tvr_prep_py(S['kmer_dat'], meta, tr)
            for k, v in S[f'prep_trunc{tr}'].items():
                assert got[k] == v, (key, tr, k)
        assert oracle.quick_tvrtel_lens_py(S['kmer_dat'], meta) == S['quick_tvrtel_lens']


def test_distance_half_pinned_against_reference_code(colorvecs, dp_golden):
    """tests/golden/dp_ref_golden.npz was produced by the reference's OWN tvr_distance / get_dist_matrix_parallel (process pool,
    random.seed, shuffle_seq, Counter identity score, clamps, float32 fill) with only aligner.score() answered by the oracle's
    NW restatement (make_golden_dp_ref.py).  The oracle's matrices -- the ones the GPU tests compare with -- must equal them."""
    import os
    from conftest import ROOT
    ref = np.load(os.path.join(ROOT, 'tests', 'golden', 'dp_ref_golden.npz'))
    for key in ('ont24_tvr_R2', 'hifi24_tvr_R2', 'ont12_tvr2000_R3', 'ont12_ms_rightfree_R3'):
        assert np.array_equal(ref[key + '|D'], dp_golden[key + '|D']), key
    cv = colorvecs['colorvecs']['human_ont']
    seqs = [c[:1000] for c in cv][:24]
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    assert np.array_equal(oracle.dist_matrix_c(seqs, P, True, True, 2, 12345), ref['ont24_tvr_R2|D'])
    ont = [c[:1500] for c in cv][:6]
    got = [oracle.tvr_distance_py(ont[0], ont[k], P, True, True, 3, rng_seed=777 + k) for k in range(1, 6)]
    assert got == list(ref['single_reseeded'])
    import random
    P2 = oracle.AlignerParams(None, (True, False))
    random.seed(4242)
    got2 = [oracle.tvr_distance_py(ont[0], ont[k], P2, False, False, 3, rng_seed=None) / 10.0 for k in range(1, 6)]
    assert got2 == list(ref['quick_compare_style_seed4242'])


def test_scan_batch_flow_pinned_against_reference_driver(kmer_tables, golden_reads):
    """tests/golden/tel_batch_golden.json.gz is the output of the reference's own get_tel_repeat_comp_parallel (process pool +
    gtrc_parallel_job, tg_tel.py:238-389; make_golden_tel.py).  The per-read restatement the GPU test compares with must equal it."""
    from test_gpu_tel_batch import _tel_inputs, digest_tel_result, oracle_flow, reference_golden, tel_cases
    gold = reference_golden()
    for tkey, filters in tel_cases():
        reads, params = _tel_inputs(kmer_tables, golden_reads, tkey, filters)
        assert digest_tel_result(oracle_flow(reads, params)) == gold[tkey], tkey


def test_scan_oracle_matches_reference_on_all_bundled_reads(kmer_tables, reads_all, scan_all_golden):
    """The whole of BASELINE configs 1-3 (83 / 200 / 13 reads): the oracle's scan functions against what the reference's own
    tg_kmer functions computed (make_golden_all.py; digests of the density counts and of the hit spans)."""
    n = 0
    for key, rs in reads_all.items():
        T = kmer_tables[key]
        KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
        KLR, CR = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
        assert len(rs) == len(scan_all_golden[key]) == {'human_ont': 83, 'human_hifi': 200, 'maize_hifi': 13}[key]
        for (name, rdat), g in zip(rs, scan_all_golden[key]):
            flip, ori, c0, c1, bc = oracle.scan_read(rdat, KL, KLR, CANON, CR, 100)
            assert bc.tolist() == g['bc'] and bool(flip) == g['flip'], name
            assert sha1_of(c0.astype(np.uint8)) == g['dens_p'] and sha1_of(c1.astype(np.uint8)) == g['dens_q'], name
            assert hits_digest(oracle.nonoverlapping_hits(ori[-6000:], KLR, ISSUB)) == (g['hits'], g['n_spans']), name
            n += 1
    assert n == 296


def _nw_by_enumeration(a, b, P):
    """Needleman-Wunsch score from the DEFINITION in Biopython's documentation, not from a recurrence: the maximum over ALL
    global alignments of the sum of the column scores, where a gap column costs the left end-gap score if the gapped
    sequence has no letter to its left, the right end-gap score if it has none to its right, the internal gap score
    otherwise (open == extend at every call site that reaches score(), tg_align.py:71-106)."""
    n, m = len(a), len(b)

    def sub(x, y):
        if P.S is not None:
            return P.S[oracle.ALPHABET.index(x), oracle.ALPHABET.index(y)]
        return P.match if x == y else P.mismatch
    best = [None]

    def walk(i, j, score):
        if i == n and j == m:
            if best[0] is None or score > best[0]:
                best[0] = score
            return
        if i < n and j < m:
            walk(i + 1, j + 1, score + sub(a[i], b[j]))
        if i < n:                                   # a[i] against a gap in b: b has consumed j of its m letters
            walk(i + 1, j, score + (P.gl if j == 0 else P.gr if j == m else P.g))
        if j < m:                                   # b[j] against a gap in a
            walk(i, j + 1, score + (P.gl if i == 0 else P.gr if i == n else P.g))
    walk(0, 0, 0.0)
    return best[0]


def test_nw_restatement_equals_exhaustive_enumeration_of_alignments():
    """Second, independent restatement of the one unpinned piece (Bio.Align.PairwiseAligner.score): brute force over every
    alignment of short sequences, all four end-gap configurations, both scoring modes."""
    rng = np.random.default_rng(11)
    letters = list('ACDEFY')
    S = oracle.scoring_matrix('C', 'tvr')
    n_cases = 0
    for trial in range(40):
        n, m = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        a, b = ''.join(rng.choice(letters, n)), ''.join(rng.choice(letters, m))
        for gb in ((True, True), (True, False), (False, True), (False, False)):
            for Sm in (S, None):
                P = oracle.AlignerParams(Sm, gb)
                assert oracle.nw_score(a, b, P) == _nw_by_enumeration(a, b, P), (a, b, gb, Sm is None)
                n_cases += 1
    assert n_cases == 320


def test_nw_properties():
    """Size-independent properties any global alignment score must have (the score stays unpinned against Biopython itself)."""
    rng = np.random.default_rng(12)
    letters = list('ACDEFHTVY')
    S = oracle.scoring_matrix('C', 'tvr')
    for trial in range(30):
        n, m = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        a, b = ''.join(rng.choice(letters, n)), ''.join(rng.choice(letters, m))
        for gb in ((True, True), (True, False)):
            for Sm in (S, None):
                P = oracle.AlignerParams(Sm, gb)
                s_ab = oracle.nw_score(a, b, P)
                assert s_ab == oracle.nw_score(b, a, P)                                   # symmetric scoring -> symmetric score
                assert s_ab == oracle.nw_score(a[::-1], b[::-1], oracle.AlignerParams(Sm, gb[::-1]))      # mirror image
                assert s_ab <= oracle.iden_score(a, a, P) / 2 + oracle.iden_score(b, b, P) / 2 or Sm is not None
                assert s_ab >= P.gl * 0 + (-4.0) * (n + m)                                # all-gap alignment is a lower bound
                assert oracle.nw_score(a, b, oracle.AlignerParams(Sm, (gb[0], False))) >= s_ab     # freeing end gaps never hurts
        P = oracle.AlignerParams(None, (True, True))
        assert oracle.nw_score(a, a, P) == 5.0 * n
