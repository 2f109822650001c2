"""GPU parity: distance path (tg_align) through the C ABI vs the CPU oracle / golden fixtures."""
import random
from itertools import combinations

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tga():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from telogator2_b200 import tg_align
    return tg_align


def _P(S, gap_bool):
    return oracle.AlignerParams(S, gap_bool)


def _cases(colorvecs):
    ont = colorvecs['colorvecs']['human_ont']
    hifi = colorvecs['colorvecs']['human_hifi']
    return ont, hifi


def _check_matrix(tga, seqs, which, gap_bool, adjust, minv, R, seed, force_generic=False):
    from telogator2_b200.dist_engine import DistEngine, Scoring
    if which is None:
        aligner = tga.get_aligner_object(None, gap_bool, 'tvr')
        P = _P(None, gap_bool)
    else:
        aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', which), gap_bool, 'tvr')
        P = _P(oracle.scoring_matrix('C', which), gap_bool)
    eng = DistEngine(Scoring.from_aligner(aligner), force_generic=force_generic)
    from telogator2_b200.dist_engine import canonical_order
    res = eng.pair_scores(seqs, adjust, minv, R, seed)
    D = eng.dist_matrix(seqs, adjust, minv, R, seed, res=res)
    res = canonical_order(res, len(seqs))
    Do, aln, rnd, cells = oracle.dist_matrix_c(seqs, P, adjust, minv, R, seed, want_scores=True)
    bad = np.nonzero(res['aln'] != aln)[0]
    assert len(bad) == 0, f'aln mismatch at pairs {bad[:10]}: gpu {res["aln"][bad[:10]]} oracle {aln[bad[:10]]}'
    if R:
        badr = np.argwhere(res['rnd'] != rnd)
        assert len(badr) == 0, f'rand mismatch at {badr[:10]}: gpu {res["rnd"][tuple(badr[:10].T)]} oracle {rnd[tuple(badr[:10].T)]}'
    assert np.array_equal(D, Do), 'distance matrix differs'
    assert eng.stats['cells'] == int(cells)
    # device epilogue + matrix fill (the path get_dist_matrix_parallel takes)
    eng2 = DistEngine(Scoring.from_aligner(aligner), force_generic=force_generic)
    st = eng2.prepare(seqs)
    if st is not None:
        Dd = eng2.matrix_device(st, adjust, minv, R, seed).cpu().numpy()
        assert np.array_equal(Dd, Do), 'device-epilogue matrix differs'
        assert eng2.stats['cells'] == int(cells) and eng2.stats['early_out'] == eng.stats['early_out']
    return eng


def test_wave_iter0_ont_golden(tga, colorvecs, dp_golden):
    ont, _ = _cases(colorvecs)
    seqs = [c[:1000] for c in ont][:24]
    eng = _check_matrix(tga, seqs, 'tvr', (True, True), True, True, 2, 12345)
    assert eng.stats['wave_jobs'] > 0 and eng.stats['generic_jobs'] == 0
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    D = tga.get_dist_matrix_parallel(seqs, aligner, True, True, 2, max_workers=3, rng_seed=12345)
    assert D.dtype == np.dtype('<f4') and np.array_equal(D, dp_golden['ont24_tvr_R2|D'])


def test_wave_iter0_hifi_golden(tga, colorvecs, dp_golden):
    _, hifi = _cases(colorvecs)
    seqs = [c[:1000] for c in hifi][:24]
    _check_matrix(tga, seqs, 'tvr', (True, True), True, True, 2, 12345)


def test_wave_iter2_shape_and_ragged(tga, colorvecs):
    ont, hifi = _cases(colorvecs)
    seqs = [c[:3000] for c in ont][:10]
    _check_matrix(tga, seqs, 'tvr', (True, True), True, True, 3, 999)
    # ragged lengths, no min_viable, odd R, mixed shapes in one packed job
    rag = [c[:137 + 211 * i] for i, c in enumerate(hifi[:9])] + ['C', 'CA', 'DDC']
    _check_matrix(tga, rag, 'tvr', (True, True), True, False, 1, 7)
    _check_matrix(tga, rag, 'consensus', (True, True), False, False, 3, 2 ** 32 + 5)


def test_wave_right_free_match_mismatch(tga, colorvecs, dp_golden):
    ont, _ = _cases(colorvecs)
    seqs = [c[:700 + 97 * i] for i, c in enumerate(ont[:12])]
    eng = _check_matrix(tga, seqs, None, (True, False), False, False, 3, 12345)
    assert eng.stats['wave_jobs'] > 0
    aligner = tga.get_aligner_object(None, (True, False), 'tvr')
    D = tga.get_dist_matrix_parallel(seqs, aligner, False, False, 3, rng_seed=12345)
    assert np.array_equal(D, dp_golden['ont12_ms_rightfree_R3|D'])
    # nucleotide alphabet (stage [4] subtel shape)
    rng = np.random.default_rng(3)
    base = ''.join(rng.choice(list('ACGT'), 900))
    nts = []
    for i in range(8):
        s = list(base)
        for p in rng.choice(len(s), 40 * i, replace=False):
            s[p] = rng.choice(list('ACGT'))
        nts.append(''.join(s)[:900 - 31 * i])
    _check_matrix(tga, nts, None, (True, False), False, False, 3, 5)


def test_generic_kernel_all_gap_modes(tga, colorvecs):
    ont, _ = _cases(colorvecs)
    seqs = [c[:300 + 41 * i] for i, c in enumerate(ont[:8])] + ['C', 'DC']
    for gap_bool in ((True, True), (True, False), (False, True), (False, False)):
        _check_matrix(tga, seqs, 'tvr', gap_bool, False, False, 2, 11, force_generic=(gap_bool[0] is True))
        _check_matrix(tga, seqs, None, gap_bool, True, False, 1, 12, force_generic=(gap_bool[0] is True))


def test_long_sequences_use_32bit_cells(tga):
    """Scores beyond 16 bits (collapse stage, untruncated consensus sequences, tg_tvr.py:375-390): the wave kernel runs
    one alignment per job in 32-bit cells, on the device-driven path and on the explicit-job path."""
    from telogator2_b200.dist_engine import DistEngine, Scoring, canonical_order
    rng = np.random.default_rng(9)
    seqs = [''.join(rng.choice(list('CCCCDEFA'), n)) for n in (6000, 5500, 7001, 5042, 300)]
    eng = _check_matrix(tga, seqs, None, (True, False), False, False, 1, 3)
    assert eng.stats['wave32_jobs'] == 10 and eng.stats['generic_jobs'] == 0
    eng = _check_matrix(tga, seqs, 'tvr', (True, True), True, True, 2, 11)
    assert eng.stats['wave32_jobs'] == 10 and eng.stats['generic_jobs'] == 0
    # explicit jobs: only the alignments that need it go to the 32-bit kernel
    aligner = tga.get_aligner_object(None, (True, False), 'tvr')
    e2 = DistEngine(Scoring.from_aligner(aligner))
    res = e2.pair_scores(seqs, False, False, 1, 3, pair_idx=np.arange(10, dtype=np.int64))
    _, aln, rnd, _ = oracle.dist_matrix_c(seqs, _P(None, (True, False)), False, False, 1, 3, want_scores=True)
    assert np.array_equal(res['aln'], aln) and np.array_equal(res['rnd'], rnd)
    assert e2.stats['wave32_jobs'] > 0 and e2.stats['wave_jobs'] > 0 and e2.stats['generic_jobs'] == 0
    eng = _check_matrix(tga, seqs[:3], None, (True, False), False, False, 1, 3, force_generic=True)
    assert eng.stats['generic_jobs'] > 0


def test_tvr_distance_and_quick_compare_follow_host_random(tga, colorvecs):
    ont, _ = _cases(colorvecs)
    a, b = ont[0][:1500], ont[1][:1400]
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    P = _P(oracle.scoring_matrix('C', 'tvr'), (True, True))
    d = tga.tvr_distance(a, b, aligner, True, True, 3, rng_seed=4242)
    assert d == oracle.tvr_distance_py(a, b, P, True, True, 3, rng_seed=4242)
    # no reseed: consumes the process-wide stream exactly like the reference (stage [7])
    random.seed(77)
    got = tga.quick_compare_tvrs(a, b)
    random.seed(77)
    want = oracle.tvr_distance_py(a, b, _P(None, (True, False)), False, False, 3, rng_seed=None) / 10.0
    assert got == want


def test_edge_cases(tga):
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    assert tga.get_dist_matrix_parallel([], aligner, True, True, 2).shape == (0, 0)
    assert np.array_equal(tga.get_dist_matrix_parallel(['CCD'], aligner, True, True, 2), np.zeros((1, 1), '<f4'))
    with pytest.raises(ValueError):
        tga.get_dist_matrix_parallel(['CCD', ''], aligner, True, True, 2)
    with pytest.raises(ValueError):
        tga.get_dist_matrix_parallel(['CCD', 'CBZ'], aligner, True, True, 2)      # letters not in the alphabet
    D = tga.get_dist_matrix_parallel(['CCDCC' * 50, 'CCDCC' * 50], aligner, True, True, 2)
    assert D[0, 1] == np.float32(1e-4) and D[0, 0] == 0            # identical sequences hit MIN_SEQ_DIST


def test_clustering_assignments_match(tga, colorvecs):
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    _, hifi = _cases(colorvecs)
    seqs = [c[:1000] for c in hifi][:40]
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    D = tga.get_dist_matrix_parallel(seqs, aligner, True, True, 2, rng_seed=31337) / 10.0
    Do = oracle.dist_matrix_c(seqs, _P(oracle.scoring_matrix('C', 'tvr'), (True, True)), True, True, 2, 31337) / 10.0
    a1 = fcluster(linkage(squareform(D), method='ward'), 0.2, 'distance')
    a2 = fcluster(linkage(squareform(Do), method='ward'), 0.2, 'distance')
    assert np.array_equal(a1, a2)


def test_mt_shuffle_kernel_direct(rng_golden):
    """tg_mt_shuffle against CPython golden permutations, seeds up to 2**63."""
    import torch
    from telogator2_b200 import _lib
    L = _lib.load()
    keys = sorted(rng_golden.keys())
    jobs = np.zeros(len(keys), dtype=_lib.SHUFFLE_JOB)
    srcs, off = [], 0
    meta = []
    for q, key in enumerate(keys):
        seed, n = (int(x) for x in key.split('|'))
        m = max(n - 1, 1)
        a = (np.arange(n) % 251).astype(np.uint8)
        b = (np.arange(m) % 241).astype(np.uint8)
        jobs[q]['seed'], jobs[q]['src_i'], jobs[q]['src_j'], jobs[q]['n'], jobs[q]['m'] = seed, off, off + n, n, m
        srcs += [a, b]
        off += n + m
        meta.append((n, m, a, b))
    src_total = off
    for q, (n, m, a, b) in enumerate(meta):
        jobs[q]['dst_off'] = off
        off += 2 * (n + m)            # R = 2 -> i, j, i, j
    pool = torch.zeros(off, dtype=torch.uint8, device='cuda')
    pool[:src_total] = torch.from_numpy(np.concatenate(srcs)).cuda()
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).cuda()
    _lib.check(L.tg_mt_shuffle(_lib.dptr(pool), _lib.dptr(d_jobs), len(keys), 2, 3000, _lib.stream_ptr(torch)))
    out = pool.cpu().numpy()
    for q, key in enumerate(keys):
        n, m, a, b = meta[q]
        perms = rng_golden[key]
        p1, p2, p3 = perms[:n], perms[n:n + m], perms[n + m:]
        d = int(jobs[q]['dst_off'])
        assert np.array_equal(out[d:d + n], a[p1]), key
        assert np.array_equal(out[d + n:d + n + m], b[p2]), key
        assert np.array_equal(out[d + n + m:d + 2 * n + m], a[p3]), key


def test_full_size_properties(tga, colorvecs):
    """BASELINE-sized iter-0 shape (N = 512 replicated colour vectors): properties that need no oracle."""
    ont, hifi = _cases(colorvecs)
    rng = np.random.default_rng(1)
    base = [c[:1000] for c in (ont + hifi) if len(c) >= 1000]
    seqs = []
    for i in range(512):
        s = list(base[i % len(base)])
        for p in rng.choice(len(s), 20, replace=False):
            s[p] = 'CVHTFDA'[rng.integers(7)]
        seqs.append(''.join(s))
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    D = tga.get_dist_matrix_parallel(seqs, aligner, True, True, 2, rng_seed=5)
    assert np.array_equal(D, D.T) and (np.diag(D) == 0).all()
    assert D[np.triu_indices(512, 1)].min() >= np.float32(1e-4) and D.max() <= 10
    # permutation consistency on a sub-block: sub-matrix recomputed alone with the matching seeds
    sub = [0, 5, 17, 300]
    P = _P(oracle.scoring_matrix('C', 'tvr'), (True, True))
    pairs = list(combinations(range(512), 2))
    index = {ij: p for p, ij in enumerate(pairs)}
    for i, j in combinations(sub, 2):
        want = oracle.tvr_distance_py(seqs[i], seqs[j], P, True, True, 2, rng_seed=5 + index[(i, j)])
        assert D[i, j] == np.float32(want)


def test_device_epilogue_untrusted_pairs_go_to_the_host(tga, colorvecs):
    """A distance whose float32 value could depend on the last bits of log() is recomputed with the host's libm.
    A huge margin makes many pairs 'untrusted'; a tiny list capacity exercises the overflow path as well."""
    from telogator2_b200.dist_engine import DistEngine, Scoring
    ont, hifi = _cases(colorvecs)
    seqs = [c[:1000] for c in (ont + hifi)[:40]]
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    Do = oracle.dist_matrix_c(seqs, _P(oracle.scoring_matrix('C', 'tvr'), (True, True)), True, True, 2, 777)
    fixed = []
    for margin, cap in ((46, 1 << 16), (12, 1 << 16), (12, 4)):
        eng = DistEngine(Scoring.from_aligner(aligner), margin_log2=margin, ambig_cap=cap)
        D = eng.matrix_device(eng.prepare(seqs), True, True, 2, 777).cpu().numpy()
        assert np.array_equal(D, Do), (margin, cap)
        fixed.append(eng.stats.get('host_fixed_pairs', 0))
    assert fixed[0] <= 2 and 4 < fixed[1] < len(seqs) * (len(seqs) - 1) // 2 and fixed[2] == len(seqs) * (len(seqs) - 1) // 2, fixed


def test_single_kernel_null_model_matches_streamed(tga, colorvecs):
    """The all-in-one shuffle kernel (fallback of the streamed seed/stream/sample pipeline) gives the same bits."""
    from telogator2_b200.dist_engine import DistEngine, Scoring, canonical_order
    ont, hifi = _cases(colorvecs)
    seqs = [c[:600 + 37 * i] for i, c in enumerate((ont + hifi)[:20])]
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    a = DistEngine(Scoring.from_aligner(aligner), single_kernel_shuffle=True).pair_scores(seqs, True, False, 3, 99)
    b = DistEngine(Scoring.from_aligner(aligner)).pair_scores(seqs, True, False, 3, 99)
    for key in ('aln', 'rnd', 'n', 'm'):
        assert np.array_equal(a[key], b[key]), key
    _, aln, rnd, _ = oracle.dist_matrix_c(seqs, _P(oracle.scoring_matrix('C', 'tvr'), (True, True)), True, False, 3, 99, want_scores=True)
    c = canonical_order(a, len(seqs))
    assert np.array_equal(c['aln'], aln) and np.array_equal(c['rnd'], rnd)


def test_unsupported_configurations_fail_loudly(tga):
    from telogator2_b200.dist_engine import Scoring
    al = tga.get_aligner_object(None, (True, True), 'msa')                  # affine gaps: MSA half only
    with pytest.raises(NotImplementedError):
        tga.get_dist_matrix_parallel(['CCD', 'CDD'], al, True, True, 1)
    m = tga.get_scoring_matrix('C', 'tvr')
    m['D', 'D'] = 4.5
    with pytest.raises(NotImplementedError):
        Scoring.from_aligner(tga.get_aligner_object(m, (True, True), 'tvr'))
    with pytest.raises(SystemExit):
        tga.get_scoring_matrix('C', 'nope')
    with pytest.raises(SystemExit):
        tga.get_aligner_object(None, (True, True), 'nope')


def test_negative_and_large_seeds(tga, colorvecs):
    ont, _ = _cases(colorvecs)
    seqs = [c[:500] for c in ont[:6]]
    for seed in (-12345, 0, 2 ** 40 + 7):
        _check_matrix(tga, seqs, 'tvr', (True, True), True, False, 2, seed)


def test_many_matrices_in_one_pass_equal_the_per_call_results(tga, colorvecs):
    """independent per-cluster matrices batched into one pair space give the per-call bits."""
    ont, hifi = _cases(colorvecs)
    allv = ont + hifi
    groups = [[c[:3000] for c in allv[0:9]], [c[:700] for c in allv[9:11]], [allv[11][:50]], [], [c[:3000] for c in allv[12:30]],
              [c[:1000 + 97 * i] for i, c in enumerate(allv[30:37])]]
    for (which, gap_bool, adjust, minv, R, seed) in (('tvr', (True, True), True, True, 3, 12345), (None, (True, False), False, False, 3, 7)):
        aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', which) if which else None, gap_bool, 'tvr')
        many = tga.get_dist_matrices_parallel(groups, aligner, adjust, minv, R, rng_seed=seed)
        assert len(many) == len(groups)
        P = _P(oracle.scoring_matrix('C', which) if which else None, gap_bool)
        for g, D in zip(groups, many):
            assert D.shape == (len(g), len(g)) and D.dtype == np.dtype('<f4')
            one = tga.get_dist_matrix_parallel(g, aligner, adjust, minv, R, rng_seed=seed)
            assert np.array_equal(D, one)
            if 2 <= len(g) <= 9:
                assert np.array_equal(D, oracle.dist_matrix_c(g, P, adjust, minv, R, seed))


@pytest.mark.parametrize('key,n', [('human_ont', 83), ('human_hifi', 101), ('maize_hifi', 13)])
def test_full_iter0_matrices_of_the_bundled_sets(tga, colorvecs, key, n):
    """D0: the whole iteration-0 matrix of every bundled set (N = 83 / 101 / 13, truncate 1000, R = 2, seed
    12345) through the drop-in call, bit-equal to the oracle, and the same ward clusters."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    seqs = [c[:1000] for c in colorvecs['colorvecs'][key]]
    assert len(seqs) == n
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    D = tga.get_dist_matrix_parallel(seqs, aligner, True, True, 2, rng_seed=12345)
    Do = oracle.dist_matrix_c(seqs, _P(oracle.scoring_matrix('C', 'tvr'), (True, True)), True, True, 2, 12345)
    assert D.dtype == np.dtype('<f4') and np.array_equal(D, Do)
    a1 = fcluster(linkage(squareform(D / 10.0), method='ward'), 0.2, 'distance')
    a2 = fcluster(linkage(squareform(Do / 10.0), method='ward'), 0.2, 'distance')
    assert np.array_equal(a1, a2)


@pytest.mark.parametrize('R', [2, 3])
def test_pipelined_batches_and_per_job_profiles_change_nothing(tga, colorvecs, R, monkeypatch):
    """Many small batches: the stream + sampling kernels of batch k run on a second stream under the alignments of batch k + 1
    (two buffer sets); null-model alignments run with per-job query profiles where the row sequences use few letters.  Both
    are scheduling / layout choices: the matrix must equal the single-stream, full-profile one and the oracle's."""
    import torch
    from telogator2_b200.dist_engine import DistEngine, Scoring
    seqs = [c[:1000] for c in colorvecs['colorvecs']['human_hifi']][:60]
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    Do = oracle.dist_matrix_c(seqs, _P(oracle.scoring_matrix('C', 'tvr'), (True, True)), True, True, R, 777)
    got = {}
    for pipe, rows in (('1', None), ('0', None), ('1', '0'), ('1', '9'), ('0', '15')):
        monkeypatch.setenv('TG_DIST_PIPELINE', pipe)
        if rows is None:
            monkeypatch.delenv('TG_WAVE_PROF_ROWS', raising=False)
        else:
            monkeypatch.setenv('TG_WAVE_PROF_ROWS', rows)
        eng = DistEngine(Scoring.from_aligner(aligner), pairs_per_batch=128)
        D = eng.matrix_device(eng.prepare(seqs), True, True, R, 777)
        torch.cuda.synchronize()
        got[(pipe, rows)] = D.cpu().numpy()
        assert np.array_equal(got[(pipe, rows)], Do), (pipe, rows)


def test_per_job_profiles_with_several_matrices(tga, colorvecs, monkeypatch):
    """Several matrices in one pair space (multi-matrix instantiation of the kernels) with the per-job query profiles forced on."""
    ont, hifi = _cases(colorvecs)
    groups = [[c[:900 + 13 * i] for i, c in enumerate(hifi[:7])], [c[:1000] for c in ont[:5]], [hifi[9][:800]], [c[:1000] for c in hifi[20:26]]]
    aligner = tga.get_aligner_object(tga.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    want = None
    for rows in ('0', '8', '13'):
        monkeypatch.setenv('TG_WAVE_PROF_ROWS', rows)
        got = tga.get_dist_matrices_parallel(groups, aligner, True, True, 3, rng_seed=4242)
        if want is None:
            want = got
            P = _P(oracle.scoring_matrix('C', 'tvr'), (True, True))
            for g, D in zip(groups, got):
                if len(g) >= 2:
                    assert np.array_equal(D, oracle.dist_matrix_c(g, P, True, True, 3, 4242))
        else:
            assert all(np.array_equal(a, b) for a, b in zip(got, want)), rows


def test_against_real_biopython_when_importable(tga, colorvecs):
    """The one unpinned piece: Bio.Align.PairwiseAligner.score.  Skipped where Biopython is absent (the build image and the GPU
    boxes: profiles/r02_pin_biopython_*.log); wherever it is importable this compares the CUDA matrices with the reference's
    tvr_distance arithmetic around the REAL aligner (tools/pin_biopython.py holds the same code for freezing golden vectors)."""
    pytest.importorskip('Bio')
    import os
    import sys
    from itertools import combinations
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import pin_biopython
    ont, hifi = _cases(colorvecs)
    for seqs, S_type, gb, adj, R in (([c[:1000] for c in ont[:20]], 'tvr', (True, True), True, 2),
                                     ([c[:700 + 61 * i] for i, c in enumerate(hifi[:14])], None, (True, False), False, 3)):
        S = tga.get_scoring_matrix('C', S_type) if S_type else None
        aligner = tga.get_aligner_object(S, gb, 'tvr')
        assert type(aligner).__module__.startswith('Bio.')
        diag = {ch: float(S[ch, ch]) for ch in tga.AMINO} if S is not None else None
        D = tga.get_dist_matrix_parallel(seqs, aligner, adj, adj, R, rng_seed=4242)
        for p, (i, j) in enumerate(combinations(range(len(seqs)), 2)):
            d, _aln, _rs = pin_biopython.tvr_distance(seqs[i], seqs[j], aligner, adj, adj, R, 4242 + p, diag)
            assert D[i, j] == np.float32(d) == D[j, i], (i, j)
