"""CPU: the host orchestration of DistEngine's device-driven path (prepare / _launch_fast / matrix_device / _pair_scores_fast:
batches, sharded ranges, several matrices in one pair space, the host fix-up of untrusted roundings) with the two C entry points
it drives -- tg_dist_batch and tg_dist_epilogue -- answered by the CPU oracle on ordinary (CPU) torch tensors.

The stand-in reads the same ctypes argument structs the CUDA library receives, so the struct layouts and the pointer
arithmetic of the host code are exercised too."""
import ctypes as C
import random

import numpy as np
import pytest
import torch

import oracle
from telogator2_b200 import _lib, dist_engine
from telogator2_b200.dist_engine import DistEngine, Scoring, NOT_COMPUTED


def _view(ptr, ctype, count):
    if not ptr or count <= 0:
        return np.zeros(0, dtype=np.dtype(ctype))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(int(count),))


def _unrank(q, n):
    r = 0
    while (r + 1) * (2 * n - (r + 1) - 1) // 2 <= q:
        r += 1
    return r, q - r * (2 * n - r - 1) // 2 + r + 1


class OracleLib:
    """tg_dist_batch / tg_dist_epilogue on host pointers, computed by the oracle."""

    def __init__(self, untrusted_every=0):
        self.untrusted_every = untrusted_every
        self.batches = 0

    def tg_dist_batch_mt_work_bytes(self, n, R, max_len):
        return 4096

    def tg_nw_wave_scratch_bytes(self, max_len):
        return 4096

    @staticmethod
    def _describe(a, q, seq_off, order):
        n, sbase, mat = a.n_seq, 0, 0
        if a.n_mat > 0:
            po = _view(a.d_mat_pair_off, C.c_int64, a.n_mat + 1)
            so = _view(a.d_mat_seq_off, C.c_int32, a.n_mat + 1)
            mat = int(np.searchsorted(po, q, side='right') - 1)
            q -= int(po[mat])
            sbase = int(so[mat])
            n = int(so[mat + 1]) - sbase
        ra, rb = _unrank(q, n)
        oa, ob = int(order[sbase + ra]), int(order[sbase + rb])
        i, j = min(oa, ob), max(oa, ob)
        li, lj = int(seq_off[i + 1] - seq_off[i]), int(seq_off[j + 1] - seq_off[j])
        if a.adjust_lens:
            ml = min(li, lj)
            if a.min_viable:
                ml = max(ml, 1000)
            li, lj = min(li, ml), min(lj, ml)
        il, jl = i - sbase, j - sbase
        return dict(i=i, j=j, n=li, m=lj, p=il * (2 * n - il - 1) // 2 + (jl - il - 1), il=il, jl=jl, mat=mat, nloc=n)

    def tg_dist_batch(self, a_ref, sc_ref, stream):
        a, sc = a_ref._obj, sc_ref._obj
        self.batches += 1
        npairs = a.q1 - a.q0
        A = sc.alphabet
        S = np.array(sc.sub, dtype=np.float64).reshape(_lib.MAXA, _lib.MAXA)[:A, :A].copy()
        seq_off = _view(a.d_seq_off, C.c_int64, a.n_seq + 1)
        order = _view(a.d_order, C.c_int32, a.n_seq)
        pool = _view(a.d_pool, C.c_uint8, int(seq_off[-1]))
        aln, rnd = _view(a.d_aln, C.c_int32, npairs), _view(a.d_rnd, C.c_int32, npairs * max(a.R, 1))
        flags, iden = _view(a.d_flags, C.c_uint8, npairs), _view(a.d_iden, C.c_double, npairs)
        shape = _view(a.d_shape, C.c_int32, 2 * npairs)
        cs = _view(a.d_diag_cs, C.c_int32, int(seq_off[-1]) + 1) if a.d_diag_cs else None
        L = oracle.lib()
        g, gl, gr = float(sc.gap), float(sc.gap_left), float(sc.gap_right)

        def score(x, y):
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
            return int(L.tgo_nw_score(x.ctypes.data, len(x), y.ctypes.data, len(y), S.ctypes.data, A, 0., 0., g, gl, gr))
        for ql in range(npairs):
            d = self._describe(a, a.q0 + ql, seq_off, order)
            si = pool[seq_off[d['i']]:seq_off[d['i']] + d['n']]
            sj = pool[seq_off[d['j']]:seq_off[d['j']] + d['m']]
            aln[ql] = score(si, sj)
            if cs is not None:
                idn = (float(cs[seq_off[d['i']] + d['n']] - cs[seq_off[d['i']]]) + float(cs[seq_off[d['j']] + d['m']] - cs[seq_off[d['j']]])) / 2.
            else:
                idn = a.match * ((d['n'] + d['m']) / 2.)
            iden[ql] = idn
            shape[2 * ql], shape[2 * ql + 1] = d['n'], d['m']
            early = float(aln[ql]) < -(0.900 * idn)
            flags[ql] = 1 if early else 0
            if not early and a.R > 0:
                random.seed(abs(a.seed0 + d['p']))
                for k in range(a.R):
                    xi = np.array(random.sample(si.tolist(), len(si)), dtype=np.uint8)
                    xj = np.array(random.sample(sj.tolist(), len(sj)), dtype=np.uint8)
                    rnd[ql * a.R + k] = score(xi, xj)
        return 0

    def tg_dist_epilogue(self, e_ref, stream):
        e = e_ref._obj
        npairs = e.q1 - e.q0
        seq_off = _view(e.d_seq_off, C.c_int64, e.n_seq + 1)
        order = _view(e.d_order, C.c_int32, e.n_seq)
        aln = _view(e.d_aln, C.c_int32, npairs)
        rnd = _view(e.d_rnd, C.c_int32, npairs * max(e.R, 1)).reshape(npairs, max(e.R, 1))[:, :e.R]
        flags, iden = _view(e.d_flags, C.c_uint8, npairs), _view(e.d_iden, C.c_double, npairs)
        stats = _view(e.d_stats, C.c_uint64, 4)
        dist = DistEngine.distances(dict(aln=aln.copy(), rnd=rnd.copy(), iden=iden.copy()), e.R).astype('<f4')

        class A:                       # describe needs the batch-style fields
            pass
        a = A()
        for f in ('n_seq', 'adjust_lens', 'min_viable', 'n_mat', 'd_mat_pair_off', 'd_mat_seq_off'):
            setattr(a, f, getattr(e, f))
        out_off = _view(e.d_mat_out_off, C.c_int64, e.n_mat) if e.n_mat > 0 else None
        total = int(out_off[-1]) + 0 if out_off is not None else 0
        for ql in range(npairs):
            d = self._describe(a, e.q0 + ql, seq_off, order)
            base = int(out_off[d['mat']]) if out_off is not None else 0
            nloc = d['nloc']
            cells = (base + d['il'] * nloc + d['jl'], base + d['jl'] * nloc + d['il'])
            val = dist[ql]
            stats[0] += d['n'] * d['m']
            if flags[ql]:
                stats[2] += 1
            else:
                stats[1] += d['n'] * d['m'] * e.R
                if self.untrusted_every and (e.q0 + ql) % self.untrusted_every == 0:
                    k = int(stats[3])
                    stats[3] += 1
                    if k < e.ambig_cap:
                        _view(e.d_ambig, C.c_int64, e.ambig_cap)[k] = e.ambig_base + ql
                    val = np.float32(123.0)                     # deliberately wrong: the host must repair it
            m = _view(e.d_matrix, C.c_float, max(cells) + 1)
            m[cells[0]] = val
            m[cells[1]] = val
        del total
        return 0


def _engine_from_scoring(sc, lib, **kw):
    """A DistEngine on CPU tensors whose C library is `lib` (DistEngine.__init__ insists on a CUDA device)."""
    eng = DistEngine.__new__(DistEngine)
    eng.torch, eng.L, eng.sc = torch, lib, sc
    eng.device = torch.device('cpu')
    eng.force_generic = False
    eng.pairs_per_batch = kw.get('pairs_per_batch', 1 << 21)
    eng.shuffle_bytes_per_batch = 8 << 30
    eng.single_kernel_shuffle = False
    eng.margin_log2, eng.ambig_cap = 46, kw.get('ambig_cap', 1 << 16)
    eng.stats = {'cells': 0, 'wave_jobs': 0, 'wave32_jobs': 0, 'generic_jobs': 0, 'launches': 0, 'pairs': 0, 'early_out': 0,
                 'h2d_bytes': 0, 'd2h_bytes': 0}
    eng._scratch = eng._counter = None
    return eng


def _engine(aligner, lib, **kw):
    return _engine_from_scoring(Scoring.from_aligner(aligner), lib, **kw)


@pytest.fixture()
def host_only(monkeypatch):
    monkeypatch.setattr(_lib, 'stream_ptr', lambda t: None)
    return monkeypatch


def _seqs(colorvecs, k, cut):
    cv = colorvecs['colorvecs']['human_ont'] + colorvecs['colorvecs']['human_hifi']
    return [c[:cut + 13 * i] for i, c in enumerate(cv[:k])]


def test_matrix_device_batches_ranges_and_fixups(host_only, colorvecs):
    from telogator2_b200 import tg_align
    seqs = _seqs(colorvecs, 9, 120)
    aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    want, aln, rnd, cells = oracle.dist_matrix_c(seqs, P, True, False, 2, 4321, want_scores=True)
    # several small batches
    lib = OracleLib()
    eng = _engine(aligner, lib, pairs_per_batch=10)
    D = eng.matrix_device(eng.prepare(seqs), True, False, 2, 4321).numpy()
    assert np.array_equal(D, want) and lib.batches == 4 and eng.stats['cells'] == int(cells)
    # sharded: three ranks' partial matrices add up; compact integer results expand to the full arrays
    monkey_ranges = dist_engine.fast_ranges
    host_only.setattr(dist_engine, 'fast_ranges', lambda n, r, w, min_block=4: monkey_ranges(n, r, w, min_block=4))
    parts = [_engine(aligner, OracleLib()) for _ in range(3)]
    tot = sum(e.matrix_device(e.prepare(seqs), True, False, 2, 4321, shard=(r, 3)) for r, e in enumerate(parts)).numpy()
    assert np.array_equal(tot, want)
    full_aln = np.full(len(aln), NOT_COMPUTED, dtype=np.int32)
    for r in range(3):
        e = _engine(aligner, OracleLib())
        res = dist_engine.canonical_order(e._pair_scores_fast(e.prepare(seqs), True, False, 2, 4321, (r, 3)), len(seqs))
        own = res['aln'] != NOT_COMPUTED
        assert not (own & (full_aln != NOT_COMPUTED)).any()
        full_aln[own] = res['aln'][own]
        assert np.array_equal(res['rnd'][own], rnd[own])
    assert np.array_equal(full_aln, aln)
    # untrusted roundings: every 3rd pair comes back wrong from the "device" and flagged; with a list that overflows, too
    for cap in (1 << 16, 2):
        e = _engine(aligner, OracleLib(untrusted_every=3), ambig_cap=cap)
        D = e.matrix_device(e.prepare(seqs), True, False, 2, 4321).numpy()
        assert np.array_equal(D, want) and e.stats['host_fixed_pairs'] > 0


def test_many_matrices_one_pair_space(host_only, colorvecs):
    from telogator2_b200 import tg_align
    cv = colorvecs['colorvecs']['human_ont']
    groups = [[c[:90 + 7 * i] for i, c in enumerate(cv[0:5])], [cv[5][:40]], [], [c[:150] for c in cv[6:10]], [c[:60] for c in cv[10:12]]]
    aligner = tg_align.get_aligner_object(None, (True, False), 'tvr')
    P = oracle.AlignerParams(None, (True, False))
    want = [oracle.dist_matrix_c(g, P, False, False, 3, 99) if len(g) >= 2 else np.zeros((len(g), len(g)), '<f4') for g in groups]
    for untrusted in (0, 2):
        e = _engine(aligner, OracleLib(untrusted_every=untrusted), pairs_per_batch=6)
        st = e.prepare(groups, many=True)
        flat = e.matrix_device(st, False, False, 3, 99).numpy()
        off = st['mat_out_off']
        for b, g in enumerate(groups):
            assert np.array_equal(flat[off[b]:off[b + 1]].reshape(len(g), len(g)), want[b]), (untrusted, b)


# ---------------------------------------------------------------------------------------------------------
# explicit-job path: tg_nw_scores_wave / _wave32 / _generic and tg_mt_shuffle answered by the oracle
# ---------------------------------------------------------------------------------------------------------
def _addr(p):
    return p.value if isinstance(p, C.c_void_p) else p


class OracleJobLib(OracleLib):
    def __init__(self):
        super().__init__()
        self.calls = {'wave': 0, 'wave32': 0, 'generic': 0, 'shuffle': 0}

    def _scores(self, kind, d_pool, d_jobs, n_jobs, sc_ref, d_scores):
        self.calls[kind] += 1
        sc = sc_ref._obj
        A = sc.alphabet
        S = np.array(sc.sub, dtype=np.float64).reshape(_lib.MAXA, _lib.MAXA)[:A, :A].copy()
        jobs = _view(_addr(d_jobs), C.c_uint8, n_jobs * _lib.NW_JOB.itemsize).view(_lib.NW_JOB)
        hi = int(max((jobs['row_off'] + jobs['n']).max(), (jobs['col_off'] + jobs['m']).max()))
        pool = _view(_addr(d_pool), C.c_uint8, hi)
        n_out = int(jobs['out'].max()) + 1
        scores = _view(_addr(d_scores), C.c_int32, n_out)
        L = oracle.lib()
        for J in jobs:
            for h in (0, 1):
                if J['out'][h] < 0:
                    continue
                x = np.ascontiguousarray(pool[J['row_off'][h]:J['row_off'][h] + J['n'][h]])
                y = np.ascontiguousarray(pool[J['col_off'][h]:J['col_off'][h] + J['m'][h]])
                scores[J['out'][h]] = int(L.tgo_nw_score(x.ctypes.data, len(x), y.ctypes.data, len(y), S.ctypes.data, A, 0., 0.,
                                                          float(sc.gap), float(sc.gap_left), float(sc.gap_right)))
        return 0

    def tg_nw_scores_wave(self, d_pool, d_jobs, n_jobs, max_n, sc_ref, d_scores, d_scratch, d_counter, stream):
        return self._scores('wave', d_pool, d_jobs, n_jobs, sc_ref, d_scores)

    def tg_nw_scores_wave32(self, d_pool, d_jobs, n_jobs, max_n, sc_ref, d_scores, d_scratch, d_counter, stream):
        return self._scores('wave32', d_pool, d_jobs, n_jobs, sc_ref, d_scores)

    def tg_nw_scores_generic(self, d_pool, d_jobs, n_jobs, max_m, sc_ref, d_scores, stream):
        return self._scores('generic', d_pool, d_jobs, n_jobs, sc_ref, d_scores)

    def tg_mt_shuffle(self, d_pool, d_jobs, n_jobs, R, max_len, stream):
        self.calls['shuffle'] += 1
        jobs = _view(_addr(d_jobs), C.c_uint8, n_jobs * _lib.SHUFFLE_JOB.itemsize).view(_lib.SHUFFLE_JOB)
        hi = int((jobs['dst_off'] + R * (jobs['n'].astype(np.int64) + jobs['m'])).max())
        pool = _view(_addr(d_pool), C.c_uint8, hi)
        for J in jobs:
            random.seed(int(J['seed']))
            dst = int(J['dst_off'])
            si = pool[J['src_i']:J['src_i'] + J['n']].tolist()
            sj = pool[J['src_j']:J['src_j'] + J['m']].tolist()
            for _ in range(R):
                for s in (si, sj):
                    pool[dst:dst + len(s)] = random.sample(s, len(s))
                    dst += len(s)
        return 0


def test_explicit_job_path_routing_and_seeds(host_only, colorvecs):
    """pair_scores(pair_idx=...): job packing, routing to the 16-bit / 32-bit / generic kernels, shuffle jobs with abs(seed + p),
    results scattered back in combinations order."""
    from telogator2_b200 import tg_align
    rng = np.random.default_rng(6)
    seqs = [''.join(rng.choice(list('CCCCDEFA'), n)) for n in (5100, 5060, 300, 200, 7)]
    n_pairs = len(seqs) * (len(seqs) - 1) // 2
    for gap_bool, seed, expect in (((True, False), -40, ('wave', 'wave32')), ((False, True), 12, ('generic',))):
        aligner = tg_align.get_aligner_object(None, gap_bool, 'tvr')
        P = oracle.AlignerParams(None, gap_bool)
        lib = OracleJobLib()
        eng = _engine(aligner, lib)
        res = eng.pair_scores(seqs, False, False, 2, seed, pair_idx=np.arange(n_pairs, dtype=np.int64))
        _, aln, rnd, _ = oracle.dist_matrix_c(seqs, P, False, False, 2, seed, want_scores=True)
        assert np.array_equal(res['aln'], aln) and np.array_equal(res['rnd'], rnd), gap_bool
        used = {k for k, v in lib.calls.items() if v and k != 'shuffle'}
        assert used == set(expect), (gap_bool, lib.calls)
        assert np.array_equal(eng.dist_matrix(seqs, False, False, 2, seed, res=res), oracle.dist_matrix_c(seqs, P, False, False, 2, seed))


def test_single_pair_api_follows_host_random(host_only, colorvecs):
    """tvr_distance / quick_compare_tvrs consume the process-wide `random` stream like the reference (tg_align.py:109-151, 379-381)."""
    from telogator2_b200 import tg_align
    host_only.setattr(tg_align, 'DistEngine', lambda sc: _engine_from_scoring(sc, OracleJobLib()))
    cv = colorvecs['colorvecs']['human_ont']
    a, b = cv[0][:400], cv[1][:350]
    aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    assert tg_align.tvr_distance(a, b, aligner, True, False, 3, rng_seed=5) == oracle.tvr_distance_py(a, b, P, True, False, 3, rng_seed=5)
    random.seed(99)
    got = [tg_align.quick_compare_tvrs(a, b), tg_align.quick_compare_tvrs(b, a)]
    P2 = oracle.AlignerParams(None, (True, False))
    random.seed(99)
    want = [oracle.tvr_distance_py(a, b, P2, False, False, 3, rng_seed=None) / 10.0, oracle.tvr_distance_py(b, a, P2, False, False, 3, rng_seed=None) / 10.0]
    assert got == want
    assert aligner.score(a, b) == oracle.nw_score(a, b, P)
