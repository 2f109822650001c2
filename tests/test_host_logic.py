"""CPU: host-side logic of the product (no CUDA calls) + C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT


def test_scoring_matrix_matches_oracle_closed_form_and_literal_replay():
    from telogator2_b200 import tg_align
    for which in ('tvr', 'consensus'):
        for canon in 'CDV':
            assert np.array_equal(np.asarray(tg_align.get_scoring_matrix(canon, which)), oracle.scoring_matrix(canon, which))
    # literal replay of tg_align.py:32-68 for all four types
    A, G, U = tg_align.AMINO, tg_align.GAP_CHR, tg_align.UNKNOWN
    for which in ('tvr', 'consensus', 'msa', 'msa_refinement'):
        m = {}
        for c1 in A:
            m[c1, c1] = 5.
            for c2 in A:
                if c1 != c2:
                    m[c1, c2] = -4.
        m[U, U] = 2.
        m[G, G] = 0.
        for c1 in A:
            m[c1, 'C'] = m['C', c1] = -4.
            m[c1, U] = m[U, c1] = -2.
            m[c1, G] = m[G, c1] = -4.
        if which == 'tvr':
            m['C', 'C'] = 0.; m['C', U] = m[U, 'C'] = -2.; m[U, G] = m[G, U] = -2.
        elif which == 'consensus':
            m['C', 'C'] = 1.; m['C', U] = m[U, 'C'] = -2.; m[U, G] = m[G, U] = -2.
        elif which == 'msa':
            m['C', 'C'] = 0.; m['C', U] = m[U, 'C'] = 1.; m[U, G] = m[G, U] = -2.
        else:
            for c1 in A:
                m[c1, G] = m[G, c1] = 0.
            m['C', 'C'] = 2.; m['C', U] = m[U, 'C'] = 1.; m[U, G] = m[G, U] = 1.
        got = tg_align.get_scoring_matrix('C', which)
        for (a, b), v in m.items():
            assert got[a, b] == v, (which, a, b)


def test_aligner_object_attributes():
    from telogator2_b200 import tg_align
    from telogator2_b200.dist_engine import Scoring
    al = tg_align.get_aligner_object(None, (True, False), 'tvr')
    sc = Scoring.from_aligner(al)
    assert (sc.gap, sc.gap_left, sc.gap_right, sc.match, sc.mismatch, sc.alphabet) == (-4, -4, 0, 5, -4, None)
    al = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (False, True), 'tvr')
    sc = Scoring.from_aligner(al)
    assert (sc.gap, sc.gap_left, sc.gap_right) == (-4, 0, -4) and sc.alphabet == 'ACDEFGHIKLMNPQRSTVWY-'
    with pytest.raises(NotImplementedError):
        Scoring.from_aligner(tg_align.get_aligner_object(None, (True, True), 'msa'))     # affine: MSA half, not score()
    assert tg_align.MAX_SEQ_DIST == 10.0 and tg_align.MIN_VIABLE_SEQ_LEN == 1000 and tg_align.UNKNOWN == 'A'


def test_pack_jobs_covers_every_alignment_once():
    from telogator2_b200.dist_engine import pack_jobs
    rng = np.random.default_rng(0)
    for allow_mixed in (True, False):
        n = rng.integers(1, 6, 101).astype(np.int32) * 100
        m = rng.integers(1, 4, 101).astype(np.int32) * 100
        ro = np.arange(101, dtype=np.int64) * 7
        co = ro + 3
        jobs = pack_jobs(ro, co, n, m, np.arange(101, dtype=np.int32), allow_mixed)
        outs = jobs['out'].reshape(-1)
        assert sorted(outs[outs >= 0].tolist()) == list(range(101))
        for h in (0, 1):
            ok = jobs['out'][:, h] >= 0
            o = jobs['out'][ok, h]
            assert np.array_equal(jobs['n'][ok, h], n[o]) and np.array_equal(jobs['m'][ok, h], m[o])
            assert np.array_equal(jobs['row_off'][ok, h], ro[o]) and np.array_equal(jobs['col_off'][ok, h], co[o])
        if not allow_mixed:
            assert (jobs['n'][:, 0] == jobs['n'][:, 1]).all() and (jobs['m'][:, 0] == jobs['m'][:, 1]).all()
        cost = np.maximum(jobs['n'][:, 0], jobs['n'][:, 1]) * np.maximum(jobs['m'][:, 0], jobs['m'][:, 1])
        assert (np.diff(cost) <= 0).all()


def test_pair_sharding_partitions_the_pair_space():
    from telogator2_b200.multi_gpu import shard_pair_indices
    for n_pairs in (0, 1, 4095, 4096, 100000):
        for world in (1, 2, 3, 8):
            parts = [shard_pair_indices(n_pairs, r, world) for r in range(world)]
            allp = np.sort(np.concatenate(parts))
            assert np.array_equal(allp, np.arange(n_pairs))
            assert all((np.diff(p) > 0).all() for p in parts if len(p) > 1)


def test_distance_epilogue_matches_python_scalar_path():
    from telogator2_b200.dist_engine import DistEngine, NOT_COMPUTED
    rng = np.random.default_rng(2)
    n = 500
    aln = rng.integers(-4000, 5000, n).astype(np.int32)
    iden = rng.integers(1000, 5000, n) / 2.
    rnd = rng.integers(-4000, 0, (n, 3)).astype(np.int32)
    aln[:5] = rnd[:5].mean(axis=1).astype(np.int32) - 1          # rand >= aln branch
    res = dict(aln=aln, rnd=rnd, iden=iden)
    early = aln < -(0.9 * iden)
    rnd[early] = NOT_COMPUTED
    d = DistEngine.distances(res, 3)
    for q in range(n):
        a, i = float(aln[q]), float(iden[q])
        if a < -(0.900 * i):
            want = 10.0
        else:
            r = np.mean([float(v) for v in rnd[q]])
            with np.errstate(all='ignore'):
                want = 10.0 if r >= a else min(-np.log((a - r) / (i - r)), 10.0)
            want = max(want, 0.0001)
        assert (np.isnan(want) and np.isnan(d[q])) or d[q] == want, q


def test_read_kmer_tsv_matches_reference(tmp_path, kmer_tables):
    from telogator2_b200 import tg_kmer
    for key, T in kmer_tables.items():
        fn = tmp_path / (key + '.tsv')
        fn.write_text(T['tsv_text'])
        meta, issub, canon = tg_kmer.read_kmer_tsv(str(fn), T['read_type'])
        assert meta[0] == T['KMER_LIST'] and meta[1] == T['KMER_COLORS'] and meta[2] == T['KMER_LETTER']
        assert [list(f) for f in meta[3]] == T['KMER_FLAGS'] and all(isinstance(f, tuple) for f in meta[3])
        assert issub == T['KMER_ISSUBSTRING'] and canon == T['CANONICAL_STRINGS']
        assert tg_kmer.get_canonical_letter(str(fn), T['read_type']) == T['canonical_letter']
    with pytest.raises(SystemExit):
        tg_kmer.read_kmer_tsv(str(tmp_path / 'missing.tsv'), 'ont')


def test_c_abi_exports_every_declared_symbol_and_table_builder(kmer_tables):
    from telogator2_b200 import _lib
    from telogator2_b200.build import build
    build()
    L = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'telogator2_b200.h')).read()
    declared = set(re.findall(r'\b(tg_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.tg_version() == 1
    # host-side table build (no GPU needed): limits are enforced, errors are reported, nothing throws
    T = kmer_tables['human_ont']
    blob = ''.join(T['KMER_LIST']).encode()
    off = np.zeros(len(T['KMER_LIST']) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(k) for k in T['KMER_LIST']])
    host = np.zeros(L.tg_kmer_table_bytes(), dtype=np.uint8)
    assert L.tg_kmer_table_build(blob, off.ctypes.data_as(ctypes.c_void_p), len(T['KMER_LIST']), None, None,
                                 host.ctypes.data_as(ctypes.c_void_p)) == 0
    hdr = host[:16].view(np.int32)
    assert hdr[0] == len(T['KMER_LIST']) and hdr[2] == 20 and hdr[3] == 4      # CCCCCC, CCCTCC and the two HOR k-mers
    bad = b'CCNTAA'
    off2 = np.array([0, 6], dtype=np.int32)
    assert L.tg_kmer_table_build(bad, off2.ctypes.data_as(ctypes.c_void_p), 1, None, None, host.ctypes.data_as(ctypes.c_void_p)) == _lib.TG_EINVAL
    assert b'non-ACGT' in L.tg_last_error()
    # compute entry points fail loudly without a device instead of falling back
    import torch
    if not torch.cuda.is_available():
        n = ctypes.c_int(0)
        assert L.tg_device_sm_count(ctypes.byref(n)) == _lib.TG_ECUDA
        with pytest.raises(RuntimeError):
            _lib.require_cuda()


def test_random_stream_capacity_covers_cpython_draw_counts():
    """tg_dist_batch pre-generates ceil((1.6 * elements + 256) / 624) blocks of MT19937 output per pair
    (mt_blocks_needed in tg_dp.cu).  Count the getrandbits() calls CPython's random.sample really makes."""
    import random

    class Counting(random.Random):
        calls = 0

        def getrandbits(self, k):
            Counting.calls += 1
            return super().getrandbits(k)

    worst = 0.0
    for n in (1, 2, 3, 17, 100, 511, 512, 513, 600, 724, 725, 1000, 1024, 1025, 1449, 2000, 3000, 4097, 5041):
        for seed in range(6):
            for R in (1, 2, 3):
                rng = Counting(seed * 7919 + n)
                Counting.calls = 0
                for _ in range(2 * R):
                    rng.sample(range(n), n)
                total = 2 * R * n
                cap = (total * 8 // 5 + 256 + 623) // 624 * 624
                assert Counting.calls <= cap, (n, seed, R, Counting.calls, cap)
                worst = max(worst, Counting.calls / cap)
    assert worst < 0.97


def _fake_tel_local(all_read_dat, offset):
    """Stand-in for tg_tel._tel_repeat_comp_local with the same result/meta structure, deterministic per read."""
    from telogator2_b200.tg_util import RC
    res = [[], [], [], 0, 0, 0, [], {}, {}]
    meta = {'tvr': [], 'interstitial': []}
    for i, (name, rdat, _q) in enumerate(all_read_dat):
        flip = (len(rdat) % 3) == 0
        my = RC(rdat) if flip else rdat
        kind = len(rdat) % 5
        if kind in (0, 1, 2):
            res[0].append([[[1, 7]], len(my) // 2, 0, 'q', name.split(' ')[0], 60, my])
            res[1].append(len(my) // 3)
            res[2].append(len(my) % 11)
            meta['tvr'].append((i + offset, flip))
        elif kind == 3:
            a, b = len(my) // 4, len(my) // 2
            res[6].append([(f'interstitial-left_{len(my)}_{name}', my[:a]), (f'interstitial-right_{len(my)}_{name}', my[b:])])
            res[7][name] = [[[0, 5]], len(my), 0, 'q', name.split(' ')[0], 60, my]
            res[8][name] = [[[2, 9]], len(my), 0, 'q', name.split(' ')[0], 60, my]
            meta['interstitial'].append((i + offset, flip))
        else:
            res[3] += 1
            res[4] += len(rdat) % 2
            res[5] += 1
    return res, meta


def test_scan_read_sharding_and_result_merge():
    """Multi-GPU scan: contiguous shares by base count; stripped per-rank results merge to the unsharded result."""
    from telogator2_b200.multi_gpu import shard_reads_by_bases, strip_tel_results, attach_tel_results
    rng = np.random.default_rng(8)
    reads = [(f'read{i} extra', ''.join(rng.choice(list('ACGT'), int(rng.integers(30, 400)))), None) for i in range(57)]
    want, _ = _fake_tel_local(reads, 0)
    for world in (2, 3, 8):
        b = shard_reads_by_bases([len(r[1]) for r in reads], world)
        assert b[0] == 0 and b[-1] == len(reads) and all(x <= y for x, y in zip(b, b[1:]))
        tot = sum(len(r[1]) for r in reads)
        assert max(sum(len(r[1]) for r in reads[b[k]:b[k + 1]]) for k in range(world)) <= tot / world + 400
        parts = [strip_tel_results(*_fake_tel_local(reads[b[k]:b[k + 1]], b[k])) for k in range(world)]
        import pickle
        assert len(pickle.dumps(parts)) < len(pickle.dumps(want))          # no read sequence travels
        assert attach_tel_results(parts, reads) == want
    assert shard_reads_by_bases([1000, 1, 1, 1], 4) == [0, 1, 1, 1, 4]
