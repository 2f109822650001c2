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
    hdr = host[:24].view(np.int32)
    assert hdr[1] == len(T['KMER_LIST']) and hdr[4] == 20 and hdr[3] == 4      # CCCCCC, CCCTCC and the two HOR k-mers
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


def test_compact_positions_map_to_matrix_cells_single_and_many():
    """Host fix-up path of the device epilogue: compact result position -> flat cells (i, j), (j, i) of the output buffer,
    for one matrix and for several matrices in one pair space, over sharded ranges."""
    from itertools import combinations
    from telogator2_b200.dist_engine import DistEngine, fast_ranges
    rng = np.random.default_rng(12)
    for sizes in ([7], [5, 0, 1, 9, 2, 6]):
        sizes = np.array(sizes, dtype=np.int64)
        many = len(sizes) > 1
        lens = [rng.integers(5, 50, size=int(n)) for n in sizes]
        seq_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        order = np.concatenate([seq_off[b] + np.argsort(-lens[b], kind='stable') for b in range(len(sizes))]).astype(np.int32)
        pair_off = np.concatenate([[0], np.cumsum(sizes * (sizes - 1) // 2)]).astype(np.int64)
        out_off = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
        st = dict(n_seq=int(sizes.sum()), order=order, n_mat=len(sizes) if many else 0, sizes=sizes, mat_pair_off=pair_off,
                  mat_seq_off=seq_off, mat_out_off=out_off)
        # brute force: sorted pair space = per matrix, combinations over the ranks; cell indices are local original indices
        want = []
        for b, n in enumerate(sizes):
            for (a, c) in combinations(range(int(n)), 2):
                i, j = order[seq_off[b] + a] - seq_off[b], order[seq_off[b] + c] - seq_off[b]
                want.append((out_off[b] + i * n + j, out_off[b] + j * n + i))
        n_pairs = len(want)
        for world in (1, 3):
            for rank in range(world):
                ranges = fast_ranges(n_pairs, rank, world, min_block=4)
                q_of_pos = np.concatenate([np.arange(b, e) for b, e in ranges]) if ranges else np.zeros(0, dtype=np.int64)
                if len(q_of_pos) == 0:
                    continue
                fij, fji = DistEngine._compact_to_flat(None, st, dict(ranges=ranges), np.arange(len(q_of_pos), dtype=np.int64))
                assert [(int(x), int(y)) for x, y in zip(fij, fji)] == [tuple(int(v) for v in want[q]) for q in q_of_pos]


def _clz32(x):
    return 32 - int(x).bit_length()


def test_warp_sampler_algorithm_equals_sequential_random_sample():
    """Python emulation of mt_sample_warp_kernel's group step (acceptance fixed point over the ballot mask, reads against the
    pre-group pool patched with the group's own earlier writes) against the sequential pool algorithm of random.sample."""
    import random

    def seq_shuffle(pool, draws):
        pool = list(pool); n = len(pool); res = []; c = 0
        for i in range(n):
            nn = n - i
            while True:
                v = draws[c]; c += 1
                r = v >> (_clz32(nn) - 16)
                if r < nn:
                    break
            res.append(pool[r]); pool[r] = pool[nn - 1]
        return res, c

    def warp_shuffle(pool, draws):
        arr = list(pool) + [None] * 8
        ln = len(pool); nn = ln; c = 0; out = [None] * ln
        while nn > 0:
            v = [draws[c + l] for l in range(32)]
            A = 0
            while True:
                Aprev = A; acc = []; rs = []; az = []
                for l in range(32):
                    a = bin(Aprev & ((1 << l) - 1)).count('1'); nl = nn - a
                    ok = False; r = 0
                    if nl > 0:
                        r = v[l] >> (_clz32(nl) - 16); ok = r < nl
                    acc.append(ok); rs.append(r); az.append(a)
                A = sum(1 << l for l in range(32) if acc[l])
                if A == Aprev:
                    break
            S = bin(A).count('1')
            c += A.bit_length() if S == nn else 32
            scr = [0] * 32
            for l in range(32):
                if (A >> l) & 1:
                    scr[az[l]] = rs[l]
            j = [scr[l] if l < S else 0 for l in range(32)]
            vj = [arr[j[l]] if l < S else 0 for l in range(32)]
            vt = [arr[nn - 1 - l] if l < S else 0 for l in range(32)]
            hit = [-1] * 32
            for l in range(S):
                tg = nn - 1 - j[l]
                if l < tg < S:
                    hit[tg] = max(hit[tg], l)
            root = [l if hit[l] < 0 else hit[l] for l in range(32)]
            while True:
                h = [hit[root[l]] for l in range(32)]
                if not any(x >= 0 for x in h):
                    break
                root = [h[l] if h[l] >= 0 else root[l] for l in range(32)]
            tail = [vt[root[l]] for l in range(32)]
            for l in range(S):
                before = [w for w in range(l) if j[w] == j[l]]
                out[ln - nn + l] = tail[before[-1]] if before else vj[l]
            for l in range(S):
                if not any(j[w] == j[l] for w in range(l + 1, S)):
                    arr[j[l]] = tail[l]
            nn -= S
        return out, c

    rng = random.Random(5)
    for trial in range(400):
        n = rng.choice([1, 2, 3, 5, 17, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 500, 1000, 1023, 1025])
        pool = [rng.randrange(256) for _ in range(n)] if trial % 2 else list(range(n))
        draws = [rng.randrange(65536) for _ in range(4 * n + 600)]
        assert seq_shuffle(pool, draws) == warp_shuffle(pool, draws), (trial, n)


def test_prefilter_bit_tricks():
    """The arithmetic of prefilter_count_kernel: the multiply that gathers bit B of four ASCII letters into a nibble, the
    2-bit code of A/C/G/T, the doubled-pattern shortcut and the jump-ahead greedy count, emulated on integers."""
    M32 = 0xffffffff
    for B in (1, 2):
        M = (1 << (28 - B)) | (1 << (29 - B - 8)) | (1 << (30 - B - 16)) | (1 << (31 - B - 24))
        for bits in range(16):
            for noise in (0x00000000, 0x41434754, 0xffffffff, 0x9e3779b9):
                w = noise & ~(0x01010101 << B) & M32
                for k in range(4):
                    if (bits >> k) & 1:
                        w |= 1 << (8 * k + B)
                assert ((((w & (0x01010101 << B)) * M) & M32) >> 28) == bits
    code = {c: ((ord(c) >> 1) & 1) | (((ord(c) >> 2) & 1) << 1) for c in 'ACGT'}
    assert sorted(code.values()) == [0, 1, 2, 3]

    def count_gpu_style(read, pat):
        k = len(pat)
        doubled = k % 2 == 0 and pat[:k // 2] == pat[k // 2:]
        n = len(read)
        half_ok = [all(i + t < n and read[i + t] == pat[t] for t in range(k // 2)) for i in range(n)] if doubled else None
        good = []
        for i in range(n - k + 1):
            m = (half_ok[i] and half_ok[i + k // 2]) if doubled else read[i:i + k] == pat
            good.append(m)
        cnt, nf = 0, 0
        for w0 in range(0, len(good), 32):                       # per 32-position word: drop bits below nf, jump by k
            g = sum(1 << i for i, m in enumerate(good[w0:w0 + 32]) if m)
            if nf > w0:
                g = 0 if nf - w0 >= 32 else g & (M32 << (nf - w0)) & M32
            while g:
                i = (g & -g).bit_length() - 1
                cnt += 1
                nf = w0 + i + k
                g = 0 if i + k >= 32 else g & (M32 << (i + k)) & M32
        return cnt
    rng = np.random.default_rng(2)
    reads = ['CCCTAA' * 50, 'A' * 40, 'CCCTAACCCTA' + 'CCCTAACCCTAA' * 3, ''.join(rng.choice(list('ACGT'), 3000)), 'AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA',
              'CCCTAA' * 7 + 'G' + 'CCCTAA' * 9]
    for pat in ('CCCTAACCCTAA', 'AAAAAAAAAAAA', 'ACACACAC', 'CCCTAAC', 'AA', 'CCCTAAACCCTAAA'):
        for r in reads:
            assert count_gpu_style(r, pat) == r.count(pat), (pat, r[:30])


def test_denoise_stencil_restatement_matches_reference_golden():
    """cv_denoise_kernel's formulation (survival of a block is local; a dropped position takes the letter of its nearest surviving
    neighbours when they match, are mergeable and close) emulated in Python against the reference's block-list algorithm."""
    import gzip
    import json
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'tvr_prep_golden.json.gz'), 'rt') as f:
        cases = json.load(f)['denoise']

    def stencil(v, rep, min_size, gap, dele, mrg):
        if not v:
            return ''
        L = len(v)
        keep = L - (len(v) - len(v.rstrip(rep)))                  # positions >= keep: the run that is never closed
        c = [v[y] if y < keep else None for y in range(L + 1)]
        surv = [False] * (L + 1)
        for y in range(keep):
            s = True
            if c[y] in dele:
                size, k = 1, y - 1
                while size < min_size and k >= 0 and c[k] == c[y]:
                    size += 1; k -= 1
                k = y + 1
                while size < min_size and k <= L and c[k] == c[y]:
                    size += 1; k += 1
                s = size >= min_size
            surv[y] = s
        out = []
        for x in range(L + 1):
            o = rep
            if surv[x]:
                o = c[x]
            elif x < keep:
                yl = next((x - k for k in range(1, gap + 1) if x - k >= 0 and surv[x - k]), None)
                yr = next((x + k for k in range(1, gap + 1) if x + k <= L and surv[x + k]), None)
                if yl is not None and yr is not None and c[yl] == c[yr] and c[yl] in mrg and yr - yl - 1 <= gap:
                    o = c[yl]
            out.append(o)
        return ''.join(out)
    for cse in cases[::3]:
        if len(cse['v']) > 3000:
            continue
        assert stencil(cse['v'], cse['replace'], cse['min_size'], cse['max_gap_fill'], cse['delete'], cse['merge']) == cse['out']


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/telogator2_b200.h must compile as C99 (no C++-isms, no torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    hdr = os.path.join(ROOT, 'include', 'telogator2_b200.h')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-Werror', '-fsyntax-only', '-x', 'c', hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    txt = open(hdr).read()
    assert '#include <torch' not in txt and 'at::' not in txt and 'std::' not in txt and 'Tensor' not in txt


def test_local_only_hides_the_process_group():
    """multi_gpu.local_only(): code that only some ranks run must see a single-process world (bench.py's rank-0 sections)."""
    from telogator2_b200 import multi_gpu
    calls = []
    real = multi_gpu._dist_state
    assert real()[2] == 1                                  # no process group in this test process
    with multi_gpu.local_only():
        assert multi_gpu._LOCAL_ONLY == 1 and multi_gpu._dist_state() == (None, 0, 1)
        with multi_gpu.local_only():
            assert multi_gpu._LOCAL_ONLY == 2
        assert multi_gpu._LOCAL_ONLY == 1
    assert multi_gpu._LOCAL_ONLY == 0
    del calls


def test_shared_column_job_mapping_covers_every_pair_once():
    """Python transcription of describe_pair + the JOBS_PHASE_A_SHCOL job selection of nw_wave_kernel: over arbitrary (even-start)
    pair ranges every pair is aligned exactly once, rows/columns are the pair's two (length-adjusted) sequences in one of the
    two orientations, and a shared job really shares its column sequence and length."""
    rng = np.random.default_rng(31)

    def unrank(q, n):
        r = 0
        while (r + 1) * (2 * n - (r + 1) - 1) // 2 <= q:
            r += 1
        return r, q - r * (2 * n - r - 1) // 2 + r + 1

    for trial in range(30):
        n = int(rng.integers(2, 14))
        lens = rng.integers(1, 40, size=n) * (50 if trial % 3 else 1)
        off = np.concatenate([[0], np.cumsum(lens)])
        order = np.argsort(-lens, kind='stable')
        adjust, minv = bool(trial % 2), bool((trial // 2) % 2)

        def describe(q):
            a, b = unrank(q, n)
            oa, ob = int(order[a]), int(order[b])
            i, j = min(oa, ob), max(oa, ob)
            li, lj = int(lens[i]), int(lens[j])
            if adjust:
                ml = min(li, lj)
                if minv:
                    ml = max(ml, 1000)
                li, lj = min(li, ml), min(lj, ml)
            return dict(off_i=int(off[i]), off_j=int(off[j]), n=li, m=lj, ra=a, a_is_i=(oa == i), i=i, j=j)

        def orient(d):
            return ((d['off_j'], d['m'], d['off_i'], d['n']) if d['a_is_i'] else (d['off_i'], d['n'], d['off_j'], d['m']))
        n_pairs = n * (n - 1) // 2
        q0 = 2 * int(rng.integers(0, n_pairs // 2 + 1))
        q1 = int(rng.integers(q0, n_pairs + 1))
        done, pend, jid, n_jobs = {}, None, 0, (q1 - q0 + 1) // 2
        while True:
            if pend is not None:
                (row, nr, col, m, out) = pend
                pend = None
                jobs = [(row, nr, col, m, out)]
            else:
                if jid >= n_jobs:
                    break
                qa = q0 + 2 * jid
                jid += 1
                da = describe(qa)
                row0, n0, col0, m0 = orient(da)
                jobs = [(row0, n0, col0, m0, qa - q0)]
                if qa + 1 < q1:
                    db = describe(qa + 1)
                    rowb, nb, colb, mb = orient(db)
                    if db['ra'] == da['ra'] and colb == col0 and mb == m0:
                        jobs.append((rowb, nb, col0, m0, qa + 1 - q0))          # shared columns
                    else:
                        pend = (rowb, nb, colb, mb, qa + 1 - q0)
            assert len({(c, m) for (_r, _n, c, m, _o) in jobs}) == 1
            for (row, nr, col, m, out) in jobs:
                assert out not in done
                done[out] = (row, nr, col, m)
        assert sorted(done) == list(range(q1 - q0))
        for out, (row, nr, col, m) in done.items():
            d = describe(q0 + out)
            assert {(row, nr), (col, m)} == {(d['off_i'], d['n']), (d['off_j'], d['m'])}


def test_bench_emits_exactly_one_json_line_on_stdout():
    """bench.py's contract: ONE JSON line on stdout; banners of libraries (NCCL prints its version there) must not join it."""
    import json
    import subprocess
    import sys
    code = ("import sys, os; sys.path.insert(0, %r); import bench; bench.claim_stdout(); print('banner'); os.system('echo child'); "
            "bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout) == {'a': 1} and r.stdout.count('\n') == 1
    assert 'banner' in r.stderr and 'child' in r.stderr


class _NumpyDev:
    """Stand-in for tg_tvr._Dev: the same five device operations answered by the oracle on numpy buffers, so that the HOST
    orchestration of tg_tvr (slicing arithmetic on (offset, length) pairs, the None-slice quirk, string assembly) runs on the CPU."""

    def paint(self, start, end, rk_off, kmer_letters, off0):
        n, K = len(off0) - 1, len(kmer_letters)
        buf = np.zeros(max(int(off0[-1]), 1), dtype=np.uint8)
        for r in range(n):
            hits = [(k, int(start[s]), int(end[s])) for k in range(K) for s in range(int(rk_off[r * K + k]), int(rk_off[r * K + k + 1]))]
            cv = oracle.colorvec_py(hits, int(off0[r + 1] - off0[r]), kmer_letters)
            buf[off0[r]:off0[r + 1]] = np.frombuffer(cv.encode('latin-1'), dtype=np.uint8)
        return buf

    def upload_strings(self, seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.int32)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        return np.frombuffer(''.join(seqs).encode('latin-1') or b'\\0', dtype=np.uint8).copy(), off[:-1].copy(), lens

    @staticmethod
    def _s(buf, o, ln, rev):
        s = buf[int(o):int(o) + int(ln)].tobytes().decode('latin-1')
        return s[::-1] if rev else s

    def boundary(self, buf, off, lens, reverse, letters, win, thresh, above, start=0):
        first, ext = [], []
        for i in range(len(lens)):
            f, e = oracle.density_boundary_pos(self._s(buf, off[i], lens[i], reverse is not None and reverse[i]), letters, win, thresh,
                                               'above' if above else 'below', start)
            first.append(-1 if f is None else f)
            ext.append(-1 if e is None else e)
        return np.array(first, dtype=np.int32), np.array(ext, dtype=np.int32)

    def lead_run(self, buf, off, lens, reverse, cap, letter):
        out = []
        for i in range(len(lens)):
            s = self._s(buf, off[i], lens[i], reverse is not None and reverse[i])
            out.append(min(len(s) - len(s.lstrip(letter)), int(cap[i])))
        return np.array(out, dtype=np.int32)

    def gather(self, src, src_off, dst_off, seg_len, total):
        dst = np.zeros(max(int(total), 1), dtype=np.uint8)
        for so, do, ln in zip(src_off, dst_off, seg_len):
            dst[int(do):int(do) + int(ln)] = src[int(so):int(so) + int(ln)]
        return dst

    def denoise(self, buf, off, lens, replace_char, min_size, max_gap_fill, chars_to_delete, chars_to_merge):
        outs = [oracle.denoise_colorvec_py(self._s(buf, off[i], lens[i], False), replace_char, min_size, max_gap_fill,
                                           chars_to_delete, chars_to_merge) for i in range(len(lens))]
        olen = np.array([len(o) for o in outs], dtype=np.int32)
        ooff = np.concatenate([[0], np.cumsum(olen)]).astype(np.int64)
        return np.frombuffer(''.join(outs).encode('latin-1') or b'\\0', dtype=np.uint8).copy(), ooff[:-1].copy(), olen

    def download(self, buf, total):
        return buf[:int(total)].tobytes()


def test_tvr_prep_host_orchestration_on_cpu(monkeypatch):
    """tg_tvr's host side (everything but the five device operations) against the reference-generated golden vectors."""
    import gzip
    import json
    from telogator2_b200 import tg_tvr
    monkeypatch.setattr(tg_tvr, '_Dev', _NumpyDev)
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'tvr_prep_golden.json.gz'), 'rt') as f:
        G = json.load(f)
    for key, S in G['sets'].items():
        K = len(S['meta_letters'])
        kd = []
        for e in S['kmer_dat']:
            hits = [[] for _ in range(K)]
            for k, a, b in e['hits']:
                hits[k].append([a, b])
            kd.append([hits, e['tlen'], 0, 'FWD', e['name'], 60, ''])
        meta = [[f'k{i}' for i in range(K)], [None] * K, S['meta_letters'], S['meta_flags']]
        for trunc in (1000, 3000):
            got = tg_tvr.tvr_prep(kd, meta, trunc)
            for k, v in S[f'prep_trunc{trunc}'].items():
                assert got[k] == v, (key, trunc, k)
        assert tg_tvr.quick_get_tvrtel_lens(kd, meta) == S['quick_tvrtel_lens']
    for c in G['fdb'][::7]:
        assert tg_tvr.find_density_boundary(c['seq'], c['letters'], c['win'], c['thresh'], thresh_dir=c['dir'], starting_coord=c['start'],
                                            use_lowest_dens=c['lowest']) == (c['left'], c['right'])
    for c in G['denoise'][::11]:
        assert tg_tvr.denoise_colorvec(c['v'], c['replace'], c['min_size'], c['max_gap_fill'], c['delete'], c['merge']) == c['out']
    with pytest.raises(IndexError):
        tg_tvr.tvr_prep([[[[[0, 5]]] + [[] for _ in range(K - 1)], 3, 0, 'FWD', 'x', 60, '']], meta, 1000)     # span beyond tlen


class _T:
    """numpy array posing as a device tensor (.cpu().numpy())."""

    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


class _OraclePackedReads:
    """Stand-in for tg_kmer.PackedReads: the device scans answered by the CPU oracle, so that tg_tel's HOST logic (phase 2,
    slices, result assembly: tg_tel.py:238-389) runs on the CPU."""

    def __init__(self, reads):
        self.reads = list(reads)
        self.n = len(reads)
        self.lens = np.array([len(r) for r in reads], dtype=np.int32)
        self.base_off = np.concatenate([[0], np.cumsum((self.lens.astype(np.int64) + 127) // 128 * 128)])

    def basecount(self, la, lb):
        return _T(np.array([[oracle.base_count(r, la), oracle.base_count(r, lb)] for r in self.reads], dtype=np.int32).reshape(-1, 2))

    def orient(self, flip):
        from conftest import rc_py
        self.reads = [rc_py(r) if f else r for r, f in zip(self.reads, flip)]

    def density_counts(self, la, lb, window):
        outs = []
        for lst in (la, lb):
            flat = np.zeros(int(self.base_off[-1]) + 16, dtype=np.uint8)
            for i, r in enumerate(self.reads):
                c = oracle.density_counts(r, lst, window)
                flat[self.base_off[i]:self.base_off[i] + len(c)] = c
            outs.append(_T(flat))
        return outs[0], outs[1]

    def counts_to_host(self, ca, cb):
        return ca.cpu().numpy(), cb.cpu().numpy()

    def split_counts(self, flat, window):
        return [flat[int(self.base_off[i]):int(self.base_off[i]) + max(int(self.lens[i]) - window, 0)] for i in range(self.n)]

    def hits(self, slice_read, slice_lo, slice_hi, kmer_list, issub):
        return [oracle.nonoverlapping_hits(self.reads[r][lo:hi], kmer_list, issub) for r, lo, hi in zip(slice_read, slice_lo, slice_hi)]


def test_scan_batch_boundary_host_logic_on_cpu(monkeypatch, kmer_tables, golden_reads):
    """tg_tel.get_tel_repeat_comp_parallel's host side against the output of the reference's own driver (tel_batch_golden)."""
    from telogator2_b200 import tg_kmer, tg_tel
    from test_gpu_tel_batch import _tel_inputs, digest_tel_result, make_gtt, reference_golden, tel_cases
    monkeypatch.setattr(tg_kmer, 'PackedReads', _OraclePackedReads)

    def dens(read, kmer_list, w):                          # the per-read function, served from the primed cache only
        hit = tg_kmer._density_cache.get((read, tuple(kmer_list), w))
        assert hit is not None, 'phase 2 asked for a density that phase 1 did not provide'
        return hit
    gold = reference_golden()
    for tkey, filters in tel_cases():
        reads, params = _tel_inputs(kmer_tables, golden_reads, tkey, filters)
        got = tg_tel.get_tel_repeat_comp_parallel(reads, params, max_workers=3, get_terminating_tl=make_gtt(dens))
        assert digest_tel_result(got) == gold[tkey], tkey


def test_host_list_helper_formats_like_the_python_code():
    """chost/tg_hostlists.c (CPython extension, formatting only): nested hit lists and read packing equal the pure-Python code."""
    import importlib
    from telogator2_b200 import build
    build.build_host_ext()
    H = importlib.import_module('telogator2_b200._tg_hostlists')
    rng = np.random.default_rng(5)
    ns, K = 7, 5
    counts = rng.integers(0, 4, ns * K)
    counts[3] = 0
    cuts = np.zeros(ns * K + 1, dtype=np.int64)
    cuts[1:] = np.cumsum(counts)
    pairs = rng.integers(0, 70000, (int(cuts[-1]), 2)).astype(np.int32)
    got = H.hit_lists(pairs, cuts, ns, K)
    pl = pairs.tolist()
    want = [[pl[cuts[q * K + k]:cuts[q * K + k + 1]] for k in range(K)] for q in range(ns)]
    assert got == want and all(type(v) is int for q in got for k in q for sp in k for v in sp)
    assert H.hit_lists(np.zeros((0, 2), dtype=np.int32), np.zeros(1, dtype=np.int64), 0, K) == []
    with pytest.raises(ValueError):
        H.hit_lists(pairs, cuts[::-1].copy(), ns, K)
    with pytest.raises(ValueError):
        H.hit_lists(pairs.astype(np.int64), cuts, ns, K)
    reads = ['ACGT' * 5, '', 'N', 'acgtX', 'ACéGT', 'A中C', 'ACGT' * 40]
    lens = np.array([len(r) for r in reads])
    padded = (lens + 128) // 128 * 128
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum(padded)
    out = np.zeros(int(off[-1]), dtype=np.uint8)
    H.pack_reads(reads, off, out, ord('N'))
    want = b''.join(r.encode('latin-1', 'replace') + b'N' * int(p - len(r)) for r, p in zip(reads, padded))
    assert out.tobytes() == want
    with pytest.raises(ValueError):
        H.pack_reads(reads, off[:-1].copy(), out, ord('N'))
    with pytest.raises(TypeError):
        H.pack_reads([b'ACGT'], off[:2].copy(), out, ord('N'))


def test_linkage_host_half_sort_and_relabel_equal_scipy():
    """tg_linkage's host half (stable sort of the merges by distance + scipy's union-find relabelling, C helper and pure-Python
    fallback) applied to the oracle's merge-ordered list gives scipy's result."""
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import squareform
    import oracle.linkage as OL
    from telogator2_b200 import build, tg_kmer, tg_linkage
    build.build_host_ext()
    rng = np.random.default_rng(8)
    for n, method in ((2, 'ward'), (9, 'single'), (40, 'ward'), (75, 'complete')):
        M = np.round(rng.random((n, n)) * 7) / 7
        M = np.minimum(M, M.T)
        np.fill_diagonal(M, 0)
        y = squareform(M, checks=False)
        saved = OL.label
        OL.label = lambda Z, n: None                      # the oracle's list before relabelling = what the device returns, sorted
        try:
            Zu = OL.linkage(y, method)
        finally:
            OL.label = saved
        want = linkage(y, method)
        Zc = np.ascontiguousarray(Zu.copy())
        tg_linkage._label(Zc, n)
        assert np.array_equal(Zc, want), (n, method)
        saved_mod = tg_kmer._hostlists_mod
        tg_kmer._hostlists_mod = None                     # pure-Python fallback
        try:
            Zp = np.ascontiguousarray(Zu.copy())
            tg_linkage._label(Zp, n)
        finally:
            tg_kmer._hostlists_mod = saved_mod
        assert np.array_equal(Zp, want), (n, method)
