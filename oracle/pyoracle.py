"""ctypes front-end of oracle/tg_oracle.c plus pure-Python restatements that use the
reference's own RNG (CPython ``random``).  TEST INFRASTRUCTURE ONLY (see package docstring).

Reference citations are relative to zstephens/telogator2.
"""
import ctypes as C
import math
import os
import random
import subprocess
from itertools import combinations

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libtg_oracle.so')
_lib = None

AMINO = ['A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M', 'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'Y']  # tg_align.py:14
GAP_CHR = '-'                     # tg_align.py:17
UNKNOWN = 'A'                     # tg_align.py:16
ALPHABET = ''.join(AMINO) + GAP_CHR
MAX_SEQ_DIST = 10.0               # tg_align.py:22
MIN_SEQ_DIST = 0.0001             # tg_align.py:24
MIN_VIABLE_SEQ_LEN = 1000         # tg_align.py:20
IDEN_SCORE_SCALAR_CHEAT = 0.900   # tg_align.py:26


def build(force=False):
    """Compile oracle/tg_oracle.c -> oracle/_build/libtg_oracle.so (gcc, no GPU needed)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'tg_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i32p, i64p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double))
        L.tgo_nw_score.restype = C.c_double
        L.tgo_nw_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_double] * 5
        L.tgo_dist_matrix.restype = C.c_int
        L.tgo_dist_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_double] * 5 + \
            [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.tgo_shuffle_stream.restype = None
        L.tgo_shuffle_stream.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_char_p), i32p, C.POINTER(C.c_void_p)]
        L.tgo_mt_base_table.argtypes = [C.c_void_p]
        L.tgo_mt_state_bytes.restype = C.c_size_t
        L.tgo_mt_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.tgo_mt_genrand.restype = C.c_uint32
        L.tgo_mt_genrand.argtypes = [C.c_void_p]
        L.tgo_finditer.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        L.tgo_base_count.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
        L.tgo_density_counts.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.tgo_nonoverlapping_hits.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.tgo_rc.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        L.tgo_scan_read.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int,
                                    C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# --------------------------------------------------------------------------------------------
# scoring matrix, closed form of the assignment replay in tg_align.py:32-68  
# --------------------------------------------------------------------------------------------
def scoring_matrix(canon_char, which_type='tvr'):
    """21x21 float64 matrix over ALPHABET for which_type in {'tvr','consensus'} (the two types that
    reach aligner.score(); 'msa'/'msa_refinement' are outside the hot path)."""
    A = len(ALPHABET)
    m = np.zeros((A, A))
    ia, ic, ig = ALPHABET.index(UNKNOWN), ALPHABET.index(canon_char), ALPHABET.index(GAP_CHR)
    for x in range(A):
        for y in range(A):
            cx, cy = ALPHABET[x], ALPHABET[y]
            if x == ig or y == ig:
                if x == ig and y == ig:
                    v = 0.
                elif ia in (x, y):
                    v = -2.
                else:
                    v = -4.
            elif x == y:
                v = 5.
                if x == ia:
                    v = -2.     # the '= 2.' at :39 is overwritten by the c1 == UNKNOWN pass of :43
                if x == ic:
                    v = {'tvr': 0., 'consensus': 1.}[which_type]
            else:
                v = -2. if ia in (x, y) else -4.
            m[x, y] = v
    if canon_char == UNKNOWN:
        raise ValueError('canonical letter may not be the unknown letter')
    return m


def encode(seq, matrix_mode):
    """str -> uint8 codes: index into ALPHABET in matrix mode, raw ord() in match/mismatch mode."""
    if matrix_mode:
        lut = np.full(256, 255, dtype=np.uint8)
        for i, ch in enumerate(ALPHABET):
            lut[ord(ch)] = i
        out = lut[np.frombuffer(seq.encode('ascii'), dtype=np.uint8)]
        if (out == 255).any():
            raise ValueError('sequence contains letters not in the alphabet')
        return out
    return np.frombuffer(seq.encode('ascii'), dtype=np.uint8).copy()


class AlignerParams:
    """The attributes of Bio.Align.PairwiseAligner that tvr_distance / score() depend on
    (tg_align.py:71-106).  S is None in match/mismatch mode."""

    def __init__(self, S=None, gap_bool=(True, True), match=5., mismatch=-4., gap=-4.):
        self.S = None if S is None else np.ascontiguousarray(S, dtype=np.float64)
        self.match, self.mismatch = float(match), float(mismatch)
        self.g = float(gap)
        self.gl = float(gap) if gap_bool[0] else 0.0
        self.gr = float(gap) if gap_bool[1] else 0.0


def nw_score(a, b, P):
    """Biopython global score restated (tg_oracle.c: tgo_nw_score). a, b: str."""
    mm = P.S is not None
    ca, cb = encode(a, mm), encode(b, mm)
    if len(ca) == 0 or len(cb) == 0:
        raise ValueError('sequence has zero length')
    return lib().tgo_nw_score(_ptr(ca), len(ca), _ptr(cb), len(cb), _ptr(P.S), 0 if P.S is None else P.S.shape[0],
                              P.match, P.mismatch, P.g, P.gl, P.gr)


def nw_score_py(a, b, P):
    """Independent full-matrix Needleman-Wunsch in pure Python (SURVEY 9-G recurrence);
    cross-checks the C row-DP on small inputs."""
    n, m = len(a), len(b)
    mm = P.S is not None

    def sub(x, y):
        if mm:
            return P.S[ALPHABET.index(x), ALPHABET.index(y)]
        return P.match if x == y else P.mismatch
    H = [[0.0] * (m + 1) for _ in range(n + 1)]
    for j in range(1, m + 1):
        H[0][j] = j * P.gl
    for i in range(1, n + 1):
        H[i][0] = i * P.gl
        for j in range(1, m + 1):
            H[i][j] = max(H[i - 1][j - 1] + sub(a[i - 1], b[j - 1]),
                          H[i - 1][j] + (P.gr if j == m else P.g),
                          H[i][j - 1] + (P.gr if i == n else P.g))
    return H[n][m]


def iden_score(seq_i, seq_j, P):
    """tg_align.py:127-134."""
    if P.S is None:
        return P.match * ((len(seq_i) + len(seq_j)) / 2.)
    d = np.diag(P.S)
    is1 = sum(d[ALPHABET.index(c)] for c in seq_i)
    is2 = sum(d[ALPHABET.index(c)] for c in seq_j)
    return (is1 + is2) / 2.


def tvr_distance_py(tvr_i, tvr_j, P, adjust_lens=True, min_viable=True, randshuffle=3, rng_seed=None, detail=None):
    """tg_align.py:109-151 restated with CPython's own `random` (the reference RNG) and the
    oracle NW score.  `detail` (dict) receives aln / rand scores when given."""
    if rng_seed is not None:
        random.seed(rng_seed)
    if adjust_lens:
        min_len = min(len(tvr_i), len(tvr_j))
        if min_viable:
            min_len = max(min_len, MIN_VIABLE_SEQ_LEN)
        seq_i, seq_j = tvr_i[:min_len], tvr_j[:min_len]
    else:
        seq_i, seq_j = tvr_i, tvr_j
    aln_score = nw_score(seq_i, seq_j, P)
    iden = iden_score(seq_i, seq_j, P)
    if detail is not None:
        detail.update(aln=aln_score, iden=iden, rand=[], n=len(seq_i), m=len(seq_j))
    if aln_score < -(IDEN_SCORE_SCALAR_CHEAT * iden):
        return MAX_SEQ_DIST
    rand_scores = []
    for _ in range(randshuffle):
        si = ''.join(random.sample(seq_i, len(seq_i)))      # tg_util.py:451-452
        sj = ''.join(random.sample(seq_j, len(seq_j)))
        rand_scores.append(nw_score(si, sj, P))
    if detail is not None:
        detail['rand'] = rand_scores
    rand_score_shuffle = np.mean(rand_scores)
    if rand_score_shuffle >= aln_score:
        dist_out = MAX_SEQ_DIST
    else:
        dist_out = min(-np.log((aln_score - rand_score_shuffle) / (iden - rand_score_shuffle)), MAX_SEQ_DIST)
    return max(dist_out, MIN_SEQ_DIST)


def dist_matrix_py(sequences, P, adjust_lens, min_viable, rand_shuffle_count, rng_seed=12345):
    """tg_align.py:154-189 (process pool removed; results are order independent)."""
    n = len(sequences)
    D = np.zeros((n, n), dtype='<f4')
    for p, (i, j) in enumerate(combinations(range(n), 2)):
        D[i, j] = D[j, i] = tvr_distance_py(sequences[i], sequences[j], P, adjust_lens, min_viable,
                                            rand_shuffle_count, rng_seed + p)
    return D


def _concat_codes(sequences, matrix_mode):
    codes = [encode(s, matrix_mode) for s in sequences]
    off = np.zeros(len(codes) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(c) for c in codes])
    flat = np.concatenate(codes) if codes else np.zeros(0, np.uint8)
    return np.ascontiguousarray(flat), off


def dist_matrix_c(sequences, P, adjust_lens, min_viable, rand_shuffle_count, rng_seed=12345,
                  pair_range=None, n_threads=0, want_scores=False):
    """The whole get_dist_matrix_parallel in C (OpenMP over pairs).  Returns D, or
    (D, aln, rnd, cells) when want_scores."""
    n = len(sequences)
    flat, off = _concat_codes(sequences, P.S is not None)
    n_pairs = n * (n - 1) // 2
    D = np.zeros((n, n), dtype='<f4')
    aln = np.zeros(max(n_pairs, 1), dtype=np.int32)
    rnd = np.zeros(max(n_pairs * max(rand_shuffle_count, 1), 1), dtype=np.int32)
    cells = C.c_double(0.0)
    pb, pe = pair_range if pair_range is not None else (0, n_pairs)
    rc = lib().tgo_dist_matrix(_ptr(flat), _ptr(off), n, _ptr(P.S), 0 if P.S is None else P.S.shape[0],
                               P.match, P.mismatch, P.g, P.gl, P.gr, int(adjust_lens), int(min_viable),
                               int(rand_shuffle_count), C.c_int64(rng_seed), pb, pe,
                               _ptr(D), _ptr(aln), _ptr(rnd), C.byref(cells), n_threads)
    if rc != 0:
        raise ValueError('sequence has zero length')
    if want_scores:
        return D, aln[:n_pairs], rnd[:n_pairs * rand_shuffle_count].reshape(n_pairs, -1) if rand_shuffle_count else rnd[:0], cells.value
    return D


def shuffle_stream(seed, seqs):
    """seed once, then shuffle_seq() every string in order -- C MT19937 (pinned vs `random`)."""
    bs = [s.encode('ascii') if isinstance(s, str) else bytes(s) for s in seqs]
    outs = [C.create_string_buffer(max(len(b), 1)) for b in bs]
    arr_in = (C.c_char_p * len(bs))(*bs)
    arr_out = (C.c_void_p * len(bs))(*[C.cast(o, C.c_void_p) for o in outs])
    lens = np.array([len(b) for b in bs], dtype=np.int32)
    lib().tgo_shuffle_stream(C.c_uint64(seed), len(bs), arr_in, lens.ctypes.data_as(C.POINTER(C.c_int32)), arr_out)
    return [o.raw[:len(b)].decode('ascii') for o, b in zip(outs, bs)]


def mt_base_table():
    t = np.zeros(624, dtype=np.uint32)
    lib().tgo_mt_base_table(_ptr(t))
    return t


# --------------------------------------------------------------------------------------------
# k-mer scan
# --------------------------------------------------------------------------------------------
def _kmer_blob(kmer_list):
    blob = ''.join(kmer_list).encode('ascii')
    off = np.zeros(len(kmer_list) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(k) for k in kmer_list])
    return blob, off


def finditer_starts(seq, kmer):
    b = seq.encode('ascii')
    out = np.zeros(len(b) // max(len(kmer), 1) + 1, dtype=np.int32)
    n = lib().tgo_finditer(b, len(b), kmer.encode('ascii'), len(kmer), _ptr(out))
    return out[:n].tolist()


def base_count(read, kmer_list):
    """get_telomere_base_count (tg_kmer.py:93-102)."""
    blob, off = _kmer_blob(kmer_list)
    b = read.encode('ascii')
    return int(lib().tgo_base_count(b, len(b), blob, _ptr(off), len(kmer_list)))


def density_counts(read, kmer_list, window):
    """integer numerators of get_telomere_kmer_density (tg_kmer.py:66-90): dens = counts / window."""
    blob, off = _kmer_blob(kmer_list)
    b = read.encode('ascii')
    cnt = np.zeros(max(len(b) - window, 0), dtype=np.int32)
    lib().tgo_density_counts(b, len(b), blob, _ptr(off), len(kmer_list), window, _ptr(cnt))
    return cnt


def kmer_issubstring(kmer_list):
    """tg_kmer.py:44-46."""
    return [[j for j in range(len(kmer_list)) if (j != i and kmer_list[i] in kmer_list[j])] for i in range(len(kmer_list))]


def nonoverlapping_hits(seq, kmer_list, issub):
    """get_nonoverlapping_kmer_hits (tg_kmer.py:173-230) -> list[K] of [[start, end], ...]."""
    blob, off = _kmer_blob(kmer_list)
    b = seq.encode('ascii')
    flat = np.array([j for l in issub for j in l], dtype=np.int32)
    ioff = np.zeros(len(issub) + 1, dtype=np.int32)
    ioff[1:] = np.cumsum([len(l) for l in issub])
    cap = len(b) + 16
    ok, os_, oe = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    cnt = np.zeros(len(kmer_list), dtype=np.int32)
    n = lib().tgo_nonoverlapping_hits(b, len(b), blob, _ptr(off), len(kmer_list), _ptr(flat), _ptr(ioff),
                                      _ptr(ok), _ptr(os_), _ptr(oe), cap, _ptr(cnt))
    assert n >= 0
    out = [[] for _ in kmer_list]
    for q in range(n):
        out[ok[q]].append([int(os_[q]), int(oe[q])])
    return out


def rc(seq):
    b = seq.encode('ascii')
    o = C.create_string_buffer(max(len(b), 1))
    if lib().tgo_rc(b, len(b), o) != 0:
        raise KeyError('non-ACGTN character')
    return o.raw[:len(b)].decode('ascii')


def scan_read(read, kmers_f, kmers_r, canon_f, canon_r, window):
    """scan half of gtrc_parallel_job (tg_tel.py:262-266,160-161) for one read, in C."""
    bf, off = _kmer_blob(kmers_f)
    br, _ = _kmer_blob(kmers_r)
    cf, coff = _kmer_blob(canon_f)
    cr, _ = _kmer_blob(canon_r)
    b = read.encode('ascii')
    L = len(b)
    ori = C.create_string_buffer(max(L, 1))
    c0 = np.zeros(max(L - window, 0), dtype=np.int32)
    c1 = np.zeros(max(L - window, 0), dtype=np.int32)
    bc = np.zeros(2, dtype=np.int32)
    fl = lib().tgo_scan_read(b, L, bf, br, _ptr(off), len(kmers_f), cf, cr, _ptr(coff), len(canon_f), window,
                             ori, _ptr(c0), _ptr(c1), _ptr(bc))
    return bool(fl), ori.raw[:L].decode('ascii'), c0, c1, bc


# =====================================================================================================
# Colour-vector preparation: restatement of the per-read preparation inside cluster_tvrs
# (tg_tvr.py:96-142) and quick_get_tvrtel_lens (tg_tvr.py:558-592).  numpy / pure Python, small inputs only.
# Pinned against the reference's own outputs: tests/golden/tvr_prep_golden.json.gz (make_golden_tvr.py).
# =====================================================================================================
UNKNOWN_WIN_SIZE, UNKNOWN_END_DENS = 100, 0.125              # tg_tvr.py:22-23
UNKNOWN_WIN_SIZE_FINE, UNKNOWN_END_DENS_FINE = 50, 0.500     # tg_tvr.py:24-25
CANON_WIN_SIZE, CANON_END_DENS = 100, 0.700                  # tg_tvr.py:27-28


def colorvec_py(hits_flat, tlen, kmer_letters):
    """tg_tvr.py:96-105: UNKNOWN everywhere, then every span painted with its k-mer's letter, k-mers in index order
    (a later k-mer overwrites an earlier one).  hits_flat: iterable of (k, start, end)."""
    v = np.full(tlen, ord(UNKNOWN), dtype=np.uint8)
    for k, s, e in sorted((tuple(int(x) for x in h) for h in hits_flat), key=lambda h: h[0]):
        v[s:e] = ord(kmer_letters[k])
    return v.tobytes().decode('latin-1')


def density_boundary_pos(seq, letters, win, thresh, direction='below', start=0):
    """Positions (first crossing or None, extreme-density position or None) of find_density_boundary (tg_tvr.py:449-479):
    dens[j] = #(seq[j+1 .. j+win] in letters, positions >= start) / win for start <= j < len - win."""
    if direction not in ('below', 'above'):
        raise ValueError('thresh_dir must be above or below')
    L = len(seq)
    if L - win <= start:
        return None, None
    a = np.frombuffer(seq.encode('latin-1'), dtype=np.uint8)
    u = np.isin(a, np.frombuffer(''.join(letters).encode('latin-1'), dtype=np.uint8)).astype(np.float64)
    u[:start] = 0
    cum = np.cumsum(u)
    j = np.arange(start, L - win)
    dens = (cum[j + win] - cum[j]) / win                      # float64 division like the reference
    if direction == 'below':
        hit = np.nonzero(dens <= thresh)[0]
        ext = int(np.argmin(dens)) if dens.min() < 1.0 else None       # first position of the strict running minimum
    else:
        hit = np.nonzero(dens >= thresh)[0]
        ext = int(np.argmax(dens)) if dens.max() > 0.0 else None
    first = int(hit[0]) + start if len(hit) else None
    ext = ext + start if ext is not None else None
    return first, ext


def find_density_boundary_py(seq, letters, win, thresh, direction='below', start=0, use_lowest=False):
    """tg_tvr.py:449-498 -> (left, right).  A None split position slices to (whole, whole), as in the reference."""
    first, ext = density_boundary_pos(seq, letters, win, thresh, direction, start)
    if first is None:
        first = ext
    p = ext if use_lowest else first
    return seq[:p], seq[p:]


def denoise_colorvec_py(v, replace_char=UNKNOWN, min_size=10, max_gap_fill=50, chars_to_delete=(), chars_to_merge=()):
    """tg_tvr.py:524-556.  Runs of equal letters of v + replace_char; the LAST run is never closed, i.e. dropped; runs
    shorter than min_size of a deletable letter are dropped; neighbouring surviving runs of the same mergeable letter
    at most max_gap_fill apart are joined (filling the gap); everything else becomes replace_char.  The output is one
    letter longer than v."""
    if len(v) == 0:
        return ''
    w = v + replace_char
    a = np.frombuffer(w.encode('latin-1'), dtype=np.uint8)
    starts = np.concatenate([[0], np.nonzero(a[1:] != a[:-1])[0] + 1])
    ends = np.concatenate([starts[1:], [len(a)]])
    runs = [(int(s), int(e), chr(a[s])) for s, e in zip(starts[:-1], ends[:-1])]          # last run dropped
    runs = [r for r in runs if not (r[1] - r[0] < min_size and r[2] in chars_to_delete)]
    out = np.full(len(a), ord(replace_char), dtype=np.uint8)
    for i, (s, e, c) in enumerate(runs):
        out[s:e] = ord(c)
        if i and runs[i - 1][2] == c and c in chars_to_merge and s - runs[i - 1][1] <= max_gap_fill:
            out[runs[i - 1][1]:s] = ord(c)
    return out.tobytes().decode('latin-1')


def _flag_letters(meta):
    kmer_list, _colors, letters, flags = meta
    sets = {f: sorted(set(letters[i] for i in range(len(kmer_list)) if f in flags[i]))
            for f in ('canonical', 'denoise', 'tvr', 'nofilt', 'dubious', 'subtel-filt')}
    canon = None
    for i in range(len(kmer_list)):
        if 'canonical' in flags[i]:
            canon = letters[i]                 # the last one wins (tg_tvr.py:67-69)
    return canon, sets


def subtel_tvr_split_py(cv, unknown_letters):
    """tg_tvr.py:117-128 / :580-591 -> (seq_left, seq_right)."""
    left, right = find_density_boundary_py(cv, unknown_letters, UNKNOWN_WIN_SIZE, UNKNOWN_END_DENS, 'below')
    rec, rest = find_density_boundary_py(left[::-1], unknown_letters, UNKNOWN_WIN_SIZE_FINE, UNKNOWN_END_DENS_FINE, 'above')
    right = rec[::-1] + right
    left = rest[::-1]
    adj = 0
    while right[adj] == UNKNOWN and adj < len(right) - 1:
        adj += 1
    return left + right[:adj], right[adj:]


def tvr_prep_py(kmer_dat, meta, tvr_truncate):
    """The lists cluster_tvrs builds before the distance matrix (tg_tvr.py:96-142)."""
    canon, sets = _flag_letters(meta)
    unk = [UNKNOWN] + sets['subtel-filt']
    out = {k: [] for k in ('all_colorvecs', 'subtel_regions', 'tvrtel_regions', 'cleaned_colorvecs', 'colorvecs_for_msa',
                           'err_end_lens', 'tvrtels_for_clustering')}
    for kd in kmer_dat:
        cv = colorvec_py(kd['hits'], kd['tlen'], meta[2])
        left, right = subtel_tvr_split_py(cv, unk)
        den = denoise_colorvec_py(right, replace_char=canon, chars_to_delete=sets['denoise'], chars_to_merge=[canon] + sets['tvr'])
        err_left, err_right = find_density_boundary_py(den[::-1], [canon], CANON_WIN_SIZE, CANON_END_DENS, 'above')
        out['all_colorvecs'].append(cv)
        out['cleaned_colorvecs'].append(left + err_right[::-1])
        out['err_end_lens'].append(len(err_left))
        out['colorvecs_for_msa'].append(err_right[::-1])
        out['subtel_regions'].append(left)
        out['tvrtel_regions'].append(right)
        out['tvrtels_for_clustering'].append(err_right[::-1][:tvr_truncate])
    return out


def quick_tvrtel_lens_py(kmer_dat, meta):
    """tg_tvr.py:558-592."""
    _canon, sets = _flag_letters(meta)
    unk = [UNKNOWN] + sets['subtel-filt']
    return [len(subtel_tvr_split_py(colorvec_py(kd['hits'], kd['tlen'], meta[2]), unk)[1]) for kd in kmer_dat]


# =====================================================================================================
# Stand-in for Bio.Align.PairwiseAligner that lets the REFERENCE's own tvr_distance / get_dist_matrix_parallel run here
# (tests/golden/make_golden_dp_ref.py): everything of tg_align.py:109-189 is then the reference's code -- length adjust,
# Counter-based identity score, early-out, random.seed + shuffle_seq order, np.mean, -log ratio, clamps, seeds per pair,
# float32 matrix fill -- and only aligner.score() (Biopython's C, absent) is answered by this oracle's NW restatement.
# =====================================================================================================
class _LetterMatrix:
    """aligner.substitution_matrix[n, n] with letter keys (tg_align.py:131-132)."""

    def __init__(self, S):
        self.S = np.ascontiguousarray(S, dtype=np.float64)

    def __getitem__(self, key):
        a, b = key
        return self.S[ALPHABET.index(a), ALPHABET.index(b)]


class ScoreStandInAligner:
    """Picklable (the reference submits it to a ProcessPoolExecutor)."""

    def __init__(self, S=None, gap_bool=(True, True), match=5., mismatch=-4., gap=-4.):
        self._args = (None if S is None else np.array(S, dtype=np.float64), tuple(gap_bool), match, mismatch, gap)
        self.substitution_matrix = None if S is None else _LetterMatrix(S)
        self.match_score = float(match)
        self.mismatch_score = float(mismatch)

    def params(self):
        S, gap_bool, match, mismatch, gap = self._args
        return AlignerParams(S, gap_bool, match, mismatch, gap)

    def score(self, a, b):
        return float(nw_score(a, b, self.params()))
