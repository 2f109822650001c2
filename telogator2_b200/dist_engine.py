"""All-vs-all TVR distance on the GPU: host orchestration (numpy/torch) over the C-ABI kernels.

Mirrors tvr_distance / get_dist_matrix_parallel of the reference (tg_align.py:109-189):
  device : NW scores (tg_nw_scores_wave / tg_nw_scores_generic), MT19937 shuffles (tg_mt_shuffle)
  host   : the float64 epilogue in the reference's expression order (identity score, early-out
           test, mean, -log ratio, clamps, float32 store) -- integers in, floats out.
"""
import ctypes as C

import os

import numpy as np

from . import _lib

MIN_VIABLE_SEQ_LEN = 1000           # tg_align.py:20
MAX_SEQ_DIST = 10.0                 # tg_align.py:22
MIN_SEQ_DIST = 0.0001               # tg_align.py:24
IDEN_SCORE_SCALAR_CHEAT = 0.900     # tg_align.py:26
NOT_COMPUTED = np.iinfo(np.int32).min
MAX_WAVE32_LEN = 20000             # shared-memory budget of the wave and sampling kernels (rows + pools per warp)


def _as_int(x, what):
    xi = int(round(float(x)))
    if float(x) != float(xi):
        raise NotImplementedError(f'{what}={x!r}: the GPU path needs integer-valued scores (all reference call sites are)')
    return xi


def _gap_attr(aligner, side, kind):
    """Biopython >= 1.86 names (open_left_gap_score) with the 1.82-1.85 fallback (left_open_gap_score),
    the same two spellings get_aligner_object tries (tg_align.py:91-97)."""
    if side == 'internal':
        names = [f'{kind}_internal_gap_score', f'internal_{kind}_gap_score']
    else:
        names = [f'{kind}_{side}_gap_score', f'{side}_{kind}_gap_score']
    for nm in names:
        try:
            return getattr(aligner, nm)
        except (AttributeError, ValueError):
            continue
    raise AttributeError(f'aligner has no {side} {kind} gap score')


class Scoring:
    """Integer view of a configured aligner: what Bio.Align.PairwiseAligner.score() depends on."""

    def __init__(self, sub, alphabet, match, mismatch, gap, gap_left, gap_right):
        self.sub = None if sub is None else np.asarray(sub)
        self.alphabet = alphabet            # str (matrix mode) or None (match/mismatch mode)
        self.match, self.mismatch = match, mismatch
        self.gap, self.gap_left, self.gap_right = gap, gap_left, gap_right

    @classmethod
    def from_aligner(cls, aligner):
        if getattr(aligner, 'mode', 'global') != 'global':
            raise NotImplementedError('only global alignment is on the hot path')
        g = {}
        for side in ('internal', 'left', 'right'):
            o, e = _gap_attr(aligner, side, 'open'), _gap_attr(aligner, side, 'extend')
            if o != e:
                raise NotImplementedError('affine gaps (open != extend) never reach aligner.score() in the reference; '
                                          'the GPU path implements the linear-gap Needleman-Wunsch score')
            g[side] = _as_int(o, f'{side} gap score')
        sm = aligner.substitution_matrix
        if sm is None:
            return cls(None, None, _as_int(aligner.match_score, 'match_score'),
                       _as_int(aligner.mismatch_score, 'mismatch_score'), g['internal'], g['left'], g['right'])
        arr = np.asarray(sm, dtype=np.float64)
        if not np.array_equal(arr, np.rint(arr)):
            raise NotImplementedError('non-integer substitution matrix')
        return cls(arr.astype(np.int64), str(sm.alphabet), None, None, g['internal'], g['left'], g['right'])

    # ---- encoding -------------------------------------------------------------------------
    def encode_all(self, sequences):
        """-> (codes uint8 concatenated, off int64[n+1], sub int64[A, A])."""
        for s in sequences:
            if len(s) == 0:
                raise ValueError('sequence has zero length')       # Biopython raises the same from score()
        raw_bytes = ''.join(sequences).encode('latin-1')
        raw = np.frombuffer(raw_bytes, dtype=np.uint8)
        off = np.zeros(len(sequences) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in sequences])
        lut = np.full(256, 255, dtype=np.uint8)
        if self.alphabet is not None:
            for i, ch in enumerate(self.alphabet):
                lut[ord(ch)] = i
            A = len(self.alphabet)
            sub = self.sub
        else:
            present = np.flatnonzero(np.bincount(raw, minlength=256))
            A = len(present)
            lut[present] = np.arange(A, dtype=np.uint8)
            sub = np.full((A, A), self.mismatch, dtype=np.int64)
            np.fill_diagonal(sub, self.match)
        if A > _lib.MAXA:
            raise NotImplementedError(f'alphabet of {A} symbols exceeds the {_lib.MAXA} the kernels support')
        tb = raw_bytes.translate(lut.tobytes())                  # bytes.translate / `in`: C speed
        if b'\xff' in tb:
            raise ValueError('sequence contains letters not in the alphabet')
        # keep only the letters that occur: the wave kernel's shared-memory profile has one row per letter, so a
        # smaller alphabet means more resident warps
        present = np.array([k for k in range(A) if bytes([k]) in tb], dtype=np.int64)
        if len(present) < A:
            remap = np.zeros(256, dtype=np.uint8)
            remap[present] = np.arange(len(present), dtype=np.uint8)
            tb = tb.translate(remap.tobytes())
            sub = np.ascontiguousarray(sub[np.ix_(present, present)])
        codes = np.frombuffer(bytearray(tb), dtype=np.uint8)
        return np.ascontiguousarray(codes), off, sub

    def c_struct(self, sub):
        sc = _lib.NwScoring()
        A = sub.shape[0]
        sc.alphabet, sc.gap, sc.gap_left, sc.gap_right = A, self.gap, self.gap_left, self.gap_right
        full = np.zeros((_lib.MAXA, _lib.MAXA), dtype=np.int16)
        full[:A, :A] = sub
        C.memmove(sc.sub, full.ctypes.data, full.nbytes)
        return sc


def pair_index_rows(n, i0, i1):
    """(I, J, P) for rows i0 <= i < i1 of itertools.combinations(range(n), 2) (tg_align.py:156)."""
    i = np.arange(i0, i1, dtype=np.int64)
    cnt = n - 1 - i
    I = np.repeat(i, cnt)
    start = i * (2 * n - i - 1) // 2
    P = np.arange(int(cnt.sum()), dtype=np.int64) + (start[0] if len(start) else 0)
    J = P - np.repeat(start, cnt) + I + 1
    return I, J, P


def pack_jobs(row_off, col_off, n, m, out, allow_mixed):
    """Group alignments two by two into tg_nw_job records, largest first.  Equal shapes are paired
    first; left-overs are paired with each other when allow_mixed, else run with a duplicated half."""
    cnt = len(n)
    if cnt == 0:
        return np.zeros(0, dtype=_lib.NW_JOB)
    order = np.lexsort((m, n))[::-1]
    ns, ms = n[order], m[order]
    new_run = np.ones(cnt, dtype=bool)
    new_run[1:] = (ns[1:] != ns[:-1]) | (ms[1:] != ms[:-1])
    run_start = np.maximum.accumulate(np.where(new_run, np.arange(cnt), 0))
    pos = np.arange(cnt) - run_start
    has_next = np.zeros(cnt, dtype=bool)
    has_next[:-1] = ~new_run[1:]
    first = np.nonzero((pos % 2 == 0) & has_next)[0]            # pairs (k, k+1) inside a run
    single = np.nonzero((pos % 2 == 0) & ~has_next)[0]          # last of an odd run
    a_idx, b_idx = [first], [first + 1]
    if allow_mixed and len(single) > 1:
        k = len(single) // 2 * 2
        a_idx.append(single[0:k:2]); b_idx.append(single[1:k:2])
        single = single[k:]
    a = np.concatenate(a_idx); b = np.concatenate(b_idx)
    jobs = np.zeros(len(a) + len(single), dtype=_lib.NW_JOB)
    ia, ib = order[a], order[b]
    k = len(a)
    jobs['row_off'][:k, 0], jobs['row_off'][:k, 1] = row_off[ia], row_off[ib]
    jobs['col_off'][:k, 0], jobs['col_off'][:k, 1] = col_off[ia], col_off[ib]
    jobs['n'][:k, 0], jobs['n'][:k, 1] = n[ia], n[ib]
    jobs['m'][:k, 0], jobs['m'][:k, 1] = m[ia], m[ib]
    jobs['out'][:k, 0], jobs['out'][:k, 1] = out[ia], out[ib]
    if len(single):
        s = order[single]
        for h in (0, 1):
            jobs['row_off'][k:, h], jobs['col_off'][k:, h] = row_off[s], col_off[s]
            jobs['n'][k:, h], jobs['m'][k:, h] = n[s], m[s]
        jobs['out'][k:, 0], jobs['out'][k:, 1] = out[s], -1
    # largest first (dynamic scheduling on the device)
    cost = np.maximum(jobs['n'][:, 0], jobs['n'][:, 1]).astype(np.int64) * np.maximum(jobs['m'][:, 0], jobs['m'][:, 1])
    return jobs[np.argsort(-cost, kind='stable')]


def choose_prof_rows(seq_mask, sub_rows, max_len, R, n_pairs):
    """Rows of the per-job query profiles of the null-model alignments (tg_dist_batch_args.prof_rows), or 0 for the full
    alphabet.  A job fits when its row sequences use fewer letters than rows; fewer rows = more resident warps for the jobs
    that fit, the others run in the full-alphabet launch.  Picks the value with the best expected issue rate (resident
    warps w hide the dependent chain roughly like w / (w + 5.3): measured 0.63 at 9 warps, 0.75 at 16)."""
    env = os.environ.get('TG_WAVE_PROF_ROWS')
    if env is not None:
        return max(int(env), 0)
    if R <= 0 or len(seq_mask) == 0 or n_pairs * R < 40000:
        return 0                    # few jobs: two launches that each leave the GPU half empty cost more than they gain (measured: 8.7 vs 5.8 ms at 1128 pairs)
    cnt = np.zeros(33, dtype=np.int64)
    um, uc = np.unique(seq_mask, return_counts=True)
    for m, c in zip(um, uc):
        cnt[bin(int(m)).count('1')] += c
    frac_le = np.cumsum(cnt) / max(cnt.sum(), 1)                 # share of sequences with <= k letters
    row_cap = (max_len + 140 + 7) // 8 * 8

    def warps(rows):
        return min(16, (226 * 1024 - 1024) // (2 * rows * 512 + 2 * row_cap + 512))

    def rate(w):
        return w / (w + 5.3)
    full = sub_rows + 1
    best, best_score = 0, rate(warps(full))
    for rows in range(4, full):
        f = float(frac_le[rows - 1]) ** (1 if R % 2 == 0 else 2)  # R odd: two pairs (two row sequences) share a job
        score = f * rate(warps(rows)) + (1.0 - f) * rate(warps(full))
        if score > best_score * 1.01:
            best, best_score = rows, score
    return best


class DistEngine:
    """Device state for one get_dist_matrix_parallel call."""

    def __init__(self, scoring, device=None, force_generic=False, pairs_per_batch=1 << 21,
                 shuffle_bytes_per_batch=8 << 30, single_kernel_shuffle=False, margin_log2=46, ambig_cap=1 << 16):
        self.torch = _lib.require_cuda()
        self.L = _lib.load()
        self.sc = scoring
        self.device = self.torch.device('cuda', self.torch.cuda.current_device()) if device is None else device
        self.force_generic = force_generic
        self.pairs_per_batch = pairs_per_batch
        self.shuffle_bytes_per_batch = shuffle_bytes_per_batch
        self.single_kernel_shuffle = single_kernel_shuffle or bool(int(__import__('os').environ.get('TG_SINGLE_KERNEL_SHUFFLE', '0')))
        self.stats = {'cells': 0, 'wave_jobs': 0, 'wave32_jobs': 0, 'generic_jobs': 0, 'launches': 0, 'pairs': 0, 'early_out': 0,
                      'h2d_bytes': 0, 'd2h_bytes': 0}
        # device epilogue: a distance is recomputed on the host when moving log() by 2^-margin_log2 (relative; 64 ulp by
        # default, far beyond any libm's error) could change its float32 value -- about 1 pair in 10^6
        self.margin_log2 = margin_log2
        self.ambig_cap = ambig_cap
        self._scratch = None
        self._counter = None
        self._side_stream = None

    # ---- device helpers --------------------------------------------------------------------
    def _to_dev(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
        self.stats['h2d_bytes'] += t.numel()
        return t.to(self.device, non_blocking=False)

    def _wave_ok_mask(self, jobs, sub):
        """(jobs the 16-bit wave kernel takes, jobs the 32-bit wave kernel takes); the rest goes to the generic kernel."""
        sc = self.sc
        none = np.zeros(len(jobs), dtype=bool)
        if self.force_generic or sc.gap_left != sc.gap or sc.gap_right < sc.gap:
            return none, none
        sp = sub - 2 * sc.gap
        if sp.min() < 0 or sp.max() > 127 or sub.shape[0] >= _lib.MAXA:
            return none, none
        N = np.maximum(jobs['n'][:, 0], jobs['n'][:, 1]).astype(np.int64)
        M = np.maximum(jobs['m'][:, 0], jobs['m'][:, 1]).astype(np.int64)
        # (the row letters of a job live in shared memory: rows beyond MAX_WAVE32_LEN go to the generic kernel in both widths)
        ok16 = (int(sp.max()) * np.minimum(N, M) <= 65535) & (N <= MAX_WAVE32_LEN)
        ok32 = ~ok16 & (N <= MAX_WAVE32_LEN)
        return ok16, ok32

    def run_jobs(self, d_pool, jobs, sub, n_out):
        """Launch the NW kernels for `jobs` (host record array); returns int32 scores on the host."""
        torch = self.torch
        d_scores = torch.full((max(n_out, 1),), NOT_COMPUTED, dtype=torch.int32, device=self.device)
        if len(jobs) == 0:
            return d_scores[:n_out].cpu().numpy()
        cstruct = self.sc.c_struct(sub)
        ok16, ok32 = self._wave_ok_mask(jobs, sub)
        gen = jobs[~(ok16 | ok32)]
        st = _lib.stream_ptr(torch)
        for wave, fn, key in ((jobs[ok16], self.L.tg_nw_scores_wave, 'wave_jobs'), (jobs[ok32], self.L.tg_nw_scores_wave32, 'wave32_jobs')):
            if not len(wave):
                continue
            max_n = int(max(wave['n'].max(), 1))
            need = self.L.tg_nw_wave_scratch_bytes(max_n)
            if self._scratch is None or self._scratch.numel() < need:
                self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
            if self._counter is None:
                self._counter = torch.zeros(2, dtype=torch.int32, device=self.device)
            d_jobs = self._to_dev(wave)
            _lib.check(fn(_lib.dptr(d_pool), _lib.dptr(d_jobs), len(wave), max_n, C.byref(cstruct),
                          _lib.dptr(d_scores), _lib.dptr(self._scratch), _lib.dptr(self._counter), st))
            self.stats[key] += len(wave)
            self.stats['launches'] += 1
        if len(gen):
            max_m = int(gen['m'].max())
            d_jobs_g = self._to_dev(gen)
            _lib.check(self.L.tg_nw_scores_generic(_lib.dptr(d_pool), _lib.dptr(d_jobs_g), len(gen), max_m, C.byref(cstruct),
                                                   _lib.dptr(d_scores), st))
            self.stats['generic_jobs'] += len(gen)
            self.stats['launches'] += 1
        out = d_scores[:n_out].cpu().numpy()
        self.stats['d2h_bytes'] += out.nbytes
        return out

    # ---- the batch path ----------------------------------------------------------------------
    # ---- device-driven path: one tg_dist_batch call per pair range, no host round trips ----------
    def _fast_ok(self, sub, max_len):
        sc = self.sc
        if self.force_generic or sc.gap_left != sc.gap or sc.gap_right < sc.gap or sub.shape[0] >= _lib.MAXA:
            return False
        sp = sub - 2 * sc.gap
        # beyond 16-bit scores (int(sp.max()) * max_len > 65535) the library switches to 32-bit cells by itself
        return sp.min() >= 0 and sp.max() <= 127 and max_len <= MAX_WAVE32_LEN

    def prepare(self, sequences, many=False):
        """Encode the sequences and make them resident on the device.  Returns the state consumed by matrix_device() /
        _pair_scores_fast(), or None if this scoring/length needs the explicit-job path (pair_scores).
        many=True: `sequences` is a list of sequence lists, one independent matrix each (the per-cluster calls of
        tg_tvr.py:147-159 / 375-390 batched into one pair space)."""
        groups = [list(g) for g in sequences] if many else [sequences]
        if many:
            # a group without a pair is never aligned upstream (zeros((n, n)), tg_align.py:158): its sequences may be
            # anything, including empty -- keep its size, not its letters
            groups = [g if len(g) >= 2 else ['A'] * len(g) for g in groups]
        flat = [s for g in groups for s in g]
        sizes = np.array([len(g) for g in groups], dtype=np.int64)
        if int((sizes * (sizes - 1) // 2).sum()) == 0:
            return None
        codes, off, sub = self.sc.encode_all(flat)
        lens = np.diff(off)
        max_len = int(lens.max())
        if not self._fast_ok(sub, max_len):
            return None
        torch = self.torch
        mat_seq_off = np.zeros(len(groups) + 1, dtype=np.int32)
        np.cumsum(sizes, out=mat_seq_off[1:])
        # each matrix's sequences sorted by length, longest first (global indices)
        order = np.concatenate([mat_seq_off[b] + np.argsort(-lens[mat_seq_off[b]:mat_seq_off[b + 1]], kind='stable')
                                for b in range(len(groups))]).astype(np.int32)
        st = dict(n_seq=len(off) - 1, n_codes=len(codes), max_len=max_len, order=order, sub=sub, off=off,
                  d_codes=torch.from_numpy(codes).to(self.device), d_off=torch.from_numpy(off).to(self.device),
                  d_order=torch.from_numpy(order).to(self.device), d_cs=None, cstruct=self.sc.c_struct(sub), pool=None,
                  n_mat=0, sizes=sizes, n_pairs=int((sizes * (sizes - 1) // 2).sum()), out_elems=int((sizes * sizes).sum()))
        if many:
            mat_pair_off = np.zeros(len(groups) + 1, dtype=np.int64)
            np.cumsum(sizes * (sizes - 1) // 2, out=mat_pair_off[1:])
            mat_out_off = np.zeros(len(groups) + 1, dtype=np.int64)
            np.cumsum(sizes * sizes, out=mat_out_off[1:])
            st.update(n_mat=len(groups), mat_seq_off=mat_seq_off, mat_pair_off=mat_pair_off, mat_out_off=mat_out_off,
                      d_mat_seq_off=torch.from_numpy(mat_seq_off).to(self.device),
                      d_mat_pair_off=torch.from_numpy(mat_pair_off).to(self.device),
                      d_mat_out_off=torch.from_numpy(mat_out_off).to(self.device))
        self.stats['h2d_bytes'] += len(codes) + off.nbytes + order.nbytes
        # letters of every sequence (bit c = code c occurs): the null-model alignments run with per-job query profiles
        if len(codes):
            bits = np.left_shift(np.uint32(1), codes.astype(np.uint32))
            mask = np.bitwise_or.reduceat(bits, np.minimum(off[:-1], len(codes) - 1)).astype(np.uint32)
            mask[lens == 0] = 0
        else:
            mask = np.zeros(len(off) - 1, dtype=np.uint32)
        st['seq_mask'] = mask
        st['d_seq_mask'] = torch.from_numpy(mask).to(self.device)
        if self.sc.alphabet is not None:
            # prefix sums of the substitution-matrix diagonal over the codes (identity score, tg_align.py:127-134)
            diag = torch.from_numpy(np.ascontiguousarray(np.diag(sub)).astype(np.int32)).to(self.device)
            cs = torch.zeros(len(codes) + 1, dtype=torch.int32, device=self.device)
            torch.cumsum(diag[st['d_codes'].long()], 0, dtype=torch.int32, out=cs[1:])
            st['d_cs'] = cs
        return st

    def _launch_fast(self, st, adjust_lens, min_viable, R, rng_seed, shard, streamed, d_matrix=None):
        """Queue every batch of this rank's pair ranges on the current stream (no host synchronisation).
        Results land in whole-call device arrays, stored range after range ("compact"); with d_matrix the float
        epilogue runs right behind each batch and fills the matrix."""
        torch = self.torch
        n_seq, n_codes, max_len = st['n_seq'], st['n_codes'], st['max_len']
        n_pairs = st['n_pairs']
        if abs(int(rng_seed)) + n_pairs >= 2 ** 63:
            # the device forms |rng_seed + p| in 64 bits (one or two init_by_array key words); CPython takes any int
            raise NotImplementedError('rng_seed beyond 63 bits is not supported on the device path')
        R = max(int(R), 0)
        slot = 2 * R * max_len
        streamed = bool(streamed and R > 0 and not self.single_kernel_shuffle)
        mt_per_pair = (self.L.tg_dist_batch_mt_work_bytes(1024, R, max_len) - 256) // 1024 if streamed else 0
        cap = self.pairs_per_batch
        if slot:
            budget = self.shuffle_bytes_per_batch
            if self.device.type == 'cuda':
                # leave room for the other whole-call arrays and for other processes on the device
                budget = max(64 << 20, min(budget, torch.cuda.mem_get_info(self.device)[0] // 2))
            cap = max(2, min(cap, budget // (slot + mt_per_pair)))
        cap -= cap % 2
        # pair ranges owned by this rank: blocks of sorted pair space dealt round-robin
        rank, world = shard if shard is not None else (0, 1)
        ranges = fast_ranges(n_pairs, rank, world)
        cap = min(cap, max(2, max((e - b for b, e in ranges), default=2) + 1))
        n_own = sum(e - b for b, e in ranges)
        range_pos = np.cumsum([0] + [e - b for b, e in ranges])
        # the batches of this rank's ranges: (first pair, one past the last pair, position in the compact result arrays)
        batches = []
        for ri, (rb, re_) in enumerate(ranges):
            q = rb
            while q < re_:
                q1 = min(re_, q + cap)
                batches.append((q, q1, int(range_pos[ri]) + (q - rb)))
                q = q1
        # TG_DIST_PIPELINE=1: with several batches the null model's stream + sampling kernels of batch k run on a second stream
        # under the real alignments of batch k + 1 (two sets of shuffle area / stream buffers / counters).  Measured on a B200 at
        # N=2048: 525-530 ms per step against 530-536 ms -- both kernels are bound by instruction issue on the same SMs, so
        # there is little to hide; off by default (it doubles the work buffers), kept for devices where the balance differs.
        pipelined = (streamed and len(batches) >= 2 and self.device.type == 'cuda' and os.environ.get('TG_DIST_PIPELINE', '0') != '0'
                     and 2 * cap * (slot + mt_per_pair) < torch.cuda.mem_get_info(self.device)[0] * 3 // 4)
        nbuf = 2 if pipelined else 1
        if st['pool'] is None or st['pool'].numel() < n_codes + nbuf * cap * slot + 64:
            st['pool'] = None
            st['pool'] = torch.empty(n_codes + nbuf * cap * slot + 64, dtype=torch.uint8, device=self.device)
            st['pool'][:n_codes] = st['d_codes']
        pool = st['pool']
        need = self.L.tg_nw_wave_scratch_bytes(max_len)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        wide = int((st['sub'] - 2 * self.sc.gap).max()) * max_len > 65535      # the library's rule (tg_dist_batch): 32-bit cells
        dev = dict(n_own=n_own, ranges=ranges, n_pairs=n_pairs, world=world, R=R, streamed=streamed, wide=wide,
                   aln=torch.empty(max(n_own, 1), dtype=torch.int32, device=self.device),
                   rnd=torch.full((max(n_own, 1), max(R, 1)), NOT_COMPUTED, dtype=torch.int32, device=self.device),
                   flags=torch.empty(max(n_own, 1), dtype=torch.uint8, device=self.device),
                   iden=torch.empty(max(n_own, 1), dtype=torch.float64, device=self.device),
                   shape=torch.empty((max(n_own, 1), 2), dtype=torch.int32, device=self.device),
                   # per buffer set: [0:2] work counter, [4] sticky stream overflow, [6] live pairs of the batch
                   counter=torch.zeros((nbuf, 8), dtype=torch.int32, device=self.device),
                   stats=torch.zeros(4, dtype=torch.int64, device=self.device),
                   ambig=torch.empty(self.ambig_cap, dtype=torch.int64, device=self.device) if d_matrix is not None else None)
        mt_bytes = self.L.tg_dist_batch_mt_work_bytes(cap, R, max_len) if streamed else 0
        mt_work = [torch.empty(mt_bytes, dtype=torch.uint8, device=self.device) for _ in range(nbuf)] if streamed else None
        d_cs = st['d_cs']
        prof_rows = 0 if wide else choose_prof_rows(st['seq_mask'], sub_rows=st['sub'].shape[0], max_len=max_len, R=R, n_pairs=min(cap, n_own))

        def batch_args(k):
            (q, q1, cpos) = batches[k]
            b = k % nbuf
            a = _lib.DistBatchArgs()
            a.d_pool, a.d_seq_off, a.d_order = pool.data_ptr(), st['d_off'].data_ptr(), st['d_order'].data_ptr()
            a.d_diag_cs = d_cs.data_ptr() if d_cs is not None else None
            a.d_aln, a.d_rnd = dev['aln'].data_ptr() + 4 * cpos, dev['rnd'].data_ptr() + 4 * max(R, 1) * cpos
            a.d_shape, a.d_flags, a.d_iden = dev['shape'].data_ptr() + 8 * cpos, dev['flags'].data_ptr() + cpos, dev['iden'].data_ptr() + 8 * cpos
            a.d_scratch, a.d_counter = self._scratch.data_ptr(), dev['counter'].data_ptr() + 32 * b
            a.d_mt_work = mt_work[b].data_ptr() if streamed else None
            a.q0, a.q1, a.seed0, a.shuf_base, a.slot = q, q1, int(rng_seed), n_codes + b * cap * slot, slot
            a.match = float(self.sc.match) if self.sc.alphabet is None else 0.0
            a.n_seq, a.adjust_lens, a.min_viable, a.R, a.max_len = n_seq, int(bool(adjust_lens)), int(bool(min_viable)), R, st['max_len']
            if st['n_mat']:
                a.n_mat, a.d_mat_pair_off, a.d_mat_seq_off = st['n_mat'], st['d_mat_pair_off'].data_ptr(), st['d_mat_seq_off'].data_ptr()
            if prof_rows:
                a.d_seq_mask, a.prof_rows = st['d_seq_mask'].data_ptr(), prof_rows
            return a

        def epilogue(k, a, stream):
            (q, q1, cpos) = batches[k]
            e = _lib.DistEpilogueArgs()
            e.d_seq_off, e.d_order = a.d_seq_off, a.d_order
            e.d_aln, e.d_rnd, e.d_flags, e.d_iden = a.d_aln, a.d_rnd, a.d_flags, a.d_iden
            e.d_matrix, e.d_ambig, e.d_stats = d_matrix.data_ptr(), dev['ambig'].data_ptr(), dev['stats'].data_ptr()
            e.q0, e.q1, e.ambig_base = q, q1, cpos
            e.n_seq, e.adjust_lens, e.min_viable, e.R = n_seq, a.adjust_lens, a.min_viable, R
            e.ambig_cap, e.margin_log2 = self.ambig_cap, self.margin_log2
            if st['n_mat']:
                e.n_mat, e.d_mat_pair_off, e.d_mat_seq_off = a.n_mat, a.d_mat_pair_off, a.d_mat_seq_off
                e.d_mat_out_off = st['d_mat_out_off'].data_ptr()
            _lib.check(self.L.tg_dist_epilogue(C.byref(e), stream))
            self.stats['launches'] += 1

        def run(a, phases, stream, reserve=0):
            a.phases, a.smem_reserve = phases, reserve
            _lib.check(self.L.tg_dist_batch(C.byref(a), C.byref(st['cstruct']), stream))

        n_launch = ((6 if streamed else 4) if R else 2) + (1 if prof_rows and R else 0)
        if not pipelined:
            stream = _lib.stream_ptr(torch)
            for k in range(len(batches)):
                a = batch_args(k)
                run(a, 0, stream)
                self.stats['launches'] += n_launch
                if d_matrix is not None:
                    epilogue(k, a, stream)
            return dev
        main = torch.cuda.current_stream(self.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(self.device)
        side = self._side_stream
        args = [batch_args(k) for k in range(len(batches))]
        ev_mt = [None] * len(batches)
        ev_b = None
        for k in range(len(batches)):
            run(args[k], 1, main.cuda_stream, reserve=-1)            # real alignments of batch k (leave room for the side stream's CTAs)
            ev_a = torch.cuda.Event()
            ev_a.record(main)
            if k > 0:
                main.wait_event(ev_mt[k - 1])
                run(args[k - 1], 4, main.cuda_stream)                # null-model alignments of batch k - 1
                if d_matrix is not None:
                    epilogue(k - 1, args[k - 1], main.cuda_stream)
                ev_b = torch.cuda.Event()
                ev_b.record(main)
            side.wait_event(ev_a)
            if ev_b is not None:
                side.wait_event(ev_b)                                 # ... so that the shuffles of batch k run under the alignments of batch k + 1
            run(args[k], 2, side.cuda_stream)
            ev_mt[k] = torch.cuda.Event()
            ev_mt[k].record(side)
            self.stats['launches'] += n_launch
        last = len(batches) - 1
        main.wait_event(ev_mt[last])
        run(args[last], 4, main.cuda_stream)
        if d_matrix is not None:
            epilogue(last, args[last], main.cuda_stream)
        dev['keepalive'] = (mt_work, args)                            # buffers the side stream used stay referenced until the caller synchronises
        return dev

    def _compact_to_flat(self, st, dev, pos):
        """compact result positions -> flat indices of D[i][j] and D[j][i] in the output buffer (all matrices back to back)."""
        starts = np.array([b for b, _ in dev['ranges']], dtype=np.int64)
        rpos = np.cumsum([0] + [e - b for b, e in dev['ranges']])
        ri = np.searchsorted(rpos, pos, side='right') - 1
        q = starts[ri] + (pos - rpos[ri])
        if st['n_mat']:
            mat = np.searchsorted(st['mat_pair_off'], q, side='right') - 1
            q = q - st['mat_pair_off'][mat]
            sbase = st['mat_seq_off'][mat].astype(np.int64)
            n = st['sizes'][mat]
            obase = st['mat_out_off'][mat]
        else:
            sbase, n, obase = 0, np.int64(st['n_seq']), 0
        t = 2.0 * n - 1.0
        r = np.floor((t - np.sqrt(t * t - 8.0 * q)) * 0.5).astype(np.int64).clip(0, np.maximum(n - 2, 0))
        first = lambda rr: rr * (2 * n - rr - 1) // 2
        r = np.where(first(r) > q, r - 1, r)
        r = np.where(first(r + 1) <= q, r + 1, r)
        bcol = q - first(r) + r + 1
        o = st['order'].astype(np.int64)
        i, j = o[sbase + r] - sbase, o[sbase + bcol] - sbase
        return obase + i * n + j, obase + j * n + i

    def matrix_device(self, st, adjust_lens, min_viable, R, rng_seed, shard=None, streamed=True):
        """float32 (n, n) distance matrix ON THE DEVICE for this rank's share of the pair space (all of it without
        shard); other entries stay 0.  tg_align.py:154-189 without the host round trips: the kernels of every batch and
        the float epilogue are queued back to back, one synchronisation at the end."""
        torch = self.torch
        D = torch.zeros(st['out_elems'], dtype=torch.float32, device=self.device)    # all matrices back to back
        dev = self._launch_fast(st, adjust_lens, min_viable, R, rng_seed, shard, streamed, d_matrix=D)
        tail = torch.cat([dev['stats'], dev['counter'].reshape(-1).to(torch.int64)]).cpu().numpy()      # the one synchronisation
        self.stats['d2h_bytes'] += tail.nbytes
        if dev['streamed'] and any(tail[4 + 8 * b + 4] != 0 for b in range(dev['counter'].shape[0])):
            # a pair drew more random numbers than its pre-generated stream holds (never observed; the stream is
            # sized > 10 sigma above the expectation): redo with the single-kernel null model
            return self.matrix_device(st, adjust_lens, min_viable, R, rng_seed, shard, streamed=False)
        n_amb = int(tail[3])
        if n_amb:
            # pairs whose float32 value could depend on the last bits of log(): recompute with the host's libm
            if n_amb <= self.ambig_cap:
                pos = np.sort(dev['ambig'][:n_amb].cpu().numpy())
            else:
                pos = np.arange(dev['n_own'], dtype=np.int64)
            for c0 in range(0, len(pos), 1 << 22):
                pc = pos[c0:c0 + (1 << 22)]
                d_pc = torch.from_numpy(pc).to(self.device)
                sub = dict(aln=dev['aln'][d_pc].cpu().numpy(), rnd=dev['rnd'][d_pc][:, :dev['R']].cpu().numpy(),
                           iden=dev['iden'][d_pc].cpu().numpy())
                d = torch.from_numpy(self.distances(sub, dev['R']).astype('<f4')).to(self.device)
                fij, fji = self._compact_to_flat(st, dev, pc)
                D[torch.from_numpy(fij).to(self.device)] = d
                D[torch.from_numpy(fji).to(self.device)] = d
            self.stats['host_fixed_pairs'] = self.stats.get('host_fixed_pairs', 0) + len(pos)
        self.stats['pairs'] += dev['n_own']
        self.stats['early_out'] += int(tail[2])
        self.stats['cells'] += int(tail[0]) + int(tail[1])
        self.stats['wave32_jobs' if dev['wide'] else 'wave_jobs'] += dev['n_own']
        if st['n_mat']:
            return D                                          # flat: matrix b = D[mat_out_off[b] : mat_out_off[b+1]].view(n_b, n_b)
        return D.view(st['n_seq'], st['n_seq'])

    def _pair_scores_fast(self, st, adjust_lens, min_viable, R, rng_seed, shard, streamed=True):
        """Integer results of the device-driven path, copied to the host (parity tests, explicit gathers)."""
        dev = self._launch_fast(st, adjust_lens, min_viable, R, rng_seed, shard, streamed)
        n_own, R = dev['n_own'], dev['R']
        if dev['streamed'] and bool(dev['counter'][:, 4].any().item()):
            return self._pair_scores_fast(st, adjust_lens, min_viable, R, rng_seed, shard, streamed=False)
        aln = dev['aln'][:n_own].cpu().numpy()
        rnd = dev['rnd'][:n_own, :R].cpu().numpy() if R else np.zeros((n_own, 0), dtype=np.int32)
        flags = dev['flags'][:n_own].cpu().numpy()
        iden = dev['iden'][:n_own].cpu().numpy()
        shape = dev['shape'][:n_own].cpu().numpy()
        self.stats['d2h_bytes'] += n_own * (4 + 4 * R + 1 + 8 + 8)
        cells = shape[:, 0].astype(np.int64) * shape[:, 1]
        self.stats['pairs'] += n_own
        self.stats['early_out'] += int(flags.sum())
        self.stats['cells'] += int(cells.sum()) + R * int(cells[flags == 0].sum())
        self.stats['wave32_jobs' if dev['wide'] else 'wave_jobs'] += n_own
        res = dict(aln=aln, rnd=np.ascontiguousarray(rnd), iden=iden, n=shape[:, 0].copy(), m=shape[:, 1].copy(), order=st['order'])
        if dev['world'] > 1:
            res['ranges'] = dev['ranges']
            res['n_pairs'] = dev['n_pairs']
        return res

    def pair_scores(self, sequences, adjust_lens, min_viable, R, rng_seed, pair_idx=None, shard=None):
        """Integer results for every pair: dict(aln, rnd[n_pairs, R], iden, n, m, order).
        rnd == NOT_COMPUTED where the reference returns early (tg_align.py:137-138).
        order is None  -> arrays follow itertools.combinations order (tg_align.py:156);
        order is given -> arrays follow combinations order over the sequences sorted by length
                          (order[rank] = sequence index); see canonical_order() / dist_matrix().
        shard=(rank, world) restricts the work to this rank's share of the pair space (multi-GPU);
        everything else stays NOT_COMPUTED / 0 and is merged by multi_gpu.gather_results."""
        if len(sequences) >= 2 and pair_idx is None:
            st = self.prepare(sequences)
            if st is not None:
                return self._pair_scores_fast(st, adjust_lens, min_viable, R, rng_seed, shard)
        if shard is not None and pair_idx is None and shard[1] > 1:
            from .multi_gpu import shard_pair_indices
            pair_idx = shard_pair_indices(len(sequences) * (len(sequences) - 1) // 2, shard[0], shard[1])
        n_seq = len(sequences)
        n_pairs = n_seq * (n_seq - 1) // 2
        codes, off, sub = self.sc.encode_all(sequences)
        lens = np.diff(off)
        aln = np.full(n_pairs, NOT_COMPUTED, dtype=np.int32)
        rnd = np.full((n_pairs, max(R, 0)), NOT_COMPUTED, dtype=np.int32)
        n_eff_all = np.zeros(n_pairs, dtype=np.int32)
        m_eff_all = np.zeros(n_pairs, dtype=np.int32)
        iden_all = np.zeros(n_pairs, dtype=np.float64)
        if n_pairs == 0:
            return dict(aln=aln, rnd=rnd, iden=iden_all, n=n_eff_all, m=m_eff_all, order=None)
        # prefix sums of the diagonal for the identity score (tg_align.py:127-134)
        diag_cs = np.zeros(len(codes) + 1, dtype=np.int64)
        np.cumsum(np.diag(sub)[codes], out=diag_cs[1:])
        d_codes = self._to_dev(codes)
        seed0 = int(rng_seed)
        if pair_idx is None:
            pair_idx = np.arange(n_pairs, dtype=np.int64)
        # row starts in combinations order, to turn pair indices into (I, J)
        ar = np.arange(n_seq, dtype=np.int64)
        row_start = ar * (2 * n_seq - ar - 1) // 2
        p = 0
        while p < len(pair_idx):
            q = min(len(pair_idx), p + self.pairs_per_batch)
            P = pair_idx[p:q]
            I = np.searchsorted(row_start, P, side='right') - 1
            I = np.minimum(I, n_seq - 2)
            J = P - row_start[I] + I + 1
            li, lj = lens[I], lens[J]
            if adjust_lens:                                   # tg_align.py:115-120
                ml = np.minimum(li, lj)
                if min_viable:
                    ml = np.maximum(ml, MIN_VIABLE_SEQ_LEN)
                li, lj = np.minimum(li, ml), np.minimum(lj, ml)
            li = li.astype(np.int32); lj = lj.astype(np.int32)
            # shuffle area budget: shrink the batch if needed
            per_pair = (li.astype(np.int64) + lj) * max(R, 0)
            if R > 0 and per_pair.sum() > self.shuffle_bytes_per_batch and len(P) > 1:
                keep = max(1, int(np.searchsorted(np.cumsum(per_pair), self.shuffle_bytes_per_batch)))
                q = p + keep
                P, I, J, li, lj = P[:keep], I[:keep], J[:keep], li[:keep], lj[:keep]
            self._batch(d_codes, len(codes), off, sub, diag_cs, I, J, P, li, lj, R, seed0, aln, rnd, iden_all)
            n_eff_all[P], m_eff_all[P] = li, lj
            p = q
        return dict(aln=aln, rnd=rnd, iden=iden_all, n=n_eff_all, m=m_eff_all, order=None)

    def _batch(self, d_codes, n_codes, off, sub, diag_cs, I, J, P, li, lj, R, seed0, aln, rnd, iden_all):
        torch = self.torch
        sc = self.sc
        npair = len(P)
        mixed_ok = True          # both kernels accept different shapes in the two halves of a job
        row_off, col_off = off[I], off[J]
        # ---- phase A: the real alignment of every pair (tg_align.py:125) ----
        jobs = pack_jobs(row_off, col_off, li, lj, np.arange(npair, dtype=np.int32), mixed_ok)
        a = self.run_jobs(d_codes, jobs, sub, npair)
        aln[P] = a
        # ---- identity score and the early-out test (tg_align.py:127-138), float64 like the reference ----
        if sc.alphabet is None:
            iden = sc.match * ((li.astype(np.int64) + lj) / 2.)
        else:
            is1 = diag_cs[row_off + li] - diag_cs[row_off]
            is2 = diag_cs[col_off + lj] - diag_cs[col_off]
            iden = (is1 + is2) / 2.
        iden_all[P] = iden
        early = a.astype(np.float64) < -(IDEN_SCORE_SCALAR_CHEAT * iden)
        cells = li.astype(np.int64) * lj
        self.stats['pairs'] += npair
        self.stats['early_out'] += int(early.sum())
        self.stats['cells'] += int(cells.sum())
        if R <= 0:
            return
        sel = np.nonzero(~early)[0]
        if len(sel) == 0:
            return
        self.stats['cells'] += int(cells[sel].sum()) * R
        # ---- phase B: null model.  seed = rng_seed + pair index (tg_align.py:157,170-172) ----
        ns, ms = li[sel], lj[sel]
        per_pair = (ns.astype(np.int64) + ms) * R
        dst = np.zeros(len(sel), dtype=np.int64)
        dst[1:] = np.cumsum(per_pair)[:-1]
        total_shuf = int(per_pair.sum())
        d_pool = torch.empty(n_codes + total_shuf, dtype=torch.uint8, device=self.device)
        d_pool[:n_codes] = d_codes
        dst += n_codes
        sj = np.zeros(len(sel), dtype=_lib.SHUFFLE_JOB)
        seeds = [abs(seed0 + int(pp)) for pp in (P[sel[0]], P[sel[-1]])]
        if max(seeds) >= 1 << 64:
            raise NotImplementedError('rng_seed + pair index beyond 64 bits')
        if seed0 >= 0:
            sj['seed'] = (P[sel] + seed0).astype(np.uint64)
        else:
            sj['seed'] = np.abs(P[sel].astype(np.int64) + seed0).astype(np.uint64)
        sj['src_i'], sj['src_j'], sj['dst_off'], sj['n'], sj['m'] = row_off[sel], col_off[sel], dst, ns, ms
        d_sj = self._to_dev(sj)
        _lib.check(self.L.tg_mt_shuffle(_lib.dptr(d_pool), _lib.dptr(d_sj), len(sel), R, int(max(ns.max(), ms.max())),
                                        _lib.stream_ptr(torch)))
        self.stats['launches'] += 1
        # R alignments per selected pair: rows at dst + k*(n+m), columns right behind
        k = np.arange(R, dtype=np.int64)
        r_off = (dst[:, None] + k[None, :] * (ns.astype(np.int64) + ms)[:, None]).reshape(-1)
        c_off = r_off + np.repeat(ns.astype(np.int64), R)
        jobs_b = pack_jobs(r_off, c_off, np.repeat(ns, R), np.repeat(ms, R),
                           np.arange(len(sel) * R, dtype=np.int32), mixed_ok)
        rs = self.run_jobs(d_pool, jobs_b, sub, len(sel) * R).reshape(len(sel), R)
        rnd[P[sel]] = rs

    # ---- float epilogue (host, float64, the reference's expression order) --------------------------
    @staticmethod
    def distances(res, R):
        """tg_align.py:137-149 for every pair at once."""
        aln = res['aln'].astype(np.float64)
        iden = res['iden']
        d = np.full(len(aln), MAX_SEQ_DIST, dtype=np.float64)
        early = aln < -(IDEN_SCORE_SCALAR_CHEAT * iden)
        live = np.nonzero(~early)[0]
        if len(live):
            if R > 0:
                rmean = np.add.reduce(res['rnd'][live].astype(np.float64), axis=1) / R      # np.mean(rand_scores)
            else:
                rmean = np.full(len(live), np.nan)                                         # np.mean([]) -> nan
            a, idn = aln[live], iden[live]
            with np.errstate(divide='ignore', invalid='ignore'):
                dd = np.minimum(-np.log((a - rmean) / (idn - rmean)), MAX_SEQ_DIST)
            dd = np.where(rmean >= a, MAX_SEQ_DIST, dd)
            d[live] = np.maximum(dd, MIN_SEQ_DIST)
        return d

    def dist_matrix(self, sequences, adjust_lens, min_viable, R, rng_seed, res=None):
        n = len(sequences)
        D = np.zeros((n, n), dtype='<f4')
        if n < 2:
            return D
        if res is None:
            res = self.pair_scores(sequences, adjust_lens, min_viable, R, rng_seed)
        res = expand_ranges(res)
        d = self.distances(res, R)
        iu = np.triu_indices(n, k=1)                 # row-major upper triangle == combinations order
        if res.get('order') is not None:
            o = res['order']
            iu = (o[iu[0]], o[iu[1]])
        D[iu] = d
        D[(iu[1], iu[0])] = d
        return D


def fast_ranges(n_pairs, rank, world, min_block=4096):
    """Pair ranges (sorted pair space) owned by `rank`: the blocks DistEngine._launch_fast deals round-robin."""
    if world <= 1:
        return [(0, n_pairs)]
    # several blocks per rank even out the cost gradient of the sorted pair space (longest sequences first); every block is
    # at least one batch of kernels, so no more of them than needed to keep a block around half a million pairs
    per_rank = max(2, min(8, (n_pairs // world) >> 19))
    blk = max(min_block, -(-n_pairs // (world * per_rank)))
    blk += blk % 2
    return [(b, min(b + blk, n_pairs)) for b in range(rank * blk, n_pairs, world * blk)]


def expand_ranges(res):
    """Compact per-rank results (see _pair_scores_fast) -> full-length arrays, NOT_COMPUTED / 0 elsewhere."""
    if 'ranges' not in res:
        return res
    n_pairs, ranges = res['n_pairs'], res['ranges']
    out = {k: v for k, v in res.items() if k not in ('ranges', 'n_pairs')}
    for key, fill in (('aln', NOT_COMPUTED), ('rnd', NOT_COMPUTED), ('iden', 0), ('n', 0), ('m', 0)):
        src = res[key]
        full = np.full((n_pairs,) + src.shape[1:], fill, dtype=src.dtype)
        pos = 0
        for (b, e) in ranges:
            full[b:e] = src[pos:pos + (e - b)]
            pos += e - b
        out[key] = full
    return out


def canonical_order(res, n_seq):
    """Re-index a result dict to itertools.combinations order over the ORIGINAL sequence indices."""
    res = expand_ranges(res)
    if res.get('order') is None:
        return res
    o = res['order'].astype(np.int64)
    a, b = np.triu_indices(n_seq, k=1)
    i, j = np.minimum(o[a], o[b]), np.maximum(o[a], o[b])
    p = i * (2 * n_seq - i - 1) // 2 + (j - i - 1)
    out = dict(order=None)
    for key in ('aln', 'rnd', 'iden', 'n', 'm'):
        arr = np.empty_like(res[key])
        arr[p] = res[key]
        out[key] = arr
    return out
