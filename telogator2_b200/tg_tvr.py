"""Colour-vector preparation of the reference's source/tg_tvr.py on the GPU.

The reference runs these per-read O(L) Python loops inside cluster_tvrs (tg_tvr.py:96-142) before every distance matrix
and inside quick_get_tvrtel_lens (tg_tvr.py:558-592).  Same names, arguments and return values here:

    find_density_boundary   tg_tvr.py:449-498      (batch form: find_density_boundary_batch)
    denoise_colorvec        tg_tvr.py:524-556      (batch form: denoise_colorvec_batch)
    quick_get_tvrtel_lens   tg_tvr.py:558-592
    tvr_prep                the lists cluster_tvrs builds at tg_tvr.py:96-142 (inline code upstream), for all reads at once

All per-character work runs in the kernels of csrc/tg_cv.cu; the host only does the reference's slicing arithmetic on
(offset, length) pairs.  No CPU fallback: without the CUDA library / a device the functions raise.
"""
import ctypes as C
from itertools import chain

import numpy as np

from . import _lib

UNKNOWN = 'A'                       # tg_align.py:16
UNKNOWN_WIN_SIZE = 100              # tg_tvr.py:22-28
UNKNOWN_END_DENS = 0.125
UNKNOWN_WIN_SIZE_FINE = 50
UNKNOWN_END_DENS_FINE = 0.500
CANON_WIN_SIZE = 100
CANON_END_DENS = 0.700


def _codes(letters):
    b = ''.join(letters).encode('latin-1')
    if len(b) != len(list(letters)) or 0xff in b:
        raise ValueError('letters must be single latin-1 characters below 0xff')
    return b


class _Dev:
    """Strings on the device: one byte buffer + per-string (offset, length) kept on the host."""

    def __init__(self):
        self.torch = _lib.require_cuda()
        self.L = _lib.load()
        self.dev = self.torch.device('cuda', self.torch.cuda.current_device())
        self._live = []      # argument tensors stay referenced until this object dies: a freed block would be handed to
                             # the next upload before the kernel that reads it has even been launched

    def up(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)
        self._live.append(t)
        return t

    def stream(self):
        return _lib.stream_ptr(self.torch)

    def upload_strings(self, seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.int32)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        raw = ''.join(seqs).encode('latin-1')
        if len(raw) != int(off[-1]):
            raise ValueError('sequences must be latin-1 strings')
        buf = self.torch.empty(max(len(raw), 1), dtype=self.torch.uint8, device=self.dev)
        if raw:
            buf[:len(raw)] = self.torch.frombuffer(bytearray(raw), dtype=self.torch.uint8).to(self.dev)
        return buf, off[:-1].copy(), lens

    def paint(self, start, end, rk_off, kmer_letters, off0):
        """Colour vectors of all reads (tg_tvr.py:96-105): spans sorted by (read, k), rk_off their (read, k) index,
        off0 the n + 1 vector offsets -> byte buffer."""
        t = self.torch
        n, K = len(off0) - 1, len(kmer_letters)
        B0 = t.empty(max(int(off0[-1]), 1), dtype=t.uint8, device=self.dev)
        if n:
            _lib.check(self.L.tg_cv_paint(_lib.dptr(self.up(start.astype(np.int32))) if len(start) else None,
                                          _lib.dptr(self.up(end.astype(np.int32))) if len(end) else None, _lib.dptr(self.up(rk_off)),
                                          _lib.dptr(self.up(np.frombuffer(_codes(kmer_letters), dtype=np.uint8).copy())) if K else None, K,
                                          _lib.dptr(self.up(off0)), n, ord(UNKNOWN), _lib.dptr(B0), self.stream()))
        return B0

    def boundary(self, buf, off, lens, reverse, letters, win, thresh, above, start=0):
        """-> (first, extreme) int32 arrays, -1 = None (tg_tvr.py:449-479)."""
        n = len(lens)
        first = self.torch.empty(max(n, 1), dtype=self.torch.int32, device=self.dev)
        ext = self.torch.empty(max(n, 1), dtype=self.torch.int32, device=self.dev)
        if n:
            lb = _codes(letters)
            d_rev = self.up(np.asarray(reverse, dtype=np.uint8)) if reverse is not None else None
            _lib.check(self.L.tg_cv_density_boundary(_lib.dptr(buf), _lib.dptr(self.up(off.astype(np.int64))),
                                                     _lib.dptr(self.up(lens.astype(np.int32))), _lib.dptr(d_rev), n, lb, len(lb),
                                                     int(win), float(thresh), int(bool(above)), int(start),
                                                     _lib.dptr(first), _lib.dptr(ext), self.stream()))
        return first[:n].cpu().numpy(), ext[:n].cpu().numpy()

    def lead_run(self, buf, off, lens, reverse, cap, letter):
        n = len(lens)
        out = self.torch.empty(max(n, 1), dtype=self.torch.int32, device=self.dev)
        if n:
            d_rev = self.up(np.asarray(reverse, dtype=np.uint8)) if reverse is not None else None
            _lib.check(self.L.tg_cv_lead_run(_lib.dptr(buf), _lib.dptr(self.up(off.astype(np.int64))), _lib.dptr(self.up(lens.astype(np.int32))),
                                             _lib.dptr(d_rev), _lib.dptr(self.up(cap.astype(np.int32))), n, ord(letter),
                                             _lib.dptr(out), self.stream()))
        return out[:n].cpu().numpy()

    def gather(self, src, src_off, dst_off, seg_len, total):
        dst = self.torch.empty(max(int(total), 1), dtype=self.torch.uint8, device=self.dev)
        keep = seg_len > 0
        if keep.any():
            _lib.check(self.L.tg_cv_gather(_lib.dptr(src), _lib.dptr(self.up(src_off[keep].astype(np.int64))),
                                           _lib.dptr(self.up(dst_off[keep].astype(np.int64))), _lib.dptr(self.up(seg_len[keep].astype(np.int32))),
                                           int(keep.sum()), _lib.dptr(dst), self.stream()))
        return dst

    def denoise(self, buf, off, lens, replace_char, min_size, max_gap_fill, chars_to_delete, chars_to_merge):
        """denoise_colorvec for non-empty strings -> (out buffer, out offsets, out lengths = len + 1)."""
        n = len(lens)
        if (lens <= 0).any():
            raise ValueError('denoise: empty string in the batch')
        trail = self.lead_run(buf, off, lens, np.ones(n, dtype=np.uint8), lens, replace_char)
        keep = (lens - trail).astype(np.int32)
        olen = (lens + 1).astype(np.int32)
        ooff = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(olen, out=ooff[1:])
        tile = self.L.tg_cv_denoise_tile()
        toff = np.zeros(n + 1, dtype=np.int64)
        np.cumsum((olen.astype(np.int64) + tile - 1) // tile, out=toff[1:])
        out = self.torch.empty(max(int(ooff[-1]), 1), dtype=self.torch.uint8, device=self.dev)
        if n:
            db, mb = _codes(chars_to_delete), _codes(chars_to_merge)
            _lib.check(self.L.tg_cv_denoise(_lib.dptr(buf), _lib.dptr(self.up(off.astype(np.int64))), _lib.dptr(self.up(lens.astype(np.int32))),
                                            _lib.dptr(self.up(keep)), _lib.dptr(self.up(ooff[:-1].copy())), _lib.dptr(self.up(toff)), n,
                                            int(toff[-1]), ord(replace_char), int(min_size), int(max_gap_fill), db, len(db), mb, len(mb),
                                            _lib.dptr(out), self.stream()))
        return out, ooff[:-1].copy(), olen

    def download(self, buf, total):
        return bytes(buf[:int(total)].cpu().numpy().tobytes()) if total else b''


# ---------------------------------------------------------------------------------------------------------
# function-level mirrors
# ---------------------------------------------------------------------------------------------------------
def _split_pos(first, ext, use_lowest_dens):
    """first_pos_below_thresh falls back to pos_with_lowest_dens (tg_tvr.py:491-492); None slices to (whole, whole)."""
    first = np.where(first < 0, ext, first)
    p = ext if use_lowest_dens else first
    return p            # -1 = None


def find_density_boundary_batch(sequences, which_letter, win_size, dens_thresh, thresh_dir='below', starting_coord=0,
                                use_lowest_dens=False):
    if thresh_dir not in ('below', 'above'):
        print('Error: find_density_boundary() thresh_dir must be above or below')
        exit(1)
    d = _Dev()
    buf, off, lens = d.upload_strings(sequences)
    first, ext = d.boundary(buf, off, lens, None, which_letter, win_size, dens_thresh, thresh_dir == 'above', starting_coord)
    p = _split_pos(first, ext, use_lowest_dens)
    return [(s[:None if q < 0 else int(q)], s[None if q < 0 else int(q):]) for s, q in zip(sequences, p)]


def find_density_boundary(sequence, which_letter, win_size, dens_thresh, thresh_dir='below', starting_coord=0,
                          use_lowest_dens=False, debug_plot=False, readname=''):
    """tg_tvr.py:449-498 (debug_plot is a matplotlib aid upstream and is ignored)."""
    return find_density_boundary_batch([sequence], which_letter, win_size, dens_thresh, thresh_dir, starting_coord,
                                       use_lowest_dens)[0]


def denoise_colorvec_batch(vs, replace_char=UNKNOWN, min_size=10, max_gap_fill=50, chars_to_delete=(), chars_to_merge=()):
    idx = [i for i, v in enumerate(vs) if len(v)]
    res = [''] * len(vs)                                   # tg_tvr.py:525-526
    if idx:
        d = _Dev()
        buf, off, lens = d.upload_strings([vs[i] for i in idx])
        out, ooff, olen = d.denoise(buf, off, lens, replace_char, min_size, max_gap_fill, chars_to_delete, chars_to_merge)
        raw = d.download(out, int(ooff[-1] + olen[-1]))
        for k, i in enumerate(idx):
            res[i] = raw[ooff[k]:ooff[k] + olen[k]].decode('latin-1')
    return res


def denoise_colorvec(v, replace_char=UNKNOWN, min_size=10, max_gap_fill=50, chars_to_delete=[], chars_to_merge=[]):
    """tg_tvr.py:524-556.  Note the upstream quirk that is kept: the result is one letter longer than v."""
    return denoise_colorvec_batch([v], replace_char, min_size, max_gap_fill, chars_to_delete, chars_to_merge)[0]


# ---------------------------------------------------------------------------------------------------------
# the per-read preparation of cluster_tvrs / quick_get_tvrtel_lens, all reads at once
# ---------------------------------------------------------------------------------------------------------
def _letter_sets(repeats_metadata):
    """tg_tvr.py:60-91."""
    [kmer_list, _kmer_colors, kmer_letters, kmer_flags] = repeats_metadata
    canonical_letter = None
    sets = {f: [] for f in ('denoise', 'tvr', 'nofilt', 'dubious', 'subtel-filt')}
    for i in range(len(kmer_list)):
        if 'canonical' in kmer_flags[i]:
            canonical_letter = kmer_letters[i]
        for f in sets:
            if f in kmer_flags[i]:
                sets[f].append(kmer_letters[i])
        if kmer_letters[i] == UNKNOWN:
            print(f'Error: character {UNKNOWN} is reserved for unknown sequence')
            exit(1)
    return canonical_letter, {f: sorted(set(v)) for f, v in sets.items()}


class _Prep:
    """Device state of one batch: colour vectors (B0) and the (subtel | tvr+tel) split of every read."""

    def __init__(self, kmer_dat, kmer_letters, subfilt_letters):
        d = self.d = _Dev()
        n = self.n = len(kmer_dat)
        K = len(kmer_letters)
        tlen = np.array([int(kd[1]) for kd in kmer_dat], dtype=np.int32)
        counts = np.zeros((n, K), dtype=np.int64)
        flat = []
        for r, kd in enumerate(kmer_dat):
            hits = kd[0]
            if len(hits) > K:
                raise IndexError('list index out of range')                  # kmer_letters[ki], tg_tvr.py:104
            counts[r, :len(hits)] = list(map(len, hits))
            flat.append(np.fromiter(chain.from_iterable(chain.from_iterable(hits)), dtype=np.int64))
        se = np.concatenate(flat) if flat else np.zeros(0, dtype=np.int64)
        start, end = se[0::2], se[1::2]
        rk_off = np.zeros(n * K + 1, dtype=np.int64)
        np.cumsum(counts.reshape(-1), out=rk_off[1:])
        if len(start):
            lim = np.repeat(tlen.astype(np.int64), counts.sum(axis=1))
            if (start < 0).any() or ((end > lim) & (end > start)).any():
                raise IndexError('list assignment index out of range')       # my_letters[j] = ..., tg_tvr.py:104
        self.len0 = tlen
        off0 = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(tlen, out=off0[1:])
        self.off0 = off0[:-1].copy()
        self.total0 = int(off0[-1])
        self.B0 = d.paint(start, end, rk_off, kmer_letters, off0)
        unk = [UNKNOWN] + list(subfilt_letters)
        L = tlen.astype(np.int64)
        # tg_tvr.py:118  (seq_left, seq_right) = find_density_boundary(cv, unknown, 100, 0.125, 'below')
        f1, e1 = d.boundary(self.B0, self.off0, tlen, None, unk, UNKNOWN_WIN_SIZE, UNKNOWN_END_DENS, False)
        p1 = _split_pos(f1, e1, False).astype(np.int64)
        none1 = p1 < 0
        left1_len = np.where(none1, L, p1)                        # seq_left = cv[:p1]   (None: whole)
        right1_off = np.where(none1, 0, p1)                       # seq_right = cv[p1:]  (None: whole)
        right1_len = L - right1_off
        # tg_tvr.py:120  find_density_boundary(seq_left[::-1], unknown, 50, 0.5, 'above')
        f2, e2 = d.boundary(self.B0, self.off0, left1_len.astype(np.int32), np.ones(n, dtype=np.uint8), unk,
                            UNKNOWN_WIN_SIZE_FINE, UNKNOWN_END_DENS_FINE, True)
        p2 = _split_pos(f2, e2, False).astype(np.int64)
        none2 = p2 < 0
        rec_off = np.where(none2, 0, left1_len - p2)              # recovered_tvr[::-1] = seq_left[len - p2:]  (None: whole)
        rec_len = left1_len - rec_off
        self.left_len = np.where(none2, left1_len, left1_len - p2)        # rest_is_subtel[::-1] = seq_left[:len - p2]
        # tg_tvr.py:121  seq_right = recovered_tvr[::-1] + seq_right  -> B2
        self.len2 = (rec_len + right1_len).astype(np.int64)
        off2 = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.len2, out=off2[1:])
        self.off2 = off2[:-1].copy()
        self.total2 = int(off2[-1])
        src_off = np.stack([self.off0 + rec_off, self.off0 + right1_off], axis=1).reshape(-1)
        dst_off = np.stack([self.off2, self.off2 + rec_len], axis=1).reshape(-1)
        seg_len = np.stack([rec_len, right1_len], axis=1).reshape(-1)
        self.B2 = d.gather(self.B0, src_off, dst_off, seg_len, self.total2)
        # tg_tvr.py:124-128  the tvr must not begin with an unknown character
        if (self.len2 <= 0).any():
            raise IndexError('string index out of range')         # seq_right[adj] on an empty string
        self.adj = d.lead_run(self.B2, self.off2, self.len2.astype(np.int32), None, (self.len2 - 1).astype(np.int32), UNKNOWN).astype(np.int64)
        self.right_off = self.off2 + self.adj                     # seq_right = seq_right[adj:]   (in B2)
        self.right_len = self.len2 - self.adj


def quick_get_tvrtel_lens(kmer_dat, repeats_metadata):
    """tg_tvr.py:558-592."""
    [kmer_list, _kmer_colors, kmer_letters, kmer_flags] = repeats_metadata
    subfilt_letters = sorted(set(kmer_letters[i] for i in range(len(kmer_list)) if 'subtel-filt' in kmer_flags[i]))
    if len(kmer_dat) == 0:
        return []
    P = _Prep(kmer_dat, kmer_letters, subfilt_letters)
    return [int(x) for x in P.right_len]


def tvr_prep(kmer_dat, repeats_metadata, tvr_truncate=3000):
    """The lists cluster_tvrs builds before the pairwise alignment (tg_tvr.py:96-142), as a dict:
    all_colorvecs, subtel_regions, tvrtel_regions, cleaned_colorvecs, colorvecs_for_msa, err_end_lens,
    tvrtels_for_clustering (the sequences handed to get_dist_matrix_parallel)."""
    [_kmer_list, _kmer_colors, kmer_letters, _kmer_flags] = repeats_metadata
    canonical_letter, sets = _letter_sets(repeats_metadata)
    if canonical_letter is None:
        print('Error: cluster_tvrs() received a kmer list that does not have any kmers marked as canonical')
        exit(1)
    keys = ('all_colorvecs', 'subtel_regions', 'tvrtel_regions', 'cleaned_colorvecs', 'colorvecs_for_msa', 'err_end_lens',
            'tvrtels_for_clustering')
    out = {k: [] for k in keys}
    n = len(kmer_dat)
    if n == 0:
        return out
    P = _Prep(kmer_dat, kmer_letters, sets['subtel-filt'])
    d = P.d
    # tg_tvr.py:130  denoise the tvr+tel section
    B3, off3, len3 = d.denoise(P.B2, P.right_off, P.right_len.astype(np.int32), canonical_letter, 10, 50, sets['denoise'],
                               [canonical_letter] + sets['tvr'])
    # tg_tvr.py:132  (err_left, err_right) = find_density_boundary(seq_right_denoise[::-1], [canonical], 100, 0.7, 'above')
    f3, e3 = d.boundary(B3, off3, len3, np.ones(n, dtype=np.uint8), [canonical_letter], CANON_WIN_SIZE, CANON_END_DENS, True)
    p3 = _split_pos(f3, e3, False).astype(np.int64)
    none3 = p3 < 0
    err_left_len = np.where(none3, len3, p3)                      # err_left = rev[:p3]            (None: whole)
    msa_len = np.where(none3, len3, len3 - p3)                    # err_right[::-1] = den[:len - p3] (None: whole)
    raw0 = d.download(P.B0, P.total0)
    raw2 = d.download(P.B2, P.total2)
    raw3 = d.download(B3, int(off3[-1] + len3[-1]))
    for i in range(n):
        cv = raw0[P.off0[i]:P.off0[i] + P.len0[i]]
        right2 = raw2[P.off2[i]:P.off2[i] + P.len2[i]]
        adj = int(P.adj[i])
        seq_left = cv[:P.left_len[i]] + right2[:adj]              # tg_tvr.py:122,127
        seq_right = right2[adj:]
        msa = raw3[off3[i]:off3[i] + msa_len[i]]
        out['all_colorvecs'].append(cv.decode('latin-1'))
        out['cleaned_colorvecs'].append((seq_left + msa).decode('latin-1'))
        out['err_end_lens'].append(int(err_left_len[i]))
        out['colorvecs_for_msa'].append(msa.decode('latin-1'))
        out['subtel_regions'].append(seq_left.decode('latin-1'))
        out['tvrtel_regions'].append(seq_right.decode('latin-1'))
        out['tvrtels_for_clustering'].append(msa[:tvr_truncate].decode('latin-1'))
    return out
