"""Drop-in for the scan functions of the reference's source/tg_kmer.py (lines 11-102, 173-230).

Per-read functions keep the reference's names, signatures and return types and run as a batch of
one on the GPU; the *_batch / ScanSession entry points are what tg_tel.get_tel_repeat_comp_parallel
uses (one launch per phase for all reads).  get_telomere_regions / wavelet_smooth / mad
(tg_kmer.py:105-170) run on the device as well, through telogator2_b200.tg_boundary.
"""
import ctypes as C
import os

import gc

import numpy as np

from . import _lib
from .tg_util import RC  # noqa: F401  (re-exported like the reference module's namespace)

INCLUDE_SUBTEL_BUFF = 500      # tg_kmer.py:8
MAX_KMERS = 64
MAX_KMER_LEN = 32


def exists_and_is_nonzero(fn):
    return os.path.isfile(fn) and os.path.getsize(fn) > 0


def _tsv_rows(fn, READ_TYPE):
    """Rows of a k-mer TSV surviving the read-type filter (tg_kmer.py:22-30, 52-60)."""
    with open(fn, 'r') as f:
        for line in f:
            if line[0] != '#' and len(line.strip()):
                splt = line.strip().split('\t')
                flags = splt[3].split(',')
                if 'pacbio_only' in flags and READ_TYPE not in ['hifi']:
                    continue
                if 'nanopore_only' in flags and READ_TYPE not in ['ont']:
                    continue
                yield splt[0], splt[1], splt[2], tuple(flags)


def read_kmer_tsv(fn, READ_TYPE):
    """tg_kmer.py:11-48: ([KMER_LIST, COLORS, LETTERS, FLAGS], KMER_ISSUBSTRING, CANONICAL_STRINGS);
    k-mers de-duplicated and sorted descending by (len, kmer, colour, letter, flags)."""
    if exists_and_is_nonzero(fn) is False:
        print('Error: ' + fn + ' not found.')
        exit(1)
    rows = list(_tsv_rows(fn, READ_TYPE))
    canonical = [r[0] for r in rows if 'canonical' in r[3]]
    dat = sorted(set((len(r[0]), r[0], r[1], r[2], r[3]) for r in rows), reverse=True)
    kmers = [d[1] for d in dat]
    meta = [kmers, [d[2] for d in dat], [d[3] for d in dat], [d[4] for d in dat]]
    issub = [[j for j in range(len(kmers)) if (j != i and kmers[i] in kmers[j])] for i in range(len(kmers))]
    return (meta, issub, canonical)


def get_canonical_letter(fn, READ_TYPE):
    """tg_kmer.py:51-63: letter of the first canonical row."""
    for (_kmer, _color, letter, flags) in _tsv_rows(fn, READ_TYPE):
        if 'canonical' in flags:
            return letter
    return None


# ------------------------------------------------------------------------------------------------
# device tables
# ------------------------------------------------------------------------------------------------
_table_cache = {}


def kmer_table(kmer_list, issub=None):
    """Device blob of the window table + k-mer sets for `kmer_list` (hits kernel; cached per device)."""
    torch = _lib.require_cuda()
    L = _lib.load()
    dev = torch.cuda.current_device()
    key = (tuple(kmer_list), None if issub is None else tuple(tuple(x) for x in issub), dev)
    t = _table_cache.get(key)
    if t is not None:
        return t
    if len(kmer_list) < 1 or len(kmer_list) > MAX_KMERS:
        raise NotImplementedError(f'{len(kmer_list)} k-mers: the GPU scan supports 1..{MAX_KMERS} per list')
    blob = ''.join(kmer_list).encode('ascii')
    off = np.zeros(len(kmer_list) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(k) for k in kmer_list])
    if issub is None:
        issub = [[j for j in range(len(kmer_list)) if (j != i and kmer_list[i] in kmer_list[j])] for i in range(len(kmer_list))]
    if issub is not None:
        edges = np.array([j for lst in issub for j in lst], dtype=np.int32)
        eoff = np.zeros(len(issub) + 1, dtype=np.int32)
        eoff[1:] = np.cumsum([len(lst) for lst in issub])
        pe, po = edges.ctypes.data_as(C.c_void_p), eoff.ctypes.data_as(C.c_void_p)
    else:
        pe = po = None
    host = np.zeros(L.tg_kmer_table_bytes(), dtype=np.uint8)
    _lib.check(L.tg_kmer_table_build(blob, off.ctypes.data_as(C.c_void_p), len(kmer_list), pe, po,
                                     host.ctypes.data_as(C.c_void_p)))
    t = torch.from_numpy(host).cuda()
    _table_cache[key] = t
    return t


def kmer_table_pair(list_a, list_b):
    """Device blob of the 7-base window table for two k-mer lists (coverage kernels: one lookup per base for both)."""
    torch = _lib.require_cuda()
    L = _lib.load()
    key = ('pair', tuple(list_a), tuple(list_b), torch.cuda.current_device())
    t = _table_cache.get(key)
    if t is not None:
        return t
    blobs, offs = [], []
    for lst in (list_a, list_b):
        if len(lst) < 1 or len(lst) > MAX_KMERS:
            raise NotImplementedError(f'{len(lst)} k-mers: the GPU scan supports 1..{MAX_KMERS} per list')
        blobs.append(''.join(lst).encode('ascii'))
        off = np.zeros(len(lst) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(k) for k in lst])
        offs.append(off)
    host = np.zeros(L.tg_kmer_table_pair_bytes(), dtype=np.uint8)
    _lib.check(L.tg_kmer_table_build_pair(blobs[0], offs[0].ctypes.data_as(C.c_void_p), len(list_a),
                                          blobs[1], offs[1].ctypes.data_as(C.c_void_p), len(list_b),
                                          host.ctypes.data_as(C.c_void_p)))
    t = torch.from_numpy(host).cuda()
    _table_cache[key] = t
    return t


# ------------------------------------------------------------------------------------------------
# reads in HBM
# ------------------------------------------------------------------------------------------------
_pinned_pool = {}


def pinned_bytes(slot, nbytes):
    """A page-locked uint8 host buffer of at least nbytes, kept per slot across calls (allocating page-locked memory costs
    ~1 ms per MB).  The caller owns it until its next request for the same slot."""
    import torch
    buf = _pinned_pool.get(slot)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, pin_memory=True)
        _pinned_pool[slot] = buf
    return buf[:nbytes]


def _hostlists():
    """The CPython helper that formats results (chost/tg_hostlists.c); None when it has not been built (pure-Python
    formatting is used then -- it is the same formatting, only slower)."""
    global _hostlists_mod
    if _hostlists_mod is False:
        try:
            from . import _tg_hostlists
            _hostlists_mod = _tg_hostlists
        except ImportError:
            _hostlists_mod = None
    return _hostlists_mod


_hostlists_mod = False

_ACGTN = np.zeros(256, dtype=bool)
_ACGTN[[ord(c) for c in 'ACGTN']] = True
ALIGN = 128


class PackedReads:
    """A batch of reads packed 2 bits/base + N-mask plane on the device (see tg_scan_pack)."""

    def __init__(self, reads):
        torch = _lib.require_cuda()
        L = _lib.load()
        self.torch, self.L = torch, L
        self.n = len(reads)
        self.lens = np.array([len(r) for r in reads], dtype=np.int32)
        padded = (self.lens.astype(np.int64) + ALIGN) // ALIGN * ALIGN          # at least one padding base per read
        self.base_off = np.zeros(self.n + 1, dtype=np.int64)
        self.base_off[1:] = np.cumsum(padded)
        total = int(self.base_off[-1])
        self.total = total
        ascii_h = pinned_bytes('ascii', max(total, 16))
        buf = ascii_h.numpy()
        H = _hostlists()
        if H is not None:
            H.pack_reads(reads if isinstance(reads, (list, tuple)) else list(reads), self.base_off, buf[:total], ord('N'))
        else:
            pad = b'N' * ALIGN
            joined = b''.join(x for i, r in enumerate(reads)
                              for x in (r.encode('latin-1', 'replace'), pad[:int(padded[i]) - len(r)]))
            if len(joined) != total:
                raise ValueError('reads must be str of single-byte characters')
            buf[:total] = np.frombuffer(joined, dtype=np.uint8)
        self._reads = reads
        self._bad_reads = None
        self.h2d_bytes = total + self.base_off.nbytes + self.lens.nbytes
        self.d_ascii = ascii_h.cuda(non_blocking=True)
        torch.cuda.current_stream().synchronize()                 # the page-locked staging buffer is reused by the next batch
        self.d_base_off = torch.from_numpy(self.base_off).cuda()
        self.d_len = torch.from_numpy(self.lens).cuda()
        self.d_pack2 = torch.zeros(total // 16 + 8, dtype=torch.int32, device='cuda')
        self.d_nmask = torch.zeros(total // 32 + 8, dtype=torch.int32, device='cuda')
        self._blk_read = None
        self.d_not_acgtn = torch.zeros(max(self.n, 1), dtype=torch.uint8, device='cuda')
        if self.n and total:
            _lib.check(L.tg_scan_pack(_lib.dptr(self.d_ascii), _lib.dptr(self.d_base_off), _lib.dptr(self.d_len), self.n,
                                      _lib.dptr(self.d_pack2), _lib.dptr(self.d_nmask), _lib.dptr(self.d_not_acgtn),
                                      _lib.stream_ptr(torch)))

    def check_rc_alphabet(self, flip):
        """RC() raises KeyError outside ACGTN (tg_util.py:433-434); mirror that for reads to be flipped."""
        if self._bad_reads is None:
            self._bad_reads = self.d_not_acgtn[:self.n].cpu().numpy().astype(bool)      # found by the pack kernel
        for i in np.nonzero(np.asarray(flip, dtype=bool) & self._bad_reads)[0]:
            RC(self._reads[i])                                   # raises the reference's KeyError

    def blk_read(self):
        """device int32: read index of every 128-base block of the padded base space."""
        if self._blk_read is None:
            torch = self.torch
            reps = torch.from_numpy(np.diff(self.base_off) // ALIGN).cuda()
            self._blk_read = torch.repeat_interleave(torch.arange(self.n, dtype=torch.int32, device='cuda'), reps)
        return self._blk_read

    def basecount(self, list_a, list_b):
        """int32 [n, 2]: get_telomere_base_count for two k-mer lists (tg_kmer.py:93-102)."""
        torch = self.torch
        out = torch.zeros((max(self.n, 1), 2), dtype=torch.int32, device='cuda')
        if self.n and self.total:
            _lib.check(self.L.tg_scan_basecount(_lib.dptr(self.d_pack2), _lib.dptr(self.d_nmask), self.total,
                                                _lib.dptr(self.blk_read()), self.n,
                                                _lib.dptr(kmer_table_pair(list_a, list_b)), _lib.dptr(out), _lib.stream_ptr(torch)))
        return out[:self.n]

    def orient(self, flip):
        """Reverse-complement the reads with flip[i] != 0 in place of tg_tel.py:265-266 (device, packed domain)."""
        torch = self.torch
        flip = np.ascontiguousarray(flip, dtype=np.uint8)
        self.check_rc_alphabet(flip)
        if not flip.any() or not self.total:
            return
        d_flip = torch.from_numpy(flip).cuda()
        p2, nm = torch.empty_like(self.d_pack2), torch.empty_like(self.d_nmask)
        _lib.check(self.L.tg_scan_orient(_lib.dptr(self.d_pack2), _lib.dptr(self.d_nmask), _lib.dptr(self.d_base_off),
                                         _lib.dptr(self.d_len), _lib.dptr(d_flip), self.n, _lib.dptr(p2), _lib.dptr(nm),
                                         _lib.stream_ptr(torch)))
        self.d_pack2, self.d_nmask = p2, nm

    def density_counts(self, list_a, list_b, window):
        """uint8 window counts for two k-mer lists, device tensors over the padded base space."""
        torch = self.torch
        ca = torch.empty(max(self.total, 16), dtype=torch.uint8, device='cuda')     # only [0, len - window) of each read is defined
        cb = torch.empty(max(self.total, 16), dtype=torch.uint8, device='cuda')
        if self.n and self.total:
            _lib.check(self.L.tg_scan_density(_lib.dptr(self.d_pack2), _lib.dptr(self.d_nmask), self.total,
                                              _lib.dptr(kmer_table_pair(list_a, list_b)), int(window), _lib.dptr(ca), _lib.dptr(cb),
                                              _lib.stream_ptr(torch)))
        return ca, cb

    def counts_to_host(self, ca, cb):
        """The two uint8 count planes -> page-locked host memory (numpy views, valid until the next call)."""
        torch = self.torch
        out = []
        for slot, t in (('cnt_a', ca), ('cnt_b', cb)):
            h = pinned_bytes(slot, t.numel())
            h.copy_(t, non_blocking=True)
            out.append(h.numpy())
        torch.cuda.current_stream().synchronize()
        return out

    def split_counts(self, flat, window):
        """host: per-read views cnt[0 : len - window) of a flat count array (numpy uint8)."""
        out = []
        for i in range(self.n):
            o = int(self.base_off[i])
            out.append(flat[o:o + max(int(self.lens[i]) - window, 0)])
        return out

    def hits_device(self, slice_read, slice_lo, slice_hi, kmer_list, issub):
        """Device part of hits(): returns (spans int32 [4, n_spans] device tensor: slice, k, start, end; unordered)."""
        torch = self.torch
        ns = len(slice_read)
        sr = np.ascontiguousarray(slice_read, dtype=np.int32)
        lo = np.ascontiguousarray(slice_lo, dtype=np.int32)
        hi = np.ascontiguousarray(slice_hi, dtype=np.int32)
        total = int((hi - lo).astype(np.int64).sum())
        if ns == 0 or total == 0:
            return torch.zeros((4, 0), dtype=torch.int32, device='cuda')
        tab = kmer_table(kmer_list, issub)
        d_sr, d_lo, d_hi = (torch.from_numpy(a).cuda() for a in (sr, lo, hi))
        cap = max(total // 4, 1024)
        while True:
            spans = torch.empty((4, cap), dtype=torch.int32, device='cuda')
            cnt = torch.zeros(1, dtype=torch.int64, device='cuda')
            _lib.check(self.L.tg_scan_hits(_lib.dptr(self.d_pack2), _lib.dptr(self.d_nmask), _lib.dptr(self.d_base_off),
                                           _lib.dptr(d_sr), _lib.dptr(d_lo), _lib.dptr(d_hi), ns,
                                           _lib.dptr(tab), _lib.dptr(spans[0]), _lib.dptr(spans[1]),
                                           _lib.dptr(spans[2]), _lib.dptr(spans[3]), cap, _lib.dptr(cnt),
                                           _lib.stream_ptr(torch)))
            n_sp = int(cnt.item())
            if n_sp <= cap:
                break
            cap = n_sp
        return spans[:, :n_sp]

    def hits(self, slice_read, slice_lo, slice_hi, kmer_list, issub):
        """get_nonoverlapping_kmer_hits for a batch of slices; returns one out_dat (list[K] of [[s, e], ...]) per slice."""
        import time
        K = len(kmer_list)
        ns = len(slice_read)
        t = [time.perf_counter()]
        d_sp = self.hits_device(slice_read, slice_lo, slice_hi, kmer_list, issub)
        t.append(time.perf_counter())
        # order by (slice, k, start) on the device, bring back (key, start, end) only
        torch = self.torch
        key = (d_sp[0].to(torch.int64) * K + d_sp[1].to(torch.int64)) * (1 << 32) + d_sp[2].to(torch.int64)
        key, order = torch.sort(key)
        pairs_d = torch.stack((d_sp[2][order], d_sp[3][order]), dim=1).contiguous()
        n_sp = int(key.numel())
        h_group = pinned_bytes('hit_group', max(n_sp, 1) * 8).view(torch.int64)[:n_sp]
        h_pairs = pinned_bytes('hit_pairs', max(n_sp, 1) * 8).view(torch.int32)[:2 * n_sp].view(n_sp, 2)
        h_group.copy_(key >> 32)
        h_pairs.copy_(pairs_d)
        group, pairs = h_group.numpy(), h_pairs.numpy()
        self.last_hits_d2h = group.nbytes + pairs.nbytes
        t.append(time.perf_counter())
        # one C-level conversion of all spans, then list slices per (slice, k); the cycle collector is paused while the
        # millions of small lists are made (it finds nothing and costs 4x)
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            H = _hostlists()
            cuts = np.ascontiguousarray(np.searchsorted(group, np.arange(ns * K + 1, dtype=np.int64)), dtype=np.int64)
            t.append(time.perf_counter())
            if H is not None:
                out = H.hit_lists(np.ascontiguousarray(pairs), cuts, ns, K)
                t.append(time.perf_counter())
            else:
                pairs = pairs.tolist()
                t.append(time.perf_counter())
                cuts = cuts.tolist()
                out = [[pairs[cuts[q * K + k]:cuts[q * K + k + 1]] for k in range(K)] for q in range(ns)]
            t.append(time.perf_counter())
            self.last_hits_phases = dict(zip(('kernel', 'sort_d2h', 'cuts', 'lists', 'rest'), np.diff(t).tolist()))
            return out
        finally:
            if gc_was_on:
                gc.enable()


# ------------------------------------------------------------------------------------------------
# the reference's per-read functions (batch of one)
# ------------------------------------------------------------------------------------------------
_density_cache = {}


def prime_density_cache(read, kmer_list, tel_window, result):
    """Lets a batch caller (tg_tel) hand pre-computed densities to code that still calls the per-read function."""
    _density_cache[(read, tuple(kmer_list), tel_window)] = result


def clear_density_cache():
    _density_cache.clear()


def get_telomere_kmer_density(read_dat, kmer_list, tel_window, mode='hifi', smoothing=False):
    """tg_kmer.py:66-90 -> (density list, all-zero list), both of length max(len(read) - tel_window, 0)."""
    if smoothing:
        raise NotImplementedError('smoothing=True (pywt) is not on the hot path; no reference caller uses it')
    if _density_cache:
        hit = _density_cache.get((read_dat, tuple(kmer_list), tel_window))
        if hit is not None:
            return hit
    pr = PackedReads([read_dat])
    ca, _cb = pr.density_counts(kmer_list, kmer_list, tel_window)
    n_out = max(len(read_dat) - tel_window, 0)
    cnt = ca[:n_out].cpu().numpy()
    dens = (cnt.astype(np.float64) / tel_window).tolist()
    return (dens, [0.0] * n_out)


def get_telomere_base_count(read_dat, kmer_list, mode='hifi'):
    """tg_kmer.py:93-102."""
    if len(read_dat) == 0:
        return 0
    pr = PackedReads([read_dat])
    return int(pr.basecount(kmer_list, kmer_list)[0, 0].item())


def get_nonoverlapping_kmer_hits(my_telseq, KMER_LIST, KMER_ISSUBSTRING, mode='hifi'):
    """tg_kmer.py:173-230 -> out_dat[k] = start-sorted list of [start, end] lists."""
    if len(my_telseq) == 0:
        return [[] for _ in KMER_LIST]
    pr = PackedReads([my_telseq])
    return pr.hits([0], [0], [len(my_telseq)], KMER_LIST, KMER_ISSUBSTRING)[0]


# ------------------------------------------------------------------------------------------------
# boundary calling (tg_kmer.py:105-170), batch of one on the device (tg_boundary.py)
# ------------------------------------------------------------------------------------------------
def get_telomere_regions(td_p_e0, td_p_e1, td_q_e0, td_q_e1, tel_window, pthresh, mode='hifi', smoothing=True):
    """tg_kmer.py:105-142 -> (p_vs_q_power float64 ndarray, [[start, end, None | 'p' | 'q'], ...])."""
    from .tg_boundary import BoundaryBatch
    lens = [len(td_p_e0), len(td_p_e1), len(td_q_e0), len(td_q_e1)]
    min_win_pos, max_win_pos = min(lens), max(lens)
    if max_win_pos < tel_window:
        return ([], [[0, 0, None]])
    power = np.zeros(max_win_pos)
    power[:min_win_pos] = (np.asarray(td_p_e0[:min_win_pos], dtype=np.float64)
                           - np.asarray(td_q_e0[:min_win_pos], dtype=np.float64))          # :114-115
    bb = BoundaryBatch([max_win_pos], tel_window)
    bb.power_from_host([power])
    if smoothing:
        bb.smooth()
    else:
        bb.no_smoothing()
    bb.regions(pthresh)
    return (bb.signal_host(0), bb.regions_host(0))


def mad(x, c=1.4826):
    """tg_kmer.py:145-149: median absolute deviation (numpy; the batch path computes it on the device)."""
    return c * np.median(np.abs(x - np.median(x)))


def wavelet_smooth(x, wavelet="db4", smoothing_level=5, fail_sigma=1e-3):
    """tg_kmer.py:152-170 for the only configuration the reference uses (db4, level 5, 1e-3); returns x itself when one
    of the early-outs applies, else the reconstruction (one sample longer than an odd-length input, as pywt.waverec)."""
    from . import tg_boundary
    if wavelet != 'db4' or smoothing_level != tg_boundary.SMOOTHING_LEVEL or fail_sigma != tg_boundary.FAIL_SIGMA:
        raise NotImplementedError('wavelet_smooth: only db4 / smoothing_level 5 / fail_sigma 1e-3 (the reference never passes others)')
    sig = np.asarray(x, dtype=np.float64)
    if len(sig) == 0:
        return x
    bb = tg_boundary.BoundaryBatch([len(sig)], 1)
    bb.power_from_host([sig])
    bb.smooth()
    if not int(bb.d_smoothed[0].item()):
        return x
    return bb.signal_host(0, trimmed=False)
