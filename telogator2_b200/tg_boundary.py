"""Telomere-boundary calling for a batch of reads on the GPU.

Device counterpart of tg_kmer.get_telomere_regions / wavelet_smooth / mad (tg_kmer.py:105-170) and of the scoring in
tg_tel.get_terminating_tl (tg_tel.py:164-233).  The reference's per-read entry points are thin wrappers (batch of one) in
telogator2_b200.tg_kmer / tg_tel; tg_tel.get_tel_repeat_comp_parallel feeds the window counts of all reads straight from
the scan kernels, so neither the densities nor the power signal ever visit the host.

PyWavelets' db4 / periodisation transform is restated (csrc/tg_bound.cu, oracle/boundary.py); its parity against the real
library is unpinned (the library is not in the reference tree and not installed), everything around it is pinned against
the reference's own code (tests/golden/boundary_golden.json.gz).
"""
import numpy as np

from . import _lib

SMOOTHING_LEVEL = 5          # wavelet_smooth defaults (tg_kmer.py:152); get_telomere_regions never overrides them
FAIL_SIGMA = 1e-3
SORT_SMEM = 4096             # BD_SORT_SMEM in csrc/tg_bound.cu
_CLS_NAMES = (None, 'p', 'q')


def _levels(n):
    """pywt.dwt_max_level(n, 8) for an int64 array."""
    q = np.maximum(n // 7, 1)
    return np.where(n >= 7, np.frexp(q.astype(np.float64))[1].astype(np.int64) - 1, 0)       # exact floor(log2 q)


def _pow2ceil(m):
    return 1 << int(max(int(m) - 1, 0)).bit_length()


class BoundaryBatch:
    """Power signals of a batch of reads in HBM; smooth() -> regions() -> terminating()."""

    def __init__(self, sig_n, window):
        torch = _lib.require_cuda()
        self.torch, self.L = torch, _lib.load()
        self.window = int(window)
        n = np.ascontiguousarray(sig_n, dtype=np.int64)
        if n.size and int(n.max()) >= 2 ** 31 - 2:
            raise NotImplementedError('reads of 2^31 bases or more')
        self.n_reads = int(n.size)
        self.sig_n = n.astype(np.int32)
        self.max_n = int(n.max()) if n.size else 0
        self.total_n = int(n.sum())
        lev = _levels(n)
        dec = (n >= self.window) & (lev + 1 >= SMOOTHING_LEVEL)        # reads wavelet_smooth decomposes
        a_sz = np.zeros_like(n)
        d_sz = np.zeros_like(n)
        m = n.copy()
        m_sig = np.zeros_like(n)                                        # len(coeff[-SMOOTHING_LEVEL])
        for l in range(1, int(lev.max()) + 1 if n.size else 1):
            act = dec & (lev >= l)
            mo = (m + 1) >> 1
            a_sz += np.where(act, (mo + 1) & ~1, 0)
            d_sz += np.where(act, mo, 0)
            m = np.where(act, mo, m)
            m_sig = np.where(act & ((l == SMOOTHING_LEVEL) | ((lev + 1 == SMOOTHING_LEVEL) & (l == lev))), mo, m_sig)
        self.x_off = np.zeros(self.n_reads + 1, dtype=np.int64)
        self.x_off[1:] = np.cumsum((n + 1) & ~1)
        self.a_off = np.zeros(self.n_reads + 1, dtype=np.int64)
        self.a_off[1:] = np.cumsum(a_sz)
        self.d_off = np.zeros(self.n_reads + 1, dtype=np.int64)
        self.d_off[1:] = np.cumsum(d_sz)
        big = np.nonzero(m_sig > SORT_SMEM)[0]
        self.sc_off = np.full(self.n_reads, -1, dtype=np.int64)
        sc_total = 0
        for r in big:
            self.sc_off[r] = sc_total
            sc_total += _pow2ceil(m_sig[r])
        with np.errstate(divide='ignore', invalid='ignore'):
            self.log_factor = np.where(n > 0, np.sqrt(2 * np.log(np.maximum(n, 1).astype(np.float64))), 0.0)   # tg_kmer.py:165
        cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()        # noqa: E731
        self.d_sig_n, self.d_x_off, self.d_a_off, self.d_d_off = cu(self.sig_n), cu(self.x_off), cu(self.a_off), cu(self.d_off)
        self.d_sc_off, self.d_log_factor = cu(self.sc_off), cu(self.log_factor)
        f64 = dict(dtype=torch.float64, device='cuda')
        self.d_X = torch.empty(max(int(self.x_off[-1]), 1), **f64)
        self.d_A = torch.empty(max(int(self.a_off[-1]), 1), **f64)
        self.d_D = torch.empty(max(int(self.d_off[-1]), 1), **f64)
        self.d_scratch = torch.empty(max(sc_total, 1), **f64)
        self.d_out_len = torch.empty(max(self.n_reads, 1), dtype=torch.int32, device='cuda')
        self.d_smoothed = torch.empty(max(self.n_reads, 1), dtype=torch.uint8, device='cuda')
        self.d_sigma = torch.empty(max(self.n_reads, 1), **f64)
        self.d_uthresh = torch.empty(max(self.n_reads, 1), **f64)
        self._smoothed = False
        self._regions = None

    @staticmethod
    def device_bytes(sig_n):
        """HBM this batch needs (3 float64 slabs of ~n each plus the region arrays)."""
        return int(np.asarray(sig_n, dtype=np.int64).sum()) * 30 + 4096

    # ---- filling X ----
    def power_from_counts(self, d_cnt_p, d_cnt_q, d_base_off):
        """X = cnt_p / window - cnt_q / window from the uint8 planes of tg_scan_density (padded base space)."""
        torch = self.torch
        _lib.check(self.L.tg_bound_power(_lib.dptr(d_cnt_p), _lib.dptr(d_cnt_q), _lib.dptr(d_base_off), _lib.dptr(self.d_sig_n),
                                         _lib.dptr(self.d_x_off), self.n_reads, self.max_n, self.total_n, self.window, _lib.dptr(self.d_X),
                                         _lib.stream_ptr(torch)))

    def power_from_host(self, signals):
        """X = the given float64 signals (one per read)."""
        torch = self.torch
        host = np.zeros(max(int(self.x_off[-1]), 1), dtype=np.float64)
        for r, s in enumerate(signals):
            host[self.x_off[r]:self.x_off[r] + len(s)] = s
        self.d_X.copy_(torch.from_numpy(host))

    # ---- the three stages ----
    def smooth(self):
        torch = self.torch
        _lib.check(self.L.tg_bound_smooth(_lib.dptr(self.d_X), _lib.dptr(self.d_A), _lib.dptr(self.d_D), _lib.dptr(self.d_x_off),
                                          _lib.dptr(self.d_a_off), _lib.dptr(self.d_d_off), _lib.dptr(self.d_sig_n), self.n_reads,
                                          self.max_n, self.total_n, self.window, SMOOTHING_LEVEL, FAIL_SIGMA, _lib.dptr(self.d_log_factor),
                                          _lib.dptr(self.d_scratch), _lib.dptr(self.d_sc_off), _lib.dptr(self.d_uthresh),
                                          _lib.dptr(self.d_sigma), _lib.dptr(self.d_out_len), _lib.dptr(self.d_smoothed),
                                          _lib.stream_ptr(torch)))
        self._smoothed = True

    def no_smoothing(self):
        """get_telomere_regions(smoothing=False): the signal stays as it is."""
        self.d_out_len.copy_(self.d_sig_n)
        self.d_smoothed.zero_()
        self._smoothed = True

    def regions(self, pthresh):
        """Region walk; leaves (reg_off, reg_count, reg_start, reg_cls, last_end) on the device.  Raises the reference's
        IndexError for a read whose trimmed signal is empty (tg_kmer.py:120)."""
        torch = self.torch
        assert self._smoothed
        nr = max(self.n_reads, 1)
        cnt = torch.zeros(nr, dtype=torch.int32, device='cuda')
        last_end = torch.zeros(nr, dtype=torch.int32, device='cuda')
        err = torch.zeros(nr, dtype=torch.uint8, device='cuda')

        def run(off, start, cls):
            _lib.check(self.L.tg_bound_regions(_lib.dptr(self.d_X), _lib.dptr(self.d_x_off), _lib.dptr(self.d_sig_n),
                                               _lib.dptr(self.d_out_len), self.n_reads, self.window, float(pthresh),
                                               None if off is None else _lib.dptr(off), _lib.dptr(cnt),
                                               None if start is None else _lib.dptr(start), None if cls is None else _lib.dptr(cls),
                                               _lib.dptr(last_end), _lib.dptr(err), _lib.stream_ptr(torch)))
        run(None, None, None)
        if self.n_reads and bool(err[:self.n_reads].any().item()):
            raise IndexError('index 0 is out of bounds for axis 0 with size 0')
        off = torch.zeros(nr + 1, dtype=torch.int64, device='cuda')
        off[1:] = torch.cumsum(cnt.to(torch.int64), 0)
        total = int(off[self.n_reads].item()) if self.n_reads else 0
        start = torch.empty(max(total, 1), dtype=torch.int32, device='cuda')
        cls = torch.empty(max(total, 1), dtype=torch.uint8, device='cuda')
        run(off, start, cls)
        self._regions = (off, cnt, start, cls, last_end)
        return self._regions

    def terminating(self, pq, min_tel_score):
        """int32 [n_reads, 4] on the host: tel_len, edge_nontel, interstitial (a, b) or (-1, -1) -- tg_tel.py:164-233."""
        torch = self.torch
        (off, cnt, start, cls, last_end) = self._regions
        out = torch.empty((max(self.n_reads, 1), 4), dtype=torch.int32, device='cuda')
        if pq not in ('p', 'q'):                 # neither branch of tg_tel.py:168/184 runs: no telomere, all regions interstitial
            raise NotImplementedError("get_terminating_tl: pq must be 'p' or 'q'")
        _lib.check(self.L.tg_bound_terminating(_lib.dptr(off), _lib.dptr(cnt), _lib.dptr(start), _lib.dptr(cls), _lib.dptr(last_end),
                                               self.n_reads, 1 if pq == 'p' else 0, self.window, float(min_tel_score),
                                               _lib.dptr(out), _lib.stream_ptr(torch)))
        return out[:self.n_reads].cpu().numpy()

    # ---- host views (per-read wrappers, tests, plots) ----
    def signal_host(self, r, trimmed=True):
        """float64 signal of read r after smooth(): [: -window] of it when trimmed (what get_telomere_regions returns)."""
        n_out = int(self.d_out_len[r].item())
        x = self.d_X[int(self.x_off[r]):int(self.x_off[r]) + n_out].cpu().numpy()
        return x[:-self.window] if trimmed else x

    def regions_host(self, r):
        """[[start, end, None | 'p' | 'q'], ...] of read r, as tg_kmer.get_telomere_regions builds them."""
        (off, cnt, start, cls, last_end) = self._regions
        lo, c = int(off[r].item()), int(cnt[r].item())
        s = start[lo:lo + c].cpu().numpy().tolist()
        k = cls[lo:lo + c].cpu().numpy().tolist()
        e = s[1:] + [int(last_end[r].item())]
        return [[s[i], e[i], _CLS_NAMES[k[i]]] for i in range(c)]


def terminating_tl_from_counts(d_cnt_p, d_cnt_q, d_base_off, lens, window, pthresh, min_tel_score, pq='q', budget_bytes=None):
    """get_terminating_tl for every read of a scanned batch -> int32 [n_reads, 4] (tel_len, edge_nontel, a, b | -1, -1).
    Reads are processed in groups that fit `budget_bytes` of HBM (default: half of what is free)."""
    torch = _lib.require_cuda()
    lens = np.asarray(lens, dtype=np.int64)
    sig_n = np.maximum(lens - int(window), 0)
    if budget_bytes is None:
        budget_bytes = torch.cuda.mem_get_info()[0] // 2
    out = np.zeros((len(lens), 4), dtype=np.int32)
    need = sig_n * 30 + 64
    r0 = 0
    while r0 < len(lens):
        r1 = r0 + max(int(np.searchsorted(np.cumsum(need[r0:]), budget_bytes, side='right')), 1)
        bb = BoundaryBatch(sig_n[r0:r1], window)
        bb.power_from_counts(d_cnt_p, d_cnt_q, d_base_off[r0:r1 + 1])
        bb.smooth()
        bb.regions(pthresh)
        out[r0:r1] = bb.terminating(pq, min_tel_score)
        del bb
        r0 = r1
    return out
