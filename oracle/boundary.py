"""CPU oracle of the telomere-boundary calling that follows the k-mer scan -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in numpy:
  * get_telomere_regions              reference source/tg_kmer.py:105-142   (pure Python: PINNED, see below)
  * mad, wavelet_smooth               reference source/tg_kmer.py:145-170   (wavelet_smooth calls pywt: PARITY UNPINNED)
  * get_terminating_tl                reference source/tg_tel.py:156-235    (pure Python: PINNED, see below)
  * pywt.wavedec / waverec / threshold for wavelet 'db4', mode 'per' ('periodization'), PyWavelets >= 1.4 as listed in the
    reference's requirements.  PyWavelets is a third-party dependency that is NOT under /root/reference and is not
    installed in this image: its published algorithm is restated here and **parity of these three functions is
    UNPINNED** (no vector of the real library is available).  What is restated:
      - dwt_max_level(n, 8) = floor(log2(n // 7)), 0 when n < 7                      (pywt/_extensions/c/common.c)
      - one periodised analysis step: the input is padded with a copy of its last sample when its length is odd, then
        out[o] = sum_j dec[j] * xp[(F/2 + 2 o - j) mod Np], o < ceil(N/2), j ascending    (c/convolution.template.c,
        downsampling_convolution_periodization: "i = F/2" start, step 2)
      - synthesis = the transpose of that orthogonal map (2 * len(cA) samples out)      (upsampling_convolution_valid_sf_periodization)
      - waverec drops the last approximation sample when len(cA) == len(cD) + 1          (pywt/_multilevel.py)
      - threshold(mode='soft'): data * max(1 - value / |data|, 0)                        (pywt/_thresholding.py)
    The db4 taps are PyWavelets' literals; tests check them against db4's defining equations (orthonormality, four vanishing
    moments; they hold to ~1e-12, the accuracy of the table), and the transform for perfect reconstruction and energy conservation.
  The two pure-Python reference functions ARE pinned: tests/golden/make_golden_boundary.py runs the reference's own
  get_telomere_regions / get_terminating_tl (imported from /root/reference) over `fake_pywt()` below and freezes their outputs;
  this restatement and the CUDA path must reproduce them exactly.

Floating point: every sum is evaluated in a fixed order with separate multiplies and adds (no fused multiply-add), so the
CUDA kernels (which use __dmul_rn/__dadd_rn in the same order) agree with this file bit for bit; against the real PyWavelets
the expected difference is rounding-order noise (~1e-16 relative), far inside the 1e-6 the north star allows.
"""
import types

import numpy as np

DB4_DEC_LO = np.array([-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
                       -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523])
DB4_DEC_HI = np.array([-0.23037781330885523, 0.7148465705525415, -0.6308807679295904, -0.02798376941698385,
                       0.18703481171888114, 0.030841381835986965, -0.032883011666982945, -0.010597401784997278])
F = 8


def dwt_max_level(n, filter_len=F):
    if filter_len <= 1 or n < filter_len - 1:
        return 0
    return int(n // (filter_len - 1)).bit_length() - 1


def dwt_per(x):
    """One periodised db4 analysis step -> (cA, cD), each ceil(len(x)/2) long."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    xp = x if n % 2 == 0 else np.concatenate([x, x[-1:]])
    npad = len(xp)
    m = npad // 2
    o = np.arange(m)
    ca = np.zeros(m)
    cd = np.zeros(m)
    for j in range(F):
        v = xp[(F // 2 + 2 * o - j) % npad]
        ca = ca + DB4_DEC_LO[j] * v
        cd = cd + DB4_DEC_HI[j] * v
    return ca, cd


def idwt_per(ca, cd):
    """Inverse of dwt_per on an even-length signal: x[n] = sum over j of the parity of n, ascending, of
    (lo[j] * cA[k] + hi[j] * cD[k]), k = ((n + j - F/2) / 2) mod len(cA)."""
    ca = np.asarray(ca, dtype=np.float64)
    cd = np.asarray(cd, dtype=np.float64)
    m = len(ca)
    if len(cd) != m:
        raise ValueError('coefficient arrays must have the same length')
    n = np.arange(2 * m)
    x = np.zeros(2 * m)
    for t in range(F // 2):
        j = 2 * t + (n & 1)
        k = ((n + j - F // 2) // 2) % m
        x = x + DB4_DEC_LO[j] * ca[k]
        x = x + DB4_DEC_HI[j] * cd[k]
    return x


def wavedec_per(x, level=None):
    """pywt.wavedec(x, 'db4', mode='per') -> [cA_n, cD_n, ..., cD_1]."""
    x = np.asarray(x, dtype=np.float64)
    if level is None:
        level = dwt_max_level(len(x))
    out = []
    a = x
    for _ in range(level):
        a, d = dwt_per(a)
        out.append(d)
    out.append(a)
    return out[::-1]


def waverec_per(coeffs):
    """pywt.waverec(coeffs, 'db4', mode='per'); even length out (one more sample than an odd-length input had)."""
    a = np.asarray(coeffs[0], dtype=np.float64)
    for d in coeffs[1:]:
        d = np.asarray(d, dtype=np.float64)
        if len(a) == len(d) + 1:
            a = a[:-1]
        a = idwt_per(a, d)
    return a


def soft_threshold(data, value):
    """pywt.threshold(data, value, mode='soft')."""
    data = np.asarray(data, dtype=np.float64)
    mag = np.absolute(data)
    with np.errstate(divide='ignore', invalid='ignore'):
        t = 1.0 - value / mag
        t = np.clip(t, 0.0, None)
        return data * t


def fake_pywt():
    """A module object exposing the three PyWavelets functions the reference calls, backed by this restatement -- lets the
    reference's own tg_kmer / tg_tel code run in the build container (golden generation)."""
    m = types.ModuleType('pywt')

    def wavedec(x, wavelet, mode='symmetric', level=None):
        assert wavelet == 'db4' and mode == 'per'
        return wavedec_per(x, level)

    def waverec(coeffs, wavelet, mode='symmetric'):
        assert wavelet == 'db4' and mode == 'per'
        return waverec_per(list(coeffs))

    def threshold(data, value, mode='soft', substitute=0):
        assert mode == 'soft'
        return soft_threshold(data, value)
    m.wavedec, m.waverec, m.threshold = wavedec, waverec, threshold
    return m


def mad(x, c=1.4826):
    """tg_kmer.py:145-149."""
    return c * np.median(np.abs(x - np.median(x)))


def wavelet_smooth(x, smoothing_level=5, fail_sigma=1e-3):
    """tg_kmer.py:152-170 for wavelet 'db4'."""
    coeff = wavedec_per(x)
    if len(coeff) < smoothing_level:
        return x
    sigma = mad(coeff[-smoothing_level])
    if sigma < fail_sigma:
        return x
    uthresh = sigma * np.sqrt(2 * np.log(len(x)))
    coeff[1:] = (soft_threshold(i, uthresh) for i in coeff[1:])
    return waverec_per(coeff)


def p_vs_q_power(td_p_e0, td_p_e1, td_q_e0, td_q_e1):
    """tg_kmer.py:106-115 (before smoothing)."""
    lens = [len(td_p_e0), len(td_p_e1), len(td_q_e0), len(td_q_e1)]
    out = np.zeros(max(lens))
    n = min(lens)
    out[:n] = np.asarray(td_p_e0[:n], dtype=np.float64) - np.asarray(td_q_e0[:n], dtype=np.float64)
    return out


def regions_of(power, pthresh):
    """tg_kmer.py:120-140: maximal runs of one class ('p' >= pthresh, 'q' <= -pthresh, None in between) -> [[start, end, class]].
    A run's end is one past its last index, except a one-wide LAST run, which ends at len(power)-1 (:138-139).  An empty
    signal raises the reference's IndexError (:120)."""
    power = np.asarray(power, dtype=np.float64)
    if len(power) == 0:
        raise IndexError('index 0 is out of bounds for axis 0 with size 0')
    cls = np.where(power >= pthresh, 1, np.where(power <= -pthresh, 2, 0))
    starts = np.concatenate([[0], np.nonzero(cls[1:] != cls[:-1])[0] + 1])
    ends = np.concatenate([starts[1:], [len(power)]])
    if starts[-1] == len(power) - 1:
        ends[-1] = len(power) - 1
    names = [None, 'p', 'q']
    return [[int(s), int(e), names[int(cls[s])]] for s, e in zip(starts, ends)]


def get_telomere_regions(td_p_e0, td_p_e1, td_q_e0, td_q_e1, tel_window, pthresh, smoothing=True):
    """tg_kmer.py:105-142."""
    lens = [len(td_p_e0), len(td_p_e1), len(td_q_e0), len(td_q_e1)]
    if max(lens) < tel_window:
        return ([], [[0, 0, None]])
    power = p_vs_q_power(td_p_e0, td_p_e1, td_q_e0, td_q_e1)
    if smoothing:
        power = wavelet_smooth(power)
    power = power[:-tel_window]
    return (power, regions_of(power, pthresh))


def terminating_tl_of_regions(tel_regions, pq, tel_window, min_tel_score):
    """tg_tel.py:164-233 on an existing region list -> (tel_len, edge_nontel, largest interstitial (a, b) or None)."""
    window_adj = tel_window // 2
    sizes = [r[1] - r[0] for r in tel_regions]
    score = [s if r[2] is not None else -s for s, r in zip(sizes, tel_regions)]
    my_tel_len = None
    edge_nontel = 0
    interstitial_regions = list(tel_regions)
    if pq == 'p':
        if tel_regions[0][2] is None:
            edge_nontel = sizes[0]
        cum = np.cumsum(score)
        max_i = int(np.argmax(cum))                    # first maximum, as tg_util.posmax
        if cum[max_i] >= min_tel_score:
            my_tel_len = int(tel_regions[max_i][1] + window_adj)
            interstitial_regions = tel_regions[max_i:]
    elif pq == 'q':
        if tel_regions[-1][2] is None:
            edge_nontel = sizes[-1]
        cum = np.cumsum(score[::-1])[::-1]
        max_i = int(np.argmax(cum))
        if cum[max_i] >= min_tel_score:
            my_tel_len = int(tel_regions[-1][1] - tel_regions[max_i][0] + window_adj)
            interstitial_regions = tel_regions[:max_i]
    largest = None
    tels = [r for r in interstitial_regions if r[2] is not None]
    if tels:
        best = None
        a = tels[0][0]
        for i in range(len(tels)):
            last = i == len(tels) - 1
            if not last:
                my_size = tels[i][1] - tels[i][0]
                next_size = tels[i + 1][1] - tels[i + 1][0]
                joined = max(my_size, next_size) > tels[i + 1][0] - tels[i][1]
            if last or not joined:
                cand = (a + window_adj, tels[i][1] + window_adj)
                key = (cand[1] - cand[0], cand)
                if best is None or key > best:
                    best = key
                if not last:
                    a = tels[i + 1][0]
        largest = best[1]
    if my_tel_len is None or my_tel_len <= 0:
        return (0, 0, largest)
    return (my_tel_len, edge_nontel, largest)


def get_terminating_tl(dens_p, dens_q, pq, tel_window, pthresh, min_tel_score):
    """tg_tel.py:156-235 from the two density tracks of one read (the all-zero second tracks have the same lengths)."""
    (_power, tel_regions) = get_telomere_regions(dens_p, dens_p, dens_q, dens_q, tel_window, pthresh)
    return terminating_tl_of_regions(tel_regions, pq, tel_window, min_tel_score)
