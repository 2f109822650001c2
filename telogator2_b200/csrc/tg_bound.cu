// Telomere-boundary calling on the device: the p-vs-q power signal of every read, its db4 wavelet
// soft-threshold smoothing, the region walk and the terminal-telomere scoring.
//
//   reference                                        here
//   tg_kmer.get_telomere_regions   tg_kmer.py:105-142   bd_power_kernel, bd_regions_kernel
//   tg_kmer.mad / wavelet_smooth   tg_kmer.py:145-170   bd_dwt_kernel, bd_sigma_kernel, bd_threshold_kernel, bd_idwt_kernel
//     (pywt.wavedec / threshold / waverec, wavelet 'db4', mode 'per': PyWavelets is not part of the reference tree; the
//      algorithm is the one restated in oracle/boundary.py, whose header lists what is pinned and what is not)
//   tg_tel.get_terminating_tl      tg_tel.py:164-233    bd_terminating_kernel
//
// Data layout (float64, one slab triple per read, offsets in elements from the host):
//   X  even(n) samples   the power signal, overwritten by the smoothed signal (n or even(n) samples, as pywt.waverec returns)
//   A  approximations a_1..a_L, a_l padded to an even length (the reconstruction of level l lands in a_{l-1}'s slot)
//   D  details d_1..d_L, contiguous
// where n = max(len(read) - window, 0), m_0 = n, m_l = ceil(m_{l-1} / 2), L = floor(log2(n / 7)).
// Every sum is evaluated in the oracle's order with separate multiplies and adds (__dmul_rn / __dadd_rn: no contraction into
// FMA), so the results are bit-identical to oracle/boundary.py.  HBM-bound: ~100 B of float64 traffic per base for all levels.
#include "tg_common.cuh"

namespace {

__constant__ double c_lo[8] = {-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
                               -0.02798376941698385,  0.6308807679295904,   0.7148465705525415,   0.23037781330885523};
__constant__ double c_hi[8] = {-0.23037781330885523,  0.7148465705525415,   -0.6308807679295904,  -0.02798376941698385,
                               0.18703481171888114,   0.030841381835986965, -0.032883011666982945, -0.010597401784997278};

constexpr int BD_TILE = 256;
constexpr int BD_SORT_SMEM = 4096;            // doubles sorted in shared memory; longer coefficient arrays use the scratch slab

struct BdParams {
    double *X, *A, *D;
    const int64_t *x_off, *a_off, *d_off;     // [n_reads + 1]
    const int32_t *sig_n;                     // n per read
    int n_reads, window, smoothing_level;
    double fail_sigma;
    const double *log_factor;                 // sqrt(2 ln n) per read (host: numpy)
    double *scratch;
    const int64_t *sc_off;                    // scratch offset per read, -1 = shared memory is enough
    double *uthresh, *sigma;
    int32_t *out_len;
    uint8_t *smoothed;
};

__device__ __forceinline__ int bd_max_level(int n)       // pywt.dwt_max_level(n, 8)
{
    return n < 7 ? 0 : 31 - __clz(n / 7);
}

__device__ __forceinline__ bool bd_decomposed(int n, const BdParams &P)      // does wavelet_smooth get past its first early-out?
{
    return n >= P.window && bd_max_level(n) + 1 >= P.smoothing_level;
}

struct LevelGeom {
    int m_in, m_out;                          // lengths of the level's input and of each of its outputs
    long long a_rel, d_rel, a_prev_rel;       // a_l, d_l, a_{l-1} inside the read's A / D slabs
};

__device__ __forceinline__ LevelGeom bd_level(int n, int lev)
{
    LevelGeom g;
    int m = n;
    long long a = 0, d = 0, ap = 0;
    for (int l = 1;; l++) {
        const int mo = (m + 1) >> 1;
        if (l == lev) {
            g.m_in = m; g.m_out = mo; g.a_rel = a; g.d_rel = d; g.a_prev_rel = ap;
            return g;
        }
        ap = a;
        a += (mo + 1) & ~1;
        d += mo;
        m = mo;
    }
}

// ---- power signal from the two window-count planes of the scan: cnt_p / W - cnt_q / W (tg_kmer.py:90, 114-115) ----
__global__ void __launch_bounds__(BD_TILE) bd_power_kernel(const uint8_t *__restrict__ cnt_p, const uint8_t *__restrict__ cnt_q,
                                                           const int64_t *__restrict__ base_off, BdParams P, int read0)
{
    // c / W for every possible count: the reference's float64 division, done once per block instead of twice per sample
    __shared__ double s_dens[256];
    s_dens[threadIdx.x] = __ddiv_rn((double)threadIdx.x, (double)P.window);
    __syncthreads();
    const int r = read0 + blockIdx.y;
    const int n = P.sig_n[r];
    const uint8_t *cp = cnt_p + base_off[r], *cq = cnt_q + base_off[r];
    double *x = P.X + P.x_off[r];
    for (int i = blockIdx.x * BD_TILE + threadIdx.x; i < n; i += gridDim.x * BD_TILE)
        x[i] = __dsub_rn(s_dens[cp[i]], s_dens[cq[i]]);
}

// ---- one periodised analysis level for every read that has it ----
__global__ void __launch_bounds__(BD_TILE) bd_dwt_kernel(BdParams P, int lev, int read0)
{
    const int r = read0 + blockIdx.y;
    const int n = P.sig_n[r];
    if (!bd_decomposed(n, P) || bd_max_level(n) < lev) return;
    const LevelGeom g = bd_level(n, lev);
    const double *in = lev == 1 ? P.X + P.x_off[r] : P.A + P.a_off[r] + g.a_prev_rel;
    double *oa = P.A + P.a_off[r] + g.a_rel, *od = P.D + P.d_off[r] + g.d_rel;
    const int np = (g.m_in + 1) & ~1;                     // an odd-length input is padded with a copy of its last sample
    // (the grid has a few blocks per read; each walks its share of the read's tiles)
    for (int o = blockIdx.x * BD_TILE + threadIdx.x; o < g.m_out; o += gridDim.x * BD_TILE) {
        double ca = 0.0, cd = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int t = 4 + 2 * o - j;
            if (t < 0) t += np; else if (t >= np) t -= np;
            if (t >= g.m_in) t = g.m_in - 1;
            const double v = in[t];
            ca = __dadd_rn(ca, __dmul_rn(c_lo[j], v));
            cd = __dadd_rn(cd, __dmul_rn(c_hi[j], v));
        }
        oa[o] = ca;
        od[o] = cd;
    }
}

__device__ void bd_bitonic_sort(double *buf, int npow)
{
    for (int k = 2; k <= npow; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double a = buf[i], b = buf[ixj];
                    if ((a > b) == ((i & k) == 0)) { buf[i] = b; buf[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ double bd_median_sorted(const double *buf, int m)      // numpy.median of m sorted values
{
    return (m & 1) ? buf[m >> 1] : __dmul_rn(__dadd_rn(buf[(m >> 1) - 1], buf[m >> 1]), 0.5);
}

// ---- sigma = mad(coeff[-smoothing_level]) and the decision to smooth (tg_kmer.py:158-167), one CTA per read ----
__global__ void __launch_bounds__(256) bd_sigma_kernel(BdParams P, int read0)
{
    __shared__ double sh[BD_SORT_SMEM];
    __shared__ double s_med;
    const int r = read0 + blockIdx.x;
    const int n = P.sig_n[r];
    if (!bd_decomposed(n, P)) {
        if (threadIdx.x == 0) { P.out_len[r] = n; P.smoothed[r] = 0; P.sigma[r] = 0.0; P.uthresh[r] = 0.0; }
        return;
    }
    const int L = bd_max_level(n), S = P.smoothing_level;
    const double *src;
    int m;
    if (L >= S) {                                         // coeff[-S] = cD_S
        const LevelGeom g = bd_level(n, S);
        src = P.D + P.d_off[r] + g.d_rel;
        m = g.m_out;
    } else {                                              // L + 1 == S: coeff[-S] = coeff[0] = cA_L
        const LevelGeom g = bd_level(n, L);
        src = P.A + P.a_off[r] + g.a_rel;
        m = g.m_out;
    }
    int npow = 1;
    while (npow < m) npow <<= 1;
    double *buf = npow <= BD_SORT_SMEM ? sh : P.scratch + P.sc_off[r];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = threadIdx.x; i < npow; i += blockDim.x) buf[i] = i < m ? src[i] : inf;
    __syncthreads();
    bd_bitonic_sort(buf, npow);
    if (threadIdx.x == 0) s_med = bd_median_sorted(buf, m);
    __syncthreads();
    const double med = s_med;
    __syncthreads();
    for (int i = threadIdx.x; i < npow; i += blockDim.x) buf[i] = i < m ? fabs(__dsub_rn(src[i], med)) : inf;
    __syncthreads();
    bd_bitonic_sort(buf, npow);
    if (threadIdx.x == 0) {
        const double sigma = __dmul_rn(1.4826, bd_median_sorted(buf, m));
        const bool smooth = !(sigma < P.fail_sigma);
        P.sigma[r] = sigma;
        P.uthresh[r] = __dmul_rn(sigma, P.log_factor[r]);
        P.smoothed[r] = smooth ? 1 : 0;
        P.out_len[r] = smooth ? ((n + 1) & ~1) : n;       // pywt.waverec returns an even number of samples
    }
}

// ---- soft threshold of every detail coefficient: d * max(1 - u / |d|, 0) (pywt.threshold, mode 'soft') ----
__global__ void __launch_bounds__(BD_TILE) bd_threshold_kernel(BdParams P, int read0)
{
    const int r = read0 + blockIdx.y;
    if (!P.smoothed[r]) return;
    double *dd = P.D + P.d_off[r];
    const long long len = P.d_off[r + 1] - P.d_off[r];
    const double u = P.uthresh[r];
    for (long long i = (long long)blockIdx.x * BD_TILE + threadIdx.x; i < len; i += (long long)gridDim.x * BD_TILE) {
        const double d = dd[i];
        double t = __dsub_rn(1.0, __ddiv_rn(u, fabs(d)));
        if (t < 0.0) t = 0.0;
        dd[i] = __dmul_rn(d, t);
    }
}

// ---- one periodised synthesis level: the transpose of bd_dwt_kernel's map ----
__global__ void __launch_bounds__(BD_TILE) bd_idwt_kernel(BdParams P, int lev, int read0)
{
    const int r = read0 + blockIdx.y;
    if (!P.smoothed[r]) return;
    const int n = P.sig_n[r];
    if (bd_max_level(n) < lev) return;
    const LevelGeom g = bd_level(n, lev);
    const int m = g.m_out;
    const double *a = P.A + P.a_off[r] + g.a_rel;         // (a reconstruction one sample longer than d_l is cut to len(d_l))
    const double *d = P.D + P.d_off[r] + g.d_rel;
    double *out = lev == 1 ? P.X + P.x_off[r] : P.A + P.a_off[r] + g.a_prev_rel;
    for (int x = blockIdx.x * BD_TILE + threadIdx.x; x < 2 * m; x += gridDim.x * BD_TILE) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int j = 2 * t + (x & 1);
            int k = (x + j - 4) >> 1;
            if (k < 0) k += m; else if (k >= m) k -= m;
            v = __dadd_rn(v, __dmul_rn(c_lo[j], a[k]));
            v = __dadd_rn(v, __dmul_rn(c_hi[j], d[k]));
        }
        out[x] = v;
    }
}

// ---- region walk (tg_kmer.py:118-140): maximal runs of one class; count mode (reg_off == nullptr) or fill mode ----
struct RegParams {
    const double *X;
    const int64_t *x_off;
    const int32_t *sig_n, *out_len;
    int n_reads, window;
    double pthresh;
    const int64_t *reg_off;
    int32_t *reg_count, *reg_start, *last_end;
    uint8_t *reg_cls, *err;
};

__device__ __forceinline__ int bd_class(double v, double thr)
{
    return v >= thr ? 1 : (v <= -thr ? 2 : 0);
}

__global__ void __launch_bounds__(256) bd_regions_kernel(RegParams P)
{
    __shared__ int s_warp[8];
    __shared__ int s_last;
    const int r = blockIdx.x;
    const int n = P.sig_n[r];
    const bool fill = P.reg_off != nullptr;
    if (n < P.window) {                                   // ([], [[0, 0, None]])   tg_kmer.py:111-112
        if (threadIdx.x == 0) {
            P.reg_count[r] = 1; P.last_end[r] = 0; P.err[r] = 0;
            if (fill) { P.reg_start[P.reg_off[r]] = 0; P.reg_cls[P.reg_off[r]] = 0; }
        }
        return;
    }
    const int plen = P.out_len[r] - P.window;             // p_vs_q_power[:-tel_window]
    if (plen <= 0) {                                      // the reference raises IndexError at p_vs_q_power[0]
        if (threadIdx.x == 0) { P.reg_count[r] = 0; P.last_end[r] = 0; P.err[r] = 1; }
        return;
    }
    const double *x = P.X + P.x_off[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_last = 0;
    int base = 0;                                         // regions before this tile
    for (int i0 = 0; i0 < plen; i0 += 256) {
        const int i = i0 + threadIdx.x;
        int cls = 0;
        bool flag = false;
        if (i < plen) {
            cls = bd_class(x[i], P.pthresh);
            flag = i == 0 || cls != bd_class(x[i - 1], P.pthresh);
        }
        const uint32_t bal = __ballot_sync(TG_FULL_MASK, flag);
        __syncthreads();                                  // (s_warp of the previous tile has been read)
        if (lane == 0) s_warp[warp] = __popc(bal);
        if (flag) atomicMax(&s_last, i);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const int c = s_warp[w];
            before += w < warp ? c : 0;
            total += c;
        }
        if (fill && flag) {
            const long long slot = P.reg_off[r] + base + before + __popc(bal & ((1u << lane) - 1u));
            P.reg_start[slot] = i;
            P.reg_cls[slot] = (uint8_t)cls;
        }
        base += total;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        P.reg_count[r] = base;
        P.last_end[r] = s_last == plen - 1 ? plen - 1 : plen;     // a one-wide last region ends at len - 1 (tg_kmer.py:138-139)
        P.err[r] = 0;
    }
}

// ---- terminal telomere length, non-telomeric edge and largest interstitial telomere region (tg_tel.py:164-233) ----
struct TermParams {
    const int64_t *reg_off;
    const int32_t *reg_count, *reg_start, *last_end;
    const uint8_t *reg_cls;
    int n_reads, pq_is_p, window;
    double min_tel_score;
    int32_t *out;                                         // [n_reads, 4]: tel_len, edge_nontel, interstitial a, b (-1 = None)
};

__global__ void __launch_bounds__(128) bd_terminating_kernel(TermParams P)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_reads) return;
    const int R = P.reg_count[r];
    int32_t *out = P.out + 4ll * r;
    if (R <= 0) { out[0] = out[1] = 0; out[2] = out[3] = -1; return; }
    const int32_t *s = P.reg_start + P.reg_off[r];
    const uint8_t *c = P.reg_cls + P.reg_off[r];
    const int last_end = P.last_end[r];
    const int adj = P.window / 2;
    auto end_of = [&](int i) { return i + 1 < R ? s[i + 1] : last_end; };
    long long best = 0, run = 0;
    int bi = 0, edge = 0, tel = 0, ilo = 0, ihi = R;
    bool has = false;
    if (P.pq_is_p) {
        if (c[0] == 0) edge = end_of(0) - s[0];
        for (int i = 0; i < R; i++) {                     // forward cumulative score, first maximum (tg_util.posmax)
            const int sz = end_of(i) - s[i];
            run += c[i] ? sz : -sz;
            if (i == 0 || run > best) { best = run; bi = i; }
        }
        if ((double)best >= P.min_tel_score) { has = true; tel = end_of(bi) + adj; ilo = bi; ihi = R; }
    } else {
        if (c[R - 1] == 0) edge = last_end - s[R - 1];
        for (int i = R - 1; i >= 0; i--) {                // reverse cumulative score, first maximum
            const int sz = end_of(i) - s[i];
            run += c[i] ? sz : -sz;
            if (i == R - 1 || run >= best) { best = run; bi = i; }
        }
        if ((double)best >= P.min_tel_score) { has = true; tel = last_end - s[bi] + adj; ilo = 0; ihi = bi; }
    }
    // telomeric regions left over, clustered while the larger neighbour is wider than the gap; widest cluster wins, ties to the later
    int prev = -1, a = 0, best_size = -1, best_a = -1, best_b = -1;
    auto finalize = [&](int lo, int hi) {
        const int size = hi - lo;
        if (size >= best_size) { best_size = size; best_a = lo + adj; best_b = hi + adj; }
    };
    for (int i = ilo; i < ihi; i++) {
        if (c[i] == 0) continue;
        if (prev < 0) a = s[i];
        else {
            const int my = end_of(prev) - s[prev], nx = end_of(i) - s[i], dist = s[i] - end_of(prev);
            if (!(max(my, nx) > dist)) { finalize(a, end_of(prev)); a = s[i]; }
        }
        prev = i;
    }
    if (prev >= 0) finalize(a, end_of(prev));
    if (!has || tel <= 0) { out[0] = 0; out[1] = 0; }
    else { out[0] = tel; out[1] = edge; }
    out[2] = best_a; out[3] = best_b;
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
// Blocks per read for a pass over `avg` elements per read (the longest has `mx`): a few tiles per block, so that reads much
// shorter than the longest one do not leave most of the grid idle; every kernel walks its read's tiles with a grid stride.
static int bd_blocks_x(long long avg, long long mx)
{
    long long b = (avg + 2 * BD_TILE - 1) / (2 * BD_TILE);
    const long long cap = (mx + BD_TILE - 1) / BD_TILE;
    if (b > cap) b = cap;
    if (b > 64) b = 64;
    return b < 1 ? 1 : (int)b;
}

TG_EXPORT int tg_bound_power(const uint8_t *d_cnt_p, const uint8_t *d_cnt_q, const int64_t *d_base_off, const int32_t *d_sig_n,
                             const int64_t *d_x_off, int n_reads, int max_n, int64_t total_n, int window, double *d_X, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_reads < 0 || window < 1 || window > 255) return tg_set_err(TG_EINVAL, "tg_bound_power: bad arguments (window 1..255)");
    if (n_reads == 0 || max_n <= 0) return TG_OK;
    const long long avg_n = total_n > 0 ? total_n / n_reads : max_n;
    BdParams P = {};
    P.X = d_X; P.x_off = d_x_off; P.sig_n = d_sig_n; P.n_reads = n_reads; P.window = window;
    for (int r0 = 0; r0 < n_reads; r0 += 65535) {
        dim3 grid(bd_blocks_x(avg_n, max_n), min(65535, n_reads - r0));
        bd_power_kernel<<<grid, BD_TILE, 0, stream>>>(d_cnt_p, d_cnt_q, d_base_off, P, r0);
    }
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_bound_smooth(double *d_X, double *d_A, double *d_D, const int64_t *d_x_off, const int64_t *d_a_off,
                              const int64_t *d_d_off, const int32_t *d_sig_n, int n_reads, int max_n, int64_t total_n, int window,
                              int smoothing_level, double fail_sigma, const double *d_log_factor, double *d_scratch,
                              const int64_t *d_sc_off, double *d_uthresh, double *d_sigma, int32_t *d_out_len,
                              uint8_t *d_smoothed, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_reads < 0 || window < 1 || smoothing_level < 1) return tg_set_err(TG_EINVAL, "tg_bound_smooth: bad arguments");
    if (n_reads == 0) return TG_OK;
    BdParams P = {};
    P.X = d_X; P.A = d_A; P.D = d_D; P.x_off = d_x_off; P.a_off = d_a_off; P.d_off = d_d_off; P.sig_n = d_sig_n;
    P.n_reads = n_reads; P.window = window; P.smoothing_level = smoothing_level; P.fail_sigma = fail_sigma;
    P.log_factor = d_log_factor; P.scratch = d_scratch; P.sc_off = d_sc_off; P.uthresh = d_uthresh; P.sigma = d_sigma;
    P.out_len = d_out_len; P.smoothed = d_smoothed;
    int max_lev = 0;
    if (max_n >= 7) { int q = max_n / 7; while (q > 1) { q >>= 1; max_lev++; } }
    const bool any = max_n >= window && max_lev + 1 >= smoothing_level;
    const long long avg_n = total_n > 0 ? total_n / n_reads : max_n;
    for (int r0 = 0; r0 < n_reads; r0 += 65535) {
        const int nr = min(65535, n_reads - r0);
        if (any)
            for (int lev = 1; lev <= max_lev; lev++) {
                dim3 grid(bd_blocks_x(avg_n >> lev, ((long long)max_n >> lev) + 1), nr);
                bd_dwt_kernel<<<grid, BD_TILE, 0, stream>>>(P, lev, r0);
            }
        bd_sigma_kernel<<<nr, 256, 0, stream>>>(P, r0);
        if (any) {
            dim3 tgrid(bd_blocks_x(avg_n, (long long)max_n + max_lev), nr);
            bd_threshold_kernel<<<tgrid, BD_TILE, 0, stream>>>(P, r0);
            for (int lev = max_lev; lev >= 1; lev--) {
                dim3 grid(bd_blocks_x(avg_n >> (lev - 1), ((long long)max_n >> (lev - 1)) + 2), nr);
                bd_idwt_kernel<<<grid, BD_TILE, 0, stream>>>(P, lev, r0);
            }
        }
    }
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_bound_regions(const double *d_X, const int64_t *d_x_off, const int32_t *d_sig_n, const int32_t *d_out_len,
                               int n_reads, int window, double pthresh, const int64_t *d_reg_off, int32_t *d_reg_count,
                               int32_t *d_reg_start, uint8_t *d_reg_cls, int32_t *d_last_end, uint8_t *d_err,
                               void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_reads < 0 || window < 1) return tg_set_err(TG_EINVAL, "tg_bound_regions: bad arguments");
    if (n_reads == 0) return TG_OK;
    RegParams P = {};
    P.X = d_X; P.x_off = d_x_off; P.sig_n = d_sig_n; P.out_len = d_out_len; P.n_reads = n_reads; P.window = window;
    P.pthresh = pthresh; P.reg_off = d_reg_off; P.reg_count = d_reg_count; P.reg_start = d_reg_start; P.reg_cls = d_reg_cls;
    P.last_end = d_last_end; P.err = d_err;
    bd_regions_kernel<<<n_reads, 256, 0, stream>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_bound_terminating(const int64_t *d_reg_off, const int32_t *d_reg_count, const int32_t *d_reg_start,
                                   const uint8_t *d_reg_cls, const int32_t *d_last_end, int n_reads, int pq_is_p, int window,
                                   double min_tel_score, int32_t *d_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_reads < 0 || window < 1) return tg_set_err(TG_EINVAL, "tg_bound_terminating: bad arguments");
    if (n_reads == 0) return TG_OK;
    TermParams P = {};
    P.reg_off = d_reg_off; P.reg_count = d_reg_count; P.reg_start = d_reg_start; P.reg_cls = d_reg_cls; P.last_end = d_last_end;
    P.n_reads = n_reads; P.pq_is_p = pq_is_p; P.window = window; P.min_tel_score = min_tel_score; P.out = d_out;
    bd_terminating_kernel<<<(n_reads + 127) / 128, 128, 0, stream>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}
