// Colour-vector preparation kernels: the per-read O(L) Python loops that cluster_tvrs and
// quick_get_tvrtel_lens run before every distance matrix (reference source/tg_tvr.py:96-142, 449-498, 524-556, 558-592).
//   cv_paint_kernel      spans -> letter string, k-mers in index order (a later k-mer overwrites)       tg_tvr.py:96-105
//   cv_boundary_kernel   windowed letter density: first threshold crossing + extreme position          tg_tvr.py:449-479
//   cv_lead_run_kernel   length of the leading run of one letter (the "adj" loop, the trailing run)    tg_tvr.py:124-126
//   cv_gather_kernel     segments -> contiguous strings (the string concatenations between the steps)
//   cv_denoise_kernel    denoise_colorvec as a local stencil                                            tg_tvr.py:524-556
// Byte work on a few tens of MB: the kernels are simple, coalesced and parallel over reads / tiles.
#include "tg_common.cuh"

namespace {

struct LetterSet { uint32_t w[8]; };
__device__ __forceinline__ bool in_set(const LetterSet &s, uint32_t c) { return (s.w[c >> 5] >> (c & 31u)) & 1u; }

// ------------------------------------------------------------------------------------------------
// paint: CTA per read.  Spans are sorted by (read, k); rk_off[read * K + k] indexes them.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cv_paint_kernel(const int32_t *span_start, const int32_t *span_end, const long long *rk_off,
                                                       const uint8_t *letters, int K, const long long *cv_off, int n_reads,
                                                       uint8_t unknown, uint8_t *cv)
{
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const long long o = cv_off[r];
        const int L = (int)(cv_off[r + 1] - o);
        for (int x = threadIdx.x; x < L; x += blockDim.x) cv[o + x] = unknown;
        __syncthreads();
        for (int k = 0; k < K; k++) {
            const long long s0 = rk_off[(long long)r * K + k], s1 = rk_off[(long long)r * K + k + 1];
            if (s0 == s1) continue;                                 // uniform over the CTA
            const uint8_t c = letters[k];
            for (long long s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
                const int a = max(span_start[s], 0), b = min(span_end[s], L);
                for (int x = a; x < b; x++) cv[o + x] = c;
            }
            __syncthreads();                                        // the next k-mer overwrites this one
        }
    }
}

// ------------------------------------------------------------------------------------------------
// density boundary: warp per string.  Virtual string x -> buf[off + x] (or buf[off + len - 1 - x] when reversed).
//   c[j] = #(letters in set among positions j+1 .. j+win, positions < start never count), start <= j < len - win
//   first   = first j with c[j] / win <= thresh (below) or >= thresh (above), float64 division like the reference
//   extreme = first j attaining the global minimum if it is < win (below: the strict running minimum starts at 1.0),
//             the global maximum if it is > 0 (above: starts at 0.0)
// -1 stands for None.
// ------------------------------------------------------------------------------------------------
struct BoundaryParams {
    const uint8_t *buf;
    const long long *off;
    const int32_t *len;
    const uint8_t *reverse;
    int n, win, above, start;
    double thresh;
    LetterSet set;
    int32_t *first, *extreme;
};

__global__ void __launch_bounds__(256) cv_boundary_kernel(const BoundaryParams P)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int sidx = wid; sidx < P.n; sidx += nw) {
        const int L = P.len[sidx];
        const bool rev = P.reverse && P.reverse[sidx];
        const uint8_t *b = P.buf + P.off[sidx];
        auto U = [&](int i) -> int {                                 // indicator of virtual position i
            if (i < P.start || i >= L) return 0;
            return in_set(P.set, rev ? b[L - 1 - i] : b[i]) ? 1 : 0;
        };
        int first = -1, ext = -1;
        const int jend = L - P.win;                                  // exclusive
        if (jend > P.start) {
            // c[start]
            int c0 = 0;
            for (int i = P.start + 1 + lane; i <= P.start + P.win; i += 32) c0 += U(i);
            for (int o = 16; o > 0; o >>= 1) c0 += __shfl_xor_sync(TG_FULL_MASK, c0, o);
            int best = P.above ? 0 : P.win;                          // strict improvement needed
            for (int j0 = P.start; j0 < jend; j0 += 32) {
                // c[j0 + l] = c[j0] + sum_{t < l} (u[j0 + t + 1 + win] - u[j0 + t + 1])
                const int e = U(j0 + lane + 1 + P.win) - U(j0 + lane + 1);
                int inc = e;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(TG_FULL_MASK, inc, o);
                    if (lane >= o) inc += v;
                }
                const int c = c0 + inc - e;                          // exclusive prefix
                const int j = j0 + lane;
                const bool valid = j < jend;
                if (first < 0) {
                    const double dens = (double)c / (double)P.win;
                    const bool hit = valid && (P.above ? dens >= P.thresh : dens <= P.thresh);
                    const unsigned m = __ballot_sync(TG_FULL_MASK, hit);
                    if (m) first = j0 + __ffs(m) - 1;
                }
                const bool better = valid && (P.above ? c > best : c < best);
                if (__any_sync(TG_FULL_MASK, better)) {
                    // best value of the chunk, lowest position on ties
                    int key = better ? (P.above ? -c : c) : INT_MAX;
                    int pos = j;
                    for (int o = 16; o > 0; o >>= 1) {
                        const int k2 = __shfl_xor_sync(TG_FULL_MASK, key, o), p2 = __shfl_xor_sync(TG_FULL_MASK, pos, o);
                        if (k2 < key || (k2 == key && p2 < pos)) { key = k2; pos = p2; }
                    }
                    best = P.above ? -key : key;
                    ext = pos;
                }
                c0 += __shfl_sync(TG_FULL_MASK, inc, 31);
            }
        }
        if (lane == 0) { P.first[sidx] = first; P.extreme[sidx] = ext; }
    }
}

// ------------------------------------------------------------------------------------------------
// leading run: warp per string.  out = min(#leading characters equal to `letter`, cap[s]) of the (virtual) string.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cv_lead_run_kernel(const uint8_t *buf, const long long *off, const int32_t *len,
                                                          const uint8_t *reverse, const int32_t *cap, int n, uint8_t letter, int32_t *out)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int s = wid; s < n; s += nw) {
        const int L = len[s];
        const bool rev = reverse && reverse[s];
        const uint8_t *b = buf + off[s];
        int run = L;
        for (int x0 = 0; x0 < L; x0 += 32) {
            const int x = x0 + lane;
            const bool diff = x < L && (rev ? b[L - 1 - x] : b[x]) != letter;
            const unsigned m = __ballot_sync(TG_FULL_MASK, diff);
            if (m) { run = x0 + __ffs(m) - 1; break; }
        }
        if (lane == 0) out[s] = min(run, cap[s]);
    }
}

// ------------------------------------------------------------------------------------------------
// gather: warp per segment, dst[dst_off + x] = src[src_off + x]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cv_gather_kernel(const uint8_t *src, const long long *src_off, const long long *dst_off,
                                                        const int32_t *len, long long n_seg, uint8_t *dst)
{
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long s = wid; s < n_seg; s += nw) {
        const uint8_t *a = src + src_off[s];
        uint8_t *d = dst + dst_off[s];
        const int L = len[s];
        for (int x = lane; x < L; x += 32) d[x] = a[x];
    }
}

// ------------------------------------------------------------------------------------------------
// denoise_colorvec as a stencil.  w = v + replace_char; runs of equal letters of w; the last run (the trailing
// replace_char letters, positions >= keep[s]) is never a block.  A block survives unless it is shorter than min_size
// and its letter is deletable.  Output position x (0 .. len):
//   x in a surviving block                              -> its letter
//   else, nearest surviving positions yl < x < yr exist, same mergeable letter, yr - yl - 1 <= max_gap_fill -> that letter
//   else                                                -> replace_char
// Tiles of DN_TILE positions with a halo of DN_HALO characters on both sides in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int DN_TILE = 1024, DN_HALO = 128, DN_THREADS = 256;

struct DenoiseParams {
    const uint8_t *buf;
    const long long *off;        // input strings
    const int32_t *len;
    const int32_t *keep;         // len - trailing run of replace_char: positions >= keep belong to the dropped last run
    const long long *out_off;    // output strings, len + 1 each
    const long long *tile_off;   // [n + 1] exclusive prefix sum of ceil((len + 1) / DN_TILE)
    int n;
    uint8_t replace;
    int min_size, max_gap_fill;
    LetterSet del, merge;
    uint8_t *out;
};

__global__ void __launch_bounds__(DN_THREADS) cv_denoise_kernel(const DenoiseParams P)
{
    __shared__ uint8_t s_c[DN_TILE + 2 * DN_HALO];     // characters, 0xff outside the string / in the dropped last run
    __shared__ uint8_t s_s[DN_TILE + 2 * DN_HALO];     // 1 = position of a surviving block
    const long long n_tiles = P.tile_off[P.n];
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // string of this tile
        int lo = 0, hi = P.n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (P.tile_off[mid] <= tile) lo = mid; else hi = mid;
        }
        const int s = lo;
        const int L = P.len[s], keep = P.keep[s];
        const uint8_t *v = P.buf + P.off[s];
        const int x0 = (int)(tile - P.tile_off[s]) * DN_TILE;          // first output position of the tile
        __syncthreads();
        for (int i = threadIdx.x; i < DN_TILE + 2 * DN_HALO; i += DN_THREADS) {
            const int y = x0 - DN_HALO + i;
            s_c[i] = (y >= 0 && y < keep) ? v[y] : (uint8_t)0xff;
        }
        __syncthreads();
        // pass 1: survival.  Only the halo part that pass 2 can look at needs it: max_gap_fill each side.
        for (int i = threadIdx.x; i < DN_TILE + 2 * DN_HALO; i += DN_THREADS) {
            const uint8_t c = s_c[i];
            uint8_t sv = 0;
            if (c != 0xff && i >= DN_HALO - P.max_gap_fill - 1 && i < DN_TILE + DN_HALO + P.max_gap_fill + 1) {
                sv = 1;
                if (in_set(P.del, c)) {
                    int size = 1;
                    for (int k = i - 1; size < P.min_size && k >= 0 && s_c[k] == c; k--) size++;
                    for (int k = i + 1; size < P.min_size && k < DN_TILE + 2 * DN_HALO && s_c[k] == c; k++) size++;
                    sv = size >= P.min_size;
                }
            }
            s_s[i] = sv;
        }
        __syncthreads();
        // pass 2: outputs
        uint8_t *out = P.out + P.out_off[s];
        for (int t = threadIdx.x; t < DN_TILE; t += DN_THREADS) {
            const int x = x0 + t;
            if (x > L) break;
            const int i = t + DN_HALO;
            uint8_t o = P.replace;
            if (s_s[i]) o = s_c[i];
            else if (x < keep) {
                int yl = -1, yr = -1;
                for (int k = 1; k <= P.max_gap_fill; k++) if (s_s[i - k]) { yl = i - k; break; }
                if (yl >= 0) {
                    for (int k = 1; k <= P.max_gap_fill; k++) if (s_s[i + k]) { yr = i + k; break; }
                    if (yr >= 0 && s_c[yl] == s_c[yr] && in_set(P.merge, s_c[yl]) && yr - yl - 1 <= P.max_gap_fill) o = s_c[yl];
                }
            }
            out[x] = o;
        }
    }
}

LetterSet make_set(const uint8_t *letters, int n)
{
    LetterSet s;
    for (int i = 0; i < 8; i++) s.w[i] = 0;
    for (int i = 0; i < n; i++) s.w[letters[i] >> 5] |= 1u << (letters[i] & 31);
    return s;
}

int grid_for(long long items_per_block_units, int sms, int per_sm)
{
    long long g = items_per_block_units;
    const long long cap = (long long)sms * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

TG_EXPORT int tg_cv_paint(const int32_t *d_span_start, const int32_t *d_span_end, const int64_t *d_rk_off,
                          const uint8_t *d_letters, int K, const int64_t *d_cv_off, int n_reads, int unknown,
                          uint8_t *d_cv, void *stream)
{
    if (n_reads <= 0) return TG_OK;
    if (!d_rk_off || !d_letters || !d_cv_off || !d_cv || K < 0) return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    cv_paint_kernel<<<grid_for(n_reads, sms, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_span_start, d_span_end, reinterpret_cast<const long long *>(d_rk_off), d_letters, K,
        reinterpret_cast<const long long *>(d_cv_off), n_reads, (uint8_t)unknown, d_cv);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_cv_density_boundary(const uint8_t *d_buf, const int64_t *d_off, const int32_t *d_len, const uint8_t *d_reverse,
                                     int n, const uint8_t *h_letters, int n_letters, int win, double thresh, int above, int start,
                                     int32_t *d_first, int32_t *d_extreme, void *stream)
{
    if (n <= 0) return TG_OK;
    if (!d_buf || !d_off || !d_len || !d_first || !d_extreme || (n_letters > 0 && !h_letters)) return tg_set_err(TG_EINVAL, "NULL pointer");
    if (win < 1 || start < 0) return tg_set_err(TG_EINVAL, "win %d / start %d out of range", win, start);
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    BoundaryParams P;
    P.buf = d_buf; P.off = reinterpret_cast<const long long *>(d_off); P.len = d_len; P.reverse = d_reverse;
    P.n = n; P.win = win; P.above = above ? 1 : 0; P.start = start; P.thresh = thresh;
    P.set = make_set(h_letters, n_letters); P.first = d_first; P.extreme = d_extreme;
    cv_boundary_kernel<<<grid_for((n + 7) / 8, sms, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_cv_lead_run(const uint8_t *d_buf, const int64_t *d_off, const int32_t *d_len, const uint8_t *d_reverse,
                             const int32_t *d_cap, int n, int letter, int32_t *d_out, void *stream)
{
    if (n <= 0) return TG_OK;
    if (!d_buf || !d_off || !d_len || !d_cap || !d_out) return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    cv_lead_run_kernel<<<grid_for((n + 7) / 8, sms, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_buf, reinterpret_cast<const long long *>(d_off), d_len, d_reverse, d_cap, n, (uint8_t)letter, d_out);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_cv_gather(const uint8_t *d_src, const int64_t *d_src_off, const int64_t *d_dst_off, const int32_t *d_len,
                           int64_t n_seg, uint8_t *d_dst, void *stream)
{
    if (n_seg <= 0) return TG_OK;
    if (!d_src || !d_src_off || !d_dst_off || !d_len || !d_dst) return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    cv_gather_kernel<<<grid_for((n_seg + 7) / 8, sms, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_src, reinterpret_cast<const long long *>(d_src_off), reinterpret_cast<const long long *>(d_dst_off), d_len, n_seg, d_dst);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_cv_denoise_tile(void) { return DN_TILE; }

TG_EXPORT int tg_cv_denoise(const uint8_t *d_buf, const int64_t *d_off, const int32_t *d_len, const int32_t *d_keep,
                            const int64_t *d_out_off, const int64_t *d_tile_off, int n, int64_t n_tiles, int replace_char,
                            int min_size, int max_gap_fill, const uint8_t *h_delete, int n_delete,
                            const uint8_t *h_merge, int n_merge, uint8_t *d_out, void *stream)
{
    if (n <= 0 || n_tiles <= 0) return TG_OK;
    if (!d_buf || !d_off || !d_len || !d_keep || !d_out_off || !d_tile_off || !d_out) return tg_set_err(TG_EINVAL, "NULL device pointer");
    if (min_size < 0 || min_size > DN_HALO / 2 || max_gap_fill < 0 || max_gap_fill > DN_HALO / 2 - 1)
        return tg_set_err(TG_ERANGE, "min_size %d / max_gap_fill %d beyond the stencil halo (%d / %d)", min_size, max_gap_fill,
                          DN_HALO / 2, DN_HALO / 2 - 1);
    if (replace_char == 0xff) return tg_set_err(TG_EINVAL, "replace_char 0xff is reserved");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    DenoiseParams P;
    P.buf = d_buf; P.off = reinterpret_cast<const long long *>(d_off); P.len = d_len; P.keep = d_keep;
    P.out_off = reinterpret_cast<const long long *>(d_out_off); P.tile_off = reinterpret_cast<const long long *>(d_tile_off);
    P.n = n; P.replace = (uint8_t)replace_char; P.min_size = min_size; P.max_gap_fill = max_gap_fill;
    P.del = make_set(h_delete, n_delete); P.merge = make_set(h_merge, n_merge); P.out = d_out;
    cv_denoise_kernel<<<grid_for(n_tiles, sms, 8), DN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}
