// Hierarchical clustering of the distance matrix on the device (second half: "scipy linkage itself").
//
// The reference hands the condensed matrix to scipy: linkage(squareform(dist_matrix), method='ward') (tg_tvr.py:166-167) and
// linkage(d_arr, method=linkage_method) with 'single' / 'ward' / 'complete' (tg_tvr.py:397-398).  Once the matrix comes from
// 8 GPUs the O(N^2) nearest-neighbour-chain loop of scipy is as long as the matrix itself (N=5793: 0.43 s against 0.52 s).
//
// Same algorithms as scipy/cluster/_hierarchy.pyx so that the merge list is bit-identical: `nn_chain` (ward, complete, average,
// weighted) and `mst_single_linkage` (single); the stable sort of the merges by distance and the union-find relabelling run
// on the host (telogator2_b200/tg_linkage.py).  One CTA owns the whole run: the algorithm is a chain of ~4 N dependent steps,
// each a reduction or an update over one row of the matrix; 1024 threads walk the row, the steps are separated by CTA
// barriers.  The matrix is kept square (float64, N^2) so that every step reads and writes whole rows; only the mirror of the
// updated row is a strided write.  All arithmetic of the Lance-Williams updates uses explicit round-to-nearest operations in
// scipy's evaluation order (no FMA contraction).
#include "tg_common.cuh"
#include <limits.h>

namespace {

constexpr int LK_THREADS = 1024;
enum { LK_WARD = 0, LK_COMPLETE = 1, LK_AVERAGE = 2, LK_WEIGHTED = 3 };

struct LkParams {
    double *D;          // [n, n] symmetric, diagonal unused
    int n, method;
    int *size;          // [n] cluster sizes (0 = merged away); MST: merged flags
    int *chain;         // [n]
    double *Z;          // [n - 1, 4]: x, y, distance, size -- in merge order, unlabelled
    double *near;       // [n] MST: distance of every point to the tree
};

// condensed upper-triangle array -> square matrix
__global__ void __launch_bounds__(256) lk_expand_kernel(const double *__restrict__ y, double *__restrict__ D, int n)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * n) return;
    const long long i = idx / n, j = idx % n;
    if (i == j) { D[idx] = 0.0; return; }
    const long long a = i < j ? i : j, b = i < j ? j : i;
    D[idx] = y[(long long)n * a - a * (a + 1) / 2 + (b - a - 1)];
}

// float32 square matrix (optionally divided by `div` in float32, tg_tvr.py:154) -> float64 square matrix
__global__ void __launch_bounds__(256) lk_widen_kernel(const float *__restrict__ M, double *__restrict__ D, long long total, float div)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const float v = div != 0.0f ? __fdiv_rn(M[idx], div) : M[idx];
    D[idx] = (double)v;
}

__device__ __forceinline__ double lk_new_dist(int method, double d_xi, double d_yi, double d_xy, int nx, int ny, int ni)
{
    switch (method) {
    case LK_WARD: {
        const double t = __ddiv_rn(1.0, (double)(nx + ny + ni));
        const double a = __dmul_rn(__dmul_rn(__dmul_rn((double)(ni + nx), t), d_xi), d_xi);
        const double b = __dmul_rn(__dmul_rn(__dmul_rn((double)(ni + ny), t), d_yi), d_yi);
        const double c = __dmul_rn(__dmul_rn(__dmul_rn((double)ni, t), d_xy), d_xy);
        return __dsqrt_rn(__dsub_rn(__dadd_rn(a, b), c));
    }
    case LK_COMPLETE:
        return d_xi > d_yi ? d_xi : d_yi;
    case LK_AVERAGE:
        return __ddiv_rn(__dadd_rn(__dmul_rn((double)nx, d_xi), __dmul_rn((double)ny, d_yi)), (double)(nx + ny));
    default:
        return __dmul_rn(0.5, __dadd_rn(d_xi, d_yi));
    }
}

// (value, index) minimum over the CTA, smallest index among equal values; result in s_val[0] / s_idx[0] after the call
__device__ __forceinline__ void lk_block_argmin(double bv, int bi, double *s_val, int *s_idx)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(TG_FULL_MASK, bv, o);
        const int oi = __shfl_down_sync(TG_FULL_MASK, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = s_val[lane];
        bi = s_idx[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(TG_FULL_MASK, bv, o);
            const int oi = __shfl_down_sync(TG_FULL_MASK, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[0] = bv; s_idx[0] = bi; }
    }
    __syncthreads();
}

// scipy _hierarchy.nn_chain
__global__ void __launch_bounds__(LK_THREADS) lk_nnchain_kernel(const LkParams P)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_y, s_stop, s_first;
    __shared__ double s_cur;
    const int n = P.n, tid = threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int chain_len = 0, first_alive = 0;
    for (int k = 0; k < n - 1; k++) {
        if (chain_len == 0) {
            if (tid == 0) {
                int p = first_alive;                       // sizes never come back from 0: the first live cluster only moves right
                while (__ldcg(P.size + p) == 0) p++;
                P.chain[0] = p;
                s_first = p;
            }
            __syncthreads();
            first_alive = s_first;
            chain_len = 1;
        }
        int x, y;
        double cur;
        for (;;) {
            x = __ldcg(P.chain + chain_len - 1);
            int y0 = -1;
            cur = inf;
            if (chain_len > 1) {                           // prefer the previous element of the chain on ties
                y0 = __ldcg(P.chain + chain_len - 2);
                cur = __ldcg(P.D + (size_t)x * n + y0);
            }
            const double *row = P.D + (size_t)x * n;
            double bv = inf;
            int bi = INT_MAX;
            for (int i = tid; i < n; i += LK_THREADS) {
                if (i == x || __ldcg(P.size + i) == 0) continue;
                const double d = __ldcg(row + i);
                if (d < bv) { bv = d; bi = i; }
            }
            lk_block_argmin(bv, bi, s_val, s_idx);
            if (tid == 0) {
                int yy = y0;
                double cc = cur;
                if (s_val[0] < cc) { cc = s_val[0]; yy = s_idx[0]; }
                const int stop = chain_len > 1 && yy == y0;
                if (!stop) P.chain[chain_len] = yy;
                s_y = yy; s_cur = cc; s_stop = stop;
            }
            __syncthreads();
            y = s_y;
            cur = s_cur;
            const int stop = s_stop;
            __syncthreads();
            if (stop) break;
            chain_len++;
        }
        chain_len -= 2;
        if (x > y) { const int t = x; x = y; y = t; }
        const int nx = __ldcg(P.size + x), ny = __ldcg(P.size + y);
        __syncthreads();
        if (tid == 0) {
            double *z = P.Z + 4ll * k;
            z[0] = (double)x; z[1] = (double)y; z[2] = cur; z[3] = (double)(nx + ny);
            P.size[x] = 0;
            P.size[y] = nx + ny;
        }
        __syncthreads();
        const double *rx = P.D + (size_t)x * n;
        double *ry = P.D + (size_t)y * n;
        for (int i = tid; i < n; i += LK_THREADS) {
            const int ni = __ldcg(P.size + i);
            if (ni == 0 || i == y) continue;
            const double nd = lk_new_dist(P.method, __ldcg(rx + i), __ldcg(ry + i), cur, nx, ny, ni);
            ry[i] = nd;
            P.D[(size_t)i * n + y] = nd;
        }
        __syncthreads();
    }
}

// scipy _hierarchy.mst_single_linkage
__global__ void __launch_bounds__(LK_THREADS) lk_mst_kernel(const LkParams P)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    const int n = P.n, tid = threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = tid; i < n; i += LK_THREADS) { P.near[i] = inf; P.size[i] = 0; }
    __syncthreads();
    int x = 0;
    for (int k = 0; k < n - 1; k++) {
        if (tid == 0) P.size[x] = 1;                       // merged[x]
        __syncthreads();
        const double *row = P.D + (size_t)x * n;
        double bv = inf;
        int bi = INT_MAX;
        for (int i = tid; i < n; i += LK_THREADS) {
            if (__ldcg(P.size + i) == 1) continue;
            const double dist = __ldcg(row + i);
            double di = P.near[i];
            if (di > dist) { di = dist; P.near[i] = di; }
            if (di < bv) { bv = di; bi = i; }
        }
        lk_block_argmin(bv, bi, s_val, s_idx);
        // (scipy keeps the previous y when nothing is below +inf; with finite distances there always is)
        const int y = s_idx[0] == INT_MAX ? x : s_idx[0];
        const double cur = s_val[0];
        __syncthreads();
        if (tid == 0) {
            double *z = P.Z + 4ll * k;
            z[0] = (double)x; z[1] = (double)y; z[2] = cur; z[3] = 0.0;
        }
        x = y;
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Large N: the same two algorithms with the row walks spread over several CTAs (one per SM at most, launched cooperatively so
// that all of them are resident).  Every CTA walks its slice of the row, publishes its (value, index) minimum, all CTAs meet
// at a grid barrier and then reduce the partial minima redundantly, so that every CTA takes the same decision and keeps its own
// copy of the chain: one barrier per chain step, one per merge.  Shared data is read with ld.cg (L2), never from L1.
// ---------------------------------------------------------------------------------------------------------------------
struct LkmParams {
    LkParams p;
    double *part_val;       // [2][ctas]
    int *part_idx;          // [2][ctas]
    unsigned *bar;          // [2]: arrival counter, generation
    int *chains;            // [ctas][n]: every CTA's own copy of the chain
    int *sizes;             // [ctas][n]: ... and of the cluster sizes (all CTAs apply the same merges: no cross-CTA ordering needed)
};

__device__ __forceinline__ void lkm_grid_barrier(unsigned *bar, int ctas)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = bar + 1;
        __threadfence();
        const unsigned g = *gen;
        if (atomicAdd(bar, 1u) == (unsigned)ctas - 1u) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// all CTAs: minimum over the published partial minima of this step (smallest index among equal values)
__device__ __forceinline__ void lkm_reduce_partials(const LkmParams &P, int parity, int ctas, double *s_val, int *s_idx)
{
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (threadIdx.x < 32) {
        double bv = inf;
        int bi = INT_MAX;
        for (int c = threadIdx.x; c < ctas; c += 32) {
            const double v = __ldcg(P.part_val + parity * ctas + c);
            const int ix = __ldcg(P.part_idx + parity * ctas + c);
            if (v < bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(TG_FULL_MASK, bv, o);
            const int oi = __shfl_down_sync(TG_FULL_MASK, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) { s_val[0] = bv; s_idx[0] = bi; }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LK_THREADS) lkm_nnchain_kernel(const LkmParams Q)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    const LkParams &P = Q.p;
    const int n = P.n, tid = threadIdx.x, ctas = gridDim.x;
    const int per = ((n + ctas - 1) / ctas + 31) & ~31;
    const int lo = blockIdx.x * per, hi = min(n, lo + per);
    int *chain = Q.chains + (size_t)blockIdx.x * n;
    int *size = Q.sizes + (size_t)blockIdx.x * n;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = tid; i < n; i += LK_THREADS) size[i] = 1;
    __syncthreads();
    int chain_len = 0, first_alive = 0, step = 0;
    for (int k = 0; k < n - 1; k++) {
        if (chain_len == 0) {
            while (size[first_alive] == 0) first_alive++;         // (every thread walks the same few entries)
            if (tid == 0) chain[0] = first_alive;
            chain_len = 1;
            __syncthreads();
        }
        int x, y;
        double cur;
        for (;;) {
            x = chain[chain_len - 1];
            int y0 = -1;
            cur = inf;
            if (chain_len > 1) {
                y0 = chain[chain_len - 2];
                cur = __ldcg(P.D + (size_t)x * n + y0);
            }
            const double *row = P.D + (size_t)x * n;
            double bv = inf;
            int bi = INT_MAX;
            for (int i = lo + tid; i < hi; i += LK_THREADS) {
                if (i == x || size[i] == 0) continue;
                const double d = __ldcg(row + i);
                if (d < bv) { bv = d; bi = i; }
            }
            lk_block_argmin(bv, bi, s_val, s_idx);
            const int parity = step & 1;
            step++;
            if (tid == 0) {
                Q.part_val[parity * ctas + blockIdx.x] = s_val[0];
                Q.part_idx[parity * ctas + blockIdx.x] = s_idx[0];
            }
            lkm_grid_barrier(Q.bar, ctas);
            lkm_reduce_partials(Q, parity, ctas, s_val, s_idx);
            y = y0;
            if (s_val[0] < cur) { cur = s_val[0]; y = s_idx[0]; }
            const bool stop = chain_len > 1 && y == y0;
            __syncthreads();                                   // (s_val / s_idx are reused by the next step)
            if (stop) break;
            if (tid == 0) chain[chain_len] = y;
            chain_len++;
            __syncthreads();
        }
        chain_len -= 2;
        if (x > y) { const int t = x; x = y; y = t; }
        const int nx = size[x], ny = size[y];
        if (blockIdx.x == 0 && tid == 0) {
            double *z = P.Z + 4ll * k;
            z[0] = (double)x; z[1] = (double)y; z[2] = cur; z[3] = (double)(nx + ny);
        }
        const double *rx = P.D + (size_t)x * n;
        double *ry = P.D + (size_t)y * n;
        for (int i = lo + tid; i < hi; i += LK_THREADS) {
            if (i == x || i == y) continue;
            const int ni = size[i];
            if (ni == 0) continue;
            const double nd = lk_new_dist(P.method, __ldcg(rx + i), __ldcg(ry + i), cur, nx, ny, ni);
            ry[i] = nd;
            P.D[(size_t)i * n + y] = nd;
        }
        __syncthreads();                                       // (every thread of the CTA has read the old sizes)
        if (tid == 0) { size[x] = 0; size[y] = nx + ny; }
        lkm_grid_barrier(Q.bar, ctas);
    }
}

__global__ void __launch_bounds__(LK_THREADS) lkm_mst_kernel(const LkmParams Q)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    const LkParams &P = Q.p;
    const int n = P.n, tid = threadIdx.x, ctas = gridDim.x;
    const int per = ((n + ctas - 1) / ctas + 31) & ~31;
    const int lo = blockIdx.x * per, hi = min(n, lo + per);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = lo + tid; i < hi; i += LK_THREADS) { P.near[i] = inf; P.size[i] = 0; }
    __syncthreads();
    int x = 0;
    for (int k = 0; k < n - 1; k++) {
        if (tid == 0 && x >= lo && x < hi) P.size[x] = 1;      // merged[x]: only the owner's walk looks at it
        __syncthreads();
        const double *row = P.D + (size_t)x * n;
        double bv = inf;
        int bi = INT_MAX;
        for (int i = lo + tid; i < hi; i += LK_THREADS) {
            if (P.size[i] == 1) continue;
            const double dist = __ldcg(row + i);
            double di = P.near[i];
            if (di > dist) { di = dist; P.near[i] = di; }
            if (di < bv) { bv = di; bi = i; }
        }
        lk_block_argmin(bv, bi, s_val, s_idx);
        const int parity = k & 1;
        if (tid == 0) {
            Q.part_val[parity * ctas + blockIdx.x] = s_val[0];
            Q.part_idx[parity * ctas + blockIdx.x] = s_idx[0];
        }
        lkm_grid_barrier(Q.bar, ctas);
        lkm_reduce_partials(Q, parity, ctas, s_val, s_idx);
        const int y = s_idx[0] == INT_MAX ? x : s_idx[0];
        const double cur = s_val[0];
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) {
            double *z = P.Z + 4ll * k;
            z[0] = (double)x; z[1] = (double)y; z[2] = cur; z[3] = 0.0;
        }
        x = y;
    }
}

static int lk_method_code(const char *m)
{
    if (!m) return -1;
    if (!strcmp(m, "ward")) return LK_WARD;
    if (!strcmp(m, "complete")) return LK_COMPLETE;
    if (!strcmp(m, "average")) return LK_AVERAGE;
    if (!strcmp(m, "weighted")) return LK_WEIGHTED;
    if (!strcmp(m, "single")) return 100;
    return -1;
}

}  // namespace

TG_EXPORT int tg_linkage_expand(const double *d_condensed, int n, double *d_square, void *stream_)
{
    if (n < 1 || !d_condensed || !d_square) return tg_set_err(TG_EINVAL, "tg_linkage_expand: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long total = (long long)n * n;
    lk_expand_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_condensed, d_square, n);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_linkage_widen(const float *d_matrix, int n, float div, double *d_square, void *stream_)
{
    if (n < 1 || !d_matrix || !d_square) return tg_set_err(TG_EINVAL, "tg_linkage_widen: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long total = (long long)n * n;
    lk_widen_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_matrix, d_square, total, div);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_linkage_run(double *d_square, int n, const char *method, int32_t *d_size, int32_t *d_chain, double *d_near,
                             double *d_Z, void *stream_)
{
    const int code = lk_method_code(method);
    if (code < 0) return tg_set_err(TG_EINVAL, "tg_linkage_run: method '%s' is not supported (ward, single, complete, average, weighted)", method ? method : "(null)");
    if (n < 2 || !d_square || !d_size || !d_chain || !d_near || !d_Z) return tg_set_err(TG_EINVAL, "tg_linkage_run: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    LkParams P;
    P.D = d_square; P.n = n; P.method = code; P.size = d_size; P.chain = d_chain; P.Z = d_Z; P.near = d_near;
    if (code == 100) lk_mst_kernel<<<1, LK_THREADS, 0, stream>>>(P);
    else lk_nnchain_kernel<<<1, LK_THREADS, 0, stream>>>(P);      // d_size must hold ones
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

/* Multi-CTA variant for large n (same results): d_work holds the partial minima, the barrier and every CTA's copy of the chain and of the cluster sizes, laid out by this call. */
TG_EXPORT size_t tg_linkage_multi_work_bytes(int n, int ctas)
{
    if (n < 2 || ctas < 1) return 0;
    return (size_t)2 * ctas * n * sizeof(int32_t) + (size_t)2 * ctas * (sizeof(double) + sizeof(int32_t)) + 64;
}

TG_EXPORT int tg_linkage_run_multi(double *d_square, int n, const char *method, int ctas, int32_t *d_size, double *d_near,
                                   double *d_Z, void *d_work, void *stream_)
{
    const int code = lk_method_code(method);
    if (code < 0) return tg_set_err(TG_EINVAL, "tg_linkage_run_multi: method '%s' is not supported", method ? method : "(null)");
    if (n < 2 || ctas < 1 || !d_square || !d_size || !d_near || !d_Z || !d_work) return tg_set_err(TG_EINVAL, "tg_linkage_run_multi: bad arguments");
    const int sms = tg_sm_count();
    if (ctas > sms) return tg_set_err(TG_EINVAL, "tg_linkage_run_multi: %d CTAs cannot be resident together on %d SMs", ctas, sms);
    cudaStream_t stream = (cudaStream_t)stream_;
    LkmParams Q;
    Q.p.D = d_square; Q.p.n = n; Q.p.method = code; Q.p.size = d_size; Q.p.chain = nullptr; Q.p.Z = d_Z; Q.p.near = d_near;
    uint8_t *w = reinterpret_cast<uint8_t *>(d_work);
    Q.part_val = reinterpret_cast<double *>(w);                       w += (size_t)2 * ctas * sizeof(double);
    Q.part_idx = reinterpret_cast<int *>(w);                          w += (size_t)2 * ctas * sizeof(int32_t);
    Q.bar = reinterpret_cast<unsigned *>(w);                          w += 64;
    Q.chains = reinterpret_cast<int *>(w);                            w += (size_t)ctas * n * sizeof(int32_t);
    Q.sizes = reinterpret_cast<int *>(w);
    TG_CUDA_CHECK(cudaMemsetAsync(Q.bar, 0, 64, stream));
    void *args[] = {&Q};
    if (code == 100) TG_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)lkm_mst_kernel, dim3(ctas), dim3(LK_THREADS), args, 0, stream));
    else TG_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)lkm_nnchain_kernel, dim3(ctas), dim3(LK_THREADS), args, 0, stream));
    return TG_OK;
}
