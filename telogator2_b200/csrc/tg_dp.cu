// Pairwise TVR distance kernels for sm_100a:
//   nw_wave_kernel     u16x2 DPX warp-wavefront Needleman-Wunsch score (two alignments per register)
//   nw_generic_kernel  int32 row-scan Needleman-Wunsch score (any gap scores / lengths)
//   mt_shuffle_kernel  CPython MT19937 + random.sample null-model generator (bit-exact)
//   int_peak_kernel    integer-ALU issue-rate microbenchmark (DP roofline denominator)
//
// Reference semantics: Bio.Align.PairwiseAligner.score() as configured in tg_align.py:71-106 and called at
// tg_align.py:125,142; random.seed / shuffle_seq at tg_align.py:110-111, tg_util.py:451-452.
//
// Score transform used by the wave kernel.  With linear gap g everywhere inside the matrix,
//     H[i][j] = max(H[i-1][j-1] + S(a_i,b_j), H[i-1][j] + g, H[i][j-1] + g)
// substitute H'[i][j] = H[i][j] - g*(i+j):
//     H'[i][j] = max(H'[i-1][j-1] + (S - 2g), H'[i-1][j], H'[i][j-1])
// so a cell costs one add and one 3-input max (VIMNMX3.U16x2 handles two alignments at once),
// all values are non-negative and non-decreasing along rows and columns, and with gap_left == g
// the whole boundary row/column of H' is zero.  Free right end gaps (gap_right >= g) are
// recovered afterwards as max over the last row / last column of H (see the epilogue).
#include "tg_common.cuh"
#include <utility>
#include <vector>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace {

constexpr int WAVE_MAX_WARPS = 16;         // warps (= concurrent jobs) per CTA, one CTA per SM
constexpr int MAXA = 32;

struct WaveParams {
    const uint8_t *pool;
    const tg_nw_job *jobs;
    long long n_jobs;
    int32_t *scores;
    uint32_t *scratch;          // per-warp border column, scratch_stride u32 each
    unsigned long long *counter;
    int scratch_stride;
    int row_cap;                // rows of the shared-memory row-letter buffer per warp
    int A;                      // alphabet size; code A is the all-zero padding letter
    int g, gr;                  // internal gap, right end gap
    int right_free;             // gr != g
    uint8_t sp[MAXA * MAXA];    // S' = S - 2g, row-major [a][b] with stride MAXA
    // implicit job modes (tg_dist_batch): jobs are derived from the pair index instead of read from memory
    const long long *seq_off;   // [n_seq + 1] offsets of the sequences in pool
    const int *order;           // [n_seq] sequence indices sorted by length, longest first
    int n_seq, adjust_lens, min_viable, R;
    long long q0, q1;           // pair range in sorted pair space
    const uint8_t *flags;       // [q1 - q0] 1 = early-out (phase B skips the pair)
    long long shuf_base, slot;  // shuffle area: pool offset and bytes per pair
    const long long *mat_pair_off; const int *mat_seq_off; int n_mat;   // several matrices in one pair space (0 = one matrix)
    // phase B with per-job alphabets: seq_mask[s] = set of letters of sequence s; a job whose row sequences use fewer than
    // prof_rows letters runs in the JOBS_PHASE_B_CPT launch (profiles of prof_rows rows: more resident warps), the others in
    // the JOBS_PHASE_B launch.  prof_rows = 0: every job runs in JOBS_PHASE_B.
    const uint32_t *seq_mask; int prof_rows;
};

// JOBS_PHASE_A_SHCOL: phase A with the two alignments of a job sharing their COLUMN sequence (pairs (a, b), (a, b+1) of the
// sorted pair space share sequence a; NW is symmetric in its two sequences when the substitution matrix is): one query profile
// per warp instead of two, so almost twice as many warps are resident.
enum { JOBS_EXPLICIT = 0, JOBS_PHASE_A = 1, JOBS_PHASE_B = 2, JOBS_PHASE_A_SHCOL = 3, JOBS_PHASE_B_CPT = 4 };

// pair q of itertools.combinations(range(n), 2) -> (a, b), a < b
__device__ __forceinline__ void unrank_pair(long long q, int n, int &a, int &b)
{
    const double t = 2.0 * n - 1.0;
    long long r = (long long)floor((t - sqrt(t * t - 8.0 * (double)q)) * 0.5);
    if (r < 0) r = 0;
    if (r > n - 2) r = n - 2;
    while (r > 0 && r * (2LL * n - r - 1) / 2 > q) r--;
    while ((r + 1) * (2LL * n - (r + 1) - 1) / 2 <= q) r++;
    a = (int)r;
    b = (int)(q - r * (2LL * n - r - 1) / 2 + r + 1);
}

struct PairDesc { long long off_i, off_j; int n, m; long long p; int i, j, mat, n_local, ra, gi; bool a_is_i; };

// sorted-space pair q -> original (i < j), length-adjusted shapes (tg_align.py:115-120) and combinations index p.
// With n_mat > 0 the pair space is the concatenation of the pair spaces of several independent matrices (the per-cluster
// calls of iterations 1/2 and of the sub-telomere stage): matrix b owns pairs [mat_pair_off[b], mat_pair_off[b+1])
// and sequences [mat_seq_off[b], mat_seq_off[b+1]); i, j, p are local to the matrix.
// MM = false compiles the matrix lookup out: the wave kernel's single-matrix instantiations keep their code generation
// (an A/B run showed the extra branch costing the dominant kernel 1 %).
template <bool MM = true, typename PT>
__device__ __forceinline__ PairDesc describe_pair(const PT &P, long long q)
{
    int a, b, nloc = P.n_seq, sbase = 0, mat = 0;
    if (MM && P.n_mat > 0) {
        int lo = 0, hi = P.n_mat;                            // last matrix with mat_pair_off[mat] <= q
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (P.mat_pair_off[mid] <= q) lo = mid; else hi = mid;
        }
        mat = lo;
        q -= P.mat_pair_off[mat];
        sbase = P.mat_seq_off[mat];
        nloc = P.mat_seq_off[mat + 1] - sbase;
    }
    unrank_pair(q, nloc, a, b);
    const int oa = P.order[sbase + a], ob = P.order[sbase + b];
    const int i = min(oa, ob), j = max(oa, ob);
    PairDesc d;
    d.off_i = P.seq_off[i];
    d.off_j = P.seq_off[j];
    int li = (int)(P.seq_off[i + 1] - d.off_i), lj = (int)(P.seq_off[j + 1] - d.off_j);
    if (P.adjust_lens) {
        int ml = min(li, lj);
        if (P.min_viable) ml = max(ml, 1000);
        li = min(li, ml);
        lj = min(lj, ml);
    }
    d.n = li; d.m = lj; d.i = i - sbase; d.j = j - sbase; d.mat = mat; d.n_local = nloc; d.gi = i;
    d.ra = a; d.a_is_i = (oa == i);            // sorted rank of the longer sequence, and whether it is the pair's first sequence
    d.p = (long long)d.i * (2LL * nloc - d.i - 1) / 2 + (d.j - d.i - 1);
    return d;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// byte k of p0 -> low half, byte k of p1 -> high half, upper bytes = replicated sign bit (= 0, S' < 128)
// WIDE (one alignment per job, 32-bit cells: sequences whose scores overflow 16 bits): byte k of p0, zero extended
template <int K, bool WIDE>
__device__ __forceinline__ uint32_t merge_scores(uint32_t p0, uint32_t p1)
{
    if (WIDE) return prmt(p0, 0u, K | 0x4440u);
    constexpr uint32_t sel = K | ((K | 8) << 4) | ((4 + K) << 8) | (((4 + K) | 8) << 12);
    return prmt(p0, p1, sel);
}

template <bool WIDE>
__device__ __forceinline__ uint32_t cell(uint32_t diag, uint32_t sc, uint32_t up, uint32_t left)
{
    return WIDE ? __vimax3_u32(diag + sc, up, left) : __vimax3_u16x2(diag + sc, up, left);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// One step of one lane: C packed cells of row r against the lane's C columns.
template <int C, bool WIDE>
struct Strip;

template <bool WIDE>
struct Strip<16, WIDE> {
    static __device__ __forceinline__ void run(uint32_t (&H)[16], uint32_t d, uint32_t l, uint32_t pa0, uint32_t pa1)
    {
        const uint4 q0 = lds128(pa0), q1 = WIDE ? q0 : lds128(pa1);
        uint32_t up;
#define TG_CELL(c, qa, qb, k) up = H[c]; l = cell<WIDE>(d, merge_scores<k, WIDE>(qa, qb), up, l); d = up; H[c] = l;
        TG_CELL(0, q0.x, q1.x, 0) TG_CELL(1, q0.x, q1.x, 1) TG_CELL(2, q0.x, q1.x, 2) TG_CELL(3, q0.x, q1.x, 3)
        TG_CELL(4, q0.y, q1.y, 0) TG_CELL(5, q0.y, q1.y, 1) TG_CELL(6, q0.y, q1.y, 2) TG_CELL(7, q0.y, q1.y, 3)
        TG_CELL(8, q0.z, q1.z, 0) TG_CELL(9, q0.z, q1.z, 1) TG_CELL(10, q0.z, q1.z, 2) TG_CELL(11, q0.z, q1.z, 3)
        TG_CELL(12, q0.w, q1.w, 0) TG_CELL(13, q0.w, q1.w, 1) TG_CELL(14, q0.w, q1.w, 2) TG_CELL(15, q0.w, q1.w, 3)
    }
};

template <bool WIDE>
struct Strip<8, WIDE> {
    static __device__ __forceinline__ void run(uint32_t (&H)[8], uint32_t d, uint32_t l, uint32_t pa0, uint32_t pa1)
    {
        const uint2 q0 = lds64(pa0), q1 = WIDE ? q0 : lds64(pa1);
        uint32_t up;
        TG_CELL(0, q0.x, q1.x, 0) TG_CELL(1, q0.x, q1.x, 1) TG_CELL(2, q0.x, q1.x, 2) TG_CELL(3, q0.x, q1.x, 3)
        TG_CELL(4, q0.y, q1.y, 0) TG_CELL(5, q0.y, q1.y, 1) TG_CELL(6, q0.y, q1.y, 2) TG_CELL(7, q0.y, q1.y, 3)
#undef TG_CELL
    }
};

// Two rows x 8 columns per step: row r0 (values a_c) and row r0 + 1 (values b_c) advance as two dependency chains
// offset by one cell, which doubles the instruction-level parallelism of the 3-input-max chain.
//   a_c = max(up_{c-1} + s0_c, up_c, a_{c-1})      b_c = max(a_{c-1} + s1_c, a_c, b_{c-1})
template <bool WIDE>
struct Strip2x8 {
    static __device__ __forceinline__ uint32_t run(uint32_t (&H)[8], uint32_t dg, uint32_t L0, uint32_t L1,
                                                   uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11)
    {
        const uint2 q00 = lds64(p00), q01 = WIDE ? q00 : lds64(p01), q10 = lds64(p10), q11 = WIDE ? q10 : lds64(p11);
        uint32_t d0 = dg, la = L0, pa = L0, lb = L1, up, a;
#define TG_CELL2(c, x0, y0, x1, y1, k)                    \
    up = H[c];                                             \
    a = cell<WIDE>(d0, merge_scores<k, WIDE>(x0, y0), up, la);         \
    lb = cell<WIDE>(pa, merge_scores<k, WIDE>(x1, y1), a, lb);         \
    d0 = up; la = a; pa = a; H[c] = lb;
        TG_CELL2(0, q00.x, q01.x, q10.x, q11.x, 0) TG_CELL2(1, q00.x, q01.x, q10.x, q11.x, 1)
        TG_CELL2(2, q00.x, q01.x, q10.x, q11.x, 2) TG_CELL2(3, q00.x, q01.x, q10.x, q11.x, 3)
        TG_CELL2(4, q00.y, q01.y, q10.y, q11.y, 0) TG_CELL2(5, q00.y, q01.y, q10.y, q11.y, 1)
        TG_CELL2(6, q00.y, q01.y, q10.y, q11.y, 2) TG_CELL2(7, q00.y, q01.y, q10.y, q11.y, 3)
#undef TG_CELL2
        return la;      // a_7: last column of row r0
    }
};

template <bool WIDE>
struct Strip2x16 {
    static __device__ __forceinline__ uint32_t run(uint32_t (&H)[16], uint32_t dg, uint32_t L0, uint32_t L1,
                                                   uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11)
    {
        const uint4 q00 = lds128(p00), q01 = WIDE ? q00 : lds128(p01), q10 = lds128(p10), q11 = WIDE ? q10 : lds128(p11);
        uint32_t d0 = dg, la = L0, pa = L0, lb = L1, up, a;
#define TG_CELL2(c, x0, y0, x1, y1, k)                    \
    up = H[c];                                             \
    a = cell<WIDE>(d0, merge_scores<k, WIDE>(x0, y0), up, la);         \
    lb = cell<WIDE>(pa, merge_scores<k, WIDE>(x1, y1), a, lb);         \
    d0 = up; la = a; pa = a; H[c] = lb;
#define TG_QUAD2(c, f)                                                                                  \
    TG_CELL2(c, q00.f, q01.f, q10.f, q11.f, 0) TG_CELL2(c + 1, q00.f, q01.f, q10.f, q11.f, 1)          \
    TG_CELL2(c + 2, q00.f, q01.f, q10.f, q11.f, 2) TG_CELL2(c + 3, q00.f, q01.f, q10.f, q11.f, 3)
        TG_QUAD2(0, x) TG_QUAD2(4, y) TG_QUAD2(8, z) TG_QUAD2(12, w)
#undef TG_QUAD2
#undef TG_CELL2
        return la;      // a_15: last column of row r0
    }
};

// Shared memory per warp: query profile [2][A+1][W] bytes, then row letters [row_cap] u16 (a0 | a1 << 8).
// V = 0: 16 columns x 1 row per step, V = 1: 8 x 1, V = 2: 8 columns x 2 rows, V = 3: 16 columns x 2 rows per step
// WIDE: one alignment per job in 32-bit cells (scores beyond 16 bits: long sequences); profile [1][A+1][W].
template <int V, int MODE, bool WIDE, bool MM>
__global__ void __launch_bounds__(WAVE_MAX_WARPS * 32, 1) nw_wave_kernel(const __grid_constant__ WaveParams P)
{
    constexpr bool SHCOL = (MODE == JOBS_PHASE_A_SHCOL);
    constexpr bool CPT = (MODE == JOBS_PHASE_B_CPT);                         // per-job alphabets: profile rows = letters that occur
    static_assert(!(CPT && WIDE), "compact profiles are a 16-bit path");
    constexpr int NH = (WIDE || SHCOL) ? 1 : 2;                              // query profiles per job
    constexpr int C = (V == 0 || V == 3) ? 16 : 8;
    constexpr int RR = (V >= 2) ? 2 : 1;
    constexpr int W = 32 * C;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int A = P.A, A1 = CPT ? P.prof_rows : P.A + 1;                     // A1 = rows of one profile in shared memory
    uint8_t *s_sp = smem;                                                    // MAXA*MAXA bytes
    constexpr int ROFF = (RR == 2) ? 64 : 0;                                 // rows[ROFF + r] = letters of row r
    const size_t per_warp = (size_t)NH * A1 * W + (size_t)P.row_cap * 2 + 512;
    uint8_t *prof = smem + MAXA * MAXA + (size_t)warp * per_warp;            // [NH][A+1][W]
    uint16_t *rows = reinterpret_cast<uint16_t *>(prof + (size_t)NH * A1 * W);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(prof + (size_t)NH * A1 * W + (size_t)P.row_cap * 2);   // 64 x uint2
    for (int i = threadIdx.x; i < MAXA * MAXA; i += blockDim.x) s_sp[i] = P.sp[i];
    __syncthreads();

    const int nwarps = blockDim.x >> 5;
    uint32_t *bord = P.scratch + (size_t)(blockIdx.x * nwarps + warp) * P.scratch_stride;
    const int g = P.g, gr = P.gr;
    const uint32_t prof_s = (uint32_t)__cvta_generic_to_shared(prof);
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    const uint32_t pb0 = prof_s + lane * C;
    const uint32_t pb1 = (WIDE || SHCOL) ? pb0 : prof_s + (uint32_t)A1 * W + lane * C;

    // SHCOL: a job whose two pairs cannot share their columns runs as two single alignments, the second one parked here
    bool pend = false;
    long long pend_row = 0, pend_col = 0;
    int pend_n = 0, pend_m = 0, pend_out = -1;
    for (;;) {
        unsigned long long jid = 0;
        if (!(SHCOL && pend)) {
            if (lane == 0) jid = atomicAdd(P.counter, 1ULL);
            jid = __shfl_sync(TG_FULL_MASK, jid, 0);
            if (jid >= (unsigned long long)P.n_jobs) break;
        }
        long long row0, row1, col0, col1;
        int n0, n1, m0, m1, out0, out1;
        uint32_t mk0 = 0u, mk1 = 0u;                                         // CPT: letters of the two row sequences
        if (SHCOL) {
            if (pend) {
                row0 = row1 = pend_row; col0 = col1 = pend_col; n0 = n1 = pend_n; m0 = m1 = pend_m; out0 = pend_out; out1 = -1;
                pend = false;
            } else {
                const long long qa = P.q0 + 2 * (long long)jid, qb = qa + 1;
                const PairDesc da = describe_pair<MM>(P, qa);
                // columns = the longer sequence (sorted rank ra, shared by neighbouring pairs), rows = its partner
                row0 = da.a_is_i ? da.off_j : da.off_i; n0 = da.a_is_i ? da.m : da.n;
                col0 = da.a_is_i ? da.off_i : da.off_j; m0 = da.a_is_i ? da.n : da.m;
                out0 = (int)(qa - P.q0);
                row1 = row0; col1 = col0; n1 = n0; m1 = m0; out1 = -1;
                if (qb < P.q1) {
                    const PairDesc db = describe_pair<MM>(P, qb);
                    const long long rowb = db.a_is_i ? db.off_j : db.off_i, colb = db.a_is_i ? db.off_i : db.off_j;
                    const int nb = db.a_is_i ? db.m : db.n, mb = db.a_is_i ? db.n : db.m;
                    if (db.ra == da.ra && db.mat == da.mat && colb == col0 && mb == m0) {
                        row1 = rowb; n1 = nb; out1 = (int)(qb - P.q0);
                    } else {
                        pend = true; pend_row = rowb; pend_col = colb; pend_n = nb; pend_m = mb; pend_out = (int)(qb - P.q0);
                    }
                }
            }
        } else if (MODE == JOBS_EXPLICIT && WIDE) {
            // item = (job, half)
            const tg_nw_job *J = P.jobs + (jid >> 1);
            const int h = (int)(jid & 1);
            if (J->out[h] < 0) continue;
            row0 = row1 = J->row_off[h]; col0 = col1 = J->col_off[h];
            n0 = n1 = J->n[h]; m0 = m1 = J->m[h];
            out0 = J->out[h]; out1 = -1;
        } else if (MODE == JOBS_EXPLICIT) {
            const tg_nw_job *J = P.jobs + jid;
            row0 = J->row_off[0]; row1 = J->row_off[1];
            col0 = J->col_off[0]; col1 = J->col_off[1];
            n0 = J->n[0]; n1 = J->n[1]; m0 = J->m[0]; m1 = J->m[1];
            out0 = J->out[0]; out1 = J->out[1];
        } else if (MODE == JOBS_PHASE_A && WIDE) {
            const long long qa = P.q0 + (long long)jid;
            const PairDesc da = describe_pair<MM>(P, qa);
            row0 = row1 = da.off_i; col0 = col1 = da.off_j; n0 = n1 = da.n; m0 = m1 = da.m;
            out0 = (int)(qa - P.q0); out1 = -1;
        } else if (MODE == JOBS_PHASE_B && WIDE) {
            const long long qa = P.q0 + (long long)(jid / P.R);
            const int ta = (int)(jid % P.R);
            if (P.flags[qa - P.q0]) continue;
            const PairDesc da = describe_pair<MM>(P, qa);
            row0 = row1 = P.shuf_base + (qa - P.q0) * P.slot + (long long)ta * (da.n + da.m); col0 = col1 = row0 + da.n;
            n0 = n1 = da.n; m0 = m1 = da.m;
            out0 = (int)((qa - P.q0) * P.R + ta); out1 = -1;
        } else if (MODE == JOBS_PHASE_A) {
            // job = sorted-space pairs q0 + 2*jid and q0 + 2*jid + 1 (tg_align.py:125)
            const long long qa = P.q0 + 2 * (long long)jid, qb = qa + 1;
            const PairDesc da = describe_pair<MM>(P, qa);
            row0 = da.off_i; col0 = da.off_j; n0 = da.n; m0 = da.m; out0 = (int)(qa - P.q0);
            if (qb < P.q1) {
                const PairDesc db = describe_pair<MM>(P, qb);
                row1 = db.off_i; col1 = db.off_j; n1 = db.n; m1 = db.m; out1 = (int)(qb - P.q0);
            } else { row1 = row0; col1 = col0; n1 = n0; m1 = m0; out1 = -1; }
        } else {
            // phase B (tg_align.py:141-142): R even -> shuffles (2t, 2t+1) of one pair share a job;
            // R odd -> shuffle t of two neighbouring pairs share a job.  Early-out pairs are skipped.
            const int R = P.R;
            long long qa, qb;
            int ta, tb;
            if ((R & 1) == 0) { qa = qb = P.q0 + (long long)(jid / (R / 2)); ta = 2 * (int)(jid % (R / 2)); tb = ta + 1; }
            else { qa = P.q0 + 2 * (long long)(jid / R); qb = qa + 1; ta = tb = (int)(jid % R); }
            const bool ha = !P.flags[qa - P.q0];
            const bool hb = (qb < P.q1) && !P.flags[qb - P.q0];
            if (!ha && !hb) continue;
            PairDesc da, db;
            if (ha) da = describe_pair<MM>(P, qa);
            if (hb) db = (qb == qa) ? da : describe_pair<MM>(P, qb);
            if (!ha) { da = db; qa = qb; ta = tb; }
            if (!hb) { db = da; qb = qa; tb = ta; }
            if (CPT || P.prof_rows > 0) {
                mk0 = P.seq_mask[da.gi]; mk1 = P.seq_mask[db.gi];
                const bool fits = __popc(mk0) < P.prof_rows && __popc(mk1) < P.prof_rows;
                if (fits != CPT) continue;                                   // the other launch takes it
            }
            row0 = P.shuf_base + (qa - P.q0) * P.slot + (long long)ta * (da.n + da.m); col0 = row0 + da.n;
            row1 = P.shuf_base + (qb - P.q0) * P.slot + (long long)tb * (db.n + db.m); col1 = row1 + db.n;
            n0 = da.n; m0 = da.m; n1 = db.n; m1 = db.m;
            out0 = ha ? (int)((qa - P.q0) * R + ta) : -1;
            out1 = hb ? (int)((qb - P.q0) * R + tb) : -1;
            if (ha && !hb) out1 = -1;
            if (!ha && hb) { out0 = out1; out1 = -1; }
        }
        const int N = max(n0, n1), M = max(m0, m1);
        const int nblk = (M + W - 1) / W;
        const int pad0 = nblk * W - m0, pad1 = nblk * W - m1;     // left padding (right-aligned columns)
        int rf0 = INT_MIN, rf1 = INT_MIN;
        uint32_t H[C];

        // row letters of both alignments; rows past the end of a sequence get the padding letter A, whose
        // all-zero profile row makes H' copy the previous row (H' is non-decreasing along a row)
        __syncwarp();
        for (int e = lane; e < N + 4 + 2 * ROFF + (ROFF ? 4 : 0); e += 32) {
            const int r = e - ROFF;
            uint32_t a0 = (r >= 0 && r < n0) ? (uint32_t)__ldg(P.pool + row0 + r) : (uint32_t)A;
            uint32_t a1 = (r >= 0 && r < n1) ? (uint32_t)__ldg(P.pool + row1 + r) : (uint32_t)A;
            if (CPT) {                                                       // letter -> its rank among the sequence's letters; padding -> the row behind them
                a0 = __popc(mk0 & ((a0 < (uint32_t)A) ? ((1u << a0) - 1u) : 0xffffffffu));
                a1 = __popc(mk1 & ((a1 < (uint32_t)A) ? ((1u << a1) - 1u) : 0xffffffffu));
            }
            rows[e] = (uint16_t)(a0 | (a1 << 8));
        }

        for (int blk = 0; blk < nblk; blk++) {
            // ---- query profile of this column block: prof[h][a][x] = S'(a, colseq_h[x]) (0 on padding) ----
            __syncwarp();
            for (int x = lane; x < W; x += 32) {
                const int X = blk * W + x;
                const int j0 = X - pad0, j1 = X - pad1;
                const int b0 = (j0 >= 0) ? (int)__ldg(P.pool + col0 + j0) : -1;
                const int b1 = (j1 >= 0) ? (int)__ldg(P.pool + col1 + j1) : -1;
                if (CPT) {
                    int ca = 0;
                    for (uint32_t mm = mk0; mm; mm &= mm - 1, ca++)
                        prof[(size_t)ca * W + x] = (b0 >= 0) ? s_sp[(__ffs(mm) - 1) * MAXA + b0] : (uint8_t)0;
                    prof[(size_t)ca * W + x] = 0;
                    ca = 0;
                    for (uint32_t mm = mk1; mm; mm &= mm - 1, ca++)
                        prof[(size_t)(A1 + ca) * W + x] = (b1 >= 0) ? s_sp[(__ffs(mm) - 1) * MAXA + b1] : (uint8_t)0;
                    prof[(size_t)(A1 + ca) * W + x] = 0;
                } else {
                for (int a = 0; a < A; a++) {
                    prof[(size_t)a * W + x] = (b0 >= 0) ? s_sp[a * MAXA + b0] : (uint8_t)0;
                    if (!WIDE && !SHCOL) prof[(size_t)(A1 + a) * W + x] = (b1 >= 0) ? s_sp[a * MAXA + b1] : (uint8_t)0;
                }
                prof[(size_t)A * W + x] = 0;
                if (!WIDE && !SHCOL) prof[(size_t)(A1 + A) * W + x] = 0;
                }
            }
            __syncwarp();

#pragma unroll
            for (int c = 0; c < C; c++) H[c] = 0u;
            if constexpr (RR == 1) {
            uint32_t prev_left = 0u;
            const int steps = N + 31;
            int r = -lane;
            uint32_t ap = (r == 0) ? lds_u16(rows_s) : 0u;        // letters of this lane's next row
            uint32_t *bw = bord + r;                              // lane 31 stores its last column here

            if (blk == 0) {
                for (int s = 0; s < steps; s++, r++, bw++) {
                    uint32_t left_in = __shfl_up_sync(TG_FULL_MASK, H[C - 1], 1);
                    if (lane == 0) left_in = 0u;
                    if ((unsigned)r < (unsigned)N) {
                        Strip<C, WIDE>::run(H, prev_left, left_in, pb0 + (ap & 0xffu) * W, pb1 + (ap >> 8) * W);
                        if (lane == 31) *bw = H[C - 1];
                        ap = lds_u16(rows_s + 2 * (r + 1));
                    } else if (r == -1) ap = lds_u16(rows_s);
                    prev_left = left_in;
                }
            } else {
                // border column of the previous block, 32 rows at a time
                uint32_t bcur = (lane < N) ? __ldcg(bord + lane) : 0u;
                uint32_t bnext = (32 + lane < N) ? __ldcg(bord + 32 + lane) : 0u;
                for (int s = 0; s < steps; s++, r++, bw++) {
                    if ((s & 31) == 0 && s > 0) {
                        bcur = bnext;
                        const int idx = s + 32 + lane;
                        bnext = (idx < N) ? __ldcg(bord + idx) : 0u;
                    }
                    uint32_t left_in = __shfl_up_sync(TG_FULL_MASK, H[C - 1], 1);
                    const uint32_t bl = __shfl_sync(TG_FULL_MASK, bcur, s & 31);
                    if (lane == 0) left_in = bl;
                    if ((unsigned)r < (unsigned)N) {
                        Strip<C, WIDE>::run(H, prev_left, left_in, pb0 + (ap & 0xffu) * W, pb1 + (ap >> 8) * W);
                        if (lane == 31) *bw = H[C - 1];
                        ap = lds_u16(rows_s + 2 * (r + 1));
                    } else if (r == -1) ap = lds_u16(rows_s);
                    prev_left = left_in;
                }
            }
            } else {
            // ---- two rows per step: this lane handles row pair k = s - lane (rows 2k, 2k + 1).  The loop body is
            //      branch free: pairs before the first row (k < 0) run on padding letters with zeroed inputs and
            //      stay zero, pairs past the last row run on padding letters and copy the last row down. ----
            const int NP = (N + 1) >> 1;
            const int steps = NP + 31;
            int k = -lane;
            uint32_t prevL1 = 0u, A7 = 0u;
            uint2 *bord2 = reinterpret_cast<uint2 *>(bord) + 32;   // bord2[k], k >= -31
            uint2 *bw = bord2 + k;
            // border column of the previous block: 32 row pairs per refill, parked in a 64-entry shared-memory ring
            uint2 bnext = (blk > 0 && lane < NP) ? __ldcg(bord2 + lane) : make_uint2(0u, 0u);
            // letters of rows 2k and 2k + 1 of both alignments: four byte loads (LSU) instead of one word and six mask/shift
            // instructions -- the loop is bound by the half-rate ALU pipe, the load/store pipe has room
            const uint32_t rk0 = rows_s + 2 * ROFF;
            uint32_t is_lane0;
            asm("mov.u32 %0, %1;" : "=r"(is_lane0) : "r"((uint32_t)(lane == 0)));      // opaque to the optimiser: keep the multiply-adds
            uint32_t not_lane0;
            asm("mov.u32 %0, %1;" : "=r"(not_lane0) : "r"((uint32_t)(lane != 0)));
            uint32_t a00 = lds_u8(rk0 + 4 * k), a01 = lds_u8(rk0 + 4 * k + 1), a10 = lds_u8(rk0 + 4 * k + 2), a11 = lds_u8(rk0 + 4 * k + 3);
            for (int s = 0; s < steps; s++, k++, bw++) {
                if ((s & 31) == 0) {
                    sts64(ring_s + 8 * ((s + lane) & 63), bnext);
                    const int idx = s + 32 + lane;
                    bnext = (blk > 0 && idx < NP) ? __ldcg(bord2 + idx) : make_uint2(0u, 0u);
                    __syncwarp();
                }
                // left neighbours: the previous lane's last column of the same row pair (one step ago).  A lane that has not
                // reached its first row pair (k < 0) receives zeros by itself: its neighbour was on padding rows too.
                uint32_t L0 = __shfl_up_sync(TG_FULL_MASK, A7, 1);
                uint32_t L1 = __shfl_up_sync(TG_FULL_MASK, H[C - 1], 1);
                const uint2 rb = lds64(ring_s + 8 * (k & 63));
                // lane 0 takes the border of the previous column block instead (it has k = s >= 0).  Written as a multiply-add
                // with a 0/1 factor so that it runs on the FMA pipe: the loop is bound by the half-rate ALU pipe, where a select lives
                L0 = rb.x * is_lane0 + L0 * not_lane0;
                L1 = rb.y * is_lane0 + L1 * not_lane0;
                using S2 = typename std::conditional<C == 16, Strip2x16<WIDE>, Strip2x8<WIDE>>::type;
                A7 = S2::run(H, prevL1, L0, L1, pb0 + a00 * W, pb1 + a01 * W, pb0 + a10 * W, pb1 + a11 * W);
                if (lane == 31) *bw = make_uint2(A7, H[C - 1]);
                a00 = lds_u8(rk0 + 4 * (k + 1)); a01 = lds_u8(rk0 + 4 * (k + 1) + 1);
                a10 = lds_u8(rk0 + 4 * (k + 1) + 2); a11 = lds_u8(rk0 + 4 * (k + 1) + 3);
                prevL1 = L1;
            }
            }
            __syncwarp();

            if (P.right_free) {
                // last row of H for this block: H[n][j] + gr*(m-j) = H'[n][j] + g*(n+j) + gr*(m-j)
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int X = blk * W + lane * C + c;
                    const int j0 = X - pad0 + 1, j1 = X - pad1 + 1;     // 1-based column
                    if (j0 >= 1) rf0 = max(rf0, (int)(WIDE ? H[c] : (H[c] & 0xffffu)) + g * (n0 + j0) + gr * (m0 - j0));
                    if (!WIDE && j1 >= 1) rf1 = max(rf1, (int)(H[c] >> 16) + g * (n1 + j1) + gr * (m1 - j1));
                }
            }
        }

        if (!P.right_free) {
            // the padding rows carried row n of each alignment down to row N: lane 31's last column is H'[n][m]
            if (lane == 31) {
                if (out0 >= 0) P.scores[out0] = (int)(WIDE ? H[C - 1] : (H[C - 1] & 0xffffu)) + g * (n0 + m0);
                if (!WIDE && out1 >= 0) P.scores[out1] = (int)(H[C - 1] >> 16) + g * (n1 + m1);
            }
        } else {
            // last column: bord[r] = H'[r+1][m]; H[i][m] + gr*(n-i) = H'[i][m] + g*(i+m) + gr*(n-i)
            const uint32_t *bcol = (RR == 2) ? bord + 64 : bord;
            for (int rr = lane; rr < N; rr += 32) {
                const uint32_t v = __ldcg(bcol + rr);
                const int i = rr + 1;
                if (i <= n0) rf0 = max(rf0, (int)(WIDE ? v : (v & 0xffffu)) + g * (i + m0) + gr * (n0 - i));
                if (!WIDE && i <= n1) rf1 = max(rf1, (int)(v >> 16) + g * (i + m1) + gr * (n1 - i));
            }
            // whole sequence against end gaps only: H[n][0] + gr*m and H[0][m] + gr*n (gap_left == g)
            rf0 = max(rf0, max(n0 * g + gr * m0, m0 * g + gr * n0));
            rf1 = max(rf1, max(n1 * g + gr * m1, m1 * g + gr * n1));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                rf0 = max(rf0, __shfl_xor_sync(TG_FULL_MASK, rf0, o));
                rf1 = max(rf1, __shfl_xor_sync(TG_FULL_MASK, rf1, o));
            }
            if (lane == 0) {
                if (out0 >= 0) P.scores[out0] = rf0;
                if (out1 >= 0) P.scores[out1] = rf1;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Generic int32 kernel: one CTA per alignment, row by row.  For row i with horizontal gap gh
// (= gap_right in the last row, else gap):
//   T[j]   = max(H[i-1][j-1] + S, H[i-1][j] + gv(j)),  T[0] = i * gap_left
//   H[i][j] = max_{k<=j} (T[k] + gh*(j-k)) = gh*j + prefixmax_k (T[k] - gh*k)
// which turns the left-neighbour dependency into a block-wide prefix max.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_THREADS = 256;
constexpr int NEG_BIG = -(1 << 29);

struct GenParams {
    const uint8_t *pool;
    const tg_nw_job *jobs;
    long long n_items;          // 2 * n_jobs (job, half)
    int32_t *scores;
    int A, g, gl, gr;
    int max_m;
    int16_t sub[MAXA * MAXA];
};

__global__ void __launch_bounds__(GEN_THREADS) nw_generic_kernel(const __grid_constant__ GenParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    int32_t *Hrow = reinterpret_cast<int32_t *>(smem);           // max_m + 1
    int32_t *U = Hrow + (P.max_m + 1);                           // max_m + 1
    __shared__ int32_t s_sub[MAXA * MAXA];
    __shared__ int32_t s_wtot[GEN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < MAXA * MAXA; i += GEN_THREADS) s_sub[i] = P.sub[i];
    __syncthreads();

    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const tg_nw_job *J = P.jobs + (item >> 1);
        const int h = (int)(item & 1);
        const int out = J->out[h];
        if (out < 0) continue;
        const int n = J->n[h], m = J->m[h];
        const uint8_t *a = P.pool + J->row_off[h], *b = P.pool + J->col_off[h];
        const int chunk = (m + 1 + GEN_THREADS - 1) / GEN_THREADS;
        const int jb = min(tid * chunk, m + 1), je = min(jb + chunk, m + 1);
        __syncthreads();
        for (int j = tid; j <= m; j += GEN_THREADS) Hrow[j] = j * P.gl;
        __syncthreads();
        for (int i = 1; i <= n; i++) {
            const int ai = __ldg(a + i - 1);
            const int gh = (i == n) ? P.gr : P.g;
            int local = NEG_BIG;
            for (int j = jb; j < je; j++) {
                int t;
                if (j == 0) t = i * P.gl;
                else {
                    const int gv = (j == m) ? P.gr : P.g;
                    t = max(Hrow[j - 1] + s_sub[ai * MAXA + __ldg(b + j - 1)], Hrow[j] + gv);
                }
                local = max(local, t - gh * j);
                U[j] = local;
            }
            // block-wide exclusive prefix max of `local`
            int inc = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(TG_FULL_MASK, inc, o);
                if (lane >= o) inc = max(inc, v);
            }
            if (lane == 31) s_wtot[warp] = inc;
            int excl = __shfl_up_sync(TG_FULL_MASK, inc, 1);
            if (lane == 0) excl = NEG_BIG;
            __syncthreads();                       // all reads of Hrow done, s_wtot visible
            for (int w = 0; w < warp; w++) excl = max(excl, s_wtot[w]);
            for (int j = jb; j < je; j++) Hrow[j] = gh * j + max(excl, U[j]);
            __syncthreads();
        }
        if (tid == 0) P.scores[out] = Hrow[m];
    }
}

// ------------------------------------------------------------------------------------------------
// MT19937 exactly as CPython seeds and consumes it.  One thread per sequence pair; the 624-word
// state lives in shared memory as mt[word][thread] (bank == thread, conflict free for any word
// index), regenerated lazily one word per draw (identical to the batch twist because the batch
// twist updates in place in the same order), so divergent rejection loops never serialise a twist.
// ------------------------------------------------------------------------------------------------
constexpr int MT_THREADS = 32;
__constant__ uint32_t c_mt_base[624];      // init_genrand(19650218)

// one thread: random.seed(seed); R x (shuffle(seq_i), shuffle(seq_j)) written back to back at dst
__device__ __forceinline__ void mt_shuffle_pair(uint8_t *pool, uint32_t *mt, uint8_t *sp, int pool_stride, int tid,
                                                uint64_t seed, long long src_i, long long src_j, int n, int m,
                                                long long dst, int R)
{
#define MT(i) mt[(i) * MT_THREADS + tid]
    // ---- random.seed(int): init_by_array(32-bit words of abs(seed)) ----
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    const int klen = key[1] ? 2 : 1;
    for (int i = 0; i < 624; i++) MT(i) = c_mt_base[i];
    {
        int i = 1, j = 0;
        uint32_t prev = MT(0);
        for (int k = 624; k; k--) {
            uint32_t v = (MT(i) ^ ((prev ^ (prev >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            MT(i) = v; prev = v;
            i++; j++;
            if (i >= 624) { MT(0) = prev; i = 1; }
            if (j >= klen) j = 0;
        }
        for (int k = 623; k; k--) {
            uint32_t v = (MT(i) ^ ((prev ^ (prev >> 30)) * 1566083941u)) - (uint32_t)i;
            MT(i) = v; prev = v;
            i++;
            if (i >= 624) { MT(0) = prev; i = 1; }
        }
        MT(0) = 0x80000000u;
    }
    int kk = 0;                                   // next state word to regenerate
    for (int rep = 0; rep < 2 * R; rep++) {
        const int len = (rep & 1) ? m : n;
        const uint8_t *src = pool + ((rep & 1) ? src_j : src_i);
        // Working array: pool = list(population) shrinks from the back while the selected items fill the
        // vacated back slots, i.e. result[i] ends up at arr[len-1-i].  Shared memory when it fits
        // (pool_stride > 0), else in place in the destination region.
        uint8_t *arr = pool_stride > 0 ? sp : pool + dst;
        for (int x = 0; x < len; x++) arr[x] = src[x];
        int i = 0;
        while (i < len) {
            // genrand_uint32, lazy twist of word kk
            const int k1 = (kk == 623) ? 0 : kk + 1;
            const int k397 = (kk >= 227) ? kk - 227 : kk + 397;
            const uint32_t y0 = (MT(kk) & 0x80000000u) | (MT(k1) & 0x7fffffffu);
            uint32_t y = MT(k397) ^ (y0 >> 1) ^ ((y0 & 1u) ? 0x9908b0dfu : 0u);
            MT(kk) = y;
            kk = k1;
            y ^= (y >> 11);
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= (y >> 18);
            // _randbelow(nn): k = nn.bit_length(); r = getrandbits(k); redraw while r >= nn
            const uint32_t nn = (uint32_t)(len - i);
            const uint32_t r = y >> __clz(nn);                 // 32 - bit_length(nn) == clz(nn)
            if (r < nn) {
                const uint8_t v = arr[r];                      // result[i] = pool[j]
                arr[r] = arr[nn - 1];                          // pool[j] = pool[n-i-1]
                arr[nn - 1] = v;
                i++;
            }
        }
        if (pool_stride > 0) {
            for (int x = 0; x < len; x++) pool[dst + x] = sp[len - 1 - x];
        } else {
            for (int x = 0; x < len / 2; x++) {
                const uint8_t a = arr[x], b = arr[len - 1 - x];
                arr[x] = b; arr[len - 1 - x] = a;
            }
        }
        dst += len;
    }
#undef MT
}

__global__ void __launch_bounds__(MT_THREADS) mt_shuffle_kernel(uint8_t *pool, const tg_shuffle_job *jobs,
                                                               long long n_jobs, int R, int pool_stride)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *mt = reinterpret_cast<uint32_t *>(smem);                 // [624][MT_THREADS]
    uint8_t *sp = smem + 624 * MT_THREADS * sizeof(uint32_t) + (size_t)threadIdx.x * pool_stride;
    for (long long job = (long long)blockIdx.x * MT_THREADS + threadIdx.x; job < n_jobs;
         job += (long long)gridDim.x * MT_THREADS) {
        const tg_shuffle_job J = jobs[job];
        mt_shuffle_pair(pool, mt, sp, pool_stride, threadIdx.x, J.seed, J.src_i, J.src_j, J.n, J.m, J.dst_off, R);
    }
}

struct BatchShuffleParams {
    uint8_t *pool;
    const long long *seq_off;
    const int *order;
    int n_seq, adjust_lens, min_viable, R;
    long long q0, q1;
    const uint8_t *flags;
    long long shuf_base, slot;
    long long seed0;
    int pool_stride;
    unsigned long long *counter;
    const long long *mat_pair_off; const int *mat_seq_off; int n_mat;   // several matrices in one pair space (0 = one matrix)
};

// null model of every non-early pair of a batch; seed = abs(rng_seed + p) (tg_align.py:110-111,157,170-172)
__global__ void __launch_bounds__(MT_THREADS) mt_shuffle_batch_kernel(const BatchShuffleParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *mt = reinterpret_cast<uint32_t *>(smem);
    uint8_t *sp = smem + 624 * MT_THREADS * sizeof(uint32_t) + (size_t)threadIdx.x * P.pool_stride;
    // dynamic scheduling, one warp fetches 32 pairs at a time (early-out pairs cost nothing but a lane)
    for (;;) {
        unsigned long long base = 0;
        if (threadIdx.x == 0) base = atomicAdd(P.counter, (unsigned long long)MT_THREADS);
        base = __shfl_sync(TG_FULL_MASK, base, 0);
        if ((long long)base >= P.q1 - P.q0) break;
        const long long ql = (long long)base + threadIdx.x;
        if (ql < P.q1 - P.q0 && !P.flags[ql]) {
            const PairDesc d = describe_pair(P, P.q0 + ql);
            const long long sv = P.seed0 + d.p;
            const uint64_t seed = (uint64_t)(sv < 0 ? -sv : sv);
            mt_shuffle_pair(P.pool, mt, sp, P.pool_stride, threadIdx.x, seed, d.off_i, d.off_j, d.n, d.m,
                            P.shuf_base + ql * P.slot, P.R);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Streamed null model (tg_dist_batch): the same bits as mt_shuffle_pair, split so that the expensive
// part -- regenerating and tempering the MT19937 state -- runs warp-parallel:
//   mt_seed_batch_kernel    thread per pair: init_by_array, state -> global (coalesced write-out)
//   mt_stream_batch_kernel  warp per pair: block twists (3 data-parallel phases) + tempering, the top
//                           16 bits of every output word go to a per-pair stream (getrandbits(k) only
//                           ever looks at the top k <= 16 bits)
//   mt_sample_warp_kernel   warp per pair: random.sample's swap-remove loop consuming the stream, 32 draws per step
// ------------------------------------------------------------------------------------------------
constexpr int MTS_PAD = 33;          // row stride of the transposing state buffer: conflict free both ways

struct StreamParams {
    uint8_t *pool;
    const long long *seq_off;
    const int *order;
    int n_seq, adjust_lens, min_viable, R;
    long long q0, q1;
    const uint8_t *flags;
    long long shuf_base, slot;
    long long seed0;
    int pool_stride;
    int cap;                         // stream capacity per pair in 16-bit draws (multiple of 624)
    uint32_t *state;                 // [pairs][624]
    uint16_t *stream;                // [pairs][cap]
    unsigned long long *counter;
    int *overflow;
    const int *live;                 // pair indices (q - q0) that survived the early-out test, in no particular order
    const int *n_live;
    const long long *mat_pair_off; const int *mat_seq_off; int n_mat;   // several matrices in one pair space (0 = one matrix)
};

__host__ __device__ __forceinline__ int mt_blocks_needed(int total_len)
{
    // expected draws <= 1.47 per element (worst just above a power of two), sd ~ sqrt(0.6 * elements)
    return (int)(((long long)total_len * 8 / 5 + 256 + 623) / 624);
}

__global__ void __launch_bounds__(MT_THREADS) mt_seed_batch_kernel(const StreamParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *mt = reinterpret_cast<uint32_t *>(smem);                 // [624][MTS_PAD]
    const int tid = threadIdx.x;
    const long long npairs = *P.n_live;
#define MTP(i) mt[(i) * MTS_PAD + tid]
    for (;;) {
        unsigned long long base = 0;
        if (tid == 0) base = atomicAdd(P.counter, (unsigned long long)MT_THREADS);
        base = __shfl_sync(TG_FULL_MASK, base, 0);
        if ((long long)base >= npairs) break;
        const bool active = (long long)base + tid < npairs;
        const long long ql = active ? P.live[base + tid] : 0;
        if (active) {
            const PairDesc d = describe_pair(P, P.q0 + ql);
            const long long sv = P.seed0 + d.p;
            const uint64_t seed = (uint64_t)(sv < 0 ? -sv : sv);
            uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
            const int klen = key[1] ? 2 : 1;
            for (int i = 0; i < 624; i++) MTP(i) = c_mt_base[i];
            int i = 1, j = 0;
            uint32_t prev = MTP(0);
            for (int k = 624; k; k--) {
                uint32_t v = (MTP(i) ^ ((prev ^ (prev >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
                MTP(i) = v; prev = v;
                i++; j++;
                if (i >= 624) { MTP(0) = prev; i = 1; }
                if (j >= klen) j = 0;
            }
            for (int k = 623; k; k--) {
                uint32_t v = (MTP(i) ^ ((prev ^ (prev >> 30)) * 1566083941u)) - (uint32_t)i;
                MTP(i) = v; prev = v;
                i++;
                if (i >= 624) { MTP(0) = prev; i = 1; }
            }
            MTP(0) = 0x80000000u;
        }
        __syncwarp();
        const uint32_t act = __ballot_sync(TG_FULL_MASK, active);
        for (int l = 0; l < MT_THREADS; l++) {
            if (!((act >> l) & 1u)) continue;
            const long long qd = __shfl_sync(TG_FULL_MASK, ql, l);
            uint32_t *dst = P.state + qd * 624;
            for (int i = tid; i < 624; i += MT_THREADS) dst[i] = mt[i * MTS_PAD + l];
        }
        __syncwarp();
    }
#undef MTP
}

constexpr int MTG_WARPS = 8;

__global__ void __launch_bounds__(MTG_WARPS * 32) mt_stream_batch_kernel(const StreamParams P)
{
    __shared__ uint32_t s_mt[MTG_WARPS][624];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *m = s_mt[warp];
    const long long npairs = *P.n_live;
    for (;;) {
        unsigned long long li = 0;
        if (lane == 0) li = atomicAdd(P.counter, 1ULL);
        li = __shfl_sync(TG_FULL_MASK, li, 0);
        if ((long long)li >= npairs) break;
        const long long ql = P.live[li];
        const PairDesc d = describe_pair(P, P.q0 + (long long)ql);
        int nblk = mt_blocks_needed(P.R * (d.n + d.m));
        if (nblk * 624 > P.cap) nblk = P.cap / 624;
        const uint32_t *src = P.state + (long long)ql * 624;
        __syncwarp();
        for (int i = lane; i < 624; i += 32) m[i] = src[i];
        __syncwarp();
        uint32_t *out = reinterpret_cast<uint32_t *>(P.stream + (long long)ql * P.cap);
        for (int blk = 0; blk < nblk; blk++) {
            // mt[kk] = mt[kk + 397 (mod 624)] ^ twist(mt[kk], mt[kk + 1]) for kk = 0 .. 623, in place and in order:
            // three ranges whose inputs are complete when the range starts, then the last word
#define TG_TWIST_RANGE(lo, hi, far)                                                                    \
    for (int i0 = (lo); i0 < (hi); i0 += 32) {                                                         \
        const int i = i0 + lane;                                                                       \
        uint32_t v = 0;                                                                                \
        if (i < (hi)) {                                                                                \
            const uint32_t y = (m[i] & 0x80000000u) | (m[i + 1] & 0x7fffffffu);                        \
            v = m[i + (far)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);                               \
        }                                                                                              \
        __syncwarp();                                                                                  \
        if (i < (hi)) m[i] = v;                                                                        \
        __syncwarp();                                                                                  \
    }
            TG_TWIST_RANGE(0, 227, 397)
            TG_TWIST_RANGE(227, 454, -227)
            TG_TWIST_RANGE(454, 623, -227)
#undef TG_TWIST_RANGE
            if (lane == 0) {
                const uint32_t y = (m[623] & 0x80000000u) | (m[0] & 0x7fffffffu);
                m[623] = m[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncwarp();
            // tempering; two consecutive 16-bit draws per 32-bit store
            for (int i = 2 * lane; i < 624; i += 64) {
                uint32_t y0 = m[i], y1 = m[i + 1];
                y0 ^= (y0 >> 11); y0 ^= (y0 << 7) & 0x9d2c5680u; y0 ^= (y0 << 15) & 0xefc60000u; y0 ^= (y0 >> 18);
                y1 ^= (y1 >> 11); y1 ^= (y1 << 7) & 0x9d2c5680u; y1 ^= (y1 << 15) & 0xefc60000u; y1 ^= (y1 >> 18);
                out[(blk * 624 + i) >> 1] = (y0 >> 16) | (y1 & 0xffff0000u);
            }
            __syncwarp();
        }
    }
}

// Warp per pair.  random.sample's loop "j = _randbelow(n - i); result[i] = pool[j]; pool[j] = pool[n - i - 1]" is
// sequential, but 32 consecutive draws can be resolved together:
//  1. acceptance: draw l is accepted iff r_l < n_l where n_l = remaining - (#accepted draws before l) -- a fixed point
//     over the ballot mask that is reached after one or two rounds (a lane's decision rarely depends on its prefix);
//  2. the accepted draws (in order s = 0 .. S-1) read pool[j_s] and the tail pool[t_s], t_s = remaining - 1 - s.  All reads
//     are issued against the pool as it was before the group; the effect of the group's earlier writes is patched in:
//       tail value of step s  = tail value of the LAST earlier step that wrote position t_s (j_s' == t_s), else pool[t_s]
//       result of step s      = tail value of the LAST earlier step with the same j, else pool[j_s]
//     (first relation via a 32-entry shared-memory atomicMax + pointer jumping, second via match.any);
//  3. per distinct j the last step writes its tail value back.
constexpr int MTW_WARPS = 8;
constexpr int MTW_EXTRA = 256;           // per-warp scratch behind the pool: 32 x u16 compaction + 32 x int hits

__global__ void __launch_bounds__(MTW_WARPS * 32) mt_sample_warp_kernel(const StreamParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *arr = smem + (size_t)warp * P.pool_stride;
    uint16_t *scr = reinterpret_cast<uint16_t *>(arr + P.pool_stride - MTW_EXTRA);
    int *hitarr = reinterpret_cast<int *>(arr + P.pool_stride - MTW_EXTRA + 64);
    const long long npairs = *P.n_live;
    const unsigned lt = (1u << lane) - 1u;
    bool ok = true;
    for (;;) {
        unsigned long long li = 0;
        if (lane == 0) li = atomicAdd(P.counter, 1ULL);
        li = __shfl_sync(TG_FULL_MASK, li, 0);
        if ((long long)li >= npairs) break;
        const long long ql = P.live[li];
        const PairDesc d = describe_pair(P, P.q0 + ql);
        int cap = mt_blocks_needed(P.R * (d.n + d.m)) * 624;
        if (cap > P.cap) cap = P.cap;
        const uint16_t *st = P.stream + ql * P.cap;
        int c = 0;                                               // draws consumed so far
        long long dst = P.shuf_base + ql * P.slot;
        for (int rep = 0; rep < 2 * P.R; rep++) {
            const int len = (rep & 1) ? d.m : d.n;
            const uint8_t *src = P.pool + ((rep & 1) ? d.off_j : d.off_i);
            uint8_t *out = P.pool + dst;
            // pool = list(population): aligned 32-bit loads + funnel shift (reads up to 7 bytes past the end of the
            // source stay inside the pool allocation)
            {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
                uint32_t *d32 = reinterpret_cast<uint32_t *>(arr);
                __syncwarp();
                for (int q = lane; q * 4 < len; q += 32) d32[q] = __funnelshift_r(__ldg(w + q), __ldg(w + q + 1), sh);
                __syncwarp();
            }
            int nn = len;                                        // remaining pool size
            while (nn > 0) {
                const int ci = min(c + lane, cap - 1);
                const uint32_t v = __ldcg(st + ci);
                // ---- 1. acceptance fixed point ----
                unsigned A = 0, Aprev;
                int a;
                uint32_t r = 0;
                do {
                    Aprev = A;
                    a = __popc(Aprev & lt);
                    const int nl = nn - a;
                    bool acc = false;
                    if (nl > 0) {
                        r = v >> (__clz(nl) - 16);               // getrandbits(nl.bit_length()): top bits of the output word
                        acc = r < (uint32_t)nl;
                    }
                    A = __ballot_sync(TG_FULL_MASK, acc);
                } while (A != Aprev);
                const int S = __popc(A);
                c += (S == nn) ? (32 - __clz(A)) : 32;           // the draws behind the last element belong to the next shuffle
                if (c > cap) ok = false;                         // stream exhausted: flag it (results are discarded)
                if ((A >> lane) & 1u) scr[a] = (uint16_t)r;
                __syncwarp();
                // ---- 2. the S swap-removes ----
                const bool act = lane < S;
                const int j = act ? (int)scr[lane] : 0;
                const int t = nn - 1 - lane;
                uint32_t vj = 0, vt = 0;
                if (act) { vj = arr[j]; vt = arr[t]; }
                const int tg = nn - 1 - j;                       // the step whose tail read sees this step's write
                const bool writes_tail = act && tg > lane && tg < S;
                const unsigned same = __match_any_sync(TG_FULL_MASK, act ? j : -1 - lane);
                uint32_t tailval = vt;
                if (__any_sync(TG_FULL_MASK, writes_tail)) {     // rare for long pools: some j falls into the group's tail window
                    hitarr[lane] = -1;
                    __syncwarp();
                    if (writes_tail) atomicMax(&hitarr[tg], lane);
                    __syncwarp();
                    const int hit = hitarr[lane];
                    int root = hit < 0 ? lane : hit;
                    for (;;) {
                        const int h = __shfl_sync(TG_FULL_MASK, hit, root);
                        if (!__any_sync(TG_FULL_MASK, h >= 0)) break;
                        if (h >= 0) root = h;
                    }
                    tailval = __shfl_sync(TG_FULL_MASK, vt, root);
                }
                const unsigned before = same & lt;
                const uint32_t tw = __shfl_sync(TG_FULL_MASK, tailval, before ? 31 - __clz(before) : lane);
                if (act) out[len - nn + lane] = (uint8_t)(before ? tw : vj);        // result[i]
                // ---- 3. write-back (every read of the group has completed: the shuffles above are warp-wide) ----
                if (act && (same >> lane) == 1u) arr[j] = (uint8_t)tailval;
                __syncwarp();
                nn -= S;
            }
            dst += len;
        }
    }
    if (!ok) atomicExch(P.overflow, 1);
}

struct FlagParams {
    const long long *seq_off;
    const int *order;
    int n_seq, adjust_lens, min_viable, R;
    long long q0, q1;
    const int32_t *aln;
    const int32_t *diag_cs;       // prefix sums of the substitution-matrix diagonal over pool codes, or NULL
    double match;                 // match score (match/mismatch mode)
    uint8_t *flags;
    double *iden;
    int32_t *shape;               // [2 * pairs] length-adjusted n, m
    int *live;                    // optional: compacted list of non-early pairs (unordered) ...
    int *n_live;                  // ... and its length
    const long long *mat_pair_off; const int *mat_seq_off; int n_mat;   // several matrices in one pair space (0 = one matrix)
};

// identity score and the early-out test of tvr_distance (tg_align.py:127-138), in float64 like the reference
__global__ void __launch_bounds__(256) dist_flags_kernel(const FlagParams P)
{
    for (long long ql = (long long)blockIdx.x * blockDim.x + threadIdx.x; ql < P.q1 - P.q0; ql += (long long)gridDim.x * blockDim.x) {
        const PairDesc d = describe_pair(P, P.q0 + ql);
        double iden;
        if (P.diag_cs) {
            const double is1 = (double)(P.diag_cs[d.off_i + d.n] - P.diag_cs[d.off_i]);
            const double is2 = (double)(P.diag_cs[d.off_j + d.m] - P.diag_cs[d.off_j]);
            iden = (is1 + is2) / 2.0;
        } else {
            iden = P.match * ((double)(d.n + d.m) / 2.0);
        }
        const double aln = (double)P.aln[ql];
        const bool early = aln < -(0.900 * iden);
        P.flags[ql] = early ? 1 : 0;
        if (P.live && !early) P.live[atomicAdd(P.n_live, 1)] = (int)ql;
        P.iden[ql] = iden;
        P.shape[2 * ql] = d.n;
        P.shape[2 * ql + 1] = d.m;
    }
}

// ------------------------------------------------------------------------------------------------
// Float epilogue (tg_align.py:137-149) + matrix fill (tg_align.py:180-186), one thread per pair.
// ------------------------------------------------------------------------------------------------
struct EpiParams {
    const long long *seq_off;
    const int *order;
    int n_seq, adjust_lens, min_viable, R;
    long long q0, q1, ambig_base;
    const int32_t *aln, *rnd;
    const uint8_t *flags;
    const double *iden;
    float *matrix;
    long long *ambig;
    unsigned long long *stats;
    int ambig_cap;
    double eps;
    const long long *mat_pair_off; const int *mat_seq_off; int n_mat;   // several matrices in one pair space (0 = one matrix)
    const long long *mat_out_off;     // offset of matrix b in `matrix` (n_mat > 0)
};

// Python's min(x, MAX_SEQ_DIST) / max(x, MIN_SEQ_DIST): a NaN first argument is kept
__device__ __forceinline__ float clamp_dist(double x)
{
    x = (10.0 < x) ? 10.0 : x;
    x = (1e-4 > x) ? 1e-4 : x;
    return __double2float_rn(x);
}

__global__ void __launch_bounds__(256) dist_epilogue_kernel(const EpiParams P)
{
    unsigned long long cells_a = 0, cells_b = 0, n_early = 0;
    for (long long ql = (long long)blockIdx.x * blockDim.x + threadIdx.x; ql < P.q1 - P.q0; ql += (long long)gridDim.x * blockDim.x) {
        const PairDesc d = describe_pair(P, P.q0 + ql);
        const unsigned long long cells = (unsigned long long)d.n * (unsigned long long)d.m;
        cells_a += cells;
        float f;
        if (P.flags[ql]) {
            f = 10.0f;
            n_early++;
        } else {
            cells_b += cells * (unsigned long long)P.R;
            const double aln = (double)P.aln[ql], iden = P.iden[ql];
            double rsum = 0.0;
            for (int k = 0; k < P.R; k++) rsum = __dadd_rn(rsum, (double)P.rnd[ql * P.R + k]);
            const double rmean = __ddiv_rn(rsum, (double)P.R);                    // np.mean; R == 0 -> nan
            if (rmean >= aln) {
                f = 10.0f;
            } else {
                const double l = log(__ddiv_rn(__dsub_rn(aln, rmean), __dsub_rn(iden, rmean)));
                f = clamp_dist(-l);
                if (l != 0.0 && isfinite(l)) {
                    const float f0 = clamp_dist(-__dmul_rn(l, 1.0 - P.eps)), f1 = clamp_dist(-__dmul_rn(l, 1.0 + P.eps));
                    if (f0 != f || f1 != f) {
                        const unsigned long long k = atomicAdd(&P.stats[3], 1ULL);
                        if (k < (unsigned long long)P.ambig_cap) P.ambig[k] = P.ambig_base + ql;
                    }
                }
            }
        }
        float *M = P.n_mat > 0 ? P.matrix + P.mat_out_off[d.mat] : P.matrix;
        M[(size_t)d.i * d.n_local + d.j] = f;
        M[(size_t)d.j * d.n_local + d.i] = f;
    }
    // block totals -> three atomics per block
    for (int o = 16; o > 0; o >>= 1) {
        cells_a += __shfl_down_sync(0xffffffffu, cells_a, o);
        cells_b += __shfl_down_sync(0xffffffffu, cells_b, o);
        n_early += __shfl_down_sync(0xffffffffu, n_early, o);
    }
    __shared__ unsigned long long s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_tot[0], cells_a);
        atomicAdd(&s_tot[1], cells_b);
        atomicAdd(&s_tot[2], n_early);
    }
    __syncthreads();
    if (threadIdx.x < 3 && s_tot[threadIdx.x]) atomicAdd(&P.stats[threadIdx.x], s_tot[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// Integer-ALU peak: 8 independent accumulator chains of (add, VIMNMX3.U16x2) per thread.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters, uint32_t seed)
{
    // eight independent chains per thread, nothing but the measured instruction inside the loop (the loop counter's add and
    // branch are 2 of 66 issue slots; the operands alternate between two register pairs so that no iteration is a no-op the
    // compiler could drop)
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed * (threadIdx.x + 1) + k;
    const uint32_t b0 = seed ^ 0x00030005u, c0 = seed | 0x00010001u, b1 = seed + 0x00070009u, c1 = seed ^ 0x01010101u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) a[k] = __vimax3_u16x2(a[k], (u & 1) ? b1 : b0, (u & 1) ? c1 : c0);       // DPX only
                else a[k] = __vimax3_u16x2(a[k] + ((u & 1) ? b1 : b0), a[(k + 1) & 7], c0);              // add + DPX (cell shape)
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
static int g_wave_variant = 3;     // 0: 16 cols x 1 row, 1: 8 x 1, 2: 8 cols x 2 rows per step; TG_WAVE_VARIANT overrides
static int g_wave_warps = 0;       // 0 = as many as shared memory allows

static bool g_wave_shcol = true;            // TG_WAVE_SHCOL=0 disables the shared-column phase A (A/B measurements)
static void wave_env()
{
    static bool done = false;
    if (done) return;
    done = true;
    if (const char *e = getenv("TG_WAVE_VARIANT")) { const int v = atoi(e); if (v >= 0 && v <= 3) g_wave_variant = v; }
    if (const char *e = getenv("TG_WAVE_WARPS")) { const int v = atoi(e); if (v >= 1 && v <= WAVE_MAX_WARPS) g_wave_warps = v; }
    if (const char *e = getenv("TG_WAVE_SHCOL")) g_wave_shcol = atoi(e) != 0;
}

TG_EXPORT size_t tg_nw_wave_scratch_bytes(int max_n)
{
    const int sms = tg_sm_count();
    const size_t stride = ((size_t)max_n + 223) / 32 * 32;
    return (size_t)(sms > 0 ? sms : 148) * WAVE_MAX_WARPS * stride * sizeof(uint32_t);
}

static int wave_envelope(const tg_nw_scoring *sc, int *max_sp_out)
{
    if (!sc) return tg_set_err(TG_EINVAL, "scoring is NULL");
    if (sc->alphabet < 1 || sc->alphabet >= MAXA) return tg_set_err(TG_EINVAL, "alphabet %d outside 1..%d", sc->alphabet, MAXA - 1);
    if (sc->gap_left != sc->gap) return tg_set_err(TG_ERANGE, "wave kernel needs gap_left == gap");
    if (sc->gap_right < sc->gap) return tg_set_err(TG_ERANGE, "wave kernel needs gap_right >= gap");
    int max_sp = 0;
    for (int a = 0; a < sc->alphabet; a++)
        for (int b = 0; b < sc->alphabet; b++) {
            const int sp = sc->sub[a * MAXA + b] - 2 * sc->gap;
            if (sp < 0 || sp > 127) return tg_set_err(TG_ERANGE, "S - 2*gap = %d outside 0..127", sp);
            if (sp > max_sp) max_sp = sp;
        }
    *max_sp_out = max_sp;
    return TG_OK;
}

TG_EXPORT int tg_nw_wave_check(const tg_nw_scoring *sc, const tg_nw_job *h_jobs, int64_t n_jobs)
{
    int max_sp = 0;
    int rc = wave_envelope(sc, &max_sp);
    if (rc != TG_OK) return rc;
    for (int64_t q = 0; q < n_jobs; q++) {
        const tg_nw_job &J = h_jobs[q];
        for (int h = 0; h < 2; h++)
            if (J.n[h] < 1 || J.m[h] < 1) return tg_set_err(TG_EINVAL, "job %lld half %d has an empty sequence", (long long)q, h);
        const long long N = J.n[0] > J.n[1] ? J.n[0] : J.n[1], M = J.m[0] > J.m[1] ? J.m[0] : J.m[1];
        if ((long long)max_sp * (N < M ? N : M) > 65535)
            return tg_set_err(TG_ERANGE, "job %lld (%lld x %lld) overflows 16-bit scores", (long long)q, N, M);
    }
    return TG_OK;
}

template <int V, int MODE, bool WIDE, bool MM = false>
static int wave_launch_t(const WaveParams &P, int grid, int warps, size_t smem, cudaStream_t st)
{
    TG_CUDA_CHECK(cudaFuncSetAttribute(nw_wave_kernel<V, MODE, WIDE, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nw_wave_kernel<V, MODE, WIDE, MM><<<grid, warps * 32, smem, st>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

template <int V, bool WIDE = false>
static int wave_launch_v(const WaveParams &P, int mode, int grid, int warps, size_t smem, cudaStream_t st)
{
    if (mode == JOBS_EXPLICIT) return wave_launch_t<V, JOBS_EXPLICIT, WIDE>(P, grid, warps, smem, st);
    if (mode == JOBS_PHASE_A_SHCOL) return wave_launch_t<3, JOBS_PHASE_A_SHCOL, false, false>(P, grid, warps, smem, st);
    if (mode == JOBS_PHASE_B_CPT) {          // compact profiles: default geometry, 16-bit cells
        if (P.n_mat > 0) return wave_launch_t<3, JOBS_PHASE_B_CPT, false, true>(P, grid, warps, smem, st);
        return wave_launch_t<3, JOBS_PHASE_B_CPT, false, false>(P, grid, warps, smem, st);
    }
    if (V == 3 && P.n_mat > 0) {              // several matrices in one pair space: default geometry only
        if (mode == JOBS_PHASE_A) return wave_launch_t<3, JOBS_PHASE_A, WIDE, true>(P, grid, warps, smem, st);
        return wave_launch_t<3, JOBS_PHASE_B, WIDE, true>(P, grid, warps, smem, st);
    }
    if (mode == JOBS_PHASE_A) return wave_launch_t<V, JOBS_PHASE_A, WIDE>(P, grid, warps, smem, st);
    return wave_launch_t<V, JOBS_PHASE_B, WIDE>(P, grid, warps, smem, st);
}

// Optional timing of the wave launches (bench.py's roofline figure): CUDA events on the launching stream around every launch.
static bool g_timing_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_timing_events;

static int wave_launch_inner(const WaveParams &P, int mode, int variant, int grid, int warps, size_t smem, cudaStream_t st, bool wide);

// wide (32-bit cells) exists for the default geometry only
static int wave_launch(const WaveParams &P, int mode, int variant, int grid, int warps, size_t smem, cudaStream_t st, bool wide = false)
{
    if (!g_timing_on) return wave_launch_inner(P, mode, variant, grid, warps, smem, st, wide);
    cudaEvent_t e0, e1;
    TG_CUDA_CHECK(cudaEventCreate(&e0));
    TG_CUDA_CHECK(cudaEventCreate(&e1));
    TG_CUDA_CHECK(cudaEventRecord(e0, st));
    const int rc = wave_launch_inner(P, mode, variant, grid, warps, smem, st, wide);
    TG_CUDA_CHECK(cudaEventRecord(e1, st));
    g_timing_events.emplace_back(e0, e1);
    return rc;
}

static int wave_launch_inner(const WaveParams &P, int mode, int variant, int grid, int warps, size_t smem, cudaStream_t st, bool wide)
{
    if (wide) return wave_launch_v<3, true>(P, mode, grid, warps, smem, st);
    if (variant == 0) return wave_launch_v<0>(P, mode, grid, warps, smem, st);
    if (variant == 1) return wave_launch_v<1>(P, mode, grid, warps, smem, st);
    if (variant == 3) return wave_launch_v<3>(P, mode, grid, warps, smem, st);
    return wave_launch_v<2>(P, mode, grid, warps, smem, st);
}

// shared set-up of the three job modes
static int wave_setup(WaveParams &P, const tg_nw_scoring *sc, int max_n, int *C_out, int *warps_out, size_t *smem_out, bool wide = false,
                      bool one_profile = false, int prof_rows = 0, int smem_reserve = 0)
{
    wave_env();
    int max_sp = 0;
    int rc = wave_envelope(sc, &max_sp);
    if (rc != TG_OK) return rc;
    if (max_n < 1) return tg_set_err(TG_EINVAL, "max_n < 1");
    P.scratch_stride = (int)(((size_t)max_n + 223) / 32 * 32);
    P.row_cap = (max_n + 140 + 7) / 8 * 8;
    P.A = sc->alphabet; P.g = sc->gap; P.gr = sc->gap_right; P.right_free = (sc->gap_right != sc->gap);
    for (int a = 0; a < MAXA; a++)
        for (int b = 0; b < MAXA; b++)
            P.sp[a * MAXA + b] = (a < sc->alphabet && b < sc->alphabet) ? (uint8_t)(sc->sub[a * MAXA + b] - 2 * sc->gap) : 0;
    const int C = (wide || g_wave_variant == 0 || g_wave_variant == 3) ? 16 : 8;
    const size_t per_warp = (size_t)((wide || one_profile) ? 1 : 2) * (prof_rows > 0 ? prof_rows : sc->alphabet + 1) * 32 * C + (size_t)P.row_cap * 2 + 512;
    int warps = (int)((226 * 1024 - smem_reserve - MAXA * MAXA) / per_warp);
    if (warps > WAVE_MAX_WARPS) warps = WAVE_MAX_WARPS;
    if (g_wave_warps && warps > g_wave_warps) warps = g_wave_warps;
    if (warps < 1) return tg_set_err(TG_ERANGE, "sequences of %d rows do not fit the wave kernel's shared memory", max_n);
    *C_out = wide ? 3 : g_wave_variant; *warps_out = warps;
    *smem_out = MAXA * MAXA + (size_t)warps * per_warp;
    return TG_OK;
}

// A launch with fewer jobs than warp slots spreads them over all SMs (fewer warps per CTA) instead of filling a few CTAs:
// a small matrix (per-cluster calls) otherwise runs on a third of the GPU.
static void wave_spread(long long n_jobs, int warps, int sms, int *grid, int *w)
{
    if (n_jobs >= (long long)sms * warps) { *grid = sms; *w = warps; return; }
    int ww = (int)((n_jobs + sms - 1) / sms);
    if (ww < 1) ww = 1;
    *w = ww;
    *grid = (int)((n_jobs + ww - 1) / ww);
}

static int nw_scores_wave(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_n,
                          const tg_nw_scoring *sc, int32_t *d_scores, void *d_scratch, int32_t *d_counter,
                          void *stream, bool wide)
{
    if (n_jobs <= 0) return TG_OK;
    if (!d_pool || !d_jobs || !d_scores || !d_scratch || !d_counter) return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    WaveParams P;
    memset(&P, 0, sizeof(P));
    int C, warps;
    size_t smem;
    int rc = wave_setup(P, sc, max_n, &C, &warps, &smem, wide);
    if (rc != TG_OK) return rc;
    P.pool = d_pool; P.jobs = d_jobs; P.n_jobs = wide ? 2 * n_jobs : n_jobs; P.scores = d_scores;
    P.scratch = reinterpret_cast<uint32_t *>(d_scratch);
    P.counter = reinterpret_cast<unsigned long long *>(d_counter);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TG_CUDA_CHECK(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    int grid, w;
    wave_spread(P.n_jobs, warps, sms, &grid, &w);
    return wave_launch(P, JOBS_EXPLICIT, C, grid, w, smem, st, wide);
}

TG_EXPORT int tg_nw_scores_wave(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_n,
                                const tg_nw_scoring *sc, int32_t *d_scores, void *d_scratch, int32_t *d_counter,
                                void *stream)
{
    return nw_scores_wave(d_pool, d_jobs, n_jobs, max_n, sc, d_scores, d_scratch, d_counter, stream, false);
}

TG_EXPORT int tg_nw_scores_wave32(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_n,
                                  const tg_nw_scoring *sc, int32_t *d_scores, void *d_scratch, int32_t *d_counter,
                                  void *stream)
{
    return nw_scores_wave(d_pool, d_jobs, n_jobs, max_n, sc, d_scores, d_scratch, d_counter, stream, true);
}

static int mt_upload_base();

TG_EXPORT size_t tg_dist_batch_mt_work_bytes(int64_t n_pairs, int R, int max_len)
{
    if (n_pairs <= 0 || R <= 0 || max_len <= 0) return 0;
    const size_t cap = (size_t)mt_blocks_needed(2 * R * max_len) * 624;
    return (size_t)n_pairs * (624 * sizeof(uint32_t) + cap * sizeof(uint16_t) + sizeof(int)) + 512;
}

TG_EXPORT int tg_dist_batch(const tg_dist_batch_args *a, const tg_nw_scoring *sc, void *stream)
{
    if (!a || !sc) return tg_set_err(TG_EINVAL, "NULL argument");
    const long long npairs = a->q1 - a->q0;
    if (npairs <= 0) return TG_OK;
    if (!a->d_pool || !a->d_seq_off || !a->d_order || !a->d_aln || !a->d_flags || !a->d_iden || !a->d_shape ||
        !a->d_scratch || !a->d_counter || (a->R > 0 && !a->d_rnd))
        return tg_set_err(TG_EINVAL, "NULL device pointer");
    if (npairs > 0x3fffffff / (a->R > 0 ? a->R : 1)) return tg_set_err(TG_ERANGE, "batch of %lld pairs is too large", npairs);
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    // phases (bit mask, 0 = all): 1 = real alignments + flags + seeding, 2 = random streams + sampling, 4 = null-model alignments.
    // A caller that pipelines batches runs phase 2 of one batch on a second stream under phase 1 of the next one.
    const int ph = a->phases ? a->phases : 7;
    int reserve = a->smem_reserve > 0 ? a->smem_reserve : 0;
    if (a->smem_reserve < 0) {               // room for one CTA of the stream kernel (static) or of the sampling kernel, whichever is larger
        const int wstride0 = (a->max_len + 15) / 16 * 16 + MTW_EXTRA;
        const int need = MTW_WARPS * wstride0 > MTG_WARPS * 624 * 4 ? MTW_WARPS * wstride0 : MTG_WARPS * 624 * 4;
        reserve = need + 2048;
    }
    WaveParams P;
    memset(&P, 0, sizeof(P));
    int C, warps;
    size_t smem;
    bool wide = false;
    {
        int max_sp = 0;
        int rc0 = wave_envelope(sc, &max_sp);
        if (rc0 != TG_OK) return rc0;
        wide = (long long)max_sp * a->max_len > 65535;          // 16-bit cells would overflow: one alignment per job, 32-bit cells
    }
    int rc = wave_setup(P, sc, a->max_len, &C, &warps, &smem, wide, false, 0, reserve);
    if (rc != TG_OK) return rc;
    if (a->n_mat > 0 && C != 3) return tg_set_err(TG_EINVAL, "several matrices per call need the default wave variant (TG_WAVE_VARIANT=3)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    P.pool = a->d_pool; P.jobs = nullptr; P.scores = a->d_aln;
    P.scratch = reinterpret_cast<uint32_t *>(a->d_scratch);
    P.counter = reinterpret_cast<unsigned long long *>(a->d_counter);
    P.seq_off = reinterpret_cast<const long long *>(a->d_seq_off); P.order = a->d_order; P.n_seq = a->n_seq;
    P.adjust_lens = a->adjust_lens; P.min_viable = a->min_viable; P.R = a->R;
    P.q0 = a->q0; P.q1 = a->q1; P.flags = a->d_flags; P.shuf_base = a->shuf_base; P.slot = a->slot;
    if (a->n_mat < 0 || (a->n_mat > 0 && (!a->d_mat_pair_off || !a->d_mat_seq_off))) return tg_set_err(TG_EINVAL, "bad matrix table");
    P.mat_pair_off = reinterpret_cast<const long long *>(a->d_mat_pair_off); P.mat_seq_off = a->d_mat_seq_off; P.n_mat = a->n_mat;
    const bool streamed = a->R > 0 && a->d_mt_work && a->max_len < 65536;
    if (ph & 1) {
    // ---- phase A: the real alignment of every pair ----
    P.n_jobs = wide ? npairs : (npairs + 1) / 2;
    TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
    bool shcol = !wide && a->n_mat == 0 && C == 3 && g_wave_shcol;
    for (int x = 0; x < sc->alphabet && shcol; x++)                    // role swap needs a symmetric substitution matrix
        for (int y = 0; y < x; y++) shcol = shcol && sc->sub[x * MAXA + y] == sc->sub[y * MAXA + x];
    if (shcol) {
        int C2, warps2;
        size_t smem2;
        rc = wave_setup(P, sc, a->max_len, &C2, &warps2, &smem2, false, true, 0, reserve);
        if (rc != TG_OK) return rc;
        int g2, w2;
        wave_spread(P.n_jobs, warps2, sms, &g2, &w2);
        rc = wave_launch(P, JOBS_PHASE_A_SHCOL, 3, g2, w2, smem2, st, false);
    } else {
        int g1, w1;
        wave_spread(P.n_jobs, warps, sms, &g1, &w1);
        rc = wave_launch(P, JOBS_PHASE_A, C, g1, w1, smem, st, wide);
    }
    if (rc != TG_OK) return rc;
    // ---- identity score + early-out flags ----
    FlagParams F;
    F.seq_off = P.seq_off; F.order = P.order; F.n_seq = P.n_seq; F.adjust_lens = P.adjust_lens; F.min_viable = P.min_viable;
    F.R = a->R; F.q0 = a->q0; F.q1 = a->q1; F.aln = a->d_aln; F.diag_cs = a->d_diag_cs; F.match = a->match;
    F.flags = a->d_flags; F.iden = a->d_iden; F.shape = a->d_shape;
    F.mat_pair_off = P.mat_pair_off; F.mat_seq_off = P.mat_seq_off; F.n_mat = P.n_mat;
    F.live = streamed ? reinterpret_cast<int *>(a->d_mt_work) : nullptr;
    F.n_live = reinterpret_cast<int *>(a->d_counter) + 6;
    if (streamed) TG_CUDA_CHECK(cudaMemsetAsync(F.n_live, 0, sizeof(int), st));
    {
        long long g = (npairs + 255) / 256;
        if (g > (long long)sms * 8) g = (long long)sms * 8;
        dist_flags_kernel<<<(int)g, 256, 0, st>>>(F);
        TG_CUDA_CHECK(cudaGetLastError());
    }
    }
    if (a->R <= 0) return TG_OK;
    // ---- null model: MT19937 shuffles of every non-early pair ----
    int stride = (a->max_len + 3) / 4;
    if ((stride & 1) == 0) stride++;
    stride *= 4;
    rc = mt_upload_base();
    if (rc != TG_OK) return rc;
    int *const n_live = reinterpret_cast<int *>(a->d_counter) + 6;
    if (streamed && (ph & 3)) {
        StreamParams S;
        S.pool = a->d_pool; S.seq_off = P.seq_off; S.order = P.order; S.n_seq = P.n_seq; S.adjust_lens = P.adjust_lens;
        S.min_viable = P.min_viable; S.R = a->R; S.q0 = a->q0; S.q1 = a->q1; S.flags = a->d_flags;
        S.shuf_base = a->shuf_base; S.slot = a->slot; S.seed0 = a->seed0; S.counter = P.counter;
        S.mat_pair_off = P.mat_pair_off; S.mat_seq_off = P.mat_seq_off; S.n_mat = P.n_mat;
        S.cap = mt_blocks_needed(2 * a->R * a->max_len) * 624;
        const size_t live_bytes = ((size_t)npairs * sizeof(int) + 255) / 256 * 256;
        S.live = reinterpret_cast<int *>(a->d_mt_work); S.n_live = n_live;
        S.state = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(a->d_mt_work) + live_bytes);
        S.stream = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(a->d_mt_work) + live_bytes + (size_t)npairs * 624 * sizeof(uint32_t));
        S.overflow = reinterpret_cast<int *>(a->d_counter) + 4;      // sticky; the caller zeroes and checks it
        const int wstride = (a->max_len + 15) / 16 * 16 + MTW_EXTRA;      // sampling: per-warp pool + scratch
        const size_t psmem = (size_t)MTW_WARPS * wstride;
        if (psmem > 200 * 1024) return tg_set_err(TG_ERANGE, "max_len %d: the sampling pool does not fit shared memory", a->max_len);
        S.pool_stride = wstride;
        const long long need32 = (npairs + MT_THREADS - 1) / MT_THREADS;
        if (ph & 1) {
        // seeding
        const size_t ssmem = (size_t)624 * MTS_PAD * sizeof(uint32_t);
        TG_CUDA_CHECK(cudaFuncSetAttribute(mt_seed_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
        TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
        {
            long long g = (long long)sms * 2;
            if (g > need32) g = need32;
            mt_seed_batch_kernel<<<(int)g, MT_THREADS, ssmem, st>>>(S);
            TG_CUDA_CHECK(cudaGetLastError());
        }
        }
        if (ph & 2) {
        // stream generation
        TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
        {
            long long g = (long long)sms * 8;
            const long long need = (npairs + MTG_WARPS - 1) / MTG_WARPS;
            if (g > need) g = need;
            mt_stream_batch_kernel<<<(int)g, MTG_WARPS * 32, 0, st>>>(S);
            TG_CUDA_CHECK(cudaGetLastError());
        }
        // sampling
        TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
        {
            TG_CUDA_CHECK(cudaFuncSetAttribute(mt_sample_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            int per_sm = (int)((200 * 1024) / (psmem + 1024));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 8) per_sm = 8;
            long long g = (long long)sms * per_sm;
            const long long need = (npairs + MTW_WARPS - 1) / MTW_WARPS;
            if (g > need) g = need;
            mt_sample_warp_kernel<<<(int)g, MTW_WARPS * 32, psmem, st>>>(S);
            TG_CUDA_CHECK(cudaGetLastError());
        }
        }
    } else if (!streamed && (ph & 2)) {
    BatchShuffleParams S;
    S.pool = a->d_pool; S.seq_off = P.seq_off; S.order = P.order; S.n_seq = P.n_seq; S.adjust_lens = P.adjust_lens;
    S.min_viable = P.min_viable; S.R = a->R; S.q0 = a->q0; S.q1 = a->q1; S.flags = a->d_flags;
    S.shuf_base = a->shuf_base; S.slot = a->slot; S.seed0 = a->seed0; S.counter = P.counter;
    S.mat_pair_off = P.mat_pair_off; S.mat_seq_off = P.mat_seq_off; S.n_mat = P.n_mat;
    size_t msmem = 624 * MT_THREADS * sizeof(uint32_t) + (size_t)MT_THREADS * stride;
    if (msmem > 200 * 1024) { stride = 0; msmem = 624 * MT_THREADS * sizeof(uint32_t); }
    S.pool_stride = stride;
    TG_CUDA_CHECK(cudaFuncSetAttribute(mt_shuffle_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    {
        int per_sm = (int)((227 * 1024) / (msmem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        long long g = (long long)sms * per_sm;
        const long long need = (npairs + MT_THREADS - 1) / MT_THREADS;
        if (g > need) g = need;
        TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
        mt_shuffle_batch_kernel<<<(int)g, MT_THREADS, msmem, st>>>(S);
        TG_CUDA_CHECK(cudaGetLastError());
    }
    }
    if (!(ph & 4)) return TG_OK;
    // ---- phase B: alignments of the shuffled pairs ----
    P.scores = a->d_rnd;
    P.n_jobs = wide ? npairs * a->R : (a->R % 2 == 0) ? npairs * (a->R / 2) : ((npairs + 1) / 2) * a->R;
    TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
    int gb, wb;
    P.seq_mask = nullptr; P.prof_rows = 0;
    if (!wide && C == 3 && a->d_seq_mask && a->prof_rows >= 2 && a->prof_rows < sc->alphabet + 1) {
        // jobs whose row sequences use fewer than prof_rows letters: compact profiles, more resident warps
        int Cc, warps_c;
        size_t smem_c;
        rc = wave_setup(P, sc, a->max_len, &Cc, &warps_c, &smem_c, false, false, a->prof_rows, reserve);
        if (rc != TG_OK) return rc;
        P.seq_mask = a->d_seq_mask; P.prof_rows = a->prof_rows;
        wave_spread(P.n_jobs, warps_c, sms, &gb, &wb);
        rc = wave_launch(P, JOBS_PHASE_B_CPT, 3, gb, wb, smem_c, st, false);
        if (rc != TG_OK) return rc;
        TG_CUDA_CHECK(cudaMemsetAsync(a->d_counter, 0, sizeof(unsigned long long), st));
    }
    wave_spread(P.n_jobs, warps, sms, &gb, &wb);
    return wave_launch(P, JOBS_PHASE_B, C, gb, wb, smem, st, wide);
}

TG_EXPORT int tg_wave_timing(int enable)
{
    g_timing_on = enable != 0;
    return TG_OK;
}

TG_EXPORT int tg_wave_timing_read(double *total_ms, int *n_launches)
{
    double tot = 0.0;
    int n = 0;
    for (auto &ev : g_timing_events) {
        float ms = 0.f;
        TG_CUDA_CHECK(cudaEventSynchronize(ev.second));
        TG_CUDA_CHECK(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
        n++;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    g_timing_events.clear();
    if (total_ms) *total_ms = tot;
    if (n_launches) *n_launches = n;
    return TG_OK;
}

TG_EXPORT int tg_nw_generic_max_m(void) { return 24000; }

TG_EXPORT int tg_nw_scores_generic(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_m,
                                   const tg_nw_scoring *sc, int32_t *d_scores, void *stream)
{
    if (n_jobs <= 0) return TG_OK;
    if (!sc || sc->alphabet < 1 || sc->alphabet > MAXA) return tg_set_err(TG_EINVAL, "bad scoring");
    if (max_m < 1 || max_m > tg_nw_generic_max_m()) return tg_set_err(TG_ERANGE, "max_m %d outside 1..%d", max_m, tg_nw_generic_max_m());
    if (!d_pool || !d_jobs || !d_scores) return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    GenParams P;
    P.pool = d_pool; P.jobs = d_jobs; P.n_items = 2 * n_jobs; P.scores = d_scores;
    P.A = sc->alphabet; P.g = sc->gap; P.gl = sc->gap_left; P.gr = sc->gap_right; P.max_m = max_m;
    for (int i = 0; i < MAXA * MAXA; i++) P.sub[i] = sc->sub[i];
    const size_t smem = 2 * (size_t)(max_m + 1) * sizeof(int32_t);
    TG_CUDA_CHECK(cudaFuncSetAttribute(nw_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((200 * 1024) / (smem + 8192));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)sms * per_sm;
    if (grid > P.n_items) grid = P.n_items;
    nw_generic_kernel<<<(int)grid, GEN_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

static void mt_base_table_host(uint32_t *mt)
{
    mt[0] = 19650218u;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

static int mt_upload_base()
{
    static thread_local int base_uploaded_dev = -1;
    int dev = 0;
    TG_CUDA_CHECK(cudaGetDevice(&dev));
    if (base_uploaded_dev != dev) {
        uint32_t base[624];
        mt_base_table_host(base);
        TG_CUDA_CHECK(cudaMemcpyToSymbol(c_mt_base, base, sizeof(base)));
        base_uploaded_dev = dev;
    }
    return TG_OK;
}

TG_EXPORT int tg_mt_shuffle(uint8_t *d_pool, const tg_shuffle_job *d_jobs, int64_t n_jobs, int R, int max_len, void *stream)
{
    if (n_jobs <= 0 || R <= 0) return TG_OK;
    if (!d_pool || !d_jobs) return tg_set_err(TG_EINVAL, "NULL device pointer");
    if (max_len < 1) return tg_set_err(TG_EINVAL, "max_len < 1");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    // per-thread pool stride: multiple of 4 with odd word count so the 32 lanes start in distinct banks
    int stride = (max_len + 3) / 4;
    if ((stride & 1) == 0) stride++;
    stride *= 4;
    size_t smem = 624 * MT_THREADS * sizeof(uint32_t) + (size_t)MT_THREADS * stride;
    if (smem > 200 * 1024) {          // long sequences: shuffle in place in global memory
        stride = 0;
        smem = 624 * MT_THREADS * sizeof(uint32_t);
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {
        const int rcb = mt_upload_base();
        if (rcb != TG_OK) return rcb;
    }
    TG_CUDA_CHECK(cudaFuncSetAttribute(mt_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)sms * per_sm;
    const long long need = (n_jobs + MT_THREADS - 1) / MT_THREADS;
    if (grid > need) grid = need;
    mt_shuffle_kernel<<<(int)grid, MT_THREADS, smem, st>>>(d_pool, d_jobs, n_jobs, R, stride);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_dist_epilogue(const tg_dist_epilogue_args *a, void *stream)
{
    if (!a) return tg_set_err(TG_EINVAL, "NULL argument");
    const long long npairs = a->q1 - a->q0;
    if (npairs <= 0) return TG_OK;
    if (!a->d_seq_off || !a->d_order || !a->d_aln || !a->d_flags || !a->d_iden || !a->d_matrix || !a->d_stats ||
        (a->R > 0 && !a->d_rnd) || (a->ambig_cap > 0 && !a->d_ambig))
        return tg_set_err(TG_EINVAL, "NULL device pointer");
    if (a->margin_log2 < 8 || a->margin_log2 > 52) return tg_set_err(TG_EINVAL, "margin_log2 %d outside [8, 52]", a->margin_log2);
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    EpiParams E;
    E.seq_off = reinterpret_cast<const long long *>(a->d_seq_off); E.order = a->d_order; E.n_seq = a->n_seq;
    E.adjust_lens = a->adjust_lens; E.min_viable = a->min_viable; E.R = a->R > 0 ? a->R : 0;
    E.q0 = a->q0; E.q1 = a->q1; E.ambig_base = a->ambig_base;
    E.aln = a->d_aln; E.rnd = a->d_rnd; E.flags = a->d_flags; E.iden = a->d_iden; E.matrix = a->d_matrix;
    E.ambig = reinterpret_cast<long long *>(a->d_ambig); E.stats = a->d_stats; E.ambig_cap = a->ambig_cap;
    E.eps = ldexp(1.0, -a->margin_log2);
    if (a->n_mat < 0 || (a->n_mat > 0 && (!a->d_mat_pair_off || !a->d_mat_seq_off || !a->d_mat_out_off))) return tg_set_err(TG_EINVAL, "bad matrix table");
    E.mat_pair_off = reinterpret_cast<const long long *>(a->d_mat_pair_off); E.mat_seq_off = a->d_mat_seq_off; E.n_mat = a->n_mat;
    E.mat_out_off = reinterpret_cast<const long long *>(a->d_mat_out_off);
    long long g = (npairs + 255) / 256;
    if (g > (long long)sms * 8) g = (long long)sms * 8;
    dist_epilogue_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(E);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_int_alu_peak(int iters, double *dpx_lane_ops_per_s, double *iadd_lane_ops_per_s)
{
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    const int grid = sms * 8, block = 256;
    uint32_t *d_out = nullptr;
    TG_CUDA_CHECK(cudaMalloc(&d_out, (size_t)grid * block * sizeof(uint32_t)));
    cudaEvent_t e0, e1;
    TG_CUDA_CHECK(cudaEventCreate(&e0));
    TG_CUDA_CHECK(cudaEventCreate(&e1));
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            TG_CUDA_CHECK(cudaEventRecord(e0));
            if (mode == 0) int_peak_kernel<0><<<grid, block>>>(d_out, iters, 12345u + rep);
            else int_peak_kernel<1><<<grid, block>>>(d_out, iters, 12345u + rep);
            TG_CUDA_CHECK(cudaEventRecord(e1));
            TG_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            TG_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        // per thread per iteration: 8 * 8 DPX ops (mode 0) or 8 * 8 (add + DPX) pairs (mode 1)
        const double ops = (double)grid * block * (double)iters * 64.0;
        res[mode] = ops / (best * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (dpx_lane_ops_per_s) *dpx_lane_ops_per_s = res[0];
    if (iadd_lane_ops_per_s) *iadd_lane_ops_per_s = res[1];
    return TG_OK;
}
