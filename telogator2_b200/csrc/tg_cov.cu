// Coverage scan for sm_100a: get_telomere_base_count (tg_kmer.py:93-102) and get_telomere_kmer_density
// (tg_kmer.py:66-90) for TWO k-mer lists in one pass over the packed reads (HBM bound by design).
//
// One thread owns 64 consecutive bases: one 128-bit load of the 2-bit plane, one 64-bit load of the N plane, then one
// shared-memory table lookup per base.  The 32-bit table entry, indexed by the next 7 bases, holds for each of the two
// lists the coverage mask (1 << L) - 1 of the longest plain k-mer starting here and a few flag bits for the k-mers a
// 7-base window cannot decide (see below).  The entry is shifted by the base's position within its group of 8 and OR-ed
// into three accumulators (coverage masks, flags) with one LOP3 each, so the union coverage cov[p] = "some selected match
// covers p" and the flag planes fall out as bit planes without any per-match loop.
//
// k-mers a window cannot decide ("special"):
//   * self-overlapping k-mers that fit the window (CCCCCC, CCCTCC): the entry flags their exact occurrences (S1).
//     re.finditer's leftmost-non-overlapping rule only matters where two occurrences of one k-mer are closer than its
//     length; elsewhere "all occurrences" is right and their coverage is the S1 plane smeared over the k-mer length.
//     Close occurrences inside one thread's chunk are resolved by that thread, across chunks by the warp (rare).
//   * k-mers longer than the window: the entry flags "its first 7 bases start here" (S2) and "bases 8.. of some long k-mer
//     start here" (T); S2(p) & T(p + 7) leaves few candidates, which are compared against the staged planes.
//   * special k-mers whose every base is covered by plain k-mers occurring inside them (the HOR k-mers) never change the
//     union coverage and are dropped.
//   * windows cut short by an N (read ends, mostly) use small tables indexed by the d < 7 bases in front of the N.
// One warp = one group: it works on its own tiles (2048 bases with halos) and only ever synchronises with itself.
// MODE_DENSITY turns the coverage plane into the uint8 window counts cnt[i] = |cov (i, i + w]| with a sliding difference
// of the entering and leaving coverage bits, eight outputs per step through a 256-entry prefix-popcount table.
#include "tg_scan_common.cuh"
#include <string.h>

namespace {

constexpr int32_t COV_MAGIC = 0x43563037;             // "CV07"
constexpr int COV_SHORT_N = 5464;                     // sum of 4^d, d = 1 .. 6, rounded up to a multiple of 4
__host__ __device__ constexpr int COV_SHORT_OFF(int d) { return ((1 << (2 * d)) - 4) / 3; }

// table entry bits (list A; list B = + 16 for the mask and S1)
constexpr uint32_t E_S1A = 1u << 7, E_T = 1u << 8, E_S2A = 1u << 15, E_S1B = 1u << 23, E_S2B = 1u << 24;

struct CovTable {
    int32_t magic, n_all, KA, n_bord, max_len, pad_;
    int32_t bord_len[2];               // per list: L > 0: its window-sized self-overlapping k-mers all have length L (smear);
                                       // 0: it has none; -1: several lengths (per-position handling)
    uint64_t pat[SC_MAXK];
    int32_t len[SC_MAXK];
    int32_t kind[SC_MAXK];             // SC_PLAIN / SC_SPECIAL / SC_BORDERED / SC_REDUNDANT; k-mers of list A come first
    int32_t bord_idx[SC_MAXBORD];      // the n_bord self-overlapping k-mers that are not redundant
    // windows cut short by an N: tab_short[COV_SHORT_OFF(d) + window of d bases], d = 1 .. 6 bases in front of the N, same
    // entry format (only k-mers of length <= d can match there)
    uint32_t tab_short[COV_SHORT_N];
    uint32_t tab[SC_TAB];
};

constexpr int COV_THREADS = 448;                      // threads per CTA: they share one copy of the lookup table (2 CTAs / SM)
constexpr int COV_GROUPS = COV_THREADS / 32;          // one warp = one group
constexpr int COV_CH = 64;                            // bases per thread
constexpr int COV_REGION = 32 * COV_CH;               // 2048 bases per tile (64 left halo, window right halo)
constexpr int COV_WORDS = COV_REGION / 32;
enum { MODE_COUNT = 0, MODE_DENSITY = 1 };

struct CovParams {
    ScPlanes pl;
    const int32_t *blk_read;           // read index of every 128-base block (MODE_COUNT)
    long long n_tiles;
    int out_chunks;                    // 64-base chunks of output per tile
    const CovTable *tab;
    int window;
    uint8_t *cnt[2];
    int32_t *counts;
};

struct alignas(16) CovGroupSmem {
    uint32_t cov[2][COV_WORDS + 4];
    uint32_t spill[2][32 + 4];
    uint32_t pk[COV_REGION / 16 + 8];          // the staged planes (slow paths read them back)
    uint32_t nm[COV_WORDS + 4];
    uint32_t defer[2], pad_[2];                // per list: chunks that need the warp-wide greedy walk, one bit per chunk
};

struct alignas(16) CovSmem {
    uint32_t tab[SC_TAB];
    uint64_t pat[SC_MAXK];
    uint8_t len[SC_MAXK], kind[SC_MAXK];
    uint8_t bsk[2][SC_MAXLIST], lgk[2][SC_MAXLIST];   // per list: window-sized self-overlapping k-mers / long k-mers
    int32_t n_bsk[2], n_lgk[2];
    uint2 lut8[256];                           // inclusive prefix popcounts of the 8 bits of a byte, one byte each
    CovGroupSmem g[COV_GROUPS];
};

__device__ __noinline__ void cov_or(uint32_t *cv, int x, int len)
{
    const uint64_t m = ((len >= 32) ? 0xffffffffull : ((1ull << len) - 1ull)) << (x & 31);
    atomicOr(&cv[x >> 5], (uint32_t)m);
    if (m >> 32) atomicOr(&cv[(x >> 5) + 1], (uint32_t)(m >> 32));
}

// popcount of bits [lo, hi] (inclusive) of a bit plane
__device__ __forceinline__ int range_popc(const uint32_t *cv, int lo, int hi)
{
    int cnt = 0;
    for (int wd = lo >> 5; wd <= (hi >> 5); wd++) {
        uint32_t v = cv[wd];
        if (wd == (lo >> 5)) v &= 0xffffffffu << (lo & 31);
        if (wd == (hi >> 5)) v &= 0xffffffffu >> (31 - (hi & 31));
        cnt += __popc(v);
    }
    return cnt;
}

// table entry of the window that starts `sh2` bits into the 64-bit register pair (lo, hi); the byte offset of the
// entry (window * 4) comes straight out of the shift
__device__ __forceinline__ uint32_t tab_at(const uint32_t *s_tab, uint32_t lo, uint32_t hi, int sh2)
{
    const uint32_t t4 = ((sh2 >= 2) ? __funnelshift_r(lo, hi, sh2 - 2) : (lo << (2 - sh2))) & (uint32_t)((SC_TAB - 1) * 4);
    return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_tab) + t4);
}

// the 32 bases that start at region position x, from the staged plane
__device__ __forceinline__ uint64_t sm_bases64(const uint32_t *pk, int x)
{
    const int pw = x >> 4, sh = (x & 15) * 2;
    uint64_t bases = ((uint64_t)pk[pw] | ((uint64_t)pk[pw + 1] << 32)) >> sh;
    if (sh) bases |= (uint64_t)pk[pw + 2] << (64 - sh);
    return bases;
}

// does k-mer (pat, len) start at region position x?  (staged planes; N anywhere in the span = no)
__device__ __noinline__ bool sm_kmer_at(const uint32_t *pk, const uint32_t *nm, int x, uint64_t pat, int len)
{
    const uint32_t nbits = __funnelshift_r(nm[x >> 5], nm[(x >> 5) + 1], x & 31);
    const uint32_t lm = (len >= 32) ? 0xffffffffu : ((1u << len) - 1u);
    const uint64_t m = (len >= 32) ? ~0ull : ((1ull << (2 * len)) - 1ull);
    return (nbits & lm) == 0u && ((sm_bases64(pk, x) ^ pat) & m) == 0ull;
}

// any flag within (bit - d, bit + d) except `bit` itself?  X: the 64 flags of my chunk, prev / next: the neighbouring words
__device__ __forceinline__ bool flags_near(uint64_t X, uint32_t prev, uint32_t next, int bit, int d)
{
    // 128-bit window [prev][X][next], my bit at index 32 + bit; d <= 32
    const int lo_i = 32 + bit - (d - 1), hi_i = 32 + bit + (d - 1);
    bool hit = false;
    for (int i = lo_i; i <= hi_i; i++) {
        if (i == 32 + bit) continue;
        const uint32_t v = (i < 32) ? (prev >> i) : (i < 96) ? (uint32_t)(X >> (i - 32)) : (next >> (i - 96));
        hit = hit || (v & 1u);
    }
    return hit;
}

// Flags X of my chunk = exact occurrences of self-overlapping k-mers of length L (prev / next: the flag words of the
// neighbouring chunks).  Returns 0: no two flags closer than L -- the coverage of my occurrences (every flag smeared over
// L positions) was OR-ed into cv[0..2];  1: close flags, all inside my chunk (the caller runs the greedy rule over them);
// 2: close flags across a chunk border (the chunk goes to the warp-wide greedy walk).
template <int LS>                             // LS = L known at compile time, or 0: use the argument
__device__ __forceinline__ int smear_flags(uint64_t X, uint32_t prev, uint32_t next, uint32_t *cv, int l_arg)
{
    const int L = LS ? LS : l_arg;
    const uint32_t pv = prev >> (32 - (L - 1)), nx = next & ((1u << (L - 1)) - 1u);
    uint64_t near_own = 0ull;
#pragma unroll
    for (int k = 1; k < L; k++) near_own |= X & (X >> k);
    if (pv | nx) {
        // V = [L-1 flags of the previous chunk][my 64][L-1 flags of the next chunk], 64 + 2 (L - 1) <= 76 bits
        const uint64_t vlo = (X << (L - 1)) | pv;
        const uint32_t vhi = (uint32_t)(X >> (64 - (L - 1))) | (nx << (L - 1));
        uint64_t nl = 0ull;
        uint32_t nh = 0u;
#pragma unroll
        for (int k = 1; k < L; k++) {
            nl |= vlo & ((vlo >> k) | ((uint64_t)vhi << (64 - k)));
            nh |= vhi & (vhi >> k);
        }
        // pairs among my own flags show up here too; a pair that involves a neighbour's flag forces the warp-wide walk
        // (the largest of my flags close to a flag of the next chunk is never the lower end of a pair of my own)
        const uint64_t own_lo = near_own << (L - 1);
        const uint32_t own_hi = (uint32_t)(near_own >> (64 - (L - 1)));
        if ((nl & ~own_lo) | (uint64_t)(nh & ~own_hi)) return 2;
    }
    if (near_own) return 1;
    uint64_t B = X;
    uint32_t over = 0u;
#pragma unroll
    for (int k = 1; k < L; k++) { B |= X << k; over |= (uint32_t)(X >> (64 - k)); }
    if ((uint32_t)B) atomicOr(&cv[0], (uint32_t)B);
    if (B >> 32) atomicOr(&cv[1], (uint32_t)(B >> 32));
    if (over) atomicOr(&cv[2], over);
    return 0;
}

// four bytes (bits sh .. sh+7 of x0..x3) -> one word
__device__ __forceinline__ uint32_t gather4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, int sh)
{
    return ((x0 >> sh) & 0xffu) | (((x1 >> sh) & 0xffu) << 8) | (((x2 >> sh) & 0xffu) << 16) | ((x3 >> sh) << 24);
}

template <int MODE>
__global__ void __launch_bounds__(COV_THREADS, 2) scan_cov_kernel(const CovParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CovSmem &S = *reinterpret_cast<CovSmem *>(smem_raw);
    const int grp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CovGroupSmem &G = S.g[grp];
    const CovTable *__restrict__ T = P.tab;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T->tab);
        uint4 *dst = reinterpret_cast<uint4 *>(S.tab);
        for (int i = threadIdx.x; i < SC_TAB / 4; i += COV_THREADS) dst[i] = src[i];
        if (threadIdx.x < SC_MAXK) {
            const int k = threadIdx.x;
            S.pat[k] = T->pat[k]; S.len[k] = (uint8_t)T->len[k]; S.kind[k] = (uint8_t)T->kind[k];
        }
        if (threadIdx.x < 2) {
            int nb = 0, nl = 0;
            for (int k = threadIdx.x ? T->KA : 0; k < (threadIdx.x ? T->n_all : T->KA); k++) {
                if (T->kind[k] != SC_SPECIAL && T->kind[k] != SC_BORDERED) continue;
                if (T->len[k] > SC_W) S.lgk[threadIdx.x][nl++] = (uint8_t)k;
                else S.bsk[threadIdx.x][nb++] = (uint8_t)k;
            }
            S.n_bsk[threadIdx.x] = nb; S.n_lgk[threadIdx.x] = nl;
        }
        if (threadIdx.x < 256) {
            uint32_t lo = 0, hi = 0;
            for (int j = 0; j < 4; j++) {
                lo |= (uint32_t)__popc(threadIdx.x & ((2 << j) - 1)) << (8 * j);
                hi |= (uint32_t)__popc(threadIdx.x & ((32 << j) - 1)) << (8 * j);
            }
            S.lut8[threadIdx.x] = make_uint2(lo, hi);
        }
    }
    __syncthreads();
    const int n_bord = T->n_bord, KA = T->KA;
    const int bord_len[2] = {T->bord_len[0], T->bord_len[1]};
    const int w = P.window;
    const int64_t total = P.pl.total;

    for (long long tile = (long long)blockIdx.x * COV_GROUPS + grp; tile < P.n_tiles; tile += (long long)gridDim.x * COV_GROUPS) {
        const int64_t a0 = tile * (int64_t)P.out_chunks * COV_CH - COV_CH;     // absolute base of region index 0
        const int64_t a = a0 + (int64_t)lane * COV_CH;
        uint32_t s1w[2][2] = {{0u, 0u}, {0u, 0u}}, s2w[2][2] = {{0u, 0u}, {0u, 0u}}, trw[2] = {0u, 0u};   // my flag planes
        __syncwarp();
        // ---- phase 1: table lookups of my 64 bases -> coverage words, spill into the next chunk, flag planes ----
        {
            uint32_t cvw[2][2] = {{0u, 0u}, {0u, 0u}}, sp[2] = {0u, 0u};
            uint32_t pk[5] = {0u, 0u, 0u, 0u, 0u}, nmw[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
            bool slow = false;
            if (a >= 0 && a < total) {
                const uint4 pv = __ldg(reinterpret_cast<const uint4 *>(P.pl.pack2 + (a >> 4)));
                const uint2 nv = __ldg(reinterpret_cast<const uint2 *>(P.pl.nmask + (a >> 5)));
                pk[0] = pv.x; pk[1] = pv.y; pk[2] = pv.z; pk[3] = pv.w; pk[4] = __ldg(P.pl.pack2 + (a >> 4) + 4);
                nmw[0] = nv.x; nmw[1] = nv.y;
                if (a + COV_CH < total) nmw[2] = __ldg(P.pl.nmask + (a >> 5) + 2);
                if ((nmw[0] | nmw[1] | (nmw[2] & ((1u << (SC_W - 1)) - 1u))) == 0u) {
                    // fast path: no N within reach of any of my windows.  16 bases per iteration (kept rolled: the loop body
                    // must stay in the instruction cache, every warp runs at its own pace)
                    uint64_t c64[2] = {0ull, 0ull}, s1_64[2] = {0ull, 0ull}, s2_64[2] = {0ull, 0ull}, tr64 = 0ull;
                    uint32_t w0 = pk[0], w1 = pk[1], w2 = pk[2], w3 = pk[3], w4 = pk[4];
#pragma unroll 1
                    for (int it = 0; it < 4; it++) {
                        uint32_t acc[2] = {0u, 0u}, f1[2] = {0u, 0u}, f2[2] = {0u, 0u};
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const int r = q & 7;
                            const uint32_t e = tab_at(S.tab, w0, w1, 2 * q);
                            const uint32_t t = e << r;
                            acc[q >> 3] |= t & (0x007f007fu << r);                // coverage masks: bits 0-6 / 16-22
                            f1[q >> 3] |= t & ((E_S1A | E_S2A | E_S1B) << r);     // three flag kinds, 8 bits apart
                            f2[q >> 3] |= t & ((E_T | E_S2B) << r);
                        }
                        const int sh = 16 * it;
                        c64[0] |= (uint64_t)((acc[0] & 0xffffu) | ((acc[1] & 0xffffu) << 8)) << sh;
                        c64[1] |= (uint64_t)((acc[0] >> 16) | ((acc[1] >> 16) << 8)) << sh;
                        if (it == 3) {
                            sp[0] = (acc[1] & 0xffffu) >> 8;
                            sp[1] = acc[1] >> 24;
                        }
                        if ((f1[0] | f1[1]) != 0u) {
                            s1_64[0] |= (uint64_t)(((f1[0] >> 7) & 0xffu) | (((f1[1] >> 7) & 0xffu) << 8)) << sh;
                            s2_64[0] |= (uint64_t)(((f1[0] >> 15) & 0xffu) | (((f1[1] >> 15) & 0xffu) << 8)) << sh;
                            s1_64[1] |= (uint64_t)(((f1[0] >> 23) & 0xffu) | (((f1[1] >> 23) & 0xffu) << 8)) << sh;
                        }
                        if ((f2[0] | f2[1]) != 0u) {
                            tr64 |= (uint64_t)(((f2[0] >> 8) & 0xffu) | (((f2[1] >> 8) & 0xffu) << 8)) << sh;
                            s2_64[1] |= (uint64_t)((f2[0] >> 24) | ((f2[1] >> 24) << 8)) << sh;
                        }
                        w0 = w1; w1 = w2; w2 = w3; w3 = w4;
                    }
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        cvw[g][0] = (uint32_t)c64[g]; cvw[g][1] = (uint32_t)(c64[g] >> 32);
                        s1w[g][0] = (uint32_t)s1_64[g]; s1w[g][1] = (uint32_t)(s1_64[g] >> 32);
                        s2w[g][0] = (uint32_t)s2_64[g]; s2w[g][1] = (uint32_t)(s2_64[g] >> 32);
                    }
                    trw[0] = (uint32_t)tr64; trw[1] = (uint32_t)(tr64 >> 32);
                } else slow = (nmw[0] & nmw[1]) != 0xffffffffu;           // some N within reach (all N: nothing to do)
            }
            // lanes with an N within reach (read ends, mostly), one at a time, all 32 lanes helping: lane j looks at bases
            // j and j + 32 of that lane's chunk; a window cut short by an N after d < 7 bases goes to the table of d-base windows
            uint32_t need = __ballot_sync(TG_FULL_MASK, slow);
            while (need) {
                const int l = __ffs(need) - 1;
                need &= need - 1;
                const uint32_t b0 = __shfl_sync(TG_FULL_MASK, pk[0], l), b1 = __shfl_sync(TG_FULL_MASK, pk[1], l),
                               b2 = __shfl_sync(TG_FULL_MASK, pk[2], l), b3 = __shfl_sync(TG_FULL_MASK, pk[3], l),
                               b4 = __shfl_sync(TG_FULL_MASK, pk[4], l);
                const uint32_t m0 = __shfl_sync(TG_FULL_MASK, nmw[0], l), m1 = __shfl_sync(TG_FULL_MASK, nmw[1], l),
                               m2 = __shfl_sync(TG_FULL_MASK, nmw[2], l);
                const bool lowhalf = lane < 16;
                const int over = lane > 32 - SC_W ? 32 - lane : 32;       // bits of my mask that stay in this word
                uint32_t carry[2] = {0u, 0u};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t lo = h ? (lowhalf ? b2 : b3) : (lowhalf ? b0 : b1);
                    const uint32_t hi = h ? (lowhalf ? b3 : b4) : (lowhalf ? b1 : b2);
                    const uint32_t nwin = __funnelshift_r(h ? m1 : m0, h ? m2 : m1, lane) & ((1u << SC_W) - 1u);
                    uint32_t e = 0u;
                    if (nwin == 0u) e = tab_at(S.tab, lo, hi, 2 * (lane & 15));
                    else if (!(nwin & 1u)) {
                        const int d = __ffs(nwin) - 1;                    // bases in front of the N: 1 .. 6
                        const uint32_t wb = __funnelshift_r(lo, hi, 2 * (lane & 15)) & ((1u << (2 * d)) - 1u);
                        e = __ldg(T->tab_short + COV_SHORT_OFF(d) + wb);
                    }
                    const uint32_t fa = __ballot_sync(TG_FULL_MASK, e & E_S1A), fb = __ballot_sync(TG_FULL_MASK, e & E_S1B),
                                   la = __ballot_sync(TG_FULL_MASK, e & E_S2A), lb = __ballot_sync(TG_FULL_MASK, e & E_S2B),
                                   tt = __ballot_sync(TG_FULL_MASK, e & E_T);
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        const uint32_t mk = (e >> (16 * g)) & 0x7fu;
                        const uint32_t wd = __reduce_or_sync(TG_FULL_MASK, mk << lane) | carry[g];
                        carry[g] = __reduce_or_sync(TG_FULL_MASK, over < 32 ? mk >> over : 0u);
                        if (lane == l) { cvw[g][h] = wd; s1w[g][h] = g ? fb : fa; s2w[g][h] = g ? lb : la; }
                    }
                    if (lane == l) trw[h] = tt;
                }
                if (lane == l) { sp[0] = carry[0]; sp[1] = carry[1]; }
            }
            *reinterpret_cast<uint4 *>(&G.pk[4 * lane]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint2 *>(&G.nm[2 * lane]) = make_uint2(nmw[0], nmw[1]);
#pragma unroll
            for (int g = 0; g < 2; g++) {
                *reinterpret_cast<uint2 *>(&G.cov[g][2 * lane]) = make_uint2(cvw[g][0], cvw[g][1]);
                G.spill[g][lane + 1] = sp[g];
            }
            if (lane < 4) {
                G.cov[0][COV_WORDS + lane] = 0u; G.cov[1][COV_WORDS + lane] = 0u;
                // planes behind the region (a long k-mer that starts in the last chunk is compared against them)
                const int64_t ax = a0 + COV_REGION + 32 * (int64_t)lane;
                const bool ok = ax >= 0 && ax < total;
                G.nm[COV_WORDS + lane] = ok ? __ldg(P.pl.nmask + (ax >> 5)) : 0xffffffffu;
                G.pk[COV_REGION / 16 + 2 * lane] = ok ? __ldg(P.pl.pack2 + (ax >> 4)) : 0u;
                G.pk[COV_REGION / 16 + 2 * lane + 1] = ok ? __ldg(P.pl.pack2 + (ax >> 4) + 1) : 0u;
            }
            if (lane < 2) { G.spill[lane][0] = 0u; G.defer[lane] = 0u; }
        }
        __syncwarp();
        // ---- phase 2a: spill of the previous chunk; special k-mers that start in my chunk ----
        bool deferred = false;
        const uint32_t tr_next = __shfl_down_sync(TG_FULL_MASK, trw[0], 1);
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const uint32_t s_in = G.spill[g][lane];
            if (s_in) atomicOr(&G.cov[g][2 * lane], s_in);
            // flag words of the neighbouring chunks.  What lies in front of the region is unknown: a flag at its very start
            // may be overlapped from there, and only the warp-wide walk looks back -- all ones force it.
            uint32_t prev = __shfl_up_sync(TG_FULL_MASK, s1w[g][1], 1), next = __shfl_down_sync(TG_FULL_MASK, s1w[g][0], 1);
            if (lane == 0) prev = (a0 > 0) ? 0xffffffffu : 0u;
            if (lane == 31) next = 0u;
            const uint64_t X = (uint64_t)s1w[g][0] | ((uint64_t)s1w[g][1] << 32);
            if (X != 0ull) {
                const int L = bord_len[g];
                if (L > 0) {
                    const int conflict = (L == 6) ? smear_flags<6>(X, prev, next, &G.cov[g][2 * lane], 6)
                                                  : smear_flags<0>(X, prev, next, &G.cov[g][2 * lane], L);
                    if (conflict == 2) {
                        atomicOr(&G.defer[g], 1u << lane);
                        deferred = true;
                    } else if (conflict == 1) {
                        // close flags, all mine: the greedy rule per k-mer, left to right
                        uint64_t B = 0ull;
                        uint32_t over = 0u;
                        const uint64_t lm = (1ull << L) - 1ull, bm = (1ull << (2 * L)) - 1ull;
                        for (int q = 0; q < S.n_bsk[g]; q++) {
                            const uint64_t pat = S.pat[S.bsk[g][q]];
                            int blocked = -1;
                            uint64_t rest = X;
                            while (rest) {
                                const int bit = __ffsll((long long)rest) - 1;
                                rest &= rest - 1;
                                if (bit < blocked || ((sm_bases64(G.pk, lane * COV_CH + bit) ^ pat) & bm) != 0ull) continue;
                                blocked = bit + L;
                                B |= lm << bit;
                                if (bit + L > 64) over |= (uint32_t)(lm >> (64 - bit));
                            }
                        }
                        if ((uint32_t)B) atomicOr(&G.cov[g][2 * lane], (uint32_t)B);
                        if (B >> 32) atomicOr(&G.cov[g][2 * lane + 1], (uint32_t)(B >> 32));
                        if (over) atomicOr(&G.cov[g][2 * lane + 2], over);
                    }
                } else {
                    // window-sized self-overlapping k-mers of several lengths: per occurrence
                    uint64_t rest = X;
                    while (rest) {
                        const int bit = __ffsll((long long)rest) - 1;
                        rest &= rest - 1;
                        for (int q = 0; q < S.n_bsk[g]; q++) {
                            const int k = S.bsk[g][q], len = S.len[k];
                            if (!sm_kmer_at(G.pk, G.nm, lane * COV_CH + bit, S.pat[k], len)) continue;
                            if (flags_near(X, prev, next, bit, SC_W)) {
                                atomicOr(&G.defer[g], 1u << lane);
                                deferred = true;
                            } else cov_or(G.cov[g], lane * COV_CH + bit, len);
                        }
                    }
                }
            }
            // long k-mers: their first 7 bases start at p and the continuation of some long k-mer at p + 7
            const uint64_t S2 = (uint64_t)s2w[g][0] | ((uint64_t)s2w[g][1] << 32);
            if (S2 != 0ull) {
                const uint64_t TR = (uint64_t)trw[0] | ((uint64_t)trw[1] << 32);
                uint64_t cand = S2 & ((TR >> SC_W) | ((uint64_t)(lane == 31 ? 0u : tr_next) << (64 - SC_W)));
                while (cand) {
                    const int x = lane * COV_CH + __ffsll((long long)cand) - 1;
                    cand &= cand - 1;
                    for (int q = 0; q < S.n_lgk[g]; q++) {
                        const int k = S.lgk[g][q], len = S.len[k];
                        if (!sm_kmer_at(G.pk, G.nm, x, S.pat[k], len)) continue;
                        if (S.kind[k] == SC_BORDERED) {                       // long AND self-overlapping: always the greedy walk
                            atomicOr(&G.defer[g], 1u << lane);
                            deferred = true;
                        } else cov_or(G.cov[g], x, len);
                    }
                }
            }
        }
        if (__any_sync(TG_FULL_MASK, deferred)) {
            // ---- phase 2b (rare): the warp walks the candidates of every self-overlapping k-mer in the deferred chunks, left
            //      to right, with re.finditer's rule.  A candidate outside the deferred chunks cannot overlap one inside. ----
            __syncwarp();
            for (int b = 0; b < n_bord; b++) {
                const int k = T->bord_idx[b];
                const int g = (k >= KA) ? 1 : 0;
                const uint32_t dmask = G.defer[g];
                if (!dmask) continue;
                const int blen = S.len[k];
                const uint64_t pat = S.pat[k];
                // lane c verifies the candidates of chunk c
                uint64_t mine = 0ull;
                if ((dmask >> lane) & 1u) {
                    uint64_t fw = (blen > SC_W) ? ((uint64_t)(g ? s2w[1][0] : s2w[0][0]) | ((uint64_t)(g ? s2w[1][1] : s2w[0][1]) << 32))
                                                : ((uint64_t)(g ? s1w[1][0] : s1w[0][0]) | ((uint64_t)(g ? s1w[1][1] : s1w[0][1]) << 32));
                    while (fw) {
                        const int bit = __ffsll((long long)fw) - 1;
                        fw &= fw - 1;
                        if (sm_kmer_at(G.pk, G.nm, lane * COV_CH + bit, pat, blen)) mine |= 1ull << bit;
                    }
                }
                uint32_t nz = __ballot_sync(TG_FULL_MASK, mine != 0ull);
                int blocked = -0x40000000;
                bool first = true;
                while (nz) {
                    const int l = __ffs(nz) - 1;
                    nz &= nz - 1;
                    uint64_t word = __shfl_sync(TG_FULL_MASK, mine, l);
                    while (word) {
                        const int x = l * COV_CH + __ffsll((long long)word) - 1;
                        word &= word - 1;
                        if (first) {
                            first = false;
                            if (x < blen - 1 && a0 + x > 0) {             // the chain may reach back beyond the region
                                const int64_t bl = sc_chain_state_before(P.pl, a0 + x, pat, blen);
                                blocked = (bl < a0) ? -0x40000000 : (int)(bl - a0);
                            }
                        }
                        if (x >= blocked) {
                            blocked = x + blen;
                            if (lane == 0) cov_or(G.cov[g], x, blen);
                        }
                    }
                }
            }
        }
        __syncwarp();
        // ---- phase 3: outputs of chunks 1 .. out_chunks ----
        const bool act = lane >= 1 && lane <= P.out_chunks && a < total;
        if (MODE == MODE_COUNT) {
            // popcount of my two coverage words, one atomic per warp when the warp lies inside one read
            const int c0 = act ? __popc(G.cov[0][2 * lane]) + __popc(G.cov[0][2 * lane + 1]) : 0;
            const int c1 = act ? __popc(G.cov[1][2 * lane]) + __popc(G.cov[1][2 * lane + 1]) : 0;
            const unsigned r1 = act ? (unsigned)P.blk_read[a >> 7] + 1u : 0u;
            const unsigned rmax = __reduce_max_sync(TG_FULL_MASK, r1);
            const unsigned rmin = __reduce_min_sync(TG_FULL_MASK, act ? r1 : 0xffffffffu);
            if (rmax != 0u) {
                if (rmin == rmax) {
                    const int s0 = __reduce_add_sync(TG_FULL_MASK, c0), s1 = __reduce_add_sync(TG_FULL_MASK, c1);
                    if (lane == 0) {
                        if (s0) atomicAdd(&P.counts[2 * (rmax - 1u)], s0);
                        if (s1) atomicAdd(&P.counts[2 * (rmax - 1u) + 1], s1);
                    }
                } else if (act) {
                    if (c0) atomicAdd(&P.counts[2 * (r1 - 1u)], c0);
                    if (c1) atomicAdd(&P.counts[2 * (r1 - 1u) + 1], c1);
                }
            }
        } else if (act) {
            const int j0 = lane * COV_CH;
            const char *lut = reinterpret_cast<const char *>(S.lut8);
#pragma unroll 1
            for (int g = 0; g < 2; g++) {
                const uint32_t *cv = G.cov[g];
                const int cnt = range_popc(cv, j0 + 1, j0 + w);
                const uint32_t a0w = cv[2 * lane], a1w = cv[2 * lane + 1], a2w = cv[2 * lane + 2];
                uint64_t sa = (uint64_t)__funnelshift_r(a0w, a1w, 1) | ((uint64_t)__funnelshift_r(a1w, a2w, 1) << 32);    // leaving bits
                const int jb = j0 + w + 1;
                const uint32_t b0w = cv[jb >> 5], b1w = cv[(jb >> 5) + 1], b2w = cv[(jb >> 5) + 2];
                uint64_t sb = (uint64_t)__funnelshift_r(b0w, b1w, jb & 31) | ((uint64_t)__funnelshift_r(b1w, b2w, jb & 31) << 32);  // entering
                uint32_t cnt4 = (uint32_t)cnt * 0x01010101u;
                uint8_t *dst = P.cnt[g] + a;
#pragma unroll 1
                for (int q = 0; q < 4; q++) {
                    uint32_t o[4];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        // 8 outputs per step: prefix popcounts of the entering and of the leaving byte
                        const uint32_t bi = ((uint32_t)sb >> (8 * h)) & 0xffu, bo = ((uint32_t)sa >> (8 * h)) & 0xffu;
                        const uint2 A = *reinterpret_cast<const uint2 *>(lut + bi * 8), B = *reinterpret_cast<const uint2 *>(lut + bo * 8);
                        const uint32_t lo = cnt4 + A.x - B.x;              // counts of outputs + 1 .. + 4 (exact mod 2^32)
                        const uint32_t hi = cnt4 + A.y - B.y;              // counts of outputs + 5 .. + 8
                        o[2 * h] = __byte_perm(lo, cnt4, 0x2104);
                        o[2 * h + 1] = __byte_perm(hi, lo, 0x2107);
                        cnt4 = __byte_perm(hi, 0u, 0x3333);
                    }
                    *reinterpret_cast<uint4 *>(dst + 16 * q) = make_uint4(o[0], o[1], o[2], o[3]);
                    sa >>= 16; sb >>= 16;
                }
            }
        }
    }
}

}  // namespace

// ================================================================================================
// host
// ================================================================================================
TG_EXPORT size_t tg_kmer_table_pair_bytes(void) { return sizeof(CovTable); }

TG_EXPORT int tg_kmer_table_build_pair(const char *kmers_a, const int32_t *off_a, int KA,
                                       const char *kmers_b, const int32_t *off_b, int KB, void *h_table)
{
    if (!h_table) return tg_set_err(TG_EINVAL, "NULL argument");
    std::vector<std::string> km;
    int rc = sc_collect_kmers(kmers_a, off_a, KA, km);
    if (rc != TG_OK) return rc;
    rc = sc_collect_kmers(kmers_b, off_b, KB, km);
    if (rc != TG_OK) return rc;
    CovTable *T = reinterpret_cast<CovTable *>(h_table);
    memset(T, 0, sizeof(CovTable));
    T->magic = COV_MAGIC; T->n_all = KA + KB; T->KA = KA;
    int n_bordered = 0;
    for (int k = 0; k < KA + KB; k++) {
        const std::string &s = km[k];
        T->pat[k] = sc_pattern(s);
        T->len[k] = (int)s.size();
        if ((int)s.size() > T->max_len) T->max_len = (int)s.size();
        const bool bord = sc_is_bordered(s);
        T->kind[k] = bord ? SC_BORDERED : ((int)s.size() > SC_W ? SC_SPECIAL : SC_PLAIN);
        if (bord && ++n_bordered > SC_MAXBORD) return tg_set_err(TG_EINVAL, "more than %d self-overlapping k-mers", SC_MAXBORD);
    }
    // A special k-mer whose every base is covered by plain k-mers of its own list occurring inside it (the HOR k-mers
    // CCCCTA|CCCTAA|CCCCTA, ...) never changes the UNION coverage: wherever it matches, those plain k-mers match too.
    for (int k = 0; k < KA + KB; k++) {
        if (T->kind[k] == SC_PLAIN) continue;
        const std::string &s = km[k];
        std::vector<char> covd(s.size(), 0);
        for (int q = (k < KA ? 0 : KA); q < (k < KA ? KA : KA + KB); q++) {
            if (T->kind[q] != SC_PLAIN) continue;
            for (size_t o = s.find(km[q]); o != std::string::npos; o = s.find(km[q], o + 1))
                for (size_t t = 0; t < km[q].size(); t++) covd[o + t] = 1;
        }
        bool all = true;
        for (char c : covd) all = all && c;
        if (all) T->kind[k] = SC_REDUNDANT;
    }
    for (int k = 0; k < KA + KB; k++) {
        if (T->kind[k] != SC_BORDERED) continue;
        T->bord_idx[T->n_bord++] = k;
        if (T->len[k] > SC_W) continue;
        int &bl = T->bord_len[k >= KA ? 1 : 0];
        if (bl == 0) bl = T->len[k];
        else if (bl != T->len[k]) bl = -1;
    }
    // entry of a window of n bases (n = SC_W: the main table; n < SC_W: an N follows, only what fits in n bases can match)
    auto entry = [&](uint32_t w, int n) {
        uint32_t e = 0;
        for (int k = 0; k < KA + KB; k++) {
            if (T->kind[k] == SC_REDUNDANT) continue;
            const bool b_list = k >= KA;
            const int len = (int)km[k].size();
            if (len > SC_W) {
                // bases 8.. of a long k-mer: "continues here" (T), whichever list
                const std::string rem = km[k].substr(SC_W);
                if (!(n < SC_W && (int)rem.size() > n) && sc_window_has_prefix(w, rem)) e |= E_T;
            }
            if (!sc_window_has_prefix(w, km[k]) || (n < SC_W && len > n)) continue;
            if (T->kind[k] == SC_PLAIN) e |= ((1u << len) - 1u) << (b_list ? 16 : 0);        // union of prefixes = the longest
            else if (len > SC_W) e |= b_list ? E_S2B : E_S2A;
            else e |= b_list ? E_S1B : E_S1A;
        }
        return e;
    };
    for (uint32_t w = 0; w < (uint32_t)SC_TAB; w++) T->tab[w] = entry(w, SC_W);
    for (int d = 1; d < SC_W; d++)
        for (uint32_t w = 0; w < (1u << (2 * d)); w++) T->tab_short[COV_SHORT_OFF(d) + w] = entry(w, d);
    return TG_OK;
}

static int launch_cov(int mode, const uint32_t *d_pack2, const uint32_t *d_nmask, int64_t total_bases, const int32_t *d_blk_read,
                      const void *table, int window, uint8_t *ca, uint8_t *cb, int32_t *counts, void *stream)
{
    if (total_bases <= 0) return TG_OK;
    if (total_bases % 128) return tg_set_err(TG_EINVAL, "total_bases must be a multiple of 128");
    if (!d_pack2 || !d_nmask || !table) return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CovParams P;
    P.pl.pack2 = d_pack2; P.pl.nmask = d_nmask; P.pl.total = total_bases;
    P.blk_read = d_blk_read;
    // 1 chunk of left halo; right halo: the window (density), or one chunk so that a long k-mer starting in the last
    // output chunk still sees its continuation flag (count)
    P.out_chunks = 32 - 1 - (mode == MODE_DENSITY ? ((window + 63) >> 6) : 1);
    const int64_t out_bases = (int64_t)P.out_chunks * COV_CH;
    P.n_tiles = (total_bases + out_bases - 1) / out_bases;
    P.tab = reinterpret_cast<const CovTable *>(table);
    P.window = window; P.cnt[0] = ca; P.cnt[1] = cb; P.counts = counts;
    const size_t smem = sizeof(CovSmem);
    long long grid = (long long)sms * 2;
    if (grid > (P.n_tiles + COV_GROUPS - 1) / COV_GROUPS) grid = (P.n_tiles + COV_GROUPS - 1) / COV_GROUPS;
    if (mode == MODE_COUNT) {
        TG_CUDA_CHECK(cudaFuncSetAttribute(scan_cov_kernel<MODE_COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        scan_cov_kernel<MODE_COUNT><<<(int)grid, COV_THREADS, smem, st>>>(P);
    } else {
        TG_CUDA_CHECK(cudaFuncSetAttribute(scan_cov_kernel<MODE_DENSITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        scan_cov_kernel<MODE_DENSITY><<<(int)grid, COV_THREADS, smem, st>>>(P);
    }
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_scan_basecount(const uint32_t *d_pack2, const uint32_t *d_nmask, int64_t total_bases,
                                const int32_t *d_blk_read, int n_reads, const void *d_table_pair, int32_t *d_counts, void *stream)
{
    if (n_reads <= 0) return TG_OK;
    if (!d_counts || !d_blk_read) return tg_set_err(TG_EINVAL, "NULL device pointer");
    TG_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * 2 * (size_t)n_reads, reinterpret_cast<cudaStream_t>(stream)));
    return launch_cov(MODE_COUNT, d_pack2, d_nmask, total_bases, d_blk_read, d_table_pair, 0, nullptr, nullptr, d_counts, stream);
}

TG_EXPORT int tg_scan_density(const uint32_t *d_pack2, const uint32_t *d_nmask, int64_t total_bases,
                              const void *d_table_pair, int window, uint8_t *d_cnt_a, uint8_t *d_cnt_b, void *stream)
{
    if (window < 1 || window > 255) return tg_set_err(TG_ERANGE, "window %d outside 1..255 (uint8 counts)", window);
    if (!d_cnt_a || !d_cnt_b) return tg_set_err(TG_EINVAL, "output pointer is NULL");
    return launch_cov(MODE_DENSITY, d_pack2, d_nmask, total_bases, nullptr, d_table_pair, window, d_cnt_a, d_cnt_b, nullptr, stream);
}
