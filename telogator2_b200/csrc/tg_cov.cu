// Coverage scan for sm_100a: get_telomere_base_count (tg_kmer.py:93-102) and get_telomere_kmer_density
// (tg_kmer.py:66-90) for TWO k-mer lists in one pass over the packed reads (HBM bound by design).
//
// One thread owns 64 consecutive bases: one 128-bit load of the 2-bit plane, one 64-bit load of the N plane, then one
// shared-memory table lookup per base.  The 32-bit table entry, indexed by the next 7 bases, holds for each of the two
// lists the coverage mask (1 << L) - 1 of the longest plain k-mer starting here (bits 0-6 / 16-22) and a "special k-mer
// might start here" flag (bit 7 / 23); the masks of 8 neighbouring bases are OR-ed into a 16-bit field per list with one
// shift and one LOP3 per base, so the union coverage (cov[p] = some selected match covers p) falls out as bit planes
// without any per-match loop.  Flagged bases (self-overlapping or long k-mers) and bases with an N within reach are rare;
// they go through per-chunk bitmaps to an exact slow path (direct comparison against the packed planes; greedy
// leftmost-non-overlapping selection per self-overlapping k-mer, one warp per k-mer).
// MODE_DENSITY then turns the coverage plane into the uint8 window counts cnt[i] = |cov (i, i + w]| with a sliding
// difference of the entering and leaving coverage bits, four outputs per step through a 16-entry prefix-sum table.
#include "tg_scan_common.cuh"
#include <string.h>

namespace {

struct CovTable {
    int32_t magic, n_all, KA, n_sp, n_bord, max_len, pad_[2];
    uint64_t pat[SC_MAXK];
    int32_t len[SC_MAXK];
    int32_t kind[SC_MAXK];             // SC_PLAIN / SC_SPECIAL / SC_BORDERED; k-mers of list A come first
    int32_t sp_idx[SC_MAXK];           // the n_sp special k-mers that are NOT self-overlapping
    int32_t bord_idx[SC_MAXBORD];      // the n_bord self-overlapping k-mers
    uint32_t tab[SC_TAB];
};
constexpr int32_t COV_MAGIC = 0x43563033;      // "CV03"
constexpr int COV_SLOT_MANY = 127;

constexpr int COV_THREADS = 512;                      // threads per CTA: they share one copy of the lookup table
constexpr int COV_GT = 128;                           // threads per group: a group works on its own tiles, own barrier
constexpr int COV_GROUPS = COV_THREADS / COV_GT;
constexpr int COV_CH = 64;                            // bases per thread
constexpr int COV_REGION = COV_GT * COV_CH;           // 8192 bases per tile (64 left halo, window right halo)
constexpr int COV_WORDS = COV_REGION / 32;
constexpr int COV_GW = COV_GT / 32;                   // warps per group
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

struct CovGroupSmem {
    uint32_t cov[2][COV_WORDS + 4];
    uint32_t spill[2][COV_GT + 4];
    uint32_t pk[COV_REGION / 16 + 8];          // the staged planes (slow paths read them back)
    uint32_t nm[COV_WORDS + 4];
    uint32_t flag[2][COV_WORDS + 2];           // per list: a special k-mer may start here (or an N is within the window)
    uint32_t defer[2][COV_GW];                 // per list: chunks holding candidates that need the greedy walk
};

struct CovSmem {
    uint32_t tab[SC_TAB];
    uint64_t pat[SC_MAXK];
    uint8_t len[SC_MAXK], kind[SC_MAXK];
    uint32_t lut[16];
    CovGroupSmem g[COV_GROUPS];
};

__device__ __forceinline__ void group_sync(int grp)
{
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(COV_GT) : "memory");
}
__device__ __forceinline__ bool group_sync_or(int grp, bool pred)
{
    int r;
    asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %3, 0; bar.red.or.pred q, %1, %2, p; selp.s32 %0, 1, 0, q; }"
                 : "=r"(r) : "r"(grp + 1), "r"(COV_GT), "r"((int)pred) : "memory");
    return r != 0;
}

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

// does k-mer (pat, len) start at region position x?  (staged planes; N anywhere in the span = no)
__device__ __noinline__ bool sm_kmer_at(const uint32_t *pk, const uint32_t *nm, int x, uint64_t pat, int len)
{
    const int pw = x >> 4, sh = (x & 15) * 2;
    uint64_t bases = ((uint64_t)pk[pw] | ((uint64_t)pk[pw + 1] << 32)) >> sh;
    if (sh) bases |= (uint64_t)pk[pw + 2] << (64 - sh);
    const uint32_t nbits = __funnelshift_r(nm[x >> 5], nm[(x >> 5) + 1], x & 31);
    const uint32_t lm = (len >= 32) ? 0xffffffffu : ((1u << len) - 1u);
    const uint64_t m = (len >= 32) ? ~0ull : ((1ull << (2 * len)) - 1ull);
    return (nbits & lm) == 0u && ((bases ^ pat) & m) == 0ull;
}

// flag bits of the positions (x - d, x + d) except x itself, one list
__device__ __forceinline__ bool flags_near(const uint32_t *fw, int x, int d)
{
    // bits [x - d + 1, x + d - 1] as a 64-bit window starting at x - 31 (d <= 32)
    const int s = x - 31;                        // may be negative: words before the region do not exist
    const int w0 = s >> 5, sh = s & 31;          // arithmetic shift: floor
    const uint32_t a = (w0 >= 0) ? fw[w0] : 0u, b = fw[w0 + 1], c = fw[w0 + 2];
    const uint64_t lo = ((uint64_t)a | ((uint64_t)b << 32)) >> sh;
    const uint64_t win = sh ? (lo | ((uint64_t)c << (64 - sh))) : lo;         // bit j <-> position s + j, position x is bit 31
    const uint64_t m = (((1ull << (2 * d - 1)) - 1ull) << (32 - d)) & ~(1ull << 31);
    return (win & m) != 0ull;
}

template <int MODE>
__global__ void __launch_bounds__(COV_THREADS, 2) scan_cov_kernel(const CovParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CovSmem &S = *reinterpret_cast<CovSmem *>(smem_raw);
    const int grp = threadIdx.x / COV_GT, tid = threadIdx.x % COV_GT, lane = tid & 31, warp = tid >> 5;
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
        if (threadIdx.x < 16) {                          // inclusive prefix popcounts of a nibble, one byte each
            uint32_t v = 0;
            for (int j = 0; j < 4; j++) v |= (uint32_t)__popc(threadIdx.x & ((2 << j) - 1)) << (8 * j);
            S.lut[threadIdx.x] = v;
        }
    }
    __syncthreads();
    const int n_bord = T->n_bord, n_all = T->n_all, KA = T->KA;
    const int w = P.window;
    const int64_t total = P.pl.total;

    for (long long tile = (long long)blockIdx.x * COV_GROUPS + grp; tile < P.n_tiles; tile += (long long)gridDim.x * COV_GROUPS) {
        const int64_t a0 = tile * (int64_t)P.out_chunks * COV_CH - COV_CH;     // absolute base of region index 0
        const int64_t a = a0 + (int64_t)tid * COV_CH;
        uint32_t fl[2][2] = {{0u, 0u}, {0u, 0u}}, nn[2] = {0u, 0u};
        group_sync(grp);
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
                    // fast path: no N within reach of any of my windows.  Per base: entry << r, then the coverage masks
                    // (bits 0-6 / 16-22) and the special flags (bit 7 / 23) of 8 neighbours share one register each.
                    uint32_t acc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, facc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < COV_CH; i++) {
                        const uint32_t e = tab_at(S.tab, pk[i >> 4], pk[(i >> 4) + 1], 2 * (i & 15));
                        const int r = i & 7;
                        const uint32_t t = e << r;
                        acc[i >> 3] |= t & (0x007f007fu << r);
                        facc[i >> 3] |= t & (0x00800080u << r);
                    }
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        uint32_t f[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) f[j] = (acc[j] >> (16 * g)) & 0xffffu;
                        cvw[g][0] = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
                        cvw[g][1] = (f[3] >> 8) | f[4] | (f[5] << 8) | (f[6] << 16) | (f[7] << 24);
                        sp[g] = f[7] >> 8;
                    }
                    if ((facc[0] | facc[1] | facc[2] | facc[3] | facc[4] | facc[5] | facc[6] | facc[7]) != 0u) {
#pragma unroll
                        for (int g = 0; g < 2; g++) {
                            uint32_t f[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) f[j] = (facc[j] >> (16 * g + 7)) & 0xffu;
                            fl[g][0] = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
                            fl[g][1] = f[4] | (f[5] << 8) | (f[6] << 16) | (f[7] << 24);
                        }
                    }
                } else slow = (nmw[0] & nmw[1]) != 0xffffffffu;           // some N within reach (all N: nothing to do)
            }
            // lanes with an N within reach (read ends, mostly), one at a time, all 32 lanes helping: lane j looks at bases
            // j and j + 32 of that lane's chunk; windows holding an N are skipped and noted
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
                uint32_t msk[2][2], fbal[2][2], nbal[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t lo = h ? (lowhalf ? b2 : b3) : (lowhalf ? b0 : b1);
                    const uint32_t hi = h ? (lowhalf ? b3 : b4) : (lowhalf ? b1 : b2);
                    const uint32_t nwin = __funnelshift_r(h ? m1 : m0, h ? m2 : m1, lane) & ((1u << SC_W) - 1u);
                    const uint32_t e = nwin ? 0u : tab_at(S.tab, lo, hi, 2 * (lane & 15));
                    msk[0][h] = e & 0x7fu;
                    msk[1][h] = (e >> 16) & 0x7fu;
                    fbal[0][h] = __ballot_sync(TG_FULL_MASK, (e >> 7) & 1u);
                    fbal[1][h] = __ballot_sync(TG_FULL_MASK, (e >> 23) & 1u);
                    nbal[h] = __ballot_sync(TG_FULL_MASK, nwin != 0u && (nwin & 1u) == 0u);
                }
                const int over = lane > 32 - SC_W ? 32 - lane : 32;       // bits of my mask that stay in this word
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t w0 = __reduce_or_sync(TG_FULL_MASK, msk[g][0] << lane);
                    const uint32_t w1 = __reduce_or_sync(TG_FULL_MASK, (over < 32 ? msk[g][0] >> over : 0u) | (msk[g][1] << lane));
                    const uint32_t w2 = __reduce_or_sync(TG_FULL_MASK, over < 32 ? msk[g][1] >> over : 0u);
                    if (lane == l) {
                        cvw[g][0] = w0; cvw[g][1] = w1; sp[g] = w2;
                        fl[g][0] = fbal[g][0]; fl[g][1] = fbal[g][1];
                    }
                }
                if (lane == l) { nn[0] = nbal[0]; nn[1] = nbal[1]; }
            }
            *reinterpret_cast<uint4 *>(&G.pk[4 * tid]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint2 *>(&G.nm[2 * tid]) = make_uint2(nmw[0], nmw[1]);
#pragma unroll
            for (int g = 0; g < 2; g++) {
                *reinterpret_cast<uint2 *>(&G.cov[g][2 * tid]) = make_uint2(cvw[g][0], cvw[g][1]);
                *reinterpret_cast<uint2 *>(&G.flag[g][2 * tid]) = make_uint2(fl[g][0] | nn[0], fl[g][1] | nn[1]);
                G.spill[g][tid + 1] = sp[g];
            }
            if (tid < 4) {
                G.cov[0][COV_WORDS + tid] = 0u; G.cov[1][COV_WORDS + tid] = 0u;
                // planes behind the region (a long k-mer that starts in the last chunk is compared against them)
                const int64_t ax = a0 + COV_REGION + 32 * (int64_t)tid;
                const bool ok = ax >= 0 && ax < total;
                G.nm[COV_WORDS + tid] = ok ? __ldg(P.pl.nmask + (ax >> 5)) : 0xffffffffu;
                G.pk[COV_REGION / 16 + 2 * tid] = ok ? __ldg(P.pl.pack2 + (ax >> 4)) : 0u;
                G.pk[COV_REGION / 16 + 2 * tid + 1] = ok ? __ldg(P.pl.pack2 + (ax >> 4) + 1) : 0u;
            }
            if (tid < 2) { G.spill[tid][0] = 0u; G.flag[tid][COV_WORDS] = 0u; G.flag[tid][COV_WORDS + 1] = 0u; }
            if (tid < 2 * COV_GW) G.defer[tid / COV_GW][tid % COV_GW] = 0u;
        }
        group_sync(grp);
        // ---- phase 2a: spill of the previous chunk; my own flagged positions ----
        bool deferred = false;
#pragma unroll 1
        for (int g = 0; g < 2; g++) {
            const uint32_t s_in = G.spill[g][tid];
            if (s_in) atomicOr(&G.cov[g][2 * tid], s_in);
            const int k0 = g ? KA : 0, k1 = g ? n_all : KA;
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                uint32_t fw = G.flag[g][2 * tid + h] & ~(h ? nn[1] : nn[0]);     // (N-near positions: below)
                while (fw) {
                    // a special k-mer of list g may start at x: compare those whose prefix the window holds.  One that does
                    // not overlap itself is covered right away; a self-overlapping one too, unless another candidate of
                    // the list is close enough to overlap it -- then re.finditer's greedy rule decides (phase 2b).
                    const int x = tid * COV_CH + h * 32 + __ffs(fw) - 1;
                    fw &= fw - 1;
                    const uint32_t e = tab_at(S.tab, G.pk[x >> 4], G.pk[(x >> 4) + 1], 2 * (x & 15));
                    const int slot = (int)((e >> (16 * g + 8)) & 0x7fu);
                    const int ka = (slot == COV_SLOT_MANY) ? k0 : k0 + slot, kb = (slot == COV_SLOT_MANY) ? k1 : ka + 1;
                    for (int k = ka; k < kb; k++) {
                        const int kind = S.kind[k], len = S.len[k];
                        if (kind != SC_SPECIAL && kind != SC_BORDERED) continue;
                        // (a k-mer that fits the window is fully decided by the entry)
                        if ((len > SC_W || slot == COV_SLOT_MANY) && !sm_kmer_at(G.pk, G.nm, x, S.pat[k], len))
                            continue;
                        if (kind == SC_BORDERED && flags_near(G.flag[g], x, len)) {
                            atomicOr(&G.defer[g][warp], 1u << lane);
                            deferred = true;
                        } else cov_or(G.cov[g], x, len);
                    }
                }
            }
        }
        {
            // an N within the window: every k-mer is compared directly, the warp takes the k-mers
            uint32_t have = __ballot_sync(TG_FULL_MASK, (nn[0] | nn[1]) != 0u);
            while (have) {
                const int l = __ffs(have) - 1;
                have &= have - 1;
                const uint32_t n0 = __shfl_sync(TG_FULL_MASK, nn[0], l), n1 = __shfl_sync(TG_FULL_MASK, nn[1], l);
                const int xb = (warp * 32 + l) * COV_CH;
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    uint32_t nw = h ? n1 : n0;
                    while (nw) {
                        const int x = xb + h * 32 + __ffs(nw) - 1;
                        nw &= nw - 1;
                        for (int k = lane; k < n_all; k += 32) {
                            const int kind = S.kind[k], g = k >= KA ? 1 : 0;
                            if (kind == SC_REDUNDANT || !sm_kmer_at(G.pk, G.nm, x, S.pat[k], S.len[k])) continue;
                            if (kind == SC_BORDERED && flags_near(G.flag[g], x, S.len[k])) {
                                atomicOr(&G.defer[g][warp], 1u << l);
                                deferred = true;
                            } else cov_or(G.cov[g], x, S.len[k]);
                        }
                    }
                }
            }
        }
        if (group_sync_or(grp, deferred)) {
            // ---- phase 2b (rare): self-overlapping k-mer b, one warp each: lanes verify the flagged positions of the
            //      deferred chunks (32 chunks per round, in position order), then the warp walks the true candidates left
            //      to right with re.finditer's rule.  Candidates outside the deferred chunks cannot overlap one inside. ----
            for (int b = warp; b < n_bord; b += COV_GW) {
                const int k = T->bord_idx[b];
                const int g = (k >= KA) ? 1 : 0;
                const int blen = S.len[k];
                const uint64_t pat = S.pat[k];
                int n_def = 0;
#pragma unroll
                for (int i = 0; i < COV_GW; i++) n_def += __popc(G.defer[g][i]);
                int blocked = -0x40000000;
                bool first = true;
                for (int r0 = 0; r0 < n_def; r0 += 32) {
                    int c = -1;                                           // my chunk: the (r0 + lane)-th deferred one
                    {
                        int rank = r0 + lane;
#pragma unroll
                        for (int i = 0; i < COV_GW; i++) {
                            const uint32_t wd = G.defer[g][i];
                            const int pc = __popc(wd);
                            if (c < 0 && rank >= 0 && rank < pc) c = i * 32 + (int)__fns(wd, 0, rank + 1);
                            rank -= pc;
                        }
                    }
                    uint64_t mine = 0ull;
                    if (c >= 0) {
#pragma unroll 1
                        for (int h = 0; h < 2; h++) {
                            uint32_t fw = G.flag[g][2 * c + h];
                            while (fw) {
                                const int bit = __ffs(fw) - 1;
                                fw &= fw - 1;
                                if (sm_kmer_at(G.pk, G.nm, c * COV_CH + h * 32 + bit, pat, blen)) mine |= 1ull << (h * 32 + bit);
                            }
                        }
                    }
                    uint32_t nz = __ballot_sync(TG_FULL_MASK, mine != 0ull);
                    while (nz) {
                        const int l = __ffs(nz) - 1;
                        nz &= nz - 1;
                        uint64_t word = __shfl_sync(TG_FULL_MASK, mine, l);
                        const int cc = __shfl_sync(TG_FULL_MASK, c, l);
                        while (word) {
                            const int x = cc * COV_CH + __ffsll((long long)word) - 1;
                            word &= word - 1;
                            if (first) {
                                first = false;
                                if (x < blen - 1 && a0 + x > 0) {         // the chain may reach back beyond the region
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
            group_sync(grp);
        }
        // ---- phase 3: outputs of chunks 1 .. out_chunks ----
        const bool act = tid >= 1 && tid <= P.out_chunks && a < total;
        if (MODE == MODE_COUNT) {
            // popcount of my two coverage words, one atomic per warp when the warp lies inside one read
            const int c0 = act ? __popc(G.cov[0][2 * tid]) + __popc(G.cov[0][2 * tid + 1]) : 0;
            const int c1 = act ? __popc(G.cov[1][2 * tid]) + __popc(G.cov[1][2 * tid + 1]) : 0;
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
            const int j0 = tid * COV_CH;
#pragma unroll 1
            for (int g = 0; g < 2; g++) {
                const uint32_t *cv = G.cov[g];
                const int cnt = range_popc(cv, j0 + 1, j0 + w);
                const uint32_t a0w = cv[2 * tid], a1w = cv[2 * tid + 1], a2w = cv[2 * tid + 2];
                const uint32_t sa[2] = {__funnelshift_r(a0w, a1w, 1), __funnelshift_r(a1w, a2w, 1)};      // leaving bits
                const int jb = j0 + w + 1;
                const uint32_t b0w = cv[jb >> 5], b1w = cv[(jb >> 5) + 1], b2w = cv[(jb >> 5) + 2];
                const uint32_t sb[2] = {__funnelshift_r(b0w, b1w, jb & 31), __funnelshift_r(b1w, b2w, jb & 31)};  // entering
                uint32_t cnt4 = (uint32_t)cnt * 0x01010101u;
                uint8_t *dst = P.cnt[g] + a;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t o[4];
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int g4 = 4 * q + h;
                        const uint32_t A = S.lut[(sb[g4 >> 3] >> (4 * (g4 & 7))) & 15u];
                        const uint32_t B = S.lut[(sa[g4 >> 3] >> (4 * (g4 & 7))) & 15u];
                        const uint32_t incl = cnt4 + A - B;            // counts of outputs 4 g4 + 1 .. 4 g4 + 4 (exact mod 2^32)
                        o[h] = __byte_perm(incl, cnt4, 0x2104);
                        cnt4 = __byte_perm(incl, 0u, 0x3333);
                    }
                    *reinterpret_cast<uint4 *>(dst + 16 * q) = make_uint4(o[0], o[1], o[2], o[3]);
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
    for (int k = 0; k < KA + KB; k++) {
        const std::string &s = km[k];
        T->pat[k] = sc_pattern(s);
        T->len[k] = (int)s.size();
        if ((int)s.size() > T->max_len) T->max_len = (int)s.size();
        const bool bord = sc_is_bordered(s);
        T->kind[k] = bord ? SC_BORDERED : ((int)s.size() > SC_W ? SC_SPECIAL : SC_PLAIN);
        if (bord && ++T->n_bord > SC_MAXBORD) return tg_set_err(TG_EINVAL, "more than %d self-overlapping k-mers", SC_MAXBORD);
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
    T->n_bord = T->n_sp = 0;
    for (int k = 0; k < KA + KB; k++) {
        if (T->kind[k] == SC_BORDERED) T->bord_idx[T->n_bord++] = k;
        else if (T->kind[k] == SC_SPECIAL) T->sp_idx[T->n_sp++] = k;
    }
    for (uint32_t w = 0; w < (uint32_t)SC_TAB; w++) {
        uint32_t e = 0;
        for (int k = 0; k < KA + KB; k++) {
            if (T->kind[k] == SC_REDUNDANT || !sc_window_has_prefix(w, km[k])) continue;
            const int sh = (k >= KA) ? 16 : 0;
            if (T->kind[k] == SC_PLAIN) e |= ((1u << km[k].size()) - 1u) << sh;      // union of prefixes = the longest
            else {
                // flag + which special k-mer (index within its list), COV_SLOT_MANY if several share the prefix
                const uint32_t slot = (uint32_t)(k >= KA ? k - KA : k);
                const uint32_t old = (e >> (sh + 8)) & 0x7fu;
                const uint32_t now = ((e >> (sh + 7)) & 1u) ? (old == slot ? slot : (uint32_t)COV_SLOT_MANY) : slot;
                e = (e & ~(0x7f00u << sh)) | (0x80u << sh) | (now << (sh + 8));
            }
        }
        T->tab[w] = e;
    }
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
    P.out_chunks = COV_GT - 1 - (mode == MODE_DENSITY ? ((window + 63) >> 6) : 0);
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
