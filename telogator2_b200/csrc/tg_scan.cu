// k-mer scan kernels for sm_100a (reference: source/tg_kmer.py); the coverage kernels live in tg_cov.cu.
//
//   pack_kernel     ASCII -> 2-bit codes + N-mask bit plane (the HBM layout of the reads)
//   orient_kernel   reverse complement of flagged reads in the packed domain (tg_tel.py:265-266), word parallel
//   hits_kernel     get_nonoverlapping_kmer_hits (tg_kmer.py:173-230), one warp per slice, fused: nothing but the packed
//                   bases is read from HBM and nothing but the spans is written
//
// hits_kernel works on SETS of k-mers (64-bit masks, K <= 64) per position, in the closed form of SURVEY 9-E:
//   R(p)  k-mers with a selected raw match starting at p: a 7-base window lookup names the set of plain k-mers; special
//         k-mers (self-overlapping: re.finditer's greedy rule; longer than the window) are flagged by the lookup and verified
//   C(q)  k-mers whose selected match covers q            = OR_d R(q - d) & {len > d}
//   S(p)  survivors: k in R(p) unless a superstring k-mer covers some position of [p, p + len_k)   (tg_kmer.py:187-202)
//   P(x)  k-mers with a surviving match ending at x; block starts Start(x) = S(x) & ~P(x), chain ends End(x) = P(x) & ~S(x)
//         (abutting matches of one k-mer collapse, tg_kmer.py:206-212)
//   sweep: the blocks opened at the latest block-start position stay open until their chain ends or the next block start
//         of any k-mer truncates them (tg_kmer.py:216-228); every closing emits one span.
// The warp streams over its slice in tiles of 128 positions; R / S, C and P live in shared-memory rings, S and the events
// lag one maximal k-mer length behind R so that nothing is computed twice.
#include "tg_scan_common.cuh"
#include <string.h>
#include <map>

namespace {

// ---- pack / orient ----------------------------------------------------------------------------------------------
// codes: A=0 C=1 T=2 G=3 ((ascii >> 1) & 3); complement = code ^ 2.
__device__ __forceinline__ uint32_t ascii_code(uint32_t c) { return (c >> 1) & 3u; }
__device__ __forceinline__ bool ascii_is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

__device__ __forceinline__ int find_read(const int64_t *off, int n, int64_t base)
{
    int lo = 0, hi = n;                      // off[lo] <= base < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= base) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ base_off,
                                                   const int32_t *__restrict__ len, int n_reads,
                                                   uint32_t *__restrict__ pack2, uint32_t *__restrict__ nmask,
                                                   uint8_t *__restrict__ not_acgtn)
{
    const int64_t total = base_off[n_reads];
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 32 < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b0 = g * 32;
        const int r = find_read(base_off, n_reads, b0);
        const int64_t rel = b0 - base_off[r];
        const int L = len[r];
        const uint4 v0 = *reinterpret_cast<const uint4 *>(ascii + b0);
        const uint4 v1 = *reinterpret_cast<const uint4 *>(ascii + b0 + 16);
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t p[2] = {0u, 0u}, m = 0u;
        bool other = false;                          // a letter RC() has no complement for (tg_util.py:41)
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
            const bool ok = ascii_is_acgt(c) && (rel + i < L);
            p[i >> 4] |= (ok ? ascii_code(c) : 0u) << ((i & 15) * 2);
            m |= (ok ? 0u : 1u) << i;
            other = other || (!ok && c != 'N' && rel + i < L);
        }
        pack2[g * 2] = p[0];
        pack2[g * 2 + 1] = p[1];
        nmask[g] = m;
        if (other && not_acgtn) not_acgtn[r] = 1;
    }
}

// 32 output positions per thread: the 32 source bases [L - 32 - q0, L - q0) as one 64-bit word, reversed pair-wise and
// complemented (code ^ 2); positions past the read stay N.
__global__ void __launch_bounds__(256) orient_kernel(const uint32_t *__restrict__ pack2, const uint32_t *__restrict__ nmask,
                                                     const int64_t *__restrict__ base_off, const int32_t *__restrict__ len,
                                                     const uint8_t *__restrict__ flip, int n_reads,
                                                     uint32_t *__restrict__ pack2_out, uint32_t *__restrict__ nmask_out)
{
    const int64_t total = base_off[n_reads];
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 32 < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b0 = g * 32;
        const int r = find_read(base_off, n_reads, b0);
        if (!flip[r]) {
            pack2_out[g * 2] = pack2[g * 2];
            pack2_out[g * 2 + 1] = pack2[g * 2 + 1];
            nmask_out[g] = nmask[g];
            continue;
        }
        const int64_t ro = base_off[r];
        const int L = len[r];
        const int s0 = L - 32 - (int)(b0 - ro);              // first source position (may be negative: outputs past the read)
        uint64_t bases = 0ull;
        uint32_t nb = 0xffffffffu;
        if (s0 > -32) {
            const int sh = s0 < 0 ? -s0 : 0;                 // source positions below 0 do not exist
            bases = sc_bases64(pack2, ro + s0 + sh) << (2 * sh);
            nb = (sc_nbits32(nmask, ro + s0 + sh) << sh) | ((1u << sh) - 1u);
        }
        // reverse the order of the 2-bit codes: bit reversal, then swap the two bits of every code back
        uint64_t rv = ((uint64_t)__brev((uint32_t)bases) << 32) | (uint64_t)__brev((uint32_t)(bases >> 32));
        rv = ((rv >> 1) & 0x5555555555555555ull) | ((rv & 0x5555555555555555ull) << 1);
        const uint32_t nr = __brev(nb);
        // spread the N bits to both bits of their code, clear those codes
        uint64_t x = (uint64_t)(~nr);
        x = (x | (x << 16)) & 0x0000ffff0000ffffull;
        x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
        x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
        x = (x | (x << 2)) & 0x3333333333333333ull;
        x = (x | (x << 1)) & 0x5555555555555555ull;
        rv = (rv ^ 0xaaaaaaaaaaaaaaaaull) & (x | (x << 1));
        pack2_out[g * 2] = (uint32_t)rv;
        pack2_out[g * 2 + 1] = (uint32_t)(rv >> 32);
        nmask_out[g] = nr;
    }
}

// ---- non-overlapping hits -----------------------------------------------------------------------------------------
constexpr int32_t HIT_MAGIC = 0x48543032;            // "HT02"
constexpr int HIT_MAXSETS = 512;                      // distinct sets of plain k-mers a window can name
constexpr int HIT_SHORT_N = 5464;                     // sum of 4^d, d = 1 .. 6, rounded up
__host__ __device__ constexpr int HIT_SHORT_OFF(int d) { return ((1 << (2 * d)) - 4) / 3; }
constexpr int HIT_SLOT_MANY = 63;
constexpr int HIT_SHORT_LEN = 8;                      // k-mers up to this length are covered by the fixed 8-step gathers

// table entry (16 bits): bits 0-8 set id, bits 9-14 special slot (HIT_SLOT_MANY: several), bit 15 "a special k-mer may start here"
struct HitTable {
    int32_t magic, K, n_sets, n_bord, max_len, n_spec, pad_[2];
    uint64_t pat[SC_MAXLIST];
    int32_t len[SC_MAXLIST];
    int32_t kind[SC_MAXLIST];                 // SC_PLAIN / SC_SPECIAL / SC_BORDERED
    int32_t bslot[SC_MAXLIST];                // index among the self-overlapping k-mers, -1 otherwise
    int32_t spec_idx[SC_MAXLIST];             // the n_spec special k-mers (entry slot -> k-mer)
    uint64_t sup[SC_MAXLIST];                 // KMER_ISSUBSTRING[k] as a bit set
    uint64_t lengt[SC_MAXLEN];                // lengt[d] = k-mers longer than d
    uint64_t bylen[SC_MAXLEN + 1];            // bylen[l] = k-mers of length l
    uint64_t longmask;                        // k-mers longer than HIT_SHORT_LEN
    uint64_t sets[HIT_MAXSETS];
    uint16_t tab_short[HIT_SHORT_N];
    uint16_t tab[SC_TAB];
};

constexpr int HIT_WARPS = 24;                         // warps per CTA (one CTA per SM, they share the table)
constexpr int HIT_T = 128;                            // positions per tile
constexpr int HIT_CAP = 256;                          // ring capacity >= HIT_T + 2 * SC_MAXLEN (+ slack), power of two
constexpr int HIT_BUF = 64;                           // spans buffered per warp before they go to global memory

struct HitParams {
    const uint32_t *pack2, *nmask;
    const int64_t *base_off;
    const int32_t *slice_read, *slice_lo, *slice_hi;
    int n_slices;
    const HitTable *tab;
    int32_t *span_slice, *span_k, *span_start, *span_end;
    long long span_cap;
    unsigned long long *span_count;
    unsigned int *next_slice;
};

struct alignas(16) HitWarpSmem {
    unsigned long long rs[HIT_CAP];           // R(p), overwritten by S(p) once decided
    unsigned long long c[HIT_CAP];            // C(q)
    unsigned long long pe[HIT_CAP];           // P(x)
    uint32_t pk[HIT_T / 16 + 4];              // staged bases [rA, rA + T + 64)
    uint32_t nm[HIT_T / 32 + 4];
    int32_t blocked[SC_MAXBORD];              // greedy state of every self-overlapping k-mer
    uint8_t list[HIT_T];                      // tile positions with a match (stage C works on these only)
    int32_t buf[HIT_BUF][3];                  // (k, start, end)
};

struct alignas(16) HitSmem {
    uint16_t tab[SC_TAB];
    unsigned long long sets[HIT_MAXSETS];
    unsigned long long sup[SC_MAXLIST], lengt[SC_MAXLEN], bylen[SC_MAXLEN + 1], pat[SC_MAXLIST];
    uint8_t len[SC_MAXLIST], kind[SC_MAXLIST], spec[SC_MAXLIST];
    int8_t bslot[SC_MAXLIST];
    uint8_t bord_k[SC_MAXBORD];
    HitWarpSmem w[HIT_WARPS];
};

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int l)
{
    return ((unsigned long long)__shfl_sync(TG_FULL_MASK, (uint32_t)(v >> 32), l) << 32) | __shfl_sync(TG_FULL_MASK, (uint32_t)v, l);
}

// does k-mer (pat, len) start at tile position u?  (staged planes; N anywhere in the span = no)
__device__ __forceinline__ bool hit_kmer_at(const uint32_t *pk, const uint32_t *nm, int u, unsigned long long pat, int len)
{
    const int pw = u >> 4, sh = (u & 15) * 2;
    unsigned long long bases = ((unsigned long long)pk[pw] | ((unsigned long long)pk[pw + 1] << 32)) >> sh;
    if (sh) bases |= (unsigned long long)pk[pw + 2] << (64 - sh);
    const uint32_t nbits = __funnelshift_r(nm[u >> 5], nm[(u >> 5) + 1], u & 31);
    const uint32_t lm = (len >= 32) ? 0xffffffffu : ((1u << len) - 1u);
    const unsigned long long m = (len >= 32) ? ~0ull : ((1ull << (2 * len)) - 1ull);
    return (nbits & lm) == 0u && ((bases ^ pat) & m) == 0ull;
}

// the span buffer of a warp goes to global memory: one atomic for the warp
__device__ __forceinline__ void flush_spans(const HitParams &P, HitWarpSmem &W, int q, int n_buf, int lane)
{
    unsigned long long base = 0ull;
    if (lane == 0) base = atomicAdd(P.span_count, (unsigned long long)n_buf);
    base = shfl64(base, 0);
    __syncwarp();
    for (int j = lane; j < n_buf; j += 32)
        if ((long long)(base + j) < P.span_cap) {
            P.span_slice[base + j] = q; P.span_k[base + j] = W.buf[j][0];
            P.span_start[base + j] = W.buf[j][1]; P.span_end[base + j] = W.buf[j][2];
        }
    __syncwarp();
}

__global__ void __launch_bounds__(HIT_WARPS * 32, 1) hits_kernel(const HitParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HitSmem &S = *reinterpret_cast<HitSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    HitWarpSmem &W = S.w[warp];
    const HitTable *__restrict__ T = P.tab;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T->tab);
        uint4 *dst = reinterpret_cast<uint4 *>(S.tab);
        for (int i = threadIdx.x; i < SC_TAB / 8; i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < HIT_MAXSETS; i += blockDim.x) S.sets[i] = T->sets[i];
        if (threadIdx.x < SC_MAXLIST) {
            const int k = threadIdx.x;
            S.sup[k] = T->sup[k]; S.pat[k] = T->pat[k]; S.len[k] = (uint8_t)T->len[k]; S.kind[k] = (uint8_t)T->kind[k];
            S.spec[k] = (uint8_t)T->spec_idx[k]; S.bslot[k] = (int8_t)T->bslot[k];
            if (k < T->K && T->bslot[k] >= 0) S.bord_k[T->bslot[k]] = (uint8_t)k;
        }
        if (threadIdx.x < SC_MAXLEN) S.lengt[threadIdx.x] = T->lengt[threadIdx.x];
        if (threadIdx.x <= SC_MAXLEN) S.bylen[threadIdx.x] = T->bylen[threadIdx.x];
    }
    __syncthreads();
    const int n_bord = T->n_bord, n_spec = T->n_spec, M = T->max_len;
    unsigned long long lg8[HIT_SHORT_LEN];
#pragma unroll
    for (int d = 0; d < HIT_SHORT_LEN; d++) lg8[d] = S.lengt[d];

    for (;;) {
        int q = 0;
        if (lane == 0) q = (int)atomicAdd(P.next_slice, 1u);
        q = __shfl_sync(TG_FULL_MASK, q, 0);
        if (q >= P.n_slices) break;
        const int len = P.slice_hi[q] - P.slice_lo[q];
        const int64_t gbase = P.base_off[P.slice_read[q]] + P.slice_lo[q];      // absolute base of slice position 0
        __syncwarp();
        for (int i = lane; i < HIT_CAP; i += 32) { W.rs[i] = 0ull; W.c[i] = 0ull; W.pe[i] = 0ull; }
        if (lane < SC_MAXBORD) W.blocked[lane] = 0;
        unsigned long long open = 0ull;                  // blocks opened at s_open that are still running (warp uniform)
        int s_open = 0, n_buf = 0;
        const int n_tiles = (len + 1 + M + HIT_T - 1) / HIT_T;
        for (int tile = 0; tile < n_tiles; tile++) {
            const int rA = tile * HIT_T;
            __syncwarp();
            // ---- stage the planes of slice positions [rA, rA + T + 64): an unaligned window, clipped to the slice ----
            if (lane < HIT_T / 16 + 4) {
                const int x0 = rA + 16 * lane;
                uint32_t v = 0u;
                if (x0 < len) {
                    const int64_t g = gbase + x0;
                    v = __funnelshift_r(P.pack2[g >> 4], P.pack2[(g >> 4) + 1], (int)(g & 15) * 2);
                }
                W.pk[lane] = v;
            } else if (lane >= 16 && lane < 16 + HIT_T / 32 + 4) {
                const int j = lane - 16, x0 = rA + 32 * j;
                uint32_t m = 0xffffffffu;
                if (x0 < len) {
                    const int64_t g = gbase + x0;
                    m = __funnelshift_r(P.nmask[g >> 5], P.nmask[(g >> 5) + 1], (int)(g & 31));
                    if (x0 + 32 > len) m |= ~((1u << (len - x0)) - 1u);
                }
                W.nm[j] = m;
            }
            __syncwarp();
            // ---- stage A: R(pos) for the T new positions; candidates of self-overlapping k-mers go to ballots ----
            uint32_t cm[HIT_T / 32];                     // per sub-tile: which self-overlapping k-mers may start at my position
#pragma unroll
            for (int i = 0; i < HIT_T / 32; i++) {
                const int u = 32 * i + lane, pos = rA + u;
                const uint32_t nwin = __funnelshift_r(W.nm[u >> 5], W.nm[(u >> 5) + 1], u & 31) & ((1u << SC_W) - 1u);
                const uint32_t win = __funnelshift_r(W.pk[u >> 4], W.pk[(u >> 4) + 1], 2 * (u & 15)) & (uint32_t)(SC_TAB - 1);
                uint32_t e = 0u;
                if (nwin == 0u) e = S.tab[win];
                else if (!(nwin & 1u)) {
                    const int d = __ffs(nwin) - 1;
                    e = __ldg(T->tab_short + HIT_SHORT_OFF(d) + (win & ((1u << (2 * d)) - 1u)));
                }
                unsigned long long r = S.sets[e & (HIT_MAXSETS - 1)];
                uint32_t cmask = 0u;
                if (e & 0x8000u) {
                    const int slot = (e >> 9) & 63;
                    const int s_lo = (slot == HIT_SLOT_MANY) ? 0 : slot, s_hi = (slot == HIT_SLOT_MANY) ? n_spec : slot + 1;
                    for (int s = s_lo; s < s_hi; s++) {
                        const int k = S.spec[s], kl = S.len[k];
                        // (a k-mer that fits the window and is named alone is decided by the entry)
                        if ((kl > SC_W || slot == HIT_SLOT_MANY || nwin != 0u) && !hit_kmer_at(W.pk, W.nm, u, S.pat[k], kl)) continue;
                        if (S.bslot[k] >= 0) cmask |= 1u << S.bslot[k];
                        else {
                            r |= 1ull << k;
                            for (int d = HIT_SHORT_LEN; d < kl; d++) atomicOr(&W.c[(pos + d) & (HIT_CAP - 1)], 1ull << k);
                        }
                    }
                }
                cm[i] = cmask;
                W.rs[pos & (HIT_CAP - 1)] = r;
            }
            // ---- self-overlapping k-mers: re.finditer's rule, left to right, the whole warp in lock step ----
            if (__any_sync(TG_FULL_MASK, (cm[0] | cm[1] | cm[2] | cm[3]) != 0u)) {
                __syncwarp();
                for (int b = 0; b < n_bord; b++) {
                    const int k = S.bord_k[b], kl = S.len[k];
                    int blocked = W.blocked[b];
#pragma unroll
                    for (int i = 0; i < HIT_T / 32; i++) {
                        uint32_t cand = __ballot_sync(TG_FULL_MASK, (cm[i] >> b) & 1u), sel = 0u;
                        while (cand) {
                            const int bit = __ffs(cand) - 1;
                            cand &= cand - 1;
                            const int pos = rA + 32 * i + bit;
                            if (pos >= blocked) { blocked = pos + kl; sel |= 1u << bit; }
                        }
                        if ((sel >> lane) & 1u) {
                            const int pos = rA + 32 * i + lane;
                            W.rs[pos & (HIT_CAP - 1)] |= 1ull << k;
                            for (int d = HIT_SHORT_LEN; d < kl; d++) atomicOr(&W.c[(pos + d) & (HIT_CAP - 1)], 1ull << k);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) W.blocked[b] = blocked;
                }
            }
            __syncwarp();
            // ---- stage B: C(q) |= OR_{d < 8} R(q - d) & {len > d}   (longer reaches were pushed above) ----
#pragma unroll
            for (int i = 0; i < HIT_T / 32; i++) {
                const int pos = rA + 32 * i + lane;
                unsigned long long c = 0ull;
#pragma unroll
                for (int d = 0; d < HIT_SHORT_LEN; d++) c |= W.rs[(pos - d) & (HIT_CAP - 1)] & lg8[d];   // (positions < 0: zeroed ring)
                if (c) W.c[pos & (HIT_CAP - 1)] |= c;
            }
            __syncwarp();
            // ---- stage C: survivors S(p), p in [rA - M, rA + T - M); each survivor announces its end in P.
            //      Positions without a match are skipped: the others are compacted so that all lanes work ----
            int n_list = 0;
#pragma unroll
            for (int i = 0; i < HIT_T / 32; i++) {
                const int p = rA - M + 32 * i + lane;
                const bool has = p >= 0 && W.rs[p & (HIT_CAP - 1)] != 0ull;
                const uint32_t bal = __ballot_sync(TG_FULL_MASK, has);
                if (has) W.list[n_list + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)(32 * i + lane);
                n_list += __popc(bal);
            }
            __syncwarp();
#pragma unroll 1
            for (int i0 = 0; i0 < n_list; i0 += 32) {
                if (i0 + lane >= n_list) continue;
                const int p = rA - M + W.list[i0 + lane];
                const unsigned long long r = W.rs[p & (HIT_CAP - 1)];
                // most deleted matches are covered by a superstring at their very first position
                const unsigned long long c0 = W.c[p & (HIT_CAP - 1)];
                unsigned long long dead = 0ull, rem = r, need = 0ull;
                while (rem) {
                    const int k = __ffsll((long long)rem) - 1;
                    rem &= rem - 1;
                    if (S.sup[k] & c0) dead |= 1ull << k; else if (S.sup[k]) need |= 1ull << k;
                }
                if (need) {
                    // coverage sets of the next positions, OR-ed up: D[t] = C(p) | .. | C(p + t)
                    unsigned long long D[HIT_SHORT_LEN];
                    D[0] = c0;
#pragma unroll
                    for (int t = 1; t < HIT_SHORT_LEN; t++) D[t] = D[t - 1] | W.c[(p + t) & (HIT_CAP - 1)];
                    while (need) {
                        const int k = __ffsll((long long)need) - 1;
                        need &= need - 1;
                        const int kl = S.len[k];
                        unsigned long long acc = D[HIT_SHORT_LEN - 1];
#pragma unroll
                        for (int t = HIT_SHORT_LEN - 2; t >= 0; t--) acc = (kl <= t + 1) ? D[t] : acc;
                        for (int d = HIT_SHORT_LEN; d < kl; d++) acc |= W.c[(p + d) & (HIT_CAP - 1)];      // long k-mers
                        if (S.sup[k] & acc) dead |= 1ull << k;
                    }
                }
                unsigned long long sv = r & ~dead;
                W.rs[p & (HIT_CAP - 1)] = sv;
                while (sv) {
                    const int k = __ffsll((long long)sv) - 1;
                    sv &= sv - 1;
                    atomicOr(&W.pe[(p + S.len[k]) & (HIT_CAP - 1)], 1ull << k);
                }
            }
            __syncwarp();
            // ---- stage D: events at x in [rA - M, rA + T - M), x <= len; one sweep, the warp in lock step ----
#pragma unroll 1
            for (int i = 0; i < HIT_T / 32; i++) {
                const int x0 = rA - M + 32 * i, x = x0 + lane;
                unsigned long long st = 0ull, en = 0ull;
                if (x >= 0 && x <= len) {
                    const unsigned long long sx = W.rs[x & (HIT_CAP - 1)], px = W.pe[x & (HIT_CAP - 1)];
                    st = sx & ~px;
                    en = px & ~sx;
                    W.pe[x & (HIT_CAP - 1)] = 0ull;
                }
                uint32_t ev = __ballot_sync(TG_FULL_MASK, (st | en) != 0ull);
                const uint32_t starts = __ballot_sync(TG_FULL_MASK, st != 0ull);
                const unsigned long long mine = st ? st : en;         // a block start closes everything: its chain ends do not matter
                while (ev) {
                    const int l = __ffs(ev) - 1;
                    ev &= ev - 1;
                    const unsigned long long v = shfl64(mine, l);
                    const bool is_start = (starts >> l) & 1u;
                    // a block start closes everything still open; otherwise the chains that end here close
                    const unsigned long long closing = is_start ? open : (v & open);
                    if (closing) {
                        const int n = __popcll(closing);
                        if (n_buf + n > HIT_BUF) { flush_spans(P, W, q, n_buf, lane); n_buf = 0; }
                        if (n == 1) {
                            if (lane == 0) {
                                W.buf[n_buf][0] = __ffsll((long long)closing) - 1;
                                W.buf[n_buf][1] = s_open;
                                W.buf[n_buf][2] = x0 + l;
                            }
                        } else {
                            for (int j = lane; j < n; j += 32) {         // lane j takes the j-th closing k-mer
                                unsigned long long m = closing;
                                for (int t = 0; t < j; t++) m &= m - 1;
                                W.buf[n_buf + j][0] = __ffsll((long long)m) - 1;
                                W.buf[n_buf + j][1] = s_open;
                                W.buf[n_buf + j][2] = x0 + l;
                            }
                        }
                        n_buf += n;
                    }
                    if (is_start) { open = v; s_open = x0 + l; } else open &= ~v;
                }
            }
            // the C ring slots the next tile will push into / OR into
            for (int i = lane; i < HIT_T; i += 32) W.c[(rA + HIT_T + M + i) & (HIT_CAP - 1)] = 0ull;
        }
        if (n_buf) flush_spans(P, W, q, n_buf, lane);
    }
}

}  // namespace

// ================================================================================================
// host: table construction
// ================================================================================================
TG_EXPORT size_t tg_kmer_table_bytes(void) { return sizeof(HitTable); }

TG_EXPORT int tg_kmer_table_build(const char *kmers_concat, const int32_t *kmer_off, int K,
                                  const int32_t *issub_edges, const int32_t *issub_off, void *h_table)
{
    if (!h_table) return tg_set_err(TG_EINVAL, "NULL argument");
    std::vector<std::string> km;
    int rc = sc_collect_kmers(kmers_concat, kmer_off, K, km);
    if (rc != TG_OK) return rc;
    HitTable *T = reinterpret_cast<HitTable *>(h_table);
    memset(T, 0, sizeof(HitTable));
    T->magic = HIT_MAGIC; T->K = K;
    for (int k = 0; k < SC_MAXLIST; k++) T->bslot[k] = -1;
    for (int k = 0; k < K; k++) {
        const int l = (int)km[k].size();
        T->pat[k] = sc_pattern(km[k]);
        T->len[k] = l;
        if (l > T->max_len) T->max_len = l;
        for (int d = 0; d < l; d++) T->lengt[d] |= 1ull << k;
        T->bylen[l] |= 1ull << k;
        if (l > HIT_SHORT_LEN) T->longmask |= 1ull << k;
        const bool bord = sc_is_bordered(km[k]);
        T->kind[k] = bord ? SC_BORDERED : (l > SC_W ? SC_SPECIAL : SC_PLAIN);
        T->bslot[k] = -1;
        if (bord) {
            if (T->n_bord >= SC_MAXBORD) return tg_set_err(TG_EINVAL, "more than %d self-overlapping k-mers", SC_MAXBORD);
            T->bslot[k] = T->n_bord++;
        }
        if (T->kind[k] != SC_PLAIN) {
            if (T->n_spec >= HIT_SLOT_MANY) return tg_set_err(TG_EINVAL, "more than %d special k-mers", HIT_SLOT_MANY);
            T->spec_idx[T->n_spec++] = k;
        }
    }
    if (issub_edges && issub_off)
        for (int k = 0; k < K; k++)
            for (int e = issub_off[k]; e < issub_off[k + 1]; e++) {
                if (issub_edges[e] < 0 || issub_edges[e] >= K) return tg_set_err(TG_EINVAL, "bad KMER_ISSUBSTRING edge");
                T->sup[k] |= 1ull << issub_edges[e];
            }
    std::map<uint64_t, int> ids;
    ids[0ull] = 0;
    T->n_sets = 1;
    bool overflow = false;
    // entry of a window of n bases (n = SC_W: the main table; n < SC_W: an N follows, only k-mers of length <= n fit)
    auto entry = [&](uint32_t w, int n) -> uint16_t {
        uint64_t set = 0;
        int slot = -1;
        for (int k = 0; k < K; k++) {
            if (!sc_window_has_prefix(w, km[k]) || (n < SC_W && (int)km[k].size() > n)) continue;
            if (T->kind[k] == SC_PLAIN) { set |= 1ull << k; continue; }
            int s = 0;
            while (T->spec_idx[s] != k) s++;
            slot = (slot < 0) ? s : HIT_SLOT_MANY;
        }
        // a plain k-mer with a plain superstring starting at the same position is deleted by it, and whatever it could
        // delete itself the superstring deletes too (substring containment is transitive): it never needs to be seen
        for (int k = 0; k < K; k++)
            if (((set >> k) & 1ull) && (T->sup[k] & set)) set &= ~(1ull << k);
        auto it = ids.find(set);
        int id;
        if (it != ids.end()) id = it->second;
        else if (T->n_sets >= HIT_MAXSETS) { overflow = true; id = 0; }
        else { id = T->n_sets; ids[set] = id; T->sets[T->n_sets++] = set; }
        return (uint16_t)(id | (slot >= 0 ? (0x8000 | (slot << 9)) : 0));
    };
    for (uint32_t w = 0; w < (uint32_t)SC_TAB; w++) T->tab[w] = entry(w, SC_W);
    for (int d = 1; d < SC_W; d++)
        for (uint32_t w = 0; w < (1u << (2 * d)); w++) T->tab_short[HIT_SHORT_OFF(d) + w] = entry(w, d);
    if (overflow) return tg_set_err(TG_ERANGE, "more than %d distinct k-mer sets per window", HIT_MAXSETS);
    return TG_OK;
}

// ================================================================================================
// launches
// ================================================================================================
TG_EXPORT int tg_scan_pack(const uint8_t *d_ascii, const int64_t *d_base_off, const int32_t *d_len, int n_reads,
                           uint32_t *d_pack2, uint32_t *d_nmask, uint8_t *d_not_acgtn, void *stream)
{
    if (n_reads <= 0) return TG_OK;
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    pack_kernel<<<sms * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_ascii, d_base_off, d_len, n_reads, d_pack2, d_nmask,
                                                                              d_not_acgtn);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

TG_EXPORT int tg_scan_orient(const uint32_t *d_pack2, const uint32_t *d_nmask, const int64_t *d_base_off,
                             const int32_t *d_len, const uint8_t *d_flip, int n_reads,
                             uint32_t *d_pack2_out, uint32_t *d_nmask_out, void *stream)
{
    if (n_reads <= 0) return TG_OK;
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    orient_kernel<<<sms * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_pack2, d_nmask, d_base_off, d_len, d_flip, n_reads,
                                                                                d_pack2_out, d_nmask_out);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}

static thread_local unsigned int *g_counter = nullptr;     // device work counter (per host thread / device)
static thread_local int g_counter_dev = -1;

static int get_counter(unsigned int **out, cudaStream_t st)
{
    int dev = 0;
    TG_CUDA_CHECK(cudaGetDevice(&dev));
    if (!g_counter || g_counter_dev != dev) {
        TG_CUDA_CHECK(cudaMalloc(&g_counter, 16));
        g_counter_dev = dev;
    }
    TG_CUDA_CHECK(cudaMemsetAsync(g_counter, 0, 16, st));
    *out = g_counter;
    return TG_OK;
}

TG_EXPORT int tg_scan_hits(const uint32_t *d_pack2, const uint32_t *d_nmask, const int64_t *d_base_off,
                           const int32_t *d_slice_read, const int32_t *d_slice_lo, const int32_t *d_slice_hi, int n_slices,
                           const void *d_table,
                           int32_t *d_span_slice, int32_t *d_span_k, int32_t *d_span_start, int32_t *d_span_end,
                           int64_t span_cap, unsigned long long *d_span_count, void *stream)
{
    if (n_slices <= 0) return TG_OK;
    if (!d_pack2 || !d_nmask || !d_base_off || !d_slice_read || !d_slice_lo || !d_slice_hi || !d_table ||
        !d_span_slice || !d_span_k || !d_span_start || !d_span_end || !d_span_count)
        return tg_set_err(TG_EINVAL, "NULL device pointer");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    HitParams P;
    P.pack2 = d_pack2; P.nmask = d_nmask; P.base_off = d_base_off;
    P.slice_read = d_slice_read; P.slice_lo = d_slice_lo; P.slice_hi = d_slice_hi; P.n_slices = n_slices;
    P.tab = reinterpret_cast<const HitTable *>(d_table);
    P.span_slice = d_span_slice; P.span_k = d_span_k; P.span_start = d_span_start; P.span_end = d_span_end;
    P.span_cap = span_cap; P.span_count = d_span_count;
    int rc = get_counter(&P.next_slice, st);
    if (rc != TG_OK) return rc;
    TG_CUDA_CHECK(cudaMemsetAsync(d_span_count, 0, sizeof(unsigned long long), st));
    const size_t smem = sizeof(HitSmem);
    TG_CUDA_CHECK(cudaFuncSetAttribute(hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = sms;
    if (grid > (n_slices + HIT_WARPS - 1) / HIT_WARPS) grid = (n_slices + HIT_WARPS - 1) / HIT_WARPS;
    hits_kernel<<<grid, HIT_WARPS * 32, smem, st>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}
