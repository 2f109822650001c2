// k-mer scan kernels for sm_100a (reference: source/tg_kmer.py).
//
//   pack_kernel        ASCII -> 2-bit codes + N-mask bit plane (the HBM layout of the reads)
//   orient_kernel      reverse complement of flagged reads in the packed domain (tg_tel.py:265-266)
//   scan_cov_kernel    union coverage of all k-mers of TWO k-mer lists in one pass:
//                        MODE_COUNT   -> get_telomere_base_count   (tg_kmer.py:93-102)
//                        MODE_DENSITY -> get_telomere_kmer_density (tg_kmer.py:66-90), uint8 window counts
//   hits_*_kernel      get_nonoverlapping_kmer_hits (tg_kmer.py:173-230)
//
// Matching: one Aho-Corasick automaton per k-mer list, built over the REVERSED k-mers.  A thread walks
// its chunk of positions right to left (after max_len-1 warm-up symbols), so the automaton reports at
// position p exactly the k-mers that START at p.  re.finditer's leftmost-non-overlapping rule only
// differs from "all occurrences" for k-mers that can overlap themselves ("bordered"); their candidate
// starts go to a bitmap that one lane per k-mer resolves greedily, left to right.  Coverage tiles are
// independent (a chain of overlapping candidates that crosses a tile start is walked back through global
// memory, exactly); the hits kernel walks the tiles of a slice in order and carries the blocking position.
#include "tg_common.cuh"
#include <string.h>
#include <vector>
#include <queue>
#include <string>

namespace {

constexpr int SC_MAXK = 128;               // two lists of up to 64 k-mers in one automaton
constexpr int SC_MAXLIST = 64;
constexpr int SC_MAXLEN = 32;
constexpr int SC_MAXSTATES = 4095;
constexpr int SC_MAXBORD = 32;
constexpr uint32_t SC_BORD_FLAG = 0x1000;   // transition entry: next state has self-overlapping k-mer outputs
constexpr uint32_t SC_OUT_FLAG = 0x2000;    // transition entry: next state has any output

// One automaton for one k-mer list (hits) or for two lists at once ("groups" 0 and 1: KMER_LIST and
// KMER_LIST_REV, or CANONICAL_STRINGS and their reverse complements).
struct KmerTable {
    int32_t K, n_states, max_len, n_bord;   // K = k-mers of all groups
    int32_t n_groups, KA, pad_[2];          // KA = k-mers of group 0 (they come first)
    int32_t len[SC_MAXK];
    int32_t bord_kmer[SC_MAXBORD];          // k-mer index of bordered slot b
    int32_t bord_len[SC_MAXBORD];
    int32_t bord_group[SC_MAXBORD];
    uint64_t bord_pat[SC_MAXBORD];          // the k-mer, 2 bits per base, first base in the low bits
    uint64_t super[SC_MAXLIST];             // KMER_ISSUBSTRING[k] as a bit mask (group 0 only)
    uint64_t lengt[SC_MAXLEN];              // lengt[d] = group-0 k-mers with len > d
    uint64_t bord_mask;                     // bordered group-0 k-mers
    uint64_t has_super_mask;                // group-0 k-mers with a non-empty super[]
    // transition entry: bits 0-11 next state | SC_BORD_FLAG | SC_OUT_FLAG | byte 2 = longest NON-bordered group-0 k-mer
    // starting at the next state | byte 3 = same for group 1 (so the common case needs no second table lookup)
    uint32_t trans[(SC_MAXSTATES + 1) * 4];
    uint32_t out_bord[SC_MAXSTATES + 1];    // bordered slots starting here
    uint64_t out_mask[SC_MAXSTATES + 1];    // group-0 k-mers starting here
};

// ---- packed read access ---------------------------------------------------------------------------
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
                                                   uint32_t *__restrict__ pack2, uint32_t *__restrict__ nmask)
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
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
            const bool ok = ascii_is_acgt(c) && (rel + i < L);
            p[i >> 4] |= (ok ? ascii_code(c) : 0u) << ((i & 15) * 2);
            m |= (ok ? 0u : 1u) << i;
        }
        pack2[g * 2] = p[0];
        pack2[g * 2 + 1] = p[1];
        nmask[g] = m;
    }
}

__device__ __forceinline__ uint32_t base_code(const uint32_t *pack2, int64_t b) { return (pack2[b >> 4] >> ((b & 15) * 2)) & 3u; }
__device__ __forceinline__ uint32_t base_isn(const uint32_t *nmask, int64_t b) { return (nmask[b >> 5] >> (b & 31)) & 1u; }

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
        const int64_t rel = b0 - ro;
        const int L = len[r];
        uint32_t p[2] = {0u, 0u}, m = 0u;
        for (int i = 0; i < 32; i++) {
            const int64_t q = rel + i;               // oriented position
            if (q < L) {
                const int64_t src = ro + (L - 1 - q);
                const uint32_t isn = base_isn(nmask, src);
                p[i >> 4] |= (isn ? 0u : (base_code(pack2, src) ^ 2u)) << ((i & 15) * 2);
                m |= isn << i;
            } else m |= 1u << i;
        }
        pack2_out[g * 2] = p[0];
        pack2_out[g * 2 + 1] = p[1];
        nmask_out[g] = m;
    }
}

// ---- shared pieces of the tile kernels -----------------------------------------------------------
__device__ __forceinline__ void smem_byte_max(uint8_t *arr, int idx, uint32_t val)
{
    uint32_t *word = reinterpret_cast<uint32_t *>(arr + (idx & ~3));
    const int sh = (idx & 3) * 8;
    uint32_t old = *word;
    for (;;) {
        if (((old >> sh) & 0xffu) >= val) return;
        const uint32_t assumed = old;
        old = atomicCAS(word, assumed, (assumed & ~(0xffu << sh)) | (val << sh));
        if (old == assumed) return;
    }
}

// Device-side view of one automaton held in shared memory.
struct SmemTable {
    const uint32_t *trans;
    const uint32_t *out_bord;
    const uint64_t *out_mask;     // only in hits mode
    int max_len, n_bord;
};

__device__ __forceinline__ uint8_t *carve(uint8_t *&p, size_t bytes)
{
    uint8_t *r = p;
    p += (bytes + 15) & ~(size_t)15;
    return r;
}

__device__ SmemTable load_table(const KmerTable *__restrict__ T, uint8_t *&sp, bool want_mask)
{
    const int ns = T->n_states;
    uint32_t *tr = reinterpret_cast<uint32_t *>(carve(sp, (size_t)ns * 4 * sizeof(uint32_t)));
    uint32_t *ob = reinterpret_cast<uint32_t *>(carve(sp, (size_t)ns * sizeof(uint32_t)));
    uint64_t *om = want_mask ? reinterpret_cast<uint64_t *>(carve(sp, (size_t)ns * sizeof(uint64_t))) : nullptr;
    for (int i = threadIdx.x; i < ns * 4; i += blockDim.x) tr[i] = T->trans[i];
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        ob[i] = T->out_bord[i];
        if (want_mask) om[i] = T->out_mask[i];
    }
    SmemTable S;
    S.trans = tr; S.out_bord = ob; S.out_mask = om;
    S.max_len = T->max_len; S.n_bord = T->n_bord;
    return S;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// ---- non-overlapping hits -------------------------------------------------------------------------
constexpr int HIT_THREADS = 128;
constexpr int HIT_CH = 16;
constexpr int HIT_TILE = HIT_THREADS * HIT_CH;      // 2048
constexpr int HIT_HALO = 32;
constexpr int HIT_SPAN = HIT_TILE + 2 * HIT_HALO;   // hm / C cover [t0 - 32, t0 + TILE + 32)
constexpr int HIT_STAGE = HIT_TILE + 2 * HIT_HALO;  // staged bases [t0, t0 + TILE + 64)

struct HitParams {
    const uint32_t *pack2, *nmask;
    const int64_t *base_off;
    const int32_t *slice_read, *slice_lo, *slice_hi;
    int n_slices;
    const int64_t *work_off;
    const KmerTable *tab;
    uint64_t *S;                 // survivors per work position
    uint32_t *B;                 // block-start bitmap per work position
    int64_t total_work;
    int32_t *span_slice, *span_k, *span_start, *span_end;
    long long span_cap;
    unsigned long long *span_count;
    unsigned int *next_slice;
};

// H1: selected raw matches -> superstring deletion -> survivors S[x] (tg_kmer.py:175-202)
__global__ void __launch_bounds__(HIT_THREADS) hits_survivors_kernel(const HitParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *sp = smem;
    SmemTable T = load_table(P.tab, sp, true);
    uint32_t *s_pack = reinterpret_cast<uint32_t *>(carve(sp, HIT_STAGE / 16 * 4));
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(carve(sp, HIT_STAGE / 32 * 4));
    unsigned long long *s_hm = reinterpret_cast<unsigned long long *>(carve(sp, (size_t)HIT_SPAN * 8));   // index = x - t0 + HALO
    uint32_t *s_sup = reinterpret_cast<uint32_t *>(carve(sp, (size_t)(HIT_SPAN / 32 + 1) * 4));             // hm has a superstring k-mer
    uint32_t *s_bm = reinterpret_cast<uint32_t *>(carve(sp, (size_t)max(T.n_bord, 1) * ((HIT_TILE + HIT_HALO) / 32) * 4));
    __shared__ int s_blocked[SC_MAXBORD];
    __shared__ int s_any[SC_MAXBORD];
    __shared__ unsigned long long s_shalo[HIT_HALO];
    __shared__ unsigned long long s_super[SC_MAXLIST], s_lengt[SC_MAXLEN];
    __shared__ int s_len[SC_MAXLIST];
    __shared__ unsigned int s_slice;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int BMW = (HIT_TILE + HIT_HALO) / 32;     // bitmap words per bordered slot: positions [t0, t0 + TILE + 32)
    if (tid < SC_MAXLIST) { s_super[tid] = P.tab->super[tid]; s_len[tid] = P.tab->len[tid]; }
    if (tid < SC_MAXLEN) s_lengt[tid] = P.tab->lengt[tid];
    const unsigned long long has_super = P.tab->has_super_mask;
    const int max_len = T.max_len;
    __syncthreads();
    unsigned long long m_sup = 0ull;                     // k-mers that are a superstring of some other k-mer
    for (int k = 0; k < SC_MAXLIST; k++) m_sup |= s_super[k];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_slice = atomicAdd(P.next_slice, 1u);
        __syncthreads();
        const int q = (int)s_slice;
        if (q >= P.n_slices) break;
        const int r = P.slice_read[q];
        const int lo = P.slice_lo[q], hi = P.slice_hi[q];
        const int len = hi - lo;
        const int64_t rbase = P.base_off[r];
        const int64_t gbase = rbase + lo;                        // absolute base index of slice position 0
        const int64_t woff = P.work_off[q];
        const int wlen = (int)(P.work_off[q + 1] - woff);        // padded to 64
        const int n_tiles = (wlen + HIT_TILE - 1) / HIT_TILE;
        if (tid < SC_MAXBORD) s_blocked[tid] = 0;
        for (int i = tid; i < HIT_HALO; i += HIT_THREADS) { s_hm[i] = 0ull; s_shalo[i] = 0ull; }   // left halos of the first tile

        for (int tile = 0; tile < n_tiles; tile++) {
            const int t0 = tile * HIT_TILE;                     // slice-relative
            __syncthreads();
            // stage slice positions [t0, t0 + STAGE): unaligned window of the read's packed planes, clipped to the slice
            for (int wq = tid; wq < HIT_STAGE / 32; wq += HIT_THREADS) {
                const int x0 = t0 + wq * 32;
                uint32_t m = 0xffffffffu, a = 0u, b = 0u;
                if (x0 < len) {
                    const int64_t g = gbase + x0;
                    const int64_t mw = g >> 5;
                    const int msh = (int)(g & 31);
                    const uint64_t mm = (uint64_t)P.nmask[mw] | ((uint64_t)P.nmask[mw + 1] << 32);
                    m = (uint32_t)(mm >> msh);
                    const int64_t pw = g >> 4;
                    const int sh = (int)(g & 15) * 2;
                    const uint32_t w0 = P.pack2[pw], w1 = P.pack2[pw + 1], w2 = P.pack2[pw + 2];
                    a = sh ? ((w0 >> sh) | (w1 << (32 - sh))) : w0;
                    b = sh ? ((w1 >> sh) | (w2 << (32 - sh))) : w1;
                    if (x0 + 32 > len) m |= ~((1u << (len - x0)) - 1u);
                }
                s_mask[wq] = m; s_pack[wq * 2] = a; s_pack[wq * 2 + 1] = b;
            }
            for (int i = tid; i < T.n_bord * BMW; i += HIT_THREADS) s_bm[i] = 0u;
            if (tid < SC_MAXBORD) s_any[tid] = 0;
            __syncthreads();

            // ---- automaton over [t0, t0 + TILE + HALO): hm = all k-mers starting at x ----
            for (int pass = 0; pass < 2; pass++) {
                int c0;
                if (pass == 0) c0 = tid * HIT_CH;
                else { if (tid >= HIT_HALO / HIT_CH) break; c0 = HIT_TILE + tid * HIT_CH; }
                uint32_t state = 0;
                for (int x = c0 + HIT_CH + max_len - 2; x >= c0; x--) {
                    uint32_t e = 0u;
                    if (x < HIT_STAGE) {
                        const uint32_t isn = (s_mask[x >> 5] >> (x & 31)) & 1u;
                        const uint32_t code = (s_pack[x >> 4] >> ((x & 15) * 2)) & 3u;
                        e = isn ? 0u : T.trans[state * 4 + code];
                    }
                    state = e & 0xfffu;
                    if (x < c0 + HIT_CH) {
                        unsigned long long hm = 0ull;
                        if (e & SC_OUT_FLAG) {
                            hm = T.out_mask[state];
                            uint32_t ob = (e & SC_BORD_FLAG) ? T.out_bord[state] : 0u;
                            while (ob) {
                                const int b = __ffs(ob) - 1;
                                ob &= ob - 1;
                                atomicOr(&s_bm[b * BMW + (x >> 5)], 1u << (x & 31));
                                s_any[b] = 1;
                            }
                        }
                        s_hm[HIT_HALO + x] = hm;
                    }
                }
            }
            __syncthreads();

            // ---- greedy resolution of bordered k-mers (one warp per k-mer); unselected candidates leave hm ----
            for (int b = warp; b < T.n_bord; b += HIT_THREADS / 32) {
                if (!s_any[b]) continue;
                const int blen = P.tab->bord_len[b];
                const unsigned long long kbit = 1ull << P.tab->bord_kmer[b];
                int blocked = s_blocked[b];
                int carry = -1;
                for (int ch = 0; ch < BMW; ch += 32) {
                    const uint32_t mine = (ch + lane < BMW) ? s_bm[b * BMW + ch + lane] : 0u;
                    uint32_t nz = __ballot_sync(TG_FULL_MASK, mine != 0u);
                    while (nz) {
                        const int l = __ffs(nz) - 1;
                        nz &= nz - 1;
                        const int wd = ch + l;
                        if (wd >= HIT_TILE / 32 && carry < 0) carry = blocked;       // state at the tile boundary
                        uint32_t word = __shfl_sync(TG_FULL_MASK, mine, l);
                        while (word) {
                            const int bit = __ffs(word) - 1;
                            word &= word - 1;
                            const int x = wd * 32 + bit;
                            const int p = t0 + x;
                            if (p >= blocked) blocked = p + blen;
                            else if (lane == 0) atomicAnd(&s_hm[HIT_HALO + x], ~kbit);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) s_blocked[b] = carry < 0 ? blocked : carry;
            }
            __syncthreads();

            // ---- bitmap of positions whose selected matches include a superstring k-mer ----
            for (int i = tid; i < HIT_SPAN; i += HIT_THREADS) {
                const uint32_t bal = __ballot_sync(TG_FULL_MASK, (s_hm[i] & m_sup) != 0ull);
                if (lane == 0) s_sup[i >> 5] = bal;
            }
            if (tid == 0) s_sup[HIT_SPAN / 32] = 0u;
            __syncthreads();

            // ---- survivors: drop (k, x) if a selected match of a superstring k-mer overlaps [x, x + len_k) ----
#pragma unroll 1
            for (int it = 0; it < HIT_CH; it++) {
                const int x = tid + it * HIT_THREADS;
                unsigned long long hm = s_hm[HIT_HALO + x];
                unsigned long long cand = hm & has_super;
                while (cand) {
                    const int k = __ffsll((long long)cand) - 1;
                    cand &= cand - 1;
                    const int lk = s_len[k];
                    // candidates s' in [x - max_len + 1, x + lk - 1] (indices shifted by HALO)
                    const int ib = HIT_HALO + x - max_len + 1, ie = HIT_HALO + x + lk - 1;     // inclusive
                    unsigned long long cover = 0ull;
                    for (int wd = ib >> 5; wd <= (ie >> 5); wd++) {
                        uint32_t v = s_sup[wd];
                        if (wd == (ib >> 5)) v &= 0xffffffffu << (ib & 31);
                        if (wd == (ie >> 5)) v &= 0xffffffffu >> (31 - (ie & 31));
                        while (v) {
                            const int i = wd * 32 + __ffs(v) - 1;
                            v &= v - 1;
                            const int d = HIT_HALO + x - i;                                    // x - s'
                            unsigned long long m = s_hm[i];
                            if (d > 0) m &= s_lengt[d];                                        // must reach position x
                            cover |= m;
                        }
                    }
                    if (cover & s_super[k]) hm &= ~(1ull << k);
                }
                if (t0 + x < wlen) P.S[woff + t0 + x] = hm;
            }
            __syncthreads();
            // the raw matches are no longer needed (except the halo of the next tile): overwrite them with the
            // survivors so that block starts can be decided here (tg_kmer.py:206-212): k starts a block at x unless
            // it also survived at x - len_k
            unsigned long long keep = 0ull;
            if (tid < HIT_HALO) keep = s_hm[HIT_TILE + tid];
            __syncthreads();
            // (each thread re-reads the survivors it just wrote; they are still in L2)
#pragma unroll 4
            for (int it = 0; it < HIT_CH; it++) {
                const int x = tid + it * HIT_THREADS;
                s_hm[HIT_HALO + x] = (t0 + x < wlen) ? P.S[woff + t0 + x] : 0ull;
            }
            if (tid < HIT_HALO) s_hm[tid] = s_shalo[tid];          // survivors of the previous tile's last positions
            __syncthreads();
#pragma unroll 1
            for (int it = 0; it < HIT_CH; it++) {
                const int x = tid + it * HIT_THREADS;
                unsigned long long sv = s_hm[HIT_HALO + x];
                bool start = false;
                while (sv) {
                    const int k = __ffsll((long long)sv) - 1;
                    sv &= sv - 1;
                    if (!((s_hm[HIT_HALO + x - s_len[k]] >> k) & 1ull)) start = true;
                }
                const uint32_t bal = __ballot_sync(TG_FULL_MASK, start);
                if (lane == 0 && t0 + x < wlen) P.B[(woff + t0 + x) >> 5] = bal;
            }
            __syncthreads();
            if (tid < HIT_HALO) {
                s_shalo[tid] = s_hm[HIT_TILE + tid];               // survivors at tile positions TILE-32 .. TILE-1
                s_hm[tid] = keep;                                  // raw left halo of the next tile
            }
        }
    }
}

// H2: one thread per word of the block-start plane: walk each block to its end, truncate at the next block start
// of any k-mer (tg_kmer.py:216-228), emit.  Only positions with a block start touch the survivor masks.
__global__ void __launch_bounds__(256) hits_emit_kernel(const HitParams P)
{
    const int64_t n_words = P.total_work >> 5;
    for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += (int64_t)gridDim.x * blockDim.x) {
        uint32_t bw = P.B[wi];
        if (!bw) continue;
        const int q = find_read(P.work_off, P.n_slices, wi << 5);
        const int64_t woff = P.work_off[q];
        const int wlen = (int)(P.work_off[q + 1] - woff);
        const int wend = wlen >> 5;
        while (bw) {
            const int bit = __ffs(bw) - 1;
            bw &= bw - 1;
            const int64_t g = (wi << 5) + bit;
            const int x = (int)(g - woff);
            // first set bit of B strictly after x (within this slice)
            int next_start = -1;
            {
                int wd = (x + 1) >> 5;
                uint32_t v = (wd < wend) ? (P.B[(woff >> 5) + wd] & (0xffffffffu << ((x + 1) & 31))) : 0u;
                while (wd < wend) {
                    if (v) { next_start = wd * 32 + __ffs(v) - 1; break; }
                    wd++;
                    if (wd < wend) v = P.B[(woff >> 5) + wd];
                }
            }
            unsigned long long s = P.S[g];
            while (s) {
                const int k = __ffsll((long long)s) - 1;
                s &= s - 1;
                const int lk = P.tab->len[k];
                if (x - lk >= 0 && ((P.S[g - lk] >> k) & 1ull)) continue;      // not a block start
                int e = x + lk;
                while (e < wlen && (next_start < 0 || e < next_start) && ((P.S[woff + e] >> k) & 1ull)) e += lk;
                if (next_start >= 0 && next_start < e) e = next_start;
                const unsigned long long slot = atomicAdd(P.span_count, 1ULL);
                if ((long long)slot < P.span_cap) {
                    P.span_slice[slot] = q; P.span_k[slot] = k; P.span_start[slot] = x; P.span_end[slot] = e;
                }
            }
        }
    }
}

}  // namespace

// ================================================================================================
// host: table construction
// ================================================================================================
TG_EXPORT size_t tg_kmer_table_bytes(void) { return sizeof(KmerTable); }

static int code_of(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'T': return 2; case 'G': return 3; default: return -1; }
}

// km[0 .. KA) = group 0, km[KA .. K) = group 1
static int build_table(const std::vector<std::string> &km, int KA, const int32_t *issub_edges, const int32_t *issub_off, KmerTable *T)
{
    const int K = (int)km.size();
    memset(T, 0, sizeof(KmerTable));
    T->K = K; T->KA = KA; T->n_groups = (KA < K) ? 2 : 1;
    int max_len = 0;
    for (int k = 0; k < K; k++) {
        const int l = (int)km[k].size();
        if (l < 1 || l > SC_MAXLEN) return tg_set_err(TG_EINVAL, "k-mer %d has length %d outside 1..%d", k, l, SC_MAXLEN);
        for (char c : km[k]) if (code_of(c) < 0) return tg_set_err(TG_EINVAL, "k-mer %d contains a non-ACGT character", k);
        T->len[k] = l;
        if (l > max_len) max_len = l;
        if (k < KA) for (int d = 0; d < l; d++) T->lengt[d] |= 1ull << k;
    }
    T->max_len = max_len;
    // bordered k-mers: some shift 0 < d < len maps the k-mer onto itself
    for (int k = 0; k < K; k++) {
        const std::string &s = km[k];
        bool bordered = false;
        for (size_t d = 1; d < s.size() && !bordered; d++) bordered = (s.compare(d, std::string::npos, s, 0, s.size() - d) == 0);
        if (bordered) {
            if (T->n_bord >= SC_MAXBORD) return tg_set_err(TG_EINVAL, "more than %d self-overlapping k-mers", SC_MAXBORD);
            const int b = T->n_bord++;
            T->bord_kmer[b] = k;
            T->bord_len[b] = (int)s.size();
            T->bord_group[b] = (k < KA) ? 0 : 1;
            uint64_t pat = 0;
            for (size_t i = 0; i < s.size(); i++) pat |= (uint64_t)code_of(s[i]) << (2 * i);
            T->bord_pat[b] = pat;
            if (k < KA) T->bord_mask |= 1ull << k;
        }
    }
    if (issub_edges && issub_off)
        for (int k = 0; k < KA; k++) {
            for (int e = issub_off[k]; e < issub_off[k + 1]; e++) {
                if (issub_edges[e] < 0 || issub_edges[e] >= KA) return tg_set_err(TG_EINVAL, "bad KMER_ISSUBSTRING edge");
                T->super[k] |= 1ull << issub_edges[e];
            }
            if (T->super[k]) T->has_super_mask |= 1ull << k;
        }
    // Aho-Corasick over the reversed k-mers; out = set of k-mers (two 64-bit halves, one per group)
    struct Node { int next[4]; int fail; uint64_t out[2]; };
    std::vector<Node> nodes(1);
    nodes[0] = Node{{-1, -1, -1, -1}, 0, {0ull, 0ull}};
    for (int k = 0; k < K; k++) {
        int cur = 0;
        for (int i = (int)km[k].size() - 1; i >= 0; i--) {
            const int c = code_of(km[k][i]);
            if (nodes[cur].next[c] < 0) {
                nodes[cur].next[c] = (int)nodes.size();
                nodes.push_back(Node{{-1, -1, -1, -1}, 0, {0ull, 0ull}});
            }
            cur = nodes[cur].next[c];
        }
        if (k < KA) nodes[cur].out[0] |= 1ull << k; else nodes[cur].out[1] |= 1ull << (k - KA);
    }
    if ((int)nodes.size() > SC_MAXSTATES) return tg_set_err(TG_EINVAL, "automaton has %d states (max %d)", (int)nodes.size(), SC_MAXSTATES);
    std::queue<int> bfs;
    for (int c = 0; c < 4; c++) {
        int v = nodes[0].next[c];
        if (v < 0) nodes[0].next[c] = 0;
        else { nodes[v].fail = 0; bfs.push(v); }
    }
    while (!bfs.empty()) {
        const int u = bfs.front(); bfs.pop();
        nodes[u].out[0] |= nodes[nodes[u].fail].out[0];
        nodes[u].out[1] |= nodes[nodes[u].fail].out[1];
        for (int c = 0; c < 4; c++) {
            const int v = nodes[u].next[c];
            if (v < 0) nodes[u].next[c] = nodes[nodes[u].fail].next[c];
            else { nodes[v].fail = nodes[nodes[u].fail].next[c]; bfs.push(v); }
        }
    }
    T->n_states = (int)nodes.size();
    std::vector<uint32_t> out_len(nodes.size(), 0u);     // byte 0: group 0, byte 1: group 1
    for (int s = 0; s < T->n_states; s++) {
        T->out_mask[s] = nodes[s].out[0];
        int best[2] = {0, 0};
        uint32_t ob = 0;
        for (int k = 0; k < K; k++) {
            const int grp = (k < KA) ? 0 : 1;
            if (!((nodes[s].out[grp] >> (grp ? k - KA : k)) & 1ull)) continue;
            bool is_bord = false;
            for (int b = 0; b < T->n_bord; b++) if (T->bord_kmer[b] == k) { ob |= 1u << b; is_bord = true; }
            if (!is_bord && T->len[k] > best[grp]) best[grp] = T->len[k];
        }
        out_len[s] = (uint32_t)best[0] | ((uint32_t)best[1] << 8);
        T->out_bord[s] = ob;
    }
    for (int s = 0; s < T->n_states; s++)
        for (int c = 0; c < 4; c++) {
            const int v = nodes[s].next[c];
            T->trans[s * 4 + c] = (uint32_t)v | (T->out_bord[v] ? SC_BORD_FLAG : 0u) |
                                  ((nodes[v].out[0] | nodes[v].out[1]) ? SC_OUT_FLAG : 0u) | (out_len[v] << 16);
        }
    return TG_OK;
}

static int collect_kmers(const char *concat, const int32_t *off, int K, std::vector<std::string> &km)
{
    if (!concat || !off) return tg_set_err(TG_EINVAL, "NULL argument");
    if (K < 1 || K > SC_MAXLIST) return tg_set_err(TG_EINVAL, "K=%d outside 1..%d", K, SC_MAXLIST);
    for (int k = 0; k < K; k++) {
        const int l = off[k + 1] - off[k];
        if (l < 1 || l > SC_MAXLEN) return tg_set_err(TG_EINVAL, "k-mer %d has length %d outside 1..%d", k, l, SC_MAXLEN);
        km.emplace_back(concat + off[k], (size_t)l);
    }
    return TG_OK;
}

TG_EXPORT int tg_kmer_table_build(const char *kmers_concat, const int32_t *kmer_off, int K,
                                  const int32_t *issub_edges, const int32_t *issub_off, void *h_table)
{
    if (!h_table) return tg_set_err(TG_EINVAL, "NULL argument");
    std::vector<std::string> km;
    int rc = collect_kmers(kmers_concat, kmer_off, K, km);
    if (rc != TG_OK) return rc;
    return build_table(km, K, issub_edges, issub_off, reinterpret_cast<KmerTable *>(h_table));
}

// ================================================================================================
// launches
// ================================================================================================
static int grid_for(int64_t items, int block, int sms, int per_sm)
{
    int64_t g = (items + block - 1) / block;
    const int64_t cap = (int64_t)sms * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// host copies of the per-list sizes are needed to size shared memory: read them back once per call
static int table_dims(const void *d_table, int *n_states, int *n_bord, cudaStream_t st)
{
    int32_t hdr[4];
    TG_CUDA_CHECK(cudaMemcpyAsync(hdr, d_table, sizeof(hdr), cudaMemcpyDeviceToHost, st));
    TG_CUDA_CHECK(cudaStreamSynchronize(st));
    *n_states = hdr[1];
    *n_bord = hdr[3];
    if (hdr[1] < 1 || hdr[1] > SC_MAXSTATES || hdr[3] < 0 || hdr[3] > SC_MAXBORD) return tg_set_err(TG_EINVAL, "corrupt k-mer table");
    return TG_OK;
}

static size_t table_smem(int n_states, bool want_mask)
{
    auto up = [](size_t b) { return (b + 15) & ~(size_t)15; };
    return up((size_t)n_states * 16) + up((size_t)n_states * 4) + (want_mask ? up((size_t)n_states * 8) : 0);
}

TG_EXPORT int tg_scan_pack(const uint8_t *d_ascii, const int64_t *d_base_off, const int32_t *d_len, int n_reads,
                           uint32_t *d_pack2, uint32_t *d_nmask, void *stream)
{
    if (n_reads <= 0) return TG_OK;
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    pack_kernel<<<sms * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_ascii, d_base_off, d_len, n_reads, d_pack2, d_nmask);
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

TG_EXPORT size_t tg_scan_hits_work_bytes(int64_t total_work_bases, int n_slices)
{
    (void)n_slices;
    const size_t t = (size_t)((total_work_bases + 63) / 64 * 64);
    return t * 8 + t / 8 + 256;
}

TG_EXPORT int tg_scan_hits(const uint32_t *d_pack2, const uint32_t *d_nmask, const int64_t *d_base_off,
                           const int32_t *d_slice_read, const int32_t *d_slice_lo, const int32_t *d_slice_hi, int n_slices,
                           const int64_t *d_slice_work_off, int64_t total_work_bases, const void *d_table, void *d_work,
                           int32_t *d_span_slice, int32_t *d_span_k, int32_t *d_span_start, int32_t *d_span_end,
                           int64_t span_cap, unsigned long long *d_span_count, void *stream)
{
    if (n_slices <= 0) return TG_OK;
    if (!d_pack2 || !d_nmask || !d_base_off || !d_slice_read || !d_slice_lo || !d_slice_hi || !d_slice_work_off || !d_table ||
        !d_work || !d_span_slice || !d_span_k || !d_span_start || !d_span_end || !d_span_count)
        return tg_set_err(TG_EINVAL, "NULL device pointer");
    if (total_work_bases % 64) return tg_set_err(TG_EINVAL, "total_work_bases must be a multiple of 64");
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int ns, nb;
    int rc = table_dims(d_table, &ns, &nb, st);
    if (rc != TG_OK) return rc;
    auto up = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t smem = table_smem(ns, true) + up(HIT_STAGE / 16 * 4) + up(HIT_STAGE / 32 * 4) + up((size_t)HIT_SPAN * 8) +
                        up((size_t)(HIT_SPAN / 32 + 1) * 4) + up((size_t)(nb > 1 ? nb : 1) * ((HIT_TILE + HIT_HALO) / 32) * 4);
    if (smem > 200 * 1024) return tg_set_err(TG_ERANGE, "k-mer table too large for shared memory (%zu bytes)", smem);
    HitParams P;
    P.pack2 = d_pack2; P.nmask = d_nmask; P.base_off = d_base_off;
    P.slice_read = d_slice_read; P.slice_lo = d_slice_lo; P.slice_hi = d_slice_hi; P.n_slices = n_slices;
    P.work_off = d_slice_work_off; P.tab = reinterpret_cast<const KmerTable *>(d_table);
    P.S = reinterpret_cast<uint64_t *>(d_work);
    P.B = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(d_work) + (size_t)total_work_bases * 8);
    P.total_work = total_work_bases;
    P.span_slice = d_span_slice; P.span_k = d_span_k; P.span_start = d_span_start; P.span_end = d_span_end;
    P.span_cap = span_cap; P.span_count = d_span_count;
    rc = get_counter(&P.next_slice, st);
    if (rc != TG_OK) return rc;
    TG_CUDA_CHECK(cudaMemsetAsync(d_span_count, 0, sizeof(unsigned long long), st));
    TG_CUDA_CHECK(cudaFuncSetAttribute(hits_survivors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220 * 1024) / (smem + 4096));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = sms * per_sm;
    if (grid > n_slices) grid = n_slices;
    hits_survivors_kernel<<<grid, HIT_THREADS, smem, st>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    const int g2 = grid_for(total_work_bases / 32, 256, sms, 16);
    hits_emit_kernel<<<g2, 256, 0, st>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}
