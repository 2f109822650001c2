// Stage [0a] telomere-read prefilter (reference telogator2.py:451-459):
//     count_fwd = read.count(KMER_INITIAL);  count_rev = read.count(KMER_INITIAL_RC)
// i.e. str.count -- the number of NON-OVERLAPPING occurrences found scanning left to right -- of two equally long
// ACGT patterns (canonical repeat twice, and its reverse complement) in every input read.  This is the one
// HBM-scale scan of the product (every read of a whole-genome run goes through it): 1 byte of ASCII per base in,
// two integers per read out.
//
// One warp per read (dynamic scheduling), 992 positions per iteration: lane l owns 32 consecutive bases (two 128-bit
// loads, the next iteration's loads are issued before this one's arithmetic); lane 31 only provides look-ahead.
//   1. bit planes: bit 1 and bit 2 of an ASCII letter are a 2-bit code of A/C/G/T (A=00 C=01 T=10 G=11); a multiply
//      gathers the four bits of a 32-bit word into its top nibble (IMAD, FMA pipe) and a funnel shift appends it.
//   2. bit-parallel matching: for pattern offset t the planes are funnel-shifted by t (shared by both patterns) and
//      mismatches accumulated with one LOP3 per plane and pattern: acc |= plane_t ^ expected_t.  ~acc = candidates for
//      32 positions at once.
//   3. candidates are rare outside telomeres; each is verified against the ASCII letters (the 2-bit code ignores the other
//      six bits, so 'c', 'N', ... can alias) and the survivors go through str.count's greedy rule with the carry
//      "next free position" kept per pattern.
// Algorithmic bytes: 1 B/base (SURVEY 8d style); the kernel is bound by HBM bandwidth.
#include "tg_common.cuh"

namespace {

constexpr int PF_WARPS = 8;
constexpr int PF_STRIDE = 31 * 32;          // positions per warp iteration
constexpr int PF_MAXK = 32;                 // pattern length limit (window must fit the 32-base look-ahead of the next lane)

struct PrefParams {
    const uint8_t *ascii;
    const long long *base_off;
    const int *len;
    int n_reads, klen;
    uint32_t x1[2][PF_MAXK], x2[2][PF_MAXK];   // expected plane values per pattern and offset: 0 or ~0
    uint8_t pat[2][PF_MAXK];
    int32_t *cnt_a, *cnt_b;
    unsigned long long *counter;
};

// bit `b` (1 or 2) of the 32 letters held in w[0..7] -> one 32-bit plane, letter i at bit i
template <int B>
__device__ __forceinline__ uint32_t plane(const uint32_t (&w)[8])
{
    // bits B, B+8, B+16, B+24 of a word -> bits 28..31 of the product (no carries: the cross terms fall on distinct lower bits)
    constexpr uint32_t M = (1u << (28 - B)) | (1u << (29 - B - 8)) | (1u << (30 - B - 16)) | (1u << (31 - B - 24));
    uint32_t acc = 0;
#pragma unroll
    for (int k = 7; k >= 0; k--) acc = __funnelshift_l((w[k] & (0x01010101u << B)) * M, acc, 4);
    return acc;
}

// KLEN > 0: pattern length known at compile time (the matching loop unrolls and the expected planes become
// constant-bank operands); KLEN == 0: any length up to PF_MAXK.
// DOUBLED: both patterns are a word repeated twice (KMER_INITIAL = canonical + canonical, telogator2.py:393-394): a match at p
// is a match of the half at p and at p + KLEN/2, which halves the matching chain.
template <int KLEN, bool DOUBLED>
__global__ void __launch_bounds__(PF_WARPS * 32) prefilter_count_kernel(const __grid_constant__ PrefParams P)
{
    const int lane = threadIdx.x & 31;
    const int klen = KLEN ? KLEN : P.klen;
    for (;;) {
        unsigned long long r64 = 0;
        if (lane == 0) r64 = atomicAdd(P.counter, 1ULL);
        r64 = __shfl_sync(TG_FULL_MASK, r64, 0);
        if (r64 >= (unsigned long long)P.n_reads) break;
        const int r = (int)r64;
        const uint8_t *base = P.ascii + P.base_off[r];
        const int L = P.len[r];
        const int last = L - klen;                           // last valid start position
        int cnt[2] = {0, 0}, nf[2] = {0, 0};
        if (last >= 0) {
            const int n_iter = (last + PF_STRIDE) / PF_STRIDE;
            const uint4 *src = reinterpret_cast<const uint4 *>(base) + 2 * lane;
            uint4 n0 = __ldcs(src), n1 = __ldcs(src + 1);
            for (int it = 0; it < n_iter; it++) {
                const uint32_t w[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                if (it + 1 < n_iter) {
                    const uint4 *nx = src + (size_t)(it + 1) * (PF_STRIDE / 16);
                    n0 = __ldcs(nx);
                    n1 = __ldcs(nx + 1);
                }
                const uint32_t b1 = plane<1>(w), b2 = plane<2>(w);
                const uint32_t b1n = __shfl_down_sync(TG_FULL_MASK, b1, 1), b2n = __shfl_down_sync(TG_FULL_MASK, b2, 1);
                uint32_t acc0 = 0, acc1 = 0;
                constexpr int CHAIN = DOUBLED ? KLEN / 2 : KLEN;
#pragma unroll
                for (int t = 0; t < (KLEN ? CHAIN : klen); t++) {
                    const uint32_t s1 = __funnelshift_r(b1, b1n, t), s2 = __funnelshift_r(b2, b2n, t);
                    acc0 |= (s1 ^ P.x1[0][t]) | (s2 ^ P.x2[0][t]);
                    acc1 |= (s1 ^ P.x1[1][t]) | (s2 ^ P.x2[1][t]);
                }
                if (DOUBLED) {
                    // acc = mismatch mask of the half word; the whole pattern also needs the half at + KLEN / 2
                    const uint32_t a0n = __shfl_down_sync(TG_FULL_MASK, acc0, 1), a1n = __shfl_down_sync(TG_FULL_MASK, acc1, 1);
                    acc0 |= __funnelshift_r(acc0, a0n, CHAIN);
                    acc1 |= __funnelshift_r(acc1, a1n, CHAIN);
                }
                // positions of this lane: it * 992 + 32 * lane + i; lane 31 is look-ahead only
                const int p0 = it * PF_STRIDE + 32 * lane;
                uint32_t ok = 0;
                if (lane < 31 && p0 <= last) ok = (last - p0 >= 31) ? 0xffffffffu : ((2u << (last - p0)) - 1u);
                uint32_t cand[2] = {~acc0 & ok, ~acc1 & ok};
                if (__any_sync(TG_FULL_MASK, (cand[0] | cand[1]) != 0u)) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        // exact check of every candidate against the letters
                        uint32_t m = cand[q], good = 0;
                        while (m) {
                            const int i = __ffs(m) - 1;
                            m &= m - 1;
                            const uint8_t *s = base + p0 + i;
                            bool eq = true;
                            for (int t = 0; t < klen; t++) eq = eq && (s[t] == P.pat[q][t]);
                            if (eq) good |= 1u << i;
                        }
                        // str.count: greedy, left to right, non-overlapping
                        unsigned lanes = __ballot_sync(TG_FULL_MASK, good != 0u);
                        while (lanes) {
                            const int l = __ffs(lanes) - 1;
                            lanes &= lanes - 1;
                            uint32_t g = __shfl_sync(TG_FULL_MASK, good, l);
                            const int pb = it * PF_STRIDE + 32 * l;
                            // jump from taken match to taken match: bits below the next free position are dropped at once
                            if (nf[q] > pb) g = (nf[q] - pb >= 32) ? 0u : (g & (0xffffffffu << (nf[q] - pb)));
                            while (g) {
                                const int i = __ffs(g) - 1;
                                cnt[q]++;
                                nf[q] = pb + i + klen;
                                g = (i + klen >= 32) ? 0u : (g & (0xffffffffu << (i + klen)));
                            }
                        }
                    }
                }
            }
        }
        if (lane == 0) { P.cnt_a[r] = cnt[0]; P.cnt_b[r] = cnt[1]; }
    }
}

}  // namespace

TG_EXPORT int tg_prefilter_tail_bytes(void) { return 2048; }

TG_EXPORT int tg_prefilter_count(const uint8_t *d_ascii, const int64_t *d_base_off, const int32_t *d_len, int n_reads,
                                 const char *kmer_a, const char *kmer_b, int klen, int32_t *d_count_a, int32_t *d_count_b,
                                 void *d_counter, void *stream)
{
    if (n_reads <= 0) return TG_OK;
    if (!d_ascii || !d_base_off || !d_len || !kmer_a || !kmer_b || !d_count_a || !d_count_b || !d_counter)
        return tg_set_err(TG_EINVAL, "NULL pointer");
    if (klen < 1 || klen > PF_MAXK) return tg_set_err(TG_ERANGE, "pattern length %d outside 1..%d", klen, PF_MAXK);
    const int sms = tg_sm_count();
    if (sms <= 0) return tg_set_err(TG_ECUDA, "no CUDA device");
    PrefParams P;
    memset(&P, 0, sizeof(P));
    const char *pats[2] = {kmer_a, kmer_b};
    for (int q = 0; q < 2; q++)
        for (int t = 0; t < klen; t++) {
            const unsigned char c = (unsigned char)pats[q][t];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return tg_set_err(TG_EINVAL, "pattern %d has a letter outside ACGT", q);
            P.pat[q][t] = c;
            P.x1[q][t] = ((c >> 1) & 1u) ? 0xffffffffu : 0u;
            P.x2[q][t] = ((c >> 2) & 1u) ? 0xffffffffu : 0u;
        }
    P.ascii = d_ascii; P.base_off = reinterpret_cast<const long long *>(d_base_off); P.len = d_len; P.n_reads = n_reads; P.klen = klen;
    P.cnt_a = d_count_a; P.cnt_b = d_count_b; P.counter = reinterpret_cast<unsigned long long *>(d_counter);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TG_CUDA_CHECK(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    long long g = ((long long)n_reads + PF_WARPS - 1) / PF_WARPS;
    if (g > (long long)sms * 8) g = (long long)sms * 8;
    bool doubled = (klen % 2) == 0;
    for (int q = 0; q < 2 && doubled; q++)
        for (int t = 0; t < klen / 2; t++) doubled = doubled && pats[q][t] == pats[q][t + klen / 2];
    if (klen == 12 && doubled) prefilter_count_kernel<12, true><<<(int)g, PF_WARPS * 32, 0, st>>>(P);        // human, C. elegans: 6-mer twice
    else if (klen == 14 && doubled) prefilter_count_kernel<14, true><<<(int)g, PF_WARPS * 32, 0, st>>>(P);   // maize: 7-mer twice
    else if (klen == 12) prefilter_count_kernel<12, false><<<(int)g, PF_WARPS * 32, 0, st>>>(P);
    else prefilter_count_kernel<0, false><<<(int)g, PF_WARPS * 32, 0, st>>>(P);
    TG_CUDA_CHECK(cudaGetLastError());
    return TG_OK;
}
