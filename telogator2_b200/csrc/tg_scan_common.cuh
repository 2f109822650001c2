// Shared pieces of the k-mer scan translation units (tg_scan.cu: pack / orient / hits, tg_cov.cu: coverage).
//
// HBM layout of the reads (tg_scan_pack): one padded base space, read r at [base_off[r], base_off[r] + len[r]),
// base_off[r] a multiple of 128 and at least ONE padding position after every read (padding carries the N bit), so a
// match can never straddle two reads and the coverage kernels may tile the base space without looking at read borders.
//   pack2[w]  bases 16w .. 16w+15, 2 bits each, first base in the low bits: A=0 C=1 T=2 G=3 ((ascii >> 1) & 3)
//   nmask[w]  bases 32w .. 32w+31, bit set = not ACGT (N, lower case, padding)
//
// Matching is table driven: the next SC_W = 7 bases of a position (14 bits) index a lookup table that answers for all
// "plain" k-mers (not longer than the window, not self-overlapping) at once.  "Special" k-mers -- longer than the window
// or self-overlapping, where re.finditer's leftmost-non-overlapping rule differs from "all occurrences" -- only raise a
// flag when their first min(len, 7) letters match; flagged positions (rare) are verified against the packed planes.
#pragma once
#include "tg_common.cuh"
#include <string>
#include <vector>

constexpr int SC_W = 7;                        // window length of the lookup tables
constexpr int SC_TAB = 1 << (2 * SC_W);        // 16384 windows
constexpr int SC_MAXK = 128;                   // two lists of up to 64 k-mers
constexpr int SC_MAXLIST = 64;
constexpr int SC_MAXLEN = 32;
constexpr int SC_MAXBORD = 32;
enum { SC_PLAIN = 0, SC_SPECIAL = 1, SC_BORDERED = 2, SC_REDUNDANT = 3 };

struct ScPlanes {
    const uint32_t *pack2, *nmask;
    int64_t total;                             // padded bases; the arrays have >= 4 words of slack behind them
};

// 32 bases starting at absolute base g (2 bits each)
__device__ __forceinline__ uint64_t sc_bases64(const uint32_t *__restrict__ pack2, int64_t g)
{
    const int64_t pw = g >> 4;
    const int sh = (int)(g & 15) * 2;
    uint64_t bits = ((uint64_t)pack2[pw] | ((uint64_t)pack2[pw + 1] << 32)) >> sh;
    if (sh) bits |= (uint64_t)pack2[pw + 2] << (64 - sh);
    return bits;
}

// N bits of the 32 bases starting at g
__device__ __forceinline__ uint32_t sc_nbits32(const uint32_t *__restrict__ nmask, int64_t g)
{
    const int64_t mw = g >> 5;
    const uint64_t v = (uint64_t)nmask[mw] | ((uint64_t)nmask[mw + 1] << 32);
    return (uint32_t)(v >> (int)(g & 31));
}

// does the k-mer (pat: 2 bits per base, first base low; len <= 32) start at absolute base g?
__device__ __forceinline__ bool sc_kmer_at(const ScPlanes &P, int64_t g, uint64_t pat, int len)
{
    if (g < 0 || g + len > P.total) return false;
    const uint32_t lm = (len >= 32) ? 0xffffffffu : ((1u << len) - 1u);
    if (sc_nbits32(P.nmask, g) & lm) return false;
    const uint64_t m = (len >= 32) ? ~0ull : ((1ull << (2 * len)) - 1ull);
    return ((sc_bases64(P.pack2, g) ^ pat) & m) == 0ull;
}

// Exact state of re.finditer's greedy rule in front of candidate c0 (absolute) of a self-overlapping k-mer: walk the
// chain of overlapping candidates back to its head (no other candidate within blen - 1 bases before it), then replay
// forward.  Returns the absolute position before which the k-mer is blocked (LLONG_MIN-ish if there is no chain).
static __device__ __noinline__ int64_t sc_chain_state_before(const ScPlanes &P, int64_t c0, uint64_t pat, int blen)
{
    int64_t head = c0, p = c0 - 1;
    while (p >= 0 && head - p < blen) {
        if (sc_kmer_at(P, p, pat, blen)) head = p;
        p--;
    }
    if (head == c0) return -(1ll << 60);
    int64_t blocked = head + blen;
    for (int64_t q = head + 1; q < c0; q++)
        if (q >= blocked && sc_kmer_at(P, q, pat, blen)) blocked = q + blen;
    return blocked;
}

// ---- host helpers (table construction) -------------------------------------------------------------------------
static inline int sc_code_of(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'T': return 2; case 'G': return 3; default: return -1; }
}

static inline bool sc_is_bordered(const std::string &s)
{
    for (size_t d = 1; d < s.size(); d++)
        if (s.compare(d, std::string::npos, s, 0, s.size() - d) == 0) return true;
    return false;
}

static inline uint64_t sc_pattern(const std::string &s)
{
    uint64_t pat = 0;
    for (size_t i = 0; i < s.size(); i++) pat |= (uint64_t)sc_code_of(s[i]) << (2 * i);
    return pat;
}

// is `s` (already classified) a prefix of window w?  Only the first min(len, SC_W) letters are compared.
static inline bool sc_window_has_prefix(uint32_t w, const std::string &s)
{
    const int n = (int)s.size() < SC_W ? (int)s.size() : SC_W;
    for (int t = 0; t < n; t++)
        if ((int)((w >> (2 * t)) & 3u) != sc_code_of(s[t])) return false;
    return true;
}

static inline int sc_collect_kmers(const char *concat, const int32_t *off, int K, std::vector<std::string> &km)
{
    if (!concat || !off) return tg_set_err(TG_EINVAL, "NULL argument");
    if (K < 1 || K > SC_MAXLIST) return tg_set_err(TG_EINVAL, "K=%d outside 1..%d", K, SC_MAXLIST);
    for (int k = 0; k < K; k++) {
        const int l = off[k + 1] - off[k];
        if (l < 1 || l > SC_MAXLEN) return tg_set_err(TG_EINVAL, "k-mer %d has length %d outside 1..%d", k, l, SC_MAXLEN);
        km.emplace_back(concat + off[k], (size_t)l);
        for (char c : km.back())
            if (sc_code_of(c) < 0) return tg_set_err(TG_EINVAL, "k-mer %d contains a non-ACGT character", k);
    }
    return TG_OK;
}
