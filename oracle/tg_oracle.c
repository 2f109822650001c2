/*
 * tg_oracle.c -- CPU restatement of Telogator2's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (telogator2_b200/) never does.
 *
 * Every function cites the reference file:line (relative to the upstream repository
 * zstephens/telogator2) that it restates.  Pinning status:
 *   - k-mer scan functions (finditer/base count/density/non-overlapping hits): PINNED.
 *     Checked against outputs of the reference's own source/tg_kmer.py (imported in the
 *     build container, fixtures under tests/golden/, generator tests/golden/make_golden.py).
 *   - MT19937 / random.sample: PINNED against CPython's stdlib `random` (the reference's RNG).
 *   - NW score (Biopython PairwiseAligner.score): PARITY UNPINNED.  The arithmetic lives in
 *     the third-party dependency biopython>=1.86 (Bio/Align/_pairwisealigner.c), which is
 *     not vendored in the reference and not installable here.  tgo_nw_score restates its
 *     published Needleman-Wunsch score path (linear gaps, left/right end-gap classes) as
 *     used from tg_align.py:71-106,125,142.
 *   - everything AROUND the score in tvr_distance / get_dist_matrix_parallel (tg_align.py:109-189: length
 *     adjustment, identity score, early-out, random.seed + shuffle order, mean, -log ratio, clamps, per-pair
 *     seeds, float32 fill): PINNED against the reference's own code, run in the build container with a stand-in
 *     aligner whose score() is tgo_nw_score (tests/golden/make_golden_dp_ref.py -> dp_ref_golden.npz).
 *   - colour-vector preparation (tg_tvr.py): PINNED (tests/golden/make_golden_tvr.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TGO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * MT19937 exactly as CPython's Modules/_randommodule.c drives it for
 *   random.seed(int)  (tg_align.py:110-113)  and  random.sample(s, len(s))  (tg_util.py:451-452)
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int mti; } tgo_mt;

static void mt_init_genrand(tgo_mt *st, uint32_t s)
{
    st->mt[0] = s;
    for (int i = 1; i < 624; i++)
        st->mt[i] = 1812433253u * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (uint32_t)i;
    st->mti = 624;
}

static void mt_init_by_array(tgo_mt *st, const uint32_t *key, int len)
{
    mt_init_genrand(st, 19650218u);
    uint32_t *mt = st->mt;
    int i = 1, j = 0, k = (624 > len) ? 624 : len;
    for (; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
}

static uint32_t mt_genrand(tgo_mt *st)
{
    uint32_t *mt = st->mt, y;
    if (st->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        st->mti = 0;
    }
    y = mt[st->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* random.seed(a) for a Python int: key = 32-bit little-endian words of abs(a), >= 1 word. */
static void mt_seed_u64(tgo_mt *st, uint64_t a)
{
    uint32_t key[2] = { (uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32) };
    mt_init_by_array(st, key, key[1] ? 2 : 1);
}

/* Random._randbelow_with_getrandbits: k = n.bit_length(); r = getrandbits(k); while r >= n: redraw */
static uint32_t mt_randbelow(tgo_mt *st, uint32_t n)
{
    int k = 32 - __builtin_clz(n);
    uint32_t r = mt_genrand(st) >> (32 - k);
    while (r >= n) r = mt_genrand(st) >> (32 - k);
    return r;
}

/* ''.join(random.sample(s, len(s))): pool branch of Random.sample (always taken when k == n). */
static void mt_shuffle_bytes(tgo_mt *st, const uint8_t *s, int n, uint8_t *out, uint8_t *pool)
{
    memcpy(pool, s, (size_t)n);
    for (int i = 0; i < n; i++) {
        uint32_t j = mt_randbelow(st, (uint32_t)(n - i));
        out[i] = pool[j];
        pool[j] = pool[n - i - 1];
    }
}

TGO_API void tgo_mt_seed(void *state, uint64_t seed) { mt_seed_u64((tgo_mt *)state, seed); }
TGO_API size_t tgo_mt_state_bytes(void) { return sizeof(tgo_mt); }
TGO_API uint32_t tgo_mt_genrand(void *state) { return mt_genrand((tgo_mt *)state); }
/* the 624-word table init_genrand(19650218) leaves behind (input of every init_by_array) */
TGO_API void tgo_mt_base_table(uint32_t *out624)
{
    tgo_mt st; mt_init_genrand(&st, 19650218u); memcpy(out624, st.mt, sizeof(st.mt));
}
/* shuffle `count` byte strings back to back from ONE stream seeded with `seed`
 * (exactly what tvr_distance does: seed once, then shuffle_seq(seq_i), shuffle_seq(seq_j), ...) */
TGO_API void tgo_shuffle_stream(uint64_t seed, int count, const uint8_t *const *seqs, const int32_t *lens,
                                uint8_t *const *outs)
{
    tgo_mt st; mt_seed_u64(&st, seed);
    int maxl = 1;
    for (int c = 0; c < count; c++) if (lens[c] > maxl) maxl = lens[c];
    uint8_t *pool = (uint8_t *)malloc((size_t)maxl);
    for (int c = 0; c < count; c++) mt_shuffle_bytes(&st, seqs[c], lens[c], outs[c], pool);
    free(pool);
}

/* ------------------------------------------------------------------------------------------
 * Needleman-Wunsch score: Biopython PairwiseAligner.score() in global mode when every
 * open gap score equals its extend gap score (tg_align.py:71-106 'tvr' -> -4/-4, end gaps 0/0
 * or -4/-4).  Row DP in doubles, pure max, horizontal moves in the last row and vertical
 * moves in the last column cost the RIGHT end-gap score; first row / column cost the LEFT one.
 * S == NULL  ->  match/mismatch compare on the raw codes (no wildcard).
 * ---------------------------------------------------------------------------------------- */
static double nw_score_d(const uint8_t *a, int n, const uint8_t *b, int m, const double *S, int A,
                         double match, double mismatch, double g, double gl, double gr, double *row)
{
#define SUB(x, y) (S ? S[(size_t)(x) * A + (y)] : ((x) == (y) ? match : mismatch))
    int i, j;
    double score = 0.0, temp, t;
    row[0] = 0.0;
    for (j = 1; j <= m; j++) row[j] = j * gl;
    for (i = 1; i < n; i++) {
        int ka = a[i - 1];
        temp = row[0];
        row[0] = i * gl;
        for (j = 1; j < m; j++) {
            score = temp + SUB(ka, b[j - 1]);
            t = row[j] + g;      if (t > score) score = t;
            t = row[j - 1] + g;  if (t > score) score = t;
            temp = row[j]; row[j] = score;
        }
        score = temp + SUB(ka, b[m - 1]);
        t = row[m] + gr;         if (t > score) score = t;
        t = row[m - 1] + g;      if (t > score) score = t;
        temp = row[m]; row[m] = score;
    }
    {
        int ka = a[n - 1];
        temp = row[0];
        row[0] = n * gl;
        for (j = 1; j < m; j++) {
            score = temp + SUB(ka, b[j - 1]);
            t = row[j] + g;      if (t > score) score = t;
            t = row[j - 1] + gr; if (t > score) score = t;
            temp = row[j]; row[j] = score;
        }
        score = temp + SUB(ka, b[m - 1]);
        t = row[m] + gr;         if (t > score) score = t;
        t = row[m - 1] + gr;     if (t > score) score = t;
    }
#undef SUB
    return score;
}

TGO_API double tgo_nw_score(const uint8_t *a, int n, const uint8_t *b, int m, const double *S, int A,
                            double match, double mismatch, double g, double gl, double gr)
{
    if (n <= 0 || m <= 0) return NAN;   /* Biopython raises ValueError: sequence has zero length */
    double *row = (double *)malloc(sizeof(double) * (size_t)(m + 1));
    double r = nw_score_d(a, n, b, m, S, A, match, mismatch, g, gl, gr, row);
    free(row);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * tvr_distance (tg_align.py:109-151) and get_dist_matrix_parallel (tg_align.py:154-189).
 * codes: all sequences concatenated (uint8 codes, index into the A x A matrix when S != NULL);
 * off[n_seq+1].  Pair p (combinations order) uses seed rng_seed + p.
 * Outputs (any may be NULL): dist [n_seq*n_seq] float32 (symmetric, zero diagonal, NOT /10),
 *   aln [n_pairs] int32, rnd [n_pairs*R] int32 (INT32_MIN where not computed),
 *   cells: sum over pairs of n'*m'*(1 + R*[not early-out]).
 * pair_begin/pair_end restrict the work to a range of pair indices (bounded CPU samples).
 * ---------------------------------------------------------------------------------------- */
TGO_API int tgo_dist_matrix(const uint8_t *codes, const int64_t *off, int n_seq,
                            const double *S, int A, double match, double mismatch,
                            double g, double gl, double gr,
                            int adjust_lens, int min_viable, int R, int64_t rng_seed,
                            int64_t pair_begin, int64_t pair_end,
                            float *dist, int32_t *aln_out, int32_t *rnd_out, double *cells_out,
                            int n_threads)
{
    const int64_t n_pairs = (int64_t)n_seq * (n_seq - 1) / 2;
    if (pair_end < 0 || pair_end > n_pairs) pair_end = n_pairs;
    if (pair_begin < 0) pair_begin = 0;
    int maxl = 1;
    for (int i = 0; i < n_seq; i++) { int l = (int)(off[i + 1] - off[i]); if (l > maxl) maxl = l; }
    double cells_total = 0.0;
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : cells_total) reduction(| : bad)
    {
        double *row = (double *)malloc(sizeof(double) * (size_t)(maxl + 1));
        uint8_t *si = (uint8_t *)malloc((size_t)maxl), *sj = (uint8_t *)malloc((size_t)maxl);
        uint8_t *pool = (uint8_t *)malloc((size_t)maxl);
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = pair_begin; p < pair_end; p++) {
            /* invert p -> (i, j) in itertools.combinations(range(n), 2) order (tg_align.py:156) */
            int64_t i = (int64_t)floor(((2.0 * n_seq - 1) - sqrt((2.0 * n_seq - 1) * (2.0 * n_seq - 1) - 8.0 * (double)p)) / 2.0);
            while (i > 0 && i * (2 * (int64_t)n_seq - i - 1) / 2 > p) i--;
            while ((i + 1) * (2 * (int64_t)n_seq - (i + 1) - 1) / 2 <= p) i++;
            int64_t j = p - i * (2 * (int64_t)n_seq - i - 1) / 2 + i + 1;
            int li = (int)(off[i + 1] - off[i]), lj = (int)(off[j + 1] - off[j]);
            const uint8_t *a = codes + off[i], *b = codes + off[j];
            if (adjust_lens) {                                   /* tg_align.py:115-120 */
                int ml = li < lj ? li : lj;
                if (min_viable && ml < 1000) ml = 1000;          /* MIN_VIABLE_SEQ_LEN, tg_align.py:20 */
                if (li > ml) li = ml;
                if (lj > ml) lj = ml;
            }
            if (li <= 0 || lj <= 0) { bad |= 1; continue; }
            double aln = nw_score_d(a, li, b, lj, S, A, match, mismatch, g, gl, gr, row);   /* :125 */
            double iden;
            if (!S) iden = match * ((li + lj) / 2.);                                          /* :130 */
            else {
                double is1 = 0, is2 = 0;                                                      /* :132-134 */
                for (int k = 0; k < li; k++) is1 += S[(size_t)a[k] * A + a[k]];
                for (int k = 0; k < lj; k++) is2 += S[(size_t)b[k] * A + b[k]];
                iden = (is1 + is2) / 2.;
            }
            double cells = (double)li * (double)lj;
            double d;
            if (aln_out) aln_out[p] = (int32_t)aln;
            if (rnd_out) for (int k = 0; k < R; k++) rnd_out[p * R + k] = INT32_MIN;
            if (aln < -(0.900 * iden)) {                                                      /* :137-138 */
                d = 10.0;
            } else {
                const int64_t sv = rng_seed + p;                  /* random.seed(int) uses abs(seed) */
                tgo_mt st; mt_seed_u64(&st, (uint64_t)(sv < 0 ? -sv : sv));                   /* :110-111, :157 */
                double sum = 0.0;
                for (int k = 0; k < R; k++) {                                                 /* :141-142 */
                    mt_shuffle_bytes(&st, a, li, si, pool);
                    mt_shuffle_bytes(&st, b, lj, sj, pool);
                    double r = nw_score_d(si, li, sj, lj, S, A, match, mismatch, g, gl, gr, row);
                    if (rnd_out) rnd_out[p * R + k] = (int32_t)r;
                    sum += r;
                }
                cells *= (1 + R);
                double rmean = (R > 0) ? sum / R : NAN;                                       /* np.mean */
                if (rmean >= aln) d = 10.0;                                                   /* :145-146 */
                else { d = -log((aln - rmean) / (iden - rmean)); if (!(d <= 10.0)) d = 10.0; }/* :148 */
                /* note: min(x, 10.0) with x = nan returns nan in Python; unreachable for R >= 1 */
                if (d < 1e-4) d = 1e-4;                                                       /* :149 */
            }
            cells_total += cells;
            if (dist) { dist[i * n_seq + j] = (float)d; dist[j * n_seq + i] = (float)d; }    /* :183 */
        }
        free(row); free(si); free(sj); free(pool);
    }
    if (cells_out) *cells_out = cells_total;
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * k-mer scan.  re.finditer(kmer, seq) for a literal pattern = leftmost, non-overlapping:
 * after a match at p the next search starts at p + len(kmer)   (tg_kmer.py:69,96,178).
 * ---------------------------------------------------------------------------------------- */
static int finditer(const char *seq, int L, const char *kmer, int klen, int32_t *starts)
{
    int n = 0, p = 0;
    if (klen <= 0) return 0;
    while (p + klen <= L) {
        if (memcmp(seq + p, kmer, (size_t)klen) == 0) { if (starts) starts[n] = p; n++; p += klen; }
        else p++;
    }
    return n;
}

TGO_API int tgo_finditer(const char *seq, int L, const char *kmer, int klen, int32_t *starts)
{
    return finditer(seq, L, kmer, klen, starts);
}

/* union coverage of all kmers' matches (tg_kmer.py:71-73, 98-100) */
static void coverage(const char *seq, int L, const char *kmers, const int32_t *koff, int K, uint8_t *cov)
{
    memset(cov, 0, (size_t)L);
    for (int k = 0; k < K; k++) {
        const char *km = kmers + koff[k];
        int klen = koff[k + 1] - koff[k], p = 0;
        if (klen <= 0) continue;
        while (p + klen <= L) {
            if (memcmp(seq + p, km, (size_t)klen) == 0) { memset(cov + p, 1, (size_t)klen); p += klen; }
            else p++;
        }
    }
}

/* get_telomere_base_count (tg_kmer.py:93-102) */
TGO_API int tgo_base_count(const char *seq, int L, const char *kmers, const int32_t *koff, int K)
{
    uint8_t *cov = (uint8_t *)malloc((size_t)(L > 0 ? L : 1));
    coverage(seq, L, kmers, koff, K, cov);
    int c = 0;
    for (int i = 0; i < L; i++) c += cov[i];
    free(cov);
    return c;
}

/* get_telomere_kmer_density (tg_kmer.py:66-90) as integer window counts:
 * cnt[i] = cum[i+w] - cum[i] = sum cov[i+1 .. i+w], i in [0, L-w);  density = cnt / w (float64).
 * Returns the number of windows written (0 when L <= w). */
TGO_API int tgo_density_counts(const char *seq, int L, const char *kmers, const int32_t *koff, int K,
                               int window, int32_t *cnt)
{
    if (L - window <= 0) return 0;
    uint8_t *cov = (uint8_t *)malloc((size_t)L);
    int64_t *cum = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
    coverage(seq, L, kmers, koff, K, cov);
    int64_t s = 0;
    for (int i = 0; i < L; i++) { s += cov[i]; cum[i] = s; }
    for (int i = 0; i < L - window; i++) cnt[i] = (int32_t)(cum[i + window] - cum[i]);
    free(cov); free(cum);
    return L - window;
}

/* get_nonoverlapping_kmer_hits (tg_kmer.py:173-230), the four phases restated literally.
 * issub / issub_off: CSR of KMER_ISSUBSTRING (tg_kmer.py:44-46).
 * Output: spans as (k, start, end) triples grouped by k ascending, start ascending;
 * span_count[k] = number of spans for k.  Returns total spans, or -1 if cap is too small. */
TGO_API int tgo_nonoverlapping_hits(const char *seq, int L, const char *kmers, const int32_t *koff, int K,
                                    const int32_t *issub, const int32_t *issub_off,
                                    int32_t *out_k, int32_t *out_s, int32_t *out_e, int cap,
                                    int32_t *span_count)
{
    int32_t **st = (int32_t **)calloc((size_t)K, sizeof(int32_t *));   /* starts */
    int32_t **en = (int32_t **)calloc((size_t)K, sizeof(int32_t *));   /* ends   */
    int *cnt = (int *)calloc((size_t)K, sizeof(int));
    uint8_t **hit = (uint8_t **)calloc((size_t)K, sizeof(uint8_t *));  /* coord_hit_dict */
    int Lz = L > 0 ? L : 1;
    for (int k = 0; k < K; k++) {                                       /* :175-183 */
        int klen = koff[k + 1] - koff[k];
        int n = finditer(seq, L, kmers + koff[k], klen, NULL);
        st[k] = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
        en[k] = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
        hit[k] = (uint8_t *)calloc((size_t)Lz, 1);
        finditer(seq, L, kmers + koff[k], klen, st[k]);
        cnt[k] = n;
        for (int h = 0; h < n; h++) { en[k][h] = st[k][h] + klen; memset(hit[k] + st[k][h], 1, (size_t)klen); }
    }
    for (int k = 0; k < K; k++) {                                       /* :187-202 */
        int w = 0;
        for (int h = 0; h < cnt[k]; h++) {
            int are_we_hit = 0;
            for (int q = issub_off[k]; q < issub_off[k + 1] && !are_we_hit; q++) {
                const uint8_t *hh = hit[issub[q]];
                for (int j = st[k][h]; j < en[k][h]; j++) if (hh[j]) { are_we_hit = 1; break; }
            }
            if (!are_we_hit) { st[k][w] = st[k][h]; en[k][w] = en[k][h]; w++; }
        }
        cnt[k] = w;
    }
    for (int k = 0; k < K; k++) {                                       /* :206-212 */
        int w = 0;
        for (int h = 0; h < cnt[k]; h++) {
            if (w > 0 && en[k][w - 1] == st[k][h]) en[k][w - 1] = en[k][h];
            else { st[k][w] = st[k][h]; en[k][w] = en[k][h]; w++; }
        }
        cnt[k] = w;
    }
    for (int ki = 0; ki < K; ki++)                                      /* :216-228 */
        for (int kj = 0; kj < K; kj++) {
            if (kj == ki) continue;
            for (int a = 0; a < cnt[ki]; a++)
                for (int b = 0; b < cnt[kj]; b++) {
                    if (st[kj][b] > en[ki][a]) break;
                    if (st[ki][a] < st[kj][b] && en[ki][a] > st[kj][b]) en[ki][a] = st[kj][b];
                }
        }
    int total = 0, rc = 0;
    for (int k = 0; k < K; k++) {
        if (span_count) span_count[k] = cnt[k];
        for (int h = 0; h < cnt[k]; h++) {
            if (total < cap) { out_k[total] = k; out_s[total] = st[k][h]; out_e[total] = en[k][h]; }
            else rc = -1;
            total++;
        }
    }
    for (int k = 0; k < K; k++) { free(st[k]); free(en[k]); free(hit[k]); }
    free(st); free(en); free(cnt); free(hit);
    return rc < 0 ? -1 : total;
}

/* RC (tg_util.py:433-434) with RC_DICT {A,C,G,T,N} (tg_util.py:41); -1 on any other character */
TGO_API int tgo_rc(const char *seq, int L, char *out)
{
    for (int i = 0; i < L; i++) {
        char c = seq[L - 1 - i], r;
        switch (c) { case 'A': r = 'T'; break; case 'C': r = 'G'; break; case 'G': r = 'C'; break;
                     case 'T': r = 'A'; break; case 'N': r = 'N'; break; default: return -1; }
        out[i] = r;
    }
    return 0;
}

/* One read through the scan part of gtrc_parallel_job (tg_tel.py:262-266, :160-161):
 * base counts fwd/rev, orientation, window counts for both strands on the oriented read.
 * Used as the timed CPU baseline for the scan (bases/s). Returns 1 if the read was RC'd. */
TGO_API int tgo_scan_read(const char *seq, int L,
                          const char *kmers_f, const char *kmers_r, const int32_t *koff, int K,
                          const char *canon_f, const char *canon_r, const int32_t *coff, int KC,
                          int window, char *oriented, int32_t *cnt_f, int32_t *cnt_r, int32_t *bc)
{
    int f = tgo_base_count(seq, L, canon_f, coff, KC);
    int r = tgo_base_count(seq, L, canon_r, coff, KC);
    if (bc) { bc[0] = f; bc[1] = r; }
    int flipped = 0;
    if (f > r) { tgo_rc(seq, L, oriented); flipped = 1; } else memcpy(oriented, seq, (size_t)L);
    tgo_density_counts(oriented, L, kmers_f, koff, K, window, cnt_f);
    tgo_density_counts(oriented, L, kmers_r, koff, K, window, cnt_r);
    return flipped;
}
