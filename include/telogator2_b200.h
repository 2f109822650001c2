/*
 * telogator2_b200.h -- C ABI of the B200-native hot path of Telogator2.
 *
 * The reference (zstephens/telogator2) has no FFI layer: its hot path is reached through the
 * Python functions of source/tg_kmer.py and source/tg_align.py.  This header declares the
 * device entry points those functions bind in telogator2_b200/ (ctypes).
 * Each entry point cites the reference interface it replaces (file:line in the reference).
 *
 * Conventions
 *   - every pointer whose name starts with d_ is a DEVICE pointer owned by the caller
 *     (PyTorch tensors in the Python host); h_ pointers are host memory.
 *   - the library never allocates memory it hands back and never frees caller memory.
 *   - every function returns 0 on success or a negative TG_E* code; it never throws / exits.
 *     tg_last_error() returns a static string describing the last failure on this thread.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); work is only enqueued.
 *   - no CPU fallback exists: without a CUDA device every compute entry point returns TG_ECUDA.
 */
#ifndef TELOGATOR2_B200_H
#define TELOGATOR2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_OK        0
#define TG_EINVAL   -1   /* bad argument / unsupported configuration          */
#define TG_ECUDA    -2   /* CUDA runtime error (see tg_last_error)            */
#define TG_ERANGE   -3   /* job outside the envelope of the requested kernel  */
#define TG_EOVERFLOW -4  /* caller-provided output buffer too small           */

const char *tg_last_error(void);
int tg_version(void);                       /* ABI version, currently 1 */
int tg_device_sm_count(int *sm_count);      /* multiprocessor count of the current device */

/* =========================================================================================
 * Pairwise TVR distance (tg_align.py)
 * ======================================================================================= */

/* One packed alignment job = up to two independent global alignments ("halves") that the
 * u16x2 kernel advances in the two 16-bit lanes of every register.  Offsets index d_pool.
 * out[h] < 0 marks an unused half (its inputs must still be valid; duplicate half 0). */
typedef struct tg_nw_job {
    int64_t row_off[2];     /* first code of the row sequence (seq_i side)    */
    int64_t col_off[2];     /* first code of the column sequence (seq_j side) */
    int32_t n[2];           /* row sequence length  (>= 1)                    */
    int32_t m[2];           /* column sequence length (>= 1)                  */
    int32_t out[2];         /* index into d_scores, or < 0                    */
    int32_t pad_[2];
} tg_nw_job;                /* 64 bytes */

/* Scoring of Bio.Align.PairwiseAligner as configured by get_aligner_object()
 * (tg_align.py:71-106) and get_scoring_matrix() (tg_align.py:32-68), integer valued.
 * sub[a * alphabet + b] is the substitution score of codes a, b (match/mismatch mode is
 * expressed as a matrix with match on the diagonal).  gap_* are the linear gap scores
 * (open == extend at every call site that reaches score()): internal, left end, right end. */
typedef struct tg_nw_scoring {
    int32_t alphabet;       /* 1..32 */
    int32_t gap, gap_left, gap_right;
    int16_t sub[32 * 32];
} tg_nw_scoring;

/* Replaces aligner.score(seq_i, seq_j) (tg_align.py:125,142), i.e. Biopython's global
 * Needleman-Wunsch score with linear gaps and left/right end-gap classes.
 *
 * tg_nw_scores_wave: warp-wavefront u16x2 DPX kernel (two alignments per register lane).
 *   Envelope (else TG_ERANGE): alphabet <= 31, gap_left == gap, gap_right >= gap,
 *   0 <= sub - 2*gap <= 127, max(sub - 2*gap) * min(max n, max m) <= 65535 over the two halves.
 *   The two halves of a job may have different shapes.
 *   d_scratch: tg_nw_wave_scratch_bytes(max_n) bytes; d_counter: one zeroed int32.
 * tg_nw_scores_generic: int32 row-scan kernel, any gap scores, half 0 and half 1 processed
 *   as separate alignments; max(m) bounded by shared memory (tg_nw_generic_max_m()). */
size_t tg_nw_wave_scratch_bytes(int max_n);
int tg_nw_wave_check(const tg_nw_scoring *sc, const tg_nw_job *h_jobs, int64_t n_jobs);
int tg_nw_scores_wave(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_n,
                      const tg_nw_scoring *sc, int32_t *d_scores, void *d_scratch, int32_t *d_counter,
                      void *stream);

/* Same jobs, 32-bit cells, one alignment at a time (each half of a job is an item): for alignments whose
 * scores leave the 16-bit envelope of tg_nw_scores_wave (tg_nw_wave_check -> TG_ERANGE), e.g. the
 * untruncated consensus sequences of the collapse stage (tg_tvr.py:375-390).  Half the throughput. */
int tg_nw_scores_wave32(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_n,
                      const tg_nw_scoring *sc, int32_t *d_scores, void *d_scratch, int32_t *d_counter,
                      void *stream);
int tg_nw_generic_max_m(void);
int tg_nw_scores_generic(const uint8_t *d_pool, const tg_nw_job *d_jobs, int64_t n_jobs, int max_m,
                         const tg_nw_scoring *sc, int32_t *d_scores, void *stream);

/* One shuffle job = the null model of one sequence pair: random.seed(seed) followed by
 * R x (shuffle_seq(seq_i), shuffle_seq(seq_j))  (tg_align.py:110-111,141-142; tg_util.py:451-452),
 * bit-exact with CPython's MT19937 + random.sample.  The 2R shuffled strings are written
 * back to back at dst_off: [i_0 (n) | j_0 (m) | i_1 | j_1 | ...]. */
typedef struct tg_shuffle_job {
    uint64_t seed;          /* abs(rng_seed + pair_index) */
    int64_t src_i, src_j;   /* offsets of the (already length-adjusted) sequences in d_pool */
    int64_t dst_off;        /* offset in d_pool of the 2R outputs */
    int32_t n, m;
} tg_shuffle_job;           /* 40 bytes */

int tg_mt_shuffle(uint8_t *d_pool, const tg_shuffle_job *d_jobs, int64_t n_jobs, int R, int max_len,
                  void *stream);

/* The whole device side of get_dist_matrix_parallel (tg_align.py:154-189) for a range of pairs,
 * enqueued without any host round trip:
 *   phase A (aligner.score of every pair, :125) -> identity score + early-out flags (:127-138)
 *   -> MT19937 null model of the surviving pairs (:110-111,141-142) -> phase B (R scores per pair).
 * Pairs are addressed in "sorted pair space": q indexes itertools.combinations over the sequences
 * sorted by length (d_order[rank] = sequence index, longest first) so neighbouring pairs have
 * similar shapes; each pair is mapped back to its original (i < j), its length-adjusted shapes
 * (:115-120) and its combinations index p (seed = abs(seed0 + p), :157,170-172) on the device.
 * Outputs are indexed by q - q0: d_aln, d_flags (1 = early-out), d_iden (float64 identity score),
 * d_shape (n, m), d_rnd[(q - q0) * R + k] (pre-filled by the caller; untouched for early-out pairs).
 * The shuffled strings live in d_pool at shuf_base + (q - q0) * slot, slot >= 2 * R * max_len.
 * Requires the wave kernel's envelope for a max_len x max_len alignment (else TG_ERANGE: use the
 * explicit-job entry points). */
typedef struct tg_dist_batch_args {
    uint8_t *d_pool;
    const int64_t *d_seq_off;       /* [n_seq + 1] */
    const int32_t *d_order;         /* [n_seq] */
    const int32_t *d_diag_cs;       /* [codes + 1] prefix sums of sub[c][c] over d_pool codes, NULL in match/mismatch mode */
    int32_t *d_aln, *d_rnd, *d_shape;
    uint8_t *d_flags;
    double *d_iden;
    void *d_scratch;                /* tg_nw_wave_scratch_bytes(max_len) */
    int32_t *d_counter;             /* 32 bytes, zeroed once by the caller; d_counter[4] is set to 1 if a random stream overflowed (results invalid: rerun with d_mt_work = NULL) */
    void *d_mt_work;                /* tg_dist_batch_mt_work_bytes(q1 - q0, R, max_len) bytes, or NULL for the single-kernel null model */
    int64_t q0, q1, seed0, shuf_base, slot;
    double match;                   /* match score, used when d_diag_cs == NULL */
    int32_t n_seq, adjust_lens, min_viable, R, max_len;
    int32_t n_mat;                  /* 0: one matrix over all n_seq sequences.  > 0: the pair space is the concatenation of the pair
                                     * spaces of n_mat independent matrices (the per-cluster calls of tg_tvr.py:147-159, 375-390 batched):
                                     * matrix b owns pairs [d_mat_pair_off[b], d_mat_pair_off[b+1]) and sequences
                                     * [d_mat_seq_off[b], d_mat_seq_off[b+1]); d_order sorts each matrix's sequences by length
                                     * (global indices); every matrix uses seeds seed0 + its own combinations index. */
    const int64_t *d_mat_pair_off;  /* [n_mat + 1] */
    const int32_t *d_mat_seq_off;   /* [n_mat + 1] */
    const uint32_t *d_seq_mask;     /* [n_seq] bit c set = code c occurs in the sequence, or NULL.  With prof_rows > 0 the null-model
                                     * alignments whose row sequences use fewer than prof_rows letters run with per-job query profiles of
                                     * prof_rows rows (smaller shared-memory footprint, more resident warps); same results. */
    int32_t prof_rows;
    int32_t phases;                 /* bit mask, 0 = everything: 1 = real alignments + early-out flags + MT19937 seeding, 2 = random streams +
                                     * sampling (the shuffled sequences), 4 = null-model alignments.  Lets the caller run phase 2 of one batch
                                     * on a second stream under phase 1 of the next batch (separate d_counter / d_mt_work / shuffle areas). */
    int32_t smem_reserve;           /* shared memory per SM the alignment kernels of this call leave free (for a co-running phase 2) */
    int32_t pad_;
} tg_dist_batch_args;

size_t tg_dist_batch_mt_work_bytes(int64_t n_pairs, int R, int max_len);
int tg_dist_batch(const tg_dist_batch_args *args, const tg_nw_scoring *sc, void *stream);

/* The float epilogue of tvr_distance (tg_align.py:137-149) and the matrix fill of
 * get_dist_matrix_parallel (tg_align.py:180-186) for the pairs [q0, q1) of a tg_dist_batch call:
 *   early-out -> 10;  rand = mean(rnd[0..R));  rand >= aln -> 10;
 *   else min(-log((aln - rand) / (iden - rand)), 10);  then max(., 1e-4);  float64 -> float32;
 *   D[i][j] = D[j][i] = dist (original sequence indices; the caller zeroes D once).
 * All float64 operations are IEEE and in the reference's order; only log() can differ from the host's
 * libm in the last bit(s).  A pair whose float32 result (clamps included) would change if log() moved by
 * 2^-margin_log2 relative is NOT trusted: its local index q - q0 is appended to d_ambig (up to ambig_cap
 * entries; d_stats[3] counts all of them) and the host recomputes that distance from the integer results
 * with its own libm, so the matrix is bit-identical to the host epilogue.
 * d_stats (4 x uint64, accumulated, zeroed by the caller): [0] cells of the real alignments,
 * [1] cells of the null-model alignments, [2] early-out pairs, [3] ambiguous pairs. */
typedef struct tg_dist_epilogue_args {
    const int64_t *d_seq_off;
    const int32_t *d_order;
    const int32_t *d_aln, *d_rnd;   /* [q1 - q0], [(q1 - q0) * R] as written by tg_dist_batch */
    const uint8_t *d_flags;
    const double *d_iden;
    float *d_matrix;                /* [n_seq * n_seq] row-major */
    int64_t *d_ambig;               /* [ambig_cap] */
    unsigned long long *d_stats;    /* [4] */
    int64_t q0, q1, ambig_base;     /* ambig entries are ambig_base + (q - q0) */
    int32_t n_seq, adjust_lens, min_viable, R, ambig_cap, margin_log2;
    int32_t n_mat, pad_;            /* as in tg_dist_batch_args; matrix b (n_b x n_b, row-major) starts at d_matrix + d_mat_out_off[b] */
    const int64_t *d_mat_pair_off;
    const int32_t *d_mat_seq_off;
    const int64_t *d_mat_out_off;   /* [n_mat] */
} tg_dist_epilogue_args;

int tg_dist_epilogue(const tg_dist_epilogue_args *args, void *stream);

/* Measurement aid (bench.py's roofline figure): while enabled, CUDA events on the launching stream bracket every
 * nw_wave_kernel launch of this process; tg_wave_timing_read waits for them, returns the summed kernel
 * milliseconds and the number of launches since the last read, and forgets them.  Not thread safe. */
int tg_wave_timing(int enable);
int tg_wave_timing_read(double *total_ms, int *n_launches);

/* Integer-ALU issue-rate microbenchmark used as the DP roofline denominator:
 * runs `iters` dependent-chain-free VIMNMX3.U16x2/IADD pairs per thread on the whole device and
 * returns lane-operations per second for the DPX op and for a plain integer add. */
int tg_int_alu_peak(int iters, double *dpx_lane_ops_per_s, double *iadd_lane_ops_per_s);

/* =========================================================================================
 * k-mer scan (tg_kmer.py)
 * ======================================================================================= */

/* Device table for one k-mer list with its KMER_ISSUBSTRING relation (the hits kernel): a lookup table over the 4^7
 * windows of 7 bases whose 16-bit entries name the SET of plain k-mers starting there (an index into a table of 64-bit
 * sets) and the special k-mer -- longer than 7 bases or self-overlapping -- that may start there; tables for windows cut
 * short by an N; per-k-mer patterns, lengths and superstring sets.  tg_kmer_table_bytes() bytes, built on the host,
 * uploaded as is.  Limits: K <= 64 k-mers, length <= 32, alphabet ACGT, <= 32 self-overlapping k-mers, <= 512 distinct
 * sets.  Replaces the per-k-mer re.finditer passes of tg_kmer.py:178. */
size_t tg_kmer_table_bytes(void);
int tg_kmer_table_build(const char *kmers_concat, const int32_t *kmer_off, int K,
                        const int32_t *issub_edges, const int32_t *issub_off, /* KMER_ISSUBSTRING in CSR form */
                        void *h_table);
/* Coverage table for TWO k-mer lists (KMER_LIST + KMER_LIST_REV, tg_tel.py:160-161; CANONICAL_STRINGS +
 * CANONICAL_STRINGS_REV, tg_tel.py:262-263): a lookup table over the 4^7 windows of 7 bases whose 32-bit entries answer, for
 * both lists, "which positions does the longest plain k-mer starting here cover" and "might a special k-mer (longer than 7
 * bases or self-overlapping) start here"; plus the packed patterns for the exact slow paths.  tg_kmer_table_pair_bytes()
 * bytes, built on the host, uploaded as is.  Replaces the per-k-mer re.finditer passes of tg_kmer.py:69,96. */
size_t tg_kmer_table_pair_bytes(void);
int tg_kmer_table_build_pair(const char *kmers_a, const int32_t *off_a, int KA,
                             const char *kmers_b, const int32_t *off_b, int KB, void *h_table);

/* Read layout: read r occupies bases [base_off[r], base_off[r] + len[r]) of a padded base
 * space; base_off[r] is a multiple of 128 and base_off[r + 1] - base_off[r] > len[r] (at least one
 * padding base, which carries the N bit, separates two reads).  ASCII input uses one byte per padded base
 * (pad bytes arbitrary).  Packed form: d_pack2[w] holds bases 16w..16w+15, 2 bits each
 * (A=0 C=1 T=2 G=3, little end first); d_nmask[w] holds 32 bases, bit set = not ACGT.
 * d_base_off has n_reads + 1 entries (the last one = total padded bases).
 * Replaces nothing in the reference (it scans Python str) -- this is the HBM layout. */
int tg_scan_pack(const uint8_t *d_ascii, const int64_t *d_base_off, const int32_t *d_len, int n_reads,
                 uint32_t *d_pack2, uint32_t *d_nmask,
                 uint8_t *d_not_acgtn, /* [n_reads] zeroed by the caller, or NULL: set to 1 for reads holding a letter RC() rejects */
                 void *stream);

/* get_telomere_base_count(read, kmer_list) (tg_kmer.py:93-102) for two k-mer lists at once
 * (CANONICAL_STRINGS and CANONICAL_STRINGS_REV, tg_tel.py:262-263): d_counts[2r + s] (zeroed here).
 * total_bases = base_off[n_reads]; d_blk_read[b] = read index of the 128-base block b of the padded base space.
 * d_pack2 / d_nmask need 4 readable words of slack behind total_bases. */
int tg_scan_basecount(const uint32_t *d_pack2, const uint32_t *d_nmask, int64_t total_bases,
                      const int32_t *d_blk_read, int n_reads, const void *d_table_pair, int32_t *d_counts, void *stream);

/* RC(read) (tg_util.py:433-434) applied in place of tg_tel.py:265-266: reads with
 * d_flip[r] != 0 are reverse-complemented into d_pack2_out / d_nmask_out, others copied. */
int tg_scan_orient(const uint32_t *d_pack2, const uint32_t *d_nmask, const int64_t *d_base_off,
                   const int32_t *d_len, const uint8_t *d_flip, int n_reads,
                   uint32_t *d_pack2_out, uint32_t *d_nmask_out, void *stream);

/* get_telomere_kmer_density(read, kmer_list, tel_window) (tg_kmer.py:66-90) for two k-mer
 * lists at once (KMER_LIST and KMER_LIST_REV, tg_tel.py:160-161).  Emits the integer window
 * counts cnt[i] = #covered positions in (i, i + window], i in [0, len - window), as uint8 at
 * d_cnt_x[base_off[r] + i] (arrays of total_bases bytes; entries outside [0, len - window) of a read are
 * unspecified); density = cnt / window in float64 on the host.  window <= 255. */
int tg_scan_density(const uint32_t *d_pack2, const uint32_t *d_nmask, int64_t total_bases,
                    const void *d_table_pair, int window, uint8_t *d_cnt_a, uint8_t *d_cnt_b, void *stream);

/* get_nonoverlapping_kmer_hits(seq[lo:hi], KMER_LIST, KMER_ISSUBSTRING) (tg_kmer.py:173-230)
 * for a batch of slices: slice q = read d_slice_read[q], positions [d_slice_lo[q], d_slice_hi[q]).
 * One warp per slice, no scratch in HBM.  Spans are appended to d_span_* (slice, k, start, end relative to lo) in no
 * particular order; *d_span_count receives the number of spans (if it exceeds span_cap the extra spans were dropped and
 * the caller must retry with a larger buffer). */
int tg_scan_hits(const uint32_t *d_pack2, const uint32_t *d_nmask, const int64_t *d_base_off,
                 const int32_t *d_slice_read, const int32_t *d_slice_lo, const int32_t *d_slice_hi, int n_slices,
                 const void *d_table,
                 int32_t *d_span_slice, int32_t *d_span_k, int32_t *d_span_start, int32_t *d_span_end,
                 int64_t span_cap, unsigned long long *d_span_count, void *stream);

/* =========================================================================================
 * Colour-vector preparation (tg_tvr.py) -- the per-read loops cluster_tvrs / quick_get_tvrtel_lens run
 * before every distance matrix.  Strings are bytes in device buffers, described by (offset, length).
 * ======================================================================================= */

/* tg_tvr.py:96-105 / :566-575.  Read r's colour vector occupies d_cv[d_cv_off[r] .. d_cv_off[r+1]): `unknown`
 * everywhere, then the spans of k-mer 0, 1, .. K-1 painted with d_letters[k] in that order (a later k-mer
 * overwrites).  Spans are sorted by (read, k): those of (r, k) are d_rk_off[r*K + k] .. d_rk_off[r*K + k + 1]. */
int tg_cv_paint(const int32_t *d_span_start, const int32_t *d_span_end, const int64_t *d_rk_off,
                const uint8_t *d_letters, int K, const int64_t *d_cv_off, int n_reads, int unknown,
                uint8_t *d_cv, void *stream);

/* The position search of find_density_boundary (tg_tvr.py:449-479) for n strings; string s is
 * d_buf[d_off[s] .. + d_len[s]), read backwards when d_reverse[s] != 0 (d_reverse may be NULL).
 * dens[j] = #(letters of the set at positions j+1 .. j+win, positions < start never count) / win for
 * start <= j < len - win.  d_first[s] = first j with dens <= thresh (above == 0) or >= thresh (above != 0);
 * d_extreme[s] = the reference's pos_with_lowest_dens (first position of the strict running minimum starting
 * at 1.0, or maximum starting at 0.0).  -1 = None. */
int tg_cv_density_boundary(const uint8_t *d_buf, const int64_t *d_off, const int32_t *d_len, const uint8_t *d_reverse,
                           int n, const uint8_t *h_letters, int n_letters, int win, double thresh, int above, int start,
                           int32_t *d_first, int32_t *d_extreme, void *stream);

/* d_out[s] = min(length of the leading run of `letter`, d_cap[s]) (tg_tvr.py:124-126 with cap = len - 1). */
int tg_cv_lead_run(const uint8_t *d_buf, const int64_t *d_off, const int32_t *d_len, const uint8_t *d_reverse,
                   const int32_t *d_cap, int n, int letter, int32_t *d_out, void *stream);

/* d_dst[d_dst_off[g] + x] = d_src[d_src_off[g] + x], x < d_len[g]: the string concatenations between the steps. */
int tg_cv_gather(const uint8_t *d_src, const int64_t *d_src_off, const int64_t *d_dst_off, const int32_t *d_len,
                 int64_t n_seg, uint8_t *d_dst, void *stream);

/* denoise_colorvec (tg_tvr.py:524-556) for n non-empty strings; output s has d_len[s] + 1 letters at d_out_off[s].
 * d_keep[s] = d_len[s] - (trailing run of replace_char), d_tile_off = exclusive prefix sum of
 * ceil((d_len[s] + 1) / tg_cv_denoise_tile()).  Limits: min_size <= 64, max_gap_fill <= 63 (TG_ERANGE). */
int tg_cv_denoise_tile(void);
int tg_cv_denoise(const uint8_t *d_buf, const int64_t *d_off, const int32_t *d_len, const int32_t *d_keep,
                  const int64_t *d_out_off, const int64_t *d_tile_off, int n, int64_t n_tiles, int replace_char,
                  int min_size, int max_gap_fill, const uint8_t *h_delete, int n_delete,
                  const uint8_t *h_merge, int n_merge, uint8_t *d_out, void *stream);

/* =========================================================================================
 * Stage [0a] telomere-read prefilter (telogator2.py:451-459)
 * ======================================================================================= */

/* d_count_a[r] = read.count(kmer_a), d_count_b[r] = read.count(kmer_b) (Python str.count: non-overlapping
 * occurrences, left to right) for every read; kmer_a / kmer_b are two ACGT patterns of the same length klen <= 32
 * (KMER_INITIAL and KMER_INITIAL_RC, telogator2.py:393-394).  Reads are ASCII in the padded layout of tg_scan_pack
 * (d_base_off multiples of 16); the buffer must extend tg_prefilter_tail_bytes() beyond the last read.
 * d_counter: 8 bytes of scratch. */
int tg_prefilter_tail_bytes(void);
int tg_prefilter_count(const uint8_t *d_ascii, const int64_t *d_base_off, const int32_t *d_len, int n_reads,
                       const char *kmer_a, const char *kmer_b, int klen, int32_t *d_count_a, int32_t *d_count_b,
                       void *d_counter, void *stream);

/* =========================================================================================
 * Telomere-boundary calling: tg_kmer.get_telomere_regions (tg_kmer.py:105-142), wavelet_smooth / mad
 * (tg_kmer.py:145-170: pywt.wavedec / threshold / waverec, wavelet 'db4', mode 'per') and tg_tel.get_terminating_tl
 * (tg_tel.py:156-235), for a batch of reads.
 *
 * Per read r: n = d_sig_n[r] power samples (len(read) - window, or 0), float64 slabs at element offsets d_x_off[r]
 * (even(n) samples), d_a_off[r] (approximations a_1..a_L, each padded to an even length) and d_d_off[r] (details d_1..d_L,
 * contiguous; d_d_off has n_reads + 1 entries), m_0 = n, m_l = ceil(m_(l-1) / 2), L = floor(log2(n / 7)).  max_n = the
 * largest n of the batch, total_n = their sum (grid sizing only: a few blocks per read walk its tiles).
 * ======================================================================================= */

/* X[r][i] = cnt_p[base_off[r] + i] / window - cnt_q[base_off[r] + i] / window, i < n: the p-vs-q power signal from the two
 * window-count planes of tg_scan_density (tg_kmer.py:90, 114-115). */
int tg_bound_power(const uint8_t *d_cnt_p, const uint8_t *d_cnt_q, const int64_t *d_base_off, const int32_t *d_sig_n,
                   const int64_t *d_x_off, int n_reads, int max_n, int64_t total_n, int window, double *d_X, void *stream);

/* wavelet_smooth(X[r]) in place for every read with n >= window (tg_kmer.py:116-117, 152-170): d_smoothed[r] = 1 and
 * d_out_len[r] = even(n) when the signal was replaced by its reconstruction, 0 and n when one of the two early-outs
 * (fewer than smoothing_level coefficient arrays; sigma < fail_sigma) left it as it was.  d_log_factor[r] = sqrt(2 ln n)
 * (host); d_sigma / d_uthresh receive mad(coeff[-smoothing_level]) and the threshold.  d_scratch / d_sc_off: sort space for
 * reads whose coeff[-smoothing_level] has more than 4096 entries (offset per read, -1 = not needed). */
int tg_bound_smooth(double *d_X, double *d_A, double *d_D, const int64_t *d_x_off, const int64_t *d_a_off,
                    const int64_t *d_d_off, const int32_t *d_sig_n, int n_reads, int max_n, int64_t total_n, int window,
                    int smoothing_level, double fail_sigma, const double *d_log_factor, double *d_scratch,
                    const int64_t *d_sc_off, double *d_uthresh, double *d_sigma, int32_t *d_out_len,
                    uint8_t *d_smoothed, void *stream);

/* Region walk over X[r][: out_len - window] (tg_kmer.py:118-140): maximal runs of one class (1 'p' >= pthresh, 2 'q' <=
 * -pthresh, 0 None).  d_reg_off == NULL: count only (d_reg_count, d_last_end, d_err); else also d_reg_start / d_reg_cls at
 * d_reg_off[r].  The end of a region is the start of the next one; the last region ends at d_last_end[r].  d_err[r] = 1
 * where the reference raises IndexError (empty signal); reads with n < window give the single region [0, 0, None]. */
int tg_bound_regions(const double *d_X, const int64_t *d_x_off, const int32_t *d_sig_n, const int32_t *d_out_len,
                     int n_reads, int window, double pthresh, const int64_t *d_reg_off, int32_t *d_reg_count,
                     int32_t *d_reg_start, uint8_t *d_reg_cls, int32_t *d_last_end, uint8_t *d_err, void *stream);

/* get_terminating_tl's scoring on the region lists (tg_tel.py:164-233): d_out[r] = (tel_len, edge_nontel, a, b) with
 * (a, b) the largest interstitial telomere region or (-1, -1). */
int tg_bound_terminating(const int64_t *d_reg_off, const int32_t *d_reg_count, const int32_t *d_reg_start,
                         const uint8_t *d_reg_cls, const int32_t *d_last_end, int n_reads, int pq_is_p, int window,
                         double min_tel_score, int32_t *d_out, void *stream);

/* =========================================================================================
 * Hierarchical clustering of the distance matrix ("scipy linkage itself"): scipy.cluster.hierarchy.linkage
 * as called at tg_tvr.py:166-167 and :397-398, methods ward / single / complete / average / weighted.
 * The merge list is bit-identical to scipy's (same nearest-neighbour-chain / MST algorithms, same float64 evaluation order).
 * ======================================================================================= */

/* d_square[i][j] = d_condensed[condensed_index(n, i, j)] (what scipy.spatial.distance.squareform(y) builds), diagonal 0. */
int tg_linkage_expand(const double *d_condensed, int n, double *d_square, void *stream);
/* d_square = float64(d_matrix / div) with the division in float32 (tg_tvr.py:154: dist_matrix /= MAX_SEQ_DIST on the float32
 * matrix), or float64(d_matrix) when div == 0: clustering straight from the device matrix of tg_dist_epilogue. */
int tg_linkage_widen(const float *d_matrix, int n, float div, double *d_square, void *stream);
/* Runs the algorithm on d_square (n x n float64, destroyed).  d_size: n int32, all ones (ward .. weighted); d_chain: n int32
 * scratch; d_near: n float64 scratch; d_Z receives n - 1 rows (x, y, distance, size) in MERGE order with original cluster
 * indices: the caller sorts them by distance (stable) and relabels them (scipy's `label`), see telogator2_b200/tg_linkage.py. */
int tg_linkage_run(double *d_square, int n, const char *method, int32_t *d_size, int32_t *d_chain, double *d_near,
                   double *d_Z, void *stream);
/* The same run with the row walks spread over `ctas` CTAs (<= SM count, launched cooperatively; grid barrier between the
 * dependent steps): same merge list, faster once a row is long (n in the tens of thousands).  d_work:
 * tg_linkage_multi_work_bytes(n, ctas) bytes of scratch. */
size_t tg_linkage_multi_work_bytes(int n, int ctas);
int tg_linkage_run_multi(double *d_square, int n, const char *method, int ctas, int32_t *d_size, double *d_near,
                         double *d_Z, void *d_work, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TELOGATOR2_B200_H */
