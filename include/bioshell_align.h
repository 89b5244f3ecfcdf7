/*
 * bioshell_align.h -- C ABI of libbioshell_align.so, the B200-native (sm_100a)
 * drop-in for the ONE data-parallel hot path of dgront/BioShell v4: all-vs-all /
 * one-vs-many pairwise global alignment (Needleman-Wunsch, Gotoh affine gaps,
 * BLOSUM/PAM matrices) with the reference aligner's exact tie-breaking.
 *
 * The reference has no FFI at all (100 % Rust); these entry points
 * are what a `bioshell-seq` `extern "C"` block would bind.  Each one names the
 * reference interface it replaces (paths relative to the reference checkout):
 *
 *   bsa_parse_ncbi_matrix  SubstitutionMatrix::ncbi_matrix_from_buffer
 *                          bioshell-seq/src/scoring/substitution_matrix.rs:96-135
 *   bsa_set_scoring        SequenceSimilarityScore::new + the gap arguments of
 *                          align_all_pairs (scoring/similarity_score.rs:67-71,
 *                          alignment/alignment_protocols.rs:83-84)
 *   bsa_load_sequences     the `&Vec<Sequence>` arguments of align_all_pairs and the
 *                          per-pair byte->index encode, similarity_score.rs:125-134
 *   bsa_align_all_pairs    align_all_pairs' t-major double loop around
 *                          GlobalAligner::align + backtrace + the identity count of
 *                          the reporter: alignment_protocols.rs:94-110,
 *                          global.rs:57-201, msa.rs:261-269
 *   bsa_all_vs_all         the `if_triangle_only = true` call of
 *                          bin/cluster_sequences.rs:175-176
 *   bsa_one_vs_many        the `if_triangle_only = false` call of
 *                          bioshell-seq/examples/needleman_wunsh.rs:108-109
 *   bsa_align_pairs_paths  GlobalAligner::align + backtrace -> AlignmentPath
 *                          (global.rs:57-201, alignment_path.rs:34-48,107-115)
 *   bsa_local_align_pairs  LocalAlignment::align + backtrace + recent_end_point (Smith-Waterman,
 *                          SURVEY.md 8f rank 4)  bioshell-seq/src/alignment/local.rs:83-284
 *   bsa_hclust             hierarchical_clustering + HierarchicalClusteringMatrix (the consumer
 *                          of the identity matrix)
 *                          bioshell-clustering/src/hierarchical/hierarchical.rs:22-80,
 *                          clustering_matrix.rs:11-74, strategies/mod.rs:25-92
 *
 * Conventions: plain pointers and sizes only.  The caller allocates every input
 * and output buffer and keeps it alive for the duration of the call; the library
 * owns device memory and streams inside bsa_ctx and never retains caller pointers
 * after return.  Calls are synchronous (results complete on return).  A bsa_ctx is
 * used by one host thread at a time.  bsa_create binds it to ONE CUDA device (one process per
 * GPU, sharded by template range with bsa_plan_shards); bsa_create_multi gives a single process
 * all GPUs behind the same calls.  There is NO CPU
 * fallback: without a usable CUDA device every compute entry point returns
 * BSA_ERR_CUDA.  No C++ exception crosses this boundary.
 */
#ifndef BIOSHELL_ALIGN_H
#define BIOSHELL_ALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSA_VERSION 1

/* return codes */
#define BSA_OK 0
#define BSA_ERR_BAD_ARG (-1)
#define BSA_ERR_UNSUPPORTED_GAPS (-2) /* need gap_open <= gap_extend <= 0 and gap_open < 0 */
#define BSA_ERR_RANGE (-3)            /* score range does not fit the kernel's integer fields */
#define BSA_ERR_CUDA (-4)
#define BSA_ERR_OOM (-5)
#define BSA_ERR_FORMAT (-6)           /* NCBI matrix text: IncorrectNCBIFormat / CantParseNCBIEntry */
#define BSA_ERR_ALPHABET (-7)         /* > 127 distinct residue byte values, or byte 255 (reference panics) */
#define BSA_ERR_EMPTY (-8)            /* empty sequence set (reference: max().unwrap() panics) */

/* flags for the alignment entry points */
#define BSA_WANT_SCORE 1u
#define BSA_WANT_IDENTICAL 2u
#define BSA_OUT_DEVICE 4u /* output pointers are device pointers on the context's GPU */
#define BSA_IN_DEVICE 8u  /* input matrix pointer is a device pointer (bsa_hclust) */

typedef struct bsa_ctx bsa_ctx;

/* Number of CUDA devices visible to this process (0 when there is none). */
int bsa_device_count(void);

/* Create a context on CUDA device `device_id`.  NULL on failure (no device). */
bsa_ctx *bsa_create(int device_id);
/*
 * ONE context over n_dev GPUs for a single-process caller (the reference's caller is one process:
 * bin/cluster_sequences.rs:173-177 calls align_all_pairs once).  n_dev == 0 takes
 * every visible device.  Every entry point below behaves as on a single-device context and
 * returns the same bytes: sequence sets and scoring are replicated to all devices; the library
 * runs one host worker thread per GPU, cuts bsa_align_all_pairs / bsa_all_vs_all /
 * bsa_one_vs_many into one cell-balanced template-range tile per GPU, and each GPU delivers its
 * tile straight into the caller's output buffers at the tile's own t-major offset (stored by the
 * kernels themselves when the buffers are page-locked, see bsa_host_alloc_pinned; staged and
 * copied otherwise).  No collective, no NCCL.
 * (BSA_MULTI_WORKERS_PER_GPU=2: guided tiles pulled from a shared counter, two children per GPU
 * taking turns on it, so that one tile's copy and the next tile's planning overlap the kernels.)
 * Pair-list calls are split into contiguous chunks of equal cells.  BSA_OUT_DEVICE is refused.
 * bsa_hclust and bsa_measure_int_peak run on the first device.
 */
bsa_ctx *bsa_create_multi(const int *device_ids, int n_dev);
/* Number of GPUs behind this context (1 for bsa_create). */
int bsa_context_devices(const bsa_ctx *ctx);
void bsa_destroy(bsa_ctx *ctx);

/* Message of the last error on this context ("" if none). ctx may be NULL for creation errors. */
const char *bsa_last_error(const bsa_ctx *ctx);

/*
 * Parse an NCBI-format substitution matrix exactly as the reference does
 * (substitution_matrix.rs:96-135): 21x21 row-major table in the file's row order
 * (A R N D C Q E G H I L K M F P S T W Y V, then X), mirrored writes, X column =
 * token n-2, X/X forced to -1; aa_index maps a residue byte to its row (every
 * byte that is not one of the 20 row letters or 'X' maps to 0).  Host only.
 */
int bsa_parse_ncbi_matrix(const char *text, size_t len, int32_t score[441], uint8_t aa_index[256]);

/* Scoring used by subsequent alignment calls.  Requires gap_open <= gap_extend <= 0, gap_open < 0. */
int bsa_set_scoring(bsa_ctx *ctx, const int32_t score[441], const uint8_t aa_index[256],
                    int32_t gap_open, int32_t gap_extend);

/*
 * Load (replace) sequence set `set_id` (0..7): n sequences, raw residue bytes packed
 * back to back, offsets[n+1].  Bytes are uploaded and encoded on the device into the
 * packed sequence store.  Residues are arbitrary bytes except 255.
 */
int bsa_load_sequences(bsa_ctx *ctx, int set_id, const uint8_t *residues_raw,
                       const uint64_t *offsets, uint32_t n);

/*
 * The general batched form of align_all_pairs (alignment_protocols.rs:94-110):
 * for t in [t_begin, t_end) and q in [0, q_counts[t]) align query q (rows) against
 * template t (columns).  q_counts == NULL means every query for every template
 * (if_triangle_only = false); q_counts[t] = t is the strict upper triangle of a set
 * against itself; the host mirror derives other values from the reference's
 * `template == query -> break` rule.  Result k of pair (q,t) is stored at
 *   k = sum_{t' in [t_begin,t)} q_counts[t'] + q          (t-major report order)
 * scores[k]      = GlobalAligner::align's return value (global.rs:143-144)
 * n_identical[k] = count_identical of the two aligned strings (msa.rs:261-269)
 * Either output may be NULL.  Returns the number of results through *n_results.
 */
int bsa_align_all_pairs(bsa_ctx *ctx, int q_set, int t_set, const uint32_t *q_counts,
                        uint32_t t_begin, uint32_t t_end, uint32_t flags, int32_t *scores,
                        uint32_t *n_identical, uint64_t *n_results);

/* Strict upper triangle of one set: k = t(t-1)/2 + q, q < t.  (cluster_sequences.rs:175-176) */
int bsa_all_vs_all(bsa_ctx *ctx, int set_id, uint32_t flags, int32_t *scores,
                   uint32_t *n_identical);

/* Every query against every database sequence: k = t*|Q| + q. (needleman_wunsh.rs:108-109) */
int bsa_one_vs_many(bsa_ctx *ctx, int q_set, int db_set, uint32_t flags, int32_t *scores,
                    uint32_t *n_identical);

/*
 * A new set from sequences of a loaded one, on the device: sequence idx[i] of src_set becomes
 * sequence i of dst_set (indices may repeat; residues are not uploaded again).  This is how a
 * pair list with a common template -- BucketClustering::sequence_identity of one candidate
 * against the representatives whose k-mer bounds are inconclusive,
 * bioshell-seq/src/sequence/bucket_clustering/bucket_clustering.rs:272-309 -- runs on the
 * forward score + identity kernels: gather the representatives, gather the candidate, then
 * bsa_align_all_pairs(gathered reps, {candidate}) (no direction store, no traceback).
 */
int bsa_gather_sequences(bsa_ctx *ctx, int src_set, int dst_set, const uint32_t *idx, uint32_t n);

/*
 * Cell-balanced contiguous template ranges for n_shards GPUs/ranks:
 * bounds[0] = 0 <= ... <= bounds[n_shards] = |T|; shard r runs
 * bsa_align_all_pairs(..., bounds[r], bounds[r+1], ...).  No collective is needed.
 */
int bsa_plan_shards(bsa_ctx *ctx, int q_set, int t_set, const uint32_t *q_counts,
                    uint32_t n_shards, uint32_t *bounds);

/*
 * Full alignments for an explicit pair list: score, path glyphs ('*' Match,
 * '-' Horizontal = gap in query, '|' Vertical = gap in template;
 * alignment_path.rs:40-47) and n_identical.  path_buf must hold
 * sum(len_q + len_t) bytes; pair p's path is path_buf[path_off[p] .. path_off[p+1])
 * after the call (path_off has n_pairs+1 entries, written by the library).
 * scores / n_identical / path_buf may be NULL.
 */
int bsa_align_pairs_paths(bsa_ctx *ctx, int q_set, int t_set, const uint32_t *q_idx,
                          const uint32_t *t_idx, uint64_t n_pairs, int32_t *scores,
                          uint32_t *n_identical, uint8_t *path_buf, uint64_t *path_off);

/*
 * Local (Smith-Waterman-Gotoh) alignments for an explicit pair list, with the reference's STOP
 * semantics and best-cell rule (local.rs:83-273).  Per pair: scores = recent_score();
 * (end_q, end_t) = recent_end_point() (1-based cell of the best score, (0,0) if the score is 0);
 * (start_q, start_t) and the path glyphs = backtrace().  Any output may be NULL; path_buf /
 * path_off as in bsa_align_pairs_paths.
 */
int bsa_local_align_pairs(bsa_ctx *ctx, int q_set, int t_set, const uint32_t *q_idx,
                          const uint32_t *t_idx, uint64_t n_pairs, int32_t *scores, uint32_t *end_q,
                          uint32_t *end_t, uint32_t *start_q, uint32_t *start_t, uint8_t *path_buf,
                          uint64_t *path_off);

/*
 * Hierarchical agglomerative clustering with the reference's exact merge order.
 * dist: n x n row-major f32 of which ONLY dist[i*n+j] with i > j is read (as
 * HierarchicalClusteringMatrix::new does, clustering_matrix.rs:14-19) and mirrored.
 * linkage: 0 single, 1 complete, 2 average, 3 median, 4 centroid, 5 Ward
 * (strategies/mod.rs:25-92).  Output, one entry per merge step s = 0..n-2:
 * mat_i[s] < mat_j[s] = the matrix indices closest_elements returned, merge_dist[s] = their
 * distance; the tree is rebuilt from them on the host (hierarchical.rs:44-75).
 * flags: BSA_IN_DEVICE if dist is a device pointer.
 */
int bsa_hclust(bsa_ctx *ctx, uint32_t n, const float *dist, int linkage, uint32_t flags,
               uint32_t *mat_i, uint32_t *mat_j, float *merge_dist);

/*
 * Page-locked host memory.  Result buffers of bsa_align_all_pairs / bsa_all_vs_all / bsa_one_vs_many
 * that are page-locked (from here, or any cudaHostAlloc'd / pinned memory) are written by the kernels
 * DIRECTLY, 8 bytes per pair over the host link while the alignment runs: no staging copy of the
 * results in device memory and no device-to-host copy at the end of the call.  Pageable buffers are
 * staged in device memory and copied when the kernels are through.  Same bytes either way.
 */
void *bsa_host_alloc_pinned(size_t bytes);
void bsa_host_free_pinned(void *p);

/* Statistics of the most recent alignment call on this context. */
typedef struct bsa_stats {
    uint64_t pairs;          /* pairs aligned */
    uint64_t cells;          /* sum of len_q*len_t (alignment_protocols.rs:104) */
    uint64_t padded_cells;   /* cells the kernels actually swept (column padding, pipeline fill) */
    double kernel_ms;        /* CUDA-event time first launch -> last kernel end */
    double total_ms;         /* host wall time of the call */
    uint32_t launches;       /* kernels launched */
    uint32_t items;          /* work items */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t fallback_pairs; /* pairs routed to the direction-store path */
    uint32_t reserved;
} bsa_stats;
int bsa_get_stats(const bsa_ctx *ctx, bsa_stats *out);

/*
 * Integer-pipe microbenchmark on the context's GPU: independent chains of the
 * kernel's own instruction mix (VIADDMNMX / VIMNMX3 / LOP3 / IADD3) on every SM.
 * Returns lane-operations per second (32 x warp instructions); this is the
 * measured denominator of the integer/DPX roofline.
 * which: 0 = the 8-op classic cell mix, 1 = VIADDMNMX only, 2 = VIMNMX3 only, 3 = LOP3 only,
 *        4 = IADD3 only, 5 = IMAD only, 6 = VIADDMNMX.S16x2 only,
 *        7 = the 7-op TAG cell mix of round 1 (VIMNMX3 + LOP3 + 2 VIADDMNMX + 3 IMAD),
 *        8 = the 6-op frame cell (2 IMAD), 9 = the 11-op K3 direction-frame cell,
 *        10 = the 5-op frame cell with column-tagged E openings (VIMNMX3 + LOP3 + 2 VIADDMNMX + 1 IMAD;
 *             what bench.py holds cfg1..cfg3 against)
 */
int bsa_measure_int_peak(bsa_ctx *ctx, int which, double *lane_ops_per_s, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* BIOSHELL_ALIGN_H */
