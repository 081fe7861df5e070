/*
 * doppelspeller_b200.h - C ABI of the B200-native DoppelSpeller hot path.
 *
 * The reference (mhaseebtariq/doppel-speller) is pure Python + numba: it has no FFI of its own.  The
 * drop-in boundary is therefore the set of Python call signatures listed in SURVEY.md 8(b); this
 * header is the C ABI a binding for those signatures talks to (ctypes stub: INTEGRATION.md and
 * doppelspeller_b200/_native.py).  Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - every function returns DS_OK (0) or a negative ds_status; ds_last_error() gives the
 *     thread-local message.  No exceptions cross the boundary.  There is no CPU fallback: without a
 *     usable sm_100 device every compute entry point fails with DS_ERR_CUDA.
 *   - data pointers may be DEVICE pointers (used in place, call is asynchronous on `stream`) or HOST
 *     pointers (staged through the library's own device workspace; when an OUTPUT pointer is a host
 *     pointer the call synchronises `stream` before it returns).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - all buffers are caller-owned; the library owns only the opaque ds_index and per-call workspace
 *     it releases (stream-ordered) before returning.
 */
#ifndef DOPPELSPELLER_B200_H
#define DOPPELSPELLER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_VERSION 100          /* 0.1.0 */
#define DS_N_WORDS 15           /* settings.py:65  NUMBER_OF_WORDS_FEATURES */
#define DS_N_FEATURES 66        /* feature_engineering.py:67  FEATURES_COUNT */
#define DS_MAX_TITLE 255        /* settings.py:68  MAX_CHARACTERS_ALLOWED_IN_THE_TITLE */
#define DS_MAX_TOP_N 1024       /* largest supported top_n */
#define DS_MAX_PEERS 31         /* most other shards whose thresholds ds_topn_local_shared reads */

typedef enum ds_status {
    DS_OK = 0,
    DS_ERR_BAD_ARG = -1,
    DS_ERR_CUDA = -2,
    DS_ERR_NO_MEMORY = -3,
    DS_ERR_UNSUPPORTED = -4,    /* e.g. n_vocab > 65535, top_n > DS_MAX_TOP_N */
    DS_ERR_TOO_FEW_ROWS = -5    /* reserved: ds_topn reports short results through out_count instead and the
                                   binding raises the reference's exception (match_maker.py:188-189) */
} ds_status;

/* ds_topn out_flags bits */
#define DS_FLAG_RESCAN 1        /* more than the retained candidates qualified: exact re-scan was taken */
#define DS_FLAG_FEW_POSITIVE 2  /* fewer than top_n positive scores: the last top_n truth rows were returned */

/* sum mode of the per-query denominator term `max_intersection_possible` (match_maker.py:197) */
#define DS_MX_PY312_COMPENSATED 0   /* CPython >= 3.12 builtin sum(): Neumaier compensated (as executed today) */
#define DS_MX_NAIVE 1               /* CPython <= 3.11 builtin sum(): plain sequential float64 */

typedef struct ds_index ds_index;

int ds_version(void);
const char *ds_last_error(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py gpu_launches) */
int64_t ds_kernel_launches(void);
/* Hands the stream-ordered allocator's cached (idle) workspace memory of `device` back to the driver.  The library
 * keeps freed per-call workspaces cached in the device's default memory pool (bounded by the environment variable
 * DS_POOL_RELEASE_THRESHOLD_MB, default unbounded) so that consecutive calls do not re-map gigabytes; a process that
 * shares the GPU with another allocator (torch's caching allocator uses cudaMalloc) calls this between phases. */
int ds_trim(int device);
/* Measurement hooks (bench.py roofline): between ds_profile_begin and ds_profile_end every launch of the
 * K1 scan kernels (k_scan, k_post) is bracketed by CUDA events on its own stream.  ds_profile_end waits
 * for them and returns the summed device time, the launch count and the (query, truth) pairs scanned. */
int ds_profile_begin(void);
int ds_profile_end(double *scan_ms, int64_t *scan_launches, double *scan_pairs);
/* the same, per kernel: index 0 = k_scan (dense row scan: the first rows and the fallbacks), 1 = k_post
 * (posting-list form of the same scan, the bulk of the rows); each argument points to two elements */
int ds_profile_end_split(double *ms, int64_t *launches, double *pairs);

/* ---------------------------------------------------------------------------------------------------
 * ds_index_create  -  replaces the truth-side half of MatchMaker.__init__ (match_maker.py:97-109:
 * `_construct_truth_data_sparse_matrix` :167-178, `_get_matrix_truth_non_zero_columns_and_values`
 * :122-133) once the host has assigned column ids (:144-147) and idf weights (:135-142).
 *
 *   n_truth, n_vocab     truth rows of THIS shard / columns (n_vocab <= 65535)
 *   t_row_ptr [n_truth+1], t_col_ids [nnz]
 *                        CSR of the truth rows; the column order inside a row is the order in which
 *                        the reference accumulates `sums_matrix_truth` (python set iteration order,
 *                        :172-174).  Rows must not contain duplicates.
 *   idf64_by_col [n_vocab]  float64 idf (query-only n-grams carry max idf, :151,:180-181); the float32
 *                        weights are derived as (float)idf64 exactly like numpy's astype (:130,:152).
 *   sums_truth_f32       optional [n_truth] precomputed `sums_matrix_truth`; NULL => computed on the
 *                        device as the sequential float32 sum in the given column order.
 *   global_row_offset    index of this shard's first row in the whole truth DB (multi-GPU sharding)
 *   n_truth_total        rows of the whole truth DB (== n_truth when not sharded)
 * ------------------------------------------------------------------------------------------------- */
int ds_index_create(ds_index **out, int device, int64_t n_truth, int32_t n_vocab, const int64_t *t_row_ptr,
                    const uint16_t *t_col_ids, const double *idf64_by_col, const float *sums_truth_f32,
                    int64_t global_row_offset, int64_t n_truth_total, void *stream);
int ds_index_destroy(ds_index *index);
/* copies the device-resident `sums_matrix_truth` (float32 [n_truth]) to `out` (host or device) */
int ds_index_get_sums(const ds_index *index, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_transform_titles  -  common.transform_title (common.py:20-47) for a batch of raw titles (SURVEY.md
 * 8(f3)): NFD + ascii-ignore, lower(), '-' -> ' ', keep [a-zA-Z0-9\s], collapse runs of ' ', strip(),
 * [:255].strip(), left-pad with '0' to 3 characters.
 *   codepoints [offsets[n_titles]], offsets [n_titles+1]   the titles as Unicode code points (UTF-32)
 *   ascii_of_cp [table_len]   for every code point below table_len the ASCII character of its canonical
 *                        decomposition (0 = none; no code point has two); code points >= table_len are
 *                        dropped.  The caller derives the table from its Unicode database
 *                        (doppelspeller_b200/common.py builds it with `unicodedata`; 8,816 entries).
 *   out_bytes [capacity offsets[n_titles] + 3 * n_titles + 1], out_offsets [n_titles+1]
 *                        compact table of the transformed titles - the input format of ds_encode_trigrams and
 *                        the ds_*_pairs entry points (letters / digits / spaces; other white space survives
 *                        exactly where the reference keeps it)
 *   out_raw_len [n_titles]   optional: len(text) before the [:255] cut (the reference logs a warning when it is
 *                        below 3 or above 255, common.py:34-45)
 * ------------------------------------------------------------------------------------------------- */
int ds_transform_titles(const uint32_t *codepoints, const int64_t *offsets, int64_t n_titles, const uint8_t *ascii_of_cp,
                        int32_t table_len, uint8_t *out_bytes, int64_t *out_offsets, int32_t *out_raw_len, int device,
                        void *stream);

/* ds_title_features  -  the per-title inputs of construct_features from a transformed-title table (bytes / offsets as
 * ds_transform_titles writes them), on the device:
 *   out_codes        uint8[total bytes] (nullable): FeatureEngineering.encode_title (feature_engineering.py:298-307,
 *                    alphabet '- a..z0..9' -> 0..37) of every title, same offsets; fails with DS_ERR_BAD_ARG on a
 *                    character outside the alphabet (the reference's encode_title fails on it too)
 *   out_word_counts  uint32[n_titles, 15] (nullable): FeatureEngineering.get_truth_words_counts (:309-319) over
 *                    common.get_words_counter (common.py:140-142) of the SAME table: the document frequency (titles
 *                    whose word SET holds the word) of each of the title's first 15 words, 0 padded.  Words = tokens of
 *                    str.split(); a word is identified by its 64-bit FNV-1a hash. */
int ds_title_features(const uint8_t *bytes, const int64_t *offsets, int64_t n_titles, uint8_t *out_codes,
                      uint32_t *out_word_counts, int device, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_encode_trigrams  -  the host half of MatchMaker.__init__ on the GPU: titles ->
 * per-title trigram SETS (common.py:150-151) -> column ids -> document frequencies over the truth sets
 * (common.py:145-147) -> idf = log(N / df), query-only trigrams weighted with the maximum idf
 * (match_maker.py:95,135-153,180-181).  Column ids are CANONICAL (rank of the trigram in ' a..z0..9' code
 * order), not the reference's PYTHONHASHSEED dependent set order; the arithmetic is the reference's.
 *   *_bytes / *_offsets   compact title tables of transform_title output (characters ' a-z0-9' only)
 *   t_row_ptr [n_truth+1], t_col_ids [capacity: total truth bytes]   truth CSR, ascending ids per title
 *   q_row_ptr [n_queries+1], q_col_ids [capacity: total query bytes]
 *   idf64_by_col, vocab_codes   [capacity: ds_encode_max_vocab() = 50,653]; vocab_codes[c] = c0*37^2+c1*37+c2
 *   out_n_vocab / out_truth_nnz / out_query_nnz   host integers; the call synchronises `stream`.
 * The outputs plug straight into ds_index_create / ds_topn (device pointers are used in place).
 * ------------------------------------------------------------------------------------------------- */
int32_t ds_encode_max_vocab(void);
int ds_encode_trigrams(const uint8_t *truth_bytes, const int64_t *truth_offsets, int64_t n_truth,
                       const uint8_t *query_bytes, const int64_t *query_offsets, int64_t n_queries,
                       int64_t *t_row_ptr, uint16_t *t_col_ids, int64_t *q_row_ptr, uint16_t *q_col_ids,
                       double *idf64_by_col, int32_t *vocab_codes, int32_t *out_n_vocab, int64_t *out_truth_nnz,
                       int64_t *out_query_nnz, int device, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_topn  -  replaces `[MatchMaker.get_closest_matches(row) for row in rows]`
 * (match_maker.py:192-203 = python sum :197 + fast_jaccard :16-50 + fast_arg_top_k :53-71), without
 * the title_id lookup of :190 (the binding maps rows to ids).
 *
 *   q_row_ptr [n_q+1], q_col_ids   CSR of the query rows' column ids (any order, no duplicates;
 *                        zero-weight columns may be present, they are no-ops like in the lil matrix)
 *   q_mx                 optional [n_q] float64 `max_intersection_possible`; NULL => computed on the
 *                        device from idf64_by_col over the ascending column ids with `mx_mode`
 *   k                    top_n
 *   out_rows [n_q*k]     GLOBAL truth row indexes, DESCENDING row index (the reference's order,
 *                        match_maker.py:71); -1 padded when fewer than k rows exist
 *   out_count [n_q]      rows written per query (== k unless the DB has fewer than k rows)
 *   out_kth_f32          optional [n_q] k-th largest float32-rounded positive score (0 when fewer)
 *   out_flags            optional [n_q] DS_FLAG_* bits
 * ------------------------------------------------------------------------------------------------- */
int ds_topn(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids,
            const double *q_mx, int32_t mx_mode, int32_t k, int64_t *out_rows, int32_t *out_count,
            float *out_kth_f32, int32_t *out_flags, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Sharded form of ds_topn (truth rows split over GPUs, one ds_index per shard).
 *   phase 1  ds_topn_local : every shard scans its rows and returns its best `m` candidates per query
 *            as (score64, global row), ordered by score descending (ties: higher row first);
 *            unused slots carry score -1 / row -1.  `m` = ds_topn_retained(k).
 *   (the caller all-gathers out_score / out_row over the shards - NCCL all_gather)
 *   phase 2  ds_topn_merge : every shard derives the global k-th key, the threshold and the final
 *            rows from the gathered [n_shards, n_q, m] candidates; queries whose candidate lists
 *            cannot prove the answer are reported in out_flags (DS_FLAG_RESCAN) with the threshold in
 *            out_threshold, and are resolved by
 *   phase 3  ds_topn_rescan : exact re-scan of the local shard for the flagged queries with the known
 *            global threshold, returning the k highest local rows (descending) that reach it.
 * ------------------------------------------------------------------------------------------------- */
int32_t ds_topn_retained(int32_t k);
int ds_topn_local(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids,
                  const double *q_mx, int32_t mx_mode, int32_t k, double *out_score, int64_t *out_row,
                  double *out_mx, void *stream);
/* ds_topn_local with the per-query thresholds shared between the shards WHILE they scan (one process per GPU, NVLink
 * peer memory, e.g. torch.distributed._symmetric_memory): `theta_own` double[n_q] is this shard's published lower bound
 * of the global k-th best score (zero it before the call, then synchronise the shards), `theta_peers[i]` is the
 * peer-mapped device address of shard i's array (n_peers <= DS_MAX_PEERS).  The selection kernel raises its pruning
 * threshold to the best bound any shard has published so far; the k-th best of any subset of the rows is a valid bound
 * and stale values only prune less, so the shards need no further synchronisation.  Results equal ds_topn_local's in
 * every slot the merge can use; without it a shard prunes with its own k-th best only and its late blocks cost as
 * much as its early ones (measured: 7 % on 8 shards of C3). */
int ds_topn_local_shared(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids,
                         const double *q_mx, int32_t mx_mode, int32_t k, double *out_score, int64_t *out_row,
                         double *out_mx, double *theta_own, const double *const *theta_peers, int32_t n_peers,
                         void *stream);
int ds_topn_merge(int32_t n_shards, int64_t n_q, int32_t k, int64_t n_truth_total, const double *all_score,
                  const int64_t *all_row, const double *q_mx /* nullable: out_mx of ds_topn_local */,
                  int64_t *out_rows, int32_t *out_count, float *out_kth_f32, double *out_threshold,
                  int32_t *out_flags, int device, void *stream);
int ds_topn_rescan(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids,
                   const double *q_mx, const double *threshold, const int32_t *flags, int32_t k,
                   int64_t *out_rows, int32_t *out_count, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_indel_ratio_u8  -  replaces fast_levenshtein_ratio (feature_engineering.py:25-63) over n pairs.
 *   a, b [n, stride]  uint8 code rows (the reference's zero padded [P,255] layout: stride 255)
 *   la, lb [n]        lengths (uint8, like the reference's arrays)
 *   out_ratio [n]     uint8 ratio exactly as the reference executes it: (100*(L-d))/L with the
 *                     uint8-wrapping distance d; 0 when la+lb == 0
 *   out_dist          optional [n] uint16 (the uint8-wrapped distance d)
 * ------------------------------------------------------------------------------------------------- */
int ds_indel_ratio_u8(const uint8_t *a, const uint8_t *b, int64_t stride, const uint8_t *la, const uint8_t *lb,
                      int64_t n, uint8_t *out_ratio, uint16_t *out_dist, void *stream);

/* Device-resident string buffers (the [n, stride] code rows above, the byte tables below) are fetched in ALIGNED 32-bit
 * words: the word that holds a buffer's last byte is read whole, so a device allocation must be readable up to the next
 * 4-byte boundary past its last byte.  Every cudaMalloc / cudaMallocAsync / torch allocation is (256-byte granules);
 * host buffers are staged into workspaces the library rounds up itself.  Checked under AddressSanitizer by the host
 * emulation of the kernels (tests/emu), which is where this contract was found to be implicit. */

/* Same arithmetic on a compact title table: pair p compares table_a[idx_a[p]] with table_b[idx_b[p]];
 * a table is (bytes, offsets[n_titles+1]); titles longer than 255 bytes are truncated like
 * FeatureEngineering.encode_title does.  This is the B200 layout (no [P,255] materialisation). */
int ds_indel_ratio_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a,
                         const uint8_t *bytes_b, const int64_t *offsets_b, int64_t n_titles_b,
                         const int32_t *idx_a, const int32_t *idx_b, int64_t n, uint8_t *out_ratio,
                         uint16_t *out_dist, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_levenshtein_ratio_pairs  -  replaces common.levenshtein_ratio (common.py:161-162) over n pairs of
 * raw byte strings: int(round(ratio*100)) with python-levenshtein's ratio = (la+lb-indel)/(la+lb)
 * (true, non-wrapping InDel distance; 1.0 for two empty strings) and Python's round-half-even.
 * Strings are given as a compact table like above.  out [n] int32 in 0..100.
 * ------------------------------------------------------------------------------------------------- */
int ds_levenshtein_ratio_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a,
                               const uint8_t *bytes_b, const int64_t *offsets_b, int64_t n_titles_b,
                               const int32_t *idx_a, const int32_t *idx_b, int64_t n, int32_t *out,
                               void *stream);

/* ---------------------------------------------------------------------------------------------------
 * The fuzzy pre-match of Prediction on the device.
 * ds_prematch_pairs        Prediction._get_levenshtein_ratio (predict.py:147-156) for n (title, candidate) pairs:
 *                          0 when _get_levenshtein_deletion_ratio (:140-145, float64, the written association) is below
 *                          `threshold`; else levenshtein_ratio (common.py:161-162); if that is <= threshold the
 *                          token-sorted ratio (common.py:165-167) instead.  Every title comes twice: as written
 *                          (bytes / offsets) and with its words sorted (`' '.join(sorted(title.split()))`, sorted once per
 *                          title by the caller).  Lengths for the filter are those of the titles as written.
 * ds_select_close_matches  predict.py:158-176 for `n_titles` test titles with `run` consecutive pairs each: out_pair[t] =
 *                          index of the pair whose ratio is > threshold and alone attains the title's maximum, -1 when no
 *                          ratio is > threshold or the maximum is attained more than once.  invalid (nullable): pairs to
 *                          leave out (e.g. padding of short candidate lists).
 * python-levenshtein's ratio is third-party code absent from the reference tree: parity of the two ratios is unpinned
 * (restated: (la + lb - indel) / (la + lb)); the cascade and the selection are pinned against the reference's own
 * Prediction methods (tests/test_oracle_vs_reference.py).
 * ------------------------------------------------------------------------------------------------- */
int ds_prematch_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, const uint8_t *sorted_a, const int64_t *sorted_offsets_a,
                      int64_t n_titles_a, const uint8_t *bytes_b, const int64_t *offsets_b, const uint8_t *sorted_b,
                      const int64_t *sorted_offsets_b, int64_t n_titles_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n,
                      int32_t threshold, int32_t *out_ratio, void *stream);
int ds_select_close_matches(const int32_t *ratios, const uint8_t *invalid, int64_t n_titles, int32_t run, int32_t threshold,
                            int64_t *out_pair, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_construct_features  -  replaces the construct_features gufunc (feature_engineering.py:69-169),
 * layout '(),(),(l),(l),(m),(),(),(n)->(n)' with l = stride, m = 15, n = 66.
 *   la, lb [P] uint8; a, b [P, stride] uint8 codes; counts [P,15] uint32 truth-word document
 *   frequencies; space_code; n_truth (uint32, number_of_truth_titles); out [P,66] float32.
 * ------------------------------------------------------------------------------------------------- */
int ds_construct_features(const uint8_t *la, const uint8_t *lb, const uint8_t *a, const uint8_t *b,
                          int64_t stride, const uint32_t *counts, uint8_t space_code, uint32_t n_truth,
                          int64_t n_pairs, float *out, void *stream);

/* Compact form: titles in (bytes, offsets) tables, truth-word counts per TRUTH TITLE [n_truth_titles,15],
 * pair p = (idx_a[p], idx_b[p]).  Same arithmetic, 373 instead of 836 bytes of traffic per pair. */
int ds_construct_features_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a,
                                const uint8_t *bytes_b, const int64_t *offsets_b, int64_t n_titles_b,
                                const uint32_t *counts_b, const int32_t *idx_a, const int32_t *idx_b,
                                uint8_t space_code, uint32_t n_truth, int64_t n_pairs, float *out,
                                void *stream);

/* ---------------------------------------------------------------------------------------------------
 * ds_gbdt_predict  -  replaces `model.predict(xgb.DMatrix(features))` (predict.py:229-233):
 * inference of a gradient-boosted tree ensemble (train.py:99-121: xgboost 0.90, depth <= 5, <= 1000 rounds)
 * over the float32 feature matrix.  Per row: psum = 0; for every tree in boosting order walk from node 0 -
 * NaN feature -> `missing` child, else feature < value (strict) ? `yes` : `no` - and psum += leaf value
 * (float32); margin = base_margin + psum; DS_GBDT_LOGISTIC applies 1 / (1 + exp(-margin)) in float32
 * (reg:logistic / binary:logistic).  Children are node indexes INSIDE their tree and follow their parent.
 *   features [n_rows, n_features] float32 row major; nodes [tree_offsets[n_trees]]; tree_offsets [n_trees+1]
 *   base_margin: logit(base_score) for the logistic objectives (0 for the default base_score 0.5)
 *   out [n_rows] float32
 * ------------------------------------------------------------------------------------------------- */
typedef struct ds_gbdt_node {
    int32_t feature;    /* column of the feature matrix this node tests; -1 = leaf */
    float value;        /* split threshold, or the leaf's weight */
    uint16_t yes, no;   /* children taken when feature < value / otherwise */
    uint16_t missing;   /* child taken when the feature is NaN (xgboost's default direction) */
    uint16_t reserved;
} ds_gbdt_node;
#define DS_GBDT_MARGIN 0
#define DS_GBDT_LOGISTIC 1
int ds_gbdt_predict(const float *features, int64_t n_rows, int32_t n_features, const ds_gbdt_node *nodes,
                    const int32_t *tree_offsets, int32_t n_trees, float base_margin, int32_t transform, float *out,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DOPPELSPELLER_B200_H */
