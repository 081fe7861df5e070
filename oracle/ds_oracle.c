/*
 * ds_oracle.c - CPU restatement of DoppelSpeller's candidate-generation + pair-scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (doppelspeller_b200/) never does.
 *
 * Plain C restatement of the reference's numba kernels *as executed* by numba 0.65.0 in the build
 * container.  Parity is pinned in tests/test_oracle_vs_reference.py
 * (runs where /root/reference exists) and by the golden vectors under tests/golden/ minted from the
 * reference's own functions by tests/golden/make_golden.py.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_WORDS 15      /* settings.py:65  NUMBER_OF_WORDS_FEATURES */
#define ORC_FEATURES 66   /* feature_engineering.py:67  6 + 4 * 15 */

int orc_version(void) { return 1; }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * fast_jaccard - doppelspeller/match_maker.py:16-50
 *   scores = zeros(N, f32); for col in query cols (ascending): scores[posting rows] += idf32(col)
 *   return f64(scores) / (f64(sums) + (mx - f64(scores)))
 * `scores` is caller-provided scratch of N floats, `out` N doubles.
 * ---------------------------------------------------------------------------------------------- */
static void jaccard_one(int64_t n_truth, double mx, const int32_t *q_cols, int64_t n_q_cols,
                        const int64_t *post_ptr, const int32_t *post_rows, const float *w32,
                        const float *sums, float *scores, double *out) {
    memset(scores, 0, (size_t)n_truth * sizeof(float));
    for (int64_t i = 0; i < n_q_cols; ++i) {
        int32_t col = q_cols[i];
        float w = w32[col];
        for (int64_t p = post_ptr[col]; p < post_ptr[col + 1]; ++p) {
            scores[post_rows[p]] += w; /* one f32 rounding per add (built with -ffp-contract=off) */
        }
    }
    for (int64_t t = 0; t < n_truth; ++t) {
        double sc = (double)scores[t];
        double inner = mx - sc;
        double den = (double)sums[t] + inner;
        out[t] = sc / den;
    }
}

/* ------------------------------------------------------------------------------------------------
 * fast_arg_top_k - doppelspeller/match_maker.py:53-71 (literal restatement of the replace-min loop)
 *   slots = zeros(k, f32); min_idx = 0; min_val = 0
 *   for value in array: if value > min_val: slots[min_idx] = f32(value); min_idx = argmin(slots);
 *                                            min_val = slots[min_idx]
 *   min_val -= f32(1e-6)   (settings.py:72, evaluated in float64)
 *   return nonzero(array >= min_val)[::-1][:k]
 * Returns the number of rows written (<= k), rows in DESCENDING index order; *kth_key gets the final
 * f32 slot minimum before the buffer is subtracted.
 * ---------------------------------------------------------------------------------------------- */
static int64_t arg_top_k(const double *array, int64_t n, int64_t k, float *slots, int64_t *out_rows,
                         float *kth_key) {
    for (int64_t i = 0; i < k; ++i) slots[i] = 0.0f;
    int64_t min_idx = 0;
    double min_val = 0.0;
    for (int64_t t = 0; t < n; ++t) {
        double value = array[t];
        if (value > min_val) {
            slots[min_idx] = (float)value;
            int64_t arg = 0;
            for (int64_t i = 1; i < k; ++i)
                if (slots[i] < slots[arg]) arg = i; /* np.argmin: first minimum */
            min_idx = arg;
            min_val = (double)slots[min_idx];
        }
    }
    if (kth_key) *kth_key = (float)min_val;
    const float buffer = 1e-6f; /* np.finfo(np.float32).resolution */
    double threshold = min_val - (double)buffer;
    int64_t count = 0;
    for (int64_t t = n - 1; t >= 0 && count < k; --t)
        if (array[t] >= threshold) out_rows[count++] = t;
    return count;
}

/* single-array entry points (used to pin the restatement against the reference's jitted functions) */
void orc_fast_jaccard(int64_t n_truth, double mx, const int32_t *q_cols, int64_t n_q_cols,
                      const int64_t *post_ptr, const int32_t *post_rows, const float *w32,
                      const float *sums, double *out) {
    float *scores = (float *)malloc((size_t)(n_truth > 0 ? n_truth : 1) * sizeof(float));
    jaccard_one(n_truth, mx, q_cols, n_q_cols, post_ptr, post_rows, w32, sums, scores, out);
    free(scores);
}

int64_t orc_fast_arg_top_k(const double *array, int64_t n, int64_t k, int64_t *out_rows, float *kth_key) {
    float *slots = (float *)malloc((size_t)(k > 0 ? k : 1) * sizeof(float));
    int64_t c = arg_top_k(array, n, k, slots, out_rows, kth_key);
    free(slots);
    return c;
}

/* ------------------------------------------------------------------------------------------------
 * `sum([idf(col) for col in cols])` - match_maker.py:197, a Python builtin sum() over Python floats.
 * As executed by CPython >= 3.12 this is Neumaier-compensated (Python/bltinmodule.c builtin_sum):
 * the first item is added to int 0 exactly, the rest go through
 *     t = s + x;  c += |s| >= |x| ? (s - t) + x : (x - t) + s;  s = t
 * and the compensation is added once at the end when it is non-zero and finite.  (CPython 3.7, the
 * version the reference pins, summed naively; `compensated == 0` selects that.)
 * ---------------------------------------------------------------------------------------------- */
double orc_py_float_sum_mode(const double *w64, const int32_t *cols, int64_t n, int compensated) {
    if (n == 0) return 0.0;
    double s = w64[cols[0]], c = 0.0;
    for (int64_t i = 1; i < n; ++i) {
        double x = w64[cols[i]];
        double t = s + x;
        if (compensated) {
            if (fabs(s) >= fabs(x)) c += (s - t) + x;
            else c += (x - t) + s;
        }
        s = t;
    }
    if (c != 0.0 && isfinite(c)) s += c;
    return s;
}

double orc_py_float_sum(const double *w64, const int32_t *cols, int64_t n) {
    return orc_py_float_sum_mode(w64, cols, n, 1);
}

/* ------------------------------------------------------------------------------------------------
 * MatchMaker.get_closest_matches for a batch of queries - match_maker.py:192-203
 *   mx = python-float sum of idf64 over the query's ascending column ids (:197)
 *   fast_jaccard (:199) -> fast_arg_top_k (:187)
 * q_mx may be NULL: then mx = orc_py_float_sum of w64[col] over q_cols (ascending, :197).
 * out_rows [n_q * k] (descending truth row, -1 padded), out_count [n_q], out_kth [n_q] (nullable).
 * OpenMP over queries (the reference itself is one python thread + numba parfors inside).
 * ---------------------------------------------------------------------------------------------- */
int orc_topn_batch(int64_t n_truth, const int64_t *post_ptr, const int32_t *post_rows, const float *w32,
                   const double *w64, const float *sums, int64_t n_q, const int64_t *q_ptr,
                   const int32_t *q_cols, const double *q_mx, int64_t k, int64_t *out_rows,
                   int64_t *out_count, float *out_kth, int n_threads) {
    if (n_threads <= 0) n_threads = orc_max_threads();
    int failed = 0;
#pragma omp parallel num_threads(n_threads)
    {
        float *scores = (float *)malloc((size_t)(n_truth > 0 ? n_truth : 1) * sizeof(float));
        double *jac = (double *)malloc((size_t)(n_truth > 0 ? n_truth : 1) * sizeof(double));
        float *slots = (float *)malloc((size_t)(k > 0 ? k : 1) * sizeof(float));
        if (!scores || !jac || !slots) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (int64_t q = 0; q < n_q; ++q) {
                const int32_t *cols = q_cols + q_ptr[q];
                int64_t n_cols = q_ptr[q + 1] - q_ptr[q];
                double mx;
                if (q_mx) {
                    mx = q_mx[q];
                } else {
                    mx = orc_py_float_sum(w64, cols, n_cols);
                }
                jaccard_one(n_truth, mx, cols, n_cols, post_ptr, post_rows, w32, sums, scores, jac);
                float kth;
                int64_t c = arg_top_k(jac, n_truth, k, slots, out_rows + q * k, &kth);
                for (int64_t i = c; i < k; ++i) out_rows[q * k + i] = -1;
                out_count[q] = c;
                if (out_kth) out_kth[q] = kth;
            }
        }
        free(scores);
        free(jac);
        free(slots);
    }
    return failed ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * sums_matrix_truth - match_maker.py:172-174: python sum() of the float32 idf values in the row's
 * n-gram SET ITERATION order => sequential float32 accumulation in the order given.
 * ---------------------------------------------------------------------------------------------- */
void orc_truth_sums(int64_t n_truth, const int64_t *t_ptr, const int32_t *t_cols_in_set_order,
                    const float *w32, float *sums) {
    for (int64_t t = 0; t < n_truth; ++t) {
        float acc = 0.0f;
        for (int64_t p = t_ptr[t]; p < t_ptr[t + 1]; ++p) acc = acc + w32[t_cols_in_set_order[p]];
        sums[t] = acc;
    }
}

/* ------------------------------------------------------------------------------------------------
 * fast_levenshtein_ratio's distance - feature_engineering.py:25-61
 *   InDel DP (match 0, mismatch +2, insert/delete +1); every cell is STORED as uint8 (:42), the
 *   candidates are formed in wide integers (numba types `uint8 + 1` as int64) -> wrap happens on store.
 *   The shorter sequence indexes the rows (:35-37) - irrelevant for the value, kept for fidelity.
 * ---------------------------------------------------------------------------------------------- */
static int indel_u8(const uint8_t *a, int la, const uint8_t *b, int lb) {
    if (la > lb) {
        const uint8_t *ts = a; a = b; b = ts;
        int tl = la; la = lb; lb = tl;
    }
    uint8_t *prev = (uint8_t *)malloc((size_t)(lb + 1) * 2);
    uint8_t *cur = prev + (lb + 1);
    for (int y = 0; y <= lb; ++y) prev[y] = (uint8_t)y;
    for (int x = 1; x <= la; ++x) {
        cur[0] = (uint8_t)x;
        for (int y = 1; y <= lb; ++y) {
            int up = prev[y] + 1, left = cur[y - 1] + 1;
            int diag = prev[y - 1] + (a[x - 1] == b[y - 1] ? 0 : 2);
            int m = up < diag ? up : diag;
            if (left < m) m = left;
            cur[y] = (uint8_t)m;
        }
        uint8_t *tmp = prev; prev = cur; cur = tmp;
    }
    int d = prev[lb];
    free(prev < cur ? prev : cur);
    return d;
}

int orc_indel_distance_u8(const uint8_t *a, int la, const uint8_t *b, int lb) { return indel_u8(a, la, b, lb); }

/* feature_engineering.py:63  `((total - d) / total) * 100` -> uint8.  Under numba fastmath the
 * expression executes as ((total - d) * 100) / total then truncates, i.e. the integer
 * (100 * (total - d)) / total  (SURVEY.md 0.8).  total == 0 raises ZeroDivisionError in the
 * reference (unreachable for real titles); here it returns 0. */
int orc_indel_ratio_u8(const uint8_t *a, int la, const uint8_t *b, int lb) {
    int total = la + lb;
    if (total == 0) return 0;
    int d = indel_u8(a, la, b, lb);
    return ((100 * (total - d)) / total) & 0xFF;
}

void orc_indel_ratio_u8_batch(const uint8_t *a, const uint8_t *b, int64_t stride, const uint8_t *la,
                              const uint8_t *lb, int64_t n, uint8_t *out_ratio, int n_threads) {
    if (n_threads <= 0) n_threads = orc_max_threads();
#pragma omp parallel for schedule(static, 256) num_threads(n_threads)
    for (int64_t i = 0; i < n; ++i)
        out_ratio[i] = (uint8_t)orc_indel_ratio_u8(a + i * stride, la[i], b + i * stride, lb[i]);
}

/* ------------------------------------------------------------------------------------------------
 * python-levenshtein 0.12.0 `ratio` (third-party, not in /root/reference; requirements.txt:9) as used
 * by common.py:161-167: ratio = (la + lb - indel) / (la + lb) with the TRUE (non-wrapping) InDel
 * distance, 1.0 when both strings are empty; levenshtein_ratio = int(round(ratio * 100)) with
 * Python's round-half-even on the float64.  Parity for this function is UNPINNED.
 * ---------------------------------------------------------------------------------------------- */
static int lcs_len(const uint8_t *a, int la, const uint8_t *b, int lb) {
    int *prev = (int *)calloc((size_t)(lb + 1) * 2, sizeof(int));
    int *cur = prev + (lb + 1);
    for (int x = 1; x <= la; ++x) {
        cur[0] = 0;
        for (int y = 1; y <= lb; ++y) {
            if (a[x - 1] == b[y - 1]) cur[y] = prev[y - 1] + 1;
            else cur[y] = prev[y] >= cur[y - 1] ? prev[y] : cur[y - 1];
        }
        int *tmp = prev; prev = cur; cur = tmp;
    }
    int r = prev[lb];
    free(prev < cur ? prev : cur);
    return r;
}

int orc_indel_distance(const uint8_t *a, int la, const uint8_t *b, int lb) {
    return la + lb - 2 * lcs_len(a, la, b, lb);
}

int orc_levenshtein_ratio(const uint8_t *a, int la, const uint8_t *b, int lb) {
    int total = la + lb;
    double ratio = 1.0;
    if (total > 0) ratio = (double)(total - orc_indel_distance(a, la, b, lb)) / (double)total;
    double scaled = ratio * 100.0;
    return (int)nearbyint(scaled); /* default rounding mode: to nearest, ties to even */
}

/* Prediction._get_levenshtein_deletion_ratio + the prefilter test of _get_levenshtein_ratio
 * (predict.py:140-151): returns 1 when ((la+lb-|la-lb|)/(la+lb))*100 < 94 (pair rejected, value 0). */
int orc_prefilter_rejects(int la, int lb) {
    int total = la + lb;
    int delta = la > lb ? la - lb : lb - la;
    double q = (double)(total - delta) / (double)total;
    double r = q * 100.0;
    return r < 94.0;
}

/* ------------------------------------------------------------------------------------------------
 * construct_features - feature_engineering.py:75-169, one (title, truth) pair.
 * title/truth are code arrays (space = space_code), counts = first 15 truth-word document frequencies.
 * ---------------------------------------------------------------------------------------------- */
void orc_construct_features(int la, int lb, const uint8_t *title, const uint8_t *truth,
                            const uint32_t *counts, uint8_t space_code, uint32_t n_truth, float *out) {
    const float nanf_ = NAN;
    int words_a = 1, words_b = 1;                                   /* :104-105 */
    for (int i = 0; i < la; ++i) words_a += (title[i] == space_code);
    for (int i = 0; i < lb; ++i) words_b += (truth[i] == space_code);
    int lev = orc_indel_ratio_u8(title, la, truth, lb);             /* :106 */

    uint8_t a_ns[256];                                              /* :108 title_wo_spaces */
    int n_ns = 0;
    for (int i = 0; i < la; ++i)
        if (title[i] != space_code) a_ns[n_ns++] = title[i];

    float best_ratios[ORC_WORDS], word_lengths[ORC_WORDS], idf_s[ORC_WORDS];
    for (int i = 0; i < ORC_WORDS; ++i) best_ratios[i] = word_lengths[i] = idf_s[i] = nanf_;   /* :121-123 */

    /* reconstructed title without the leading space of :115; trailing space dropped at the end (:161) */
    uint8_t *recon = (uint8_t *)malloc((size_t)(lb + 2 * ORC_WORDS + 8 + 256));
    int n_recon = 0;

    int n_words = 0, last = 0;
    for (int pos = 0; pos <= lb && n_words < ORC_WORDS; ++pos) {    /* :110-114 first 15 separators */
        int is_sep = (pos == lb) || (truth[pos] == space_code);
        if (!is_sep) continue;
        const uint8_t *word = truth + last;                         /* :128-131 */
        int wl = pos - last;
        last = pos + 1;

        int best_ratio = 0;                                         /* :136-149 */
        const uint8_t *best_match = &space_code;
        int best_len = 1;
        for (int i = 0; i < n_ns; ++i) {
            int pl = (i + wl <= n_ns) ? wl : (n_ns - i);
            if (pl == 0) break;
            int r = orc_indel_ratio_u8(a_ns + i, pl, word, wl);
            if (r > best_ratio) { best_ratio = r; best_match = a_ns + i; best_len = pl; }
        }
        best_ratios[n_words] = (float)best_ratio;                   /* :151-153 */
        word_lengths[n_words] = (float)wl;
        idf_s[n_words] = (float)log((double)n_truth / (double)counts[n_words]);
        memcpy(recon + n_recon, best_match, (size_t)best_len);      /* :154-155 */
        n_recon += best_len;
        recon[n_recon++] = space_code;
        ++n_words;
    }
    if (n_recon > 0) --n_recon; /* drop the trailing space; the leading one was never added (:161) */

    /* :158  ranks = 1 + ((nanmax(idf_s) - idf_s) / truth_number_of_words).  As executed under
     * fastmath nanmax returns NaN as soon as one slot is NaN => all ranks NaN unless
     * all 15 word slots are filled.  float32 subtraction, then float64 divide / add, cast to float32. */
    float ranks[ORC_WORDS];
    if (n_words < ORC_WORDS) {
        for (int i = 0; i < ORC_WORDS; ++i) ranks[i] = nanf_;
    } else {
        float mxv = idf_s[0];
        for (int i = 1; i < ORC_WORDS; ++i)
            if (idf_s[i] > mxv) mxv = idf_s[i];
        for (int i = 0; i < ORC_WORDS; ++i) {
            float diff = mxv - idf_s[i];
            double q = (double)diff / (double)words_b;
            ranks[i] = (float)(1.0 + q);
        }
    }
    int recon_lev = orc_indel_ratio_u8(recon, n_recon, truth, lb);  /* :161-162 */
    free(recon);

    out[0] = (float)la; out[1] = (float)lb; out[2] = (float)words_a; out[3] = (float)words_b;   /* :164-169 */
    out[4] = (float)lev; out[5] = (float)recon_lev;
    memcpy(out + 6, best_ratios, sizeof(best_ratios));
    memcpy(out + 6 + ORC_WORDS, word_lengths, sizeof(word_lengths));
    memcpy(out + 6 + 2 * ORC_WORDS, idf_s, sizeof(idf_s));
    memcpy(out + 6 + 3 * ORC_WORDS, ranks, sizeof(ranks));
}

/* gufunc-style batch over P pairs in the reference's [P,255] layout (feature_engineering.py:69-80). */
void orc_construct_features_batch(const uint8_t *la, const uint8_t *lb, const uint8_t *title,
                                  const uint8_t *truth, int64_t stride, const uint32_t *counts,
                                  uint8_t space_code, uint32_t n_truth, int64_t n_pairs, float *out,
                                  int n_threads) {
    if (n_threads <= 0) n_threads = orc_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int64_t p = 0; p < n_pairs; ++p)
        orc_construct_features(la[p], lb[p], title + p * stride, truth + p * stride, counts + p * ORC_WORDS,
                               space_code, n_truth, out + p * ORC_FEATURES);
}
