/* TEST INFRASTRUCTURE: a plain C99 client of include/doppelspeller_b200.h - no Python, no torch, no C++.
 *
 * Proves that the drop-in boundary is a C ABI a maintainer can bind from any language: the header compiles as ISO C
 * (-std=c99 -pedantic -Wall -Werror), the library links with nothing but itself, and host buffers in / host buffers
 * out give the reference's answers:
 *   - MatchMaker top-n on a six-title truth DB (match_maker.py:16-71,192-203; rows from the CPU oracle),
 *   - fast_levenshtein_ratio("coolblu bv", "coolblue bv") = 95 (feature_engineering.py:25-63; docstring encoding :28-29),
 *   - construct_features of the same pair: [10, 11, 2, 2, 95, 90], best word ratios [87, 100], word lengths [8, 2]
 *     (feature_engineering.py:75-169; the values SURVEY.md 8(c) probed on the reference).
 * Exit status: 0 = every answer right; 3 = the library refused to compute because no GPU is usable (DS_ERR_CUDA: there is
 * no CPU fallback); 1 = a wrong answer or an unexpected status.  Driven by tests/test_library_abi.py. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "doppelspeller_b200.h"

static int check(int status, const char *what) {
    if (status == DS_OK) return 0;
    printf("%s: status %d: %s\n", what, status, ds_last_error());
    return status == DS_ERR_CUDA ? 3 : 1;
}

int main(void) {
    /* six truth titles over six trigram columns; two query titles */
    const double idf64[6] = {1.5, 0.7, 2.2, 0.3, 1.1, 0.9};
    const int64_t t_row_ptr[7] = {0, 2, 5, 7, 9, 12, 15};
    const uint16_t t_col_ids[15] = {0, 1, 1, 2, 3, 0, 3, 2, 4, 3, 4, 5, 0, 1, 2};
    const int64_t q_row_ptr[3] = {0, 3, 5};
    const uint16_t q_col_ids[5] = {0, 1, 2, 3, 5};
    const int64_t want_rows[2][3] = {{5, 1, 0}, {4, 2, 1}};
    const float want_kth[2] = {0.5f, 0.0731707364320755f};
    /* "coolblu bv" / "coolblue bv" in the reference's character codes */
    uint8_t a[255], b[255];
    const uint8_t code_a[10] = {4, 16, 16, 13, 3, 13, 22, 1, 3, 23};
    const uint8_t code_b[11] = {4, 16, 16, 13, 3, 13, 22, 6, 1, 3, 23};
    const uint8_t la = 10, lb = 11;
    uint32_t counts[DS_N_WORDS] = {1, 2145};
    const float want_basic[8] = {10, 11, 2, 2, 95, 90, 87, 100};

    ds_index *index = NULL;
    int64_t rows[2][3];
    int32_t count[2], flags[2];
    float kth[2], features[DS_N_FEATURES];
    uint8_t ratio = 0;
    uint16_t dist = 0;
    int rc, q, j;

    printf("ds_version %d, %d features, top_n <= %d\n", ds_version(), DS_N_FEATURES, DS_MAX_TOP_N);
    if (ds_version() != DS_VERSION) return 1;

    rc = check(ds_index_create(&index, 0, 6, 6, t_row_ptr, t_col_ids, idf64, NULL, 0, 6, NULL), "ds_index_create");
    if (rc) return rc;
    rc = check(ds_topn(index, 2, q_row_ptr, q_col_ids, NULL, DS_MX_PY312_COMPENSATED, 3, &rows[0][0], count, kth, flags, NULL),
               "ds_topn");
    if (rc) return rc;
    for (q = 0; q < 2; q++) {
        printf("query %d: rows %lld %lld %lld, count %d, kth %.9g\n", q, (long long)rows[q][0], (long long)rows[q][1],
               (long long)rows[q][2], (int)count[q], (double)kth[q]);
        if (count[q] != 3 || kth[q] != want_kth[q]) return 1;
        for (j = 0; j < 3; j++)
            if (rows[q][j] != want_rows[q][j]) return 1;
    }
    if (check(ds_index_destroy(index), "ds_index_destroy")) return 1;

    memset(a, 0, sizeof a);
    memset(b, 0, sizeof b);
    memcpy(a, code_a, sizeof code_a);
    memcpy(b, code_b, sizeof code_b);
    rc = check(ds_indel_ratio_u8(a, b, DS_MAX_TITLE, &la, &lb, 1, &ratio, &dist, NULL), "ds_indel_ratio_u8");
    if (rc) return rc;
    printf("fast_levenshtein_ratio %d (distance %d)\n", (int)ratio, (int)dist);
    if (ratio != 95 || dist != 1) return 1;

    rc = check(ds_construct_features(&la, &lb, a, b, DS_MAX_TITLE, counts, 1, 30000, 1, features, NULL), "ds_construct_features");
    if (rc) return rc;
    for (j = 0; j < 8; j++)
        if (features[j] != want_basic[j]) return 1;
    if (features[21] != 8.0f || features[22] != 2.0f || !isnan(features[8]) || !isnan(features[65])) return 1;
    if (fabs(features[36] - 10.308952331542969) > 1e-5 || fabs(features[37] - 2.6380579471588135) > 1e-5) return 1;
    printf("construct_features ok\n");
    return 0;
}
