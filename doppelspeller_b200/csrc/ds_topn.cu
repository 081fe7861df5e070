// ds_topn.cu - K1: IDF-weighted trigram Jaccard scan of the truth index fused with the reference's
// top-n selection rule, plus the index build and the sharded local / merge / rescan phases.
//
// Reference semantics reproduced bit-for-bit (paths relative to /root/reference):
//   score    doppelspeller/match_maker.py:16-50   fast_jaccard
//            sc = f32 sequential sum, in ASCENDING column id order, of idf32 over the columns the
//            query and the truth row share; s64 = f64(sc) / (f64(sums) + (mx - f64(sc)))
//   select   doppelspeller/match_maker.py:53-71   fast_arg_top_k
//            T = k-th largest f32(s64) over positive scores (0 when fewer than k),
//            thr = f64(T) - f64(f32(1e-6)); result = the k HIGHEST row indexes with s64 >= thr,
//            in descending row order.
//   mx       doppelspeller/match_maker.py:197     python sum() over the ascending column ids
//
// Design (B200): two forms of the same scan.  k_post (the bulk of the rows) walks the inverted index: per
// block of 2,048 row positions and column id the rows holding the column; one warp = one query x a run of
// blocks, one float32 accumulator per row in shared memory, the query's columns walked in ascending id
// order (see the kernel).  k_scan (the first 4,096 rows that seed the thresholds, the fixed-threshold and
// bounded-memory fallbacks, small or non-monotone indexes) streams the rows themselves: they live in HBM as
// 8-byte chunks of four ascending u16 column ids
// (sentinel padded); a CTA owns a tile of up to 32 queries whose columns are scattered once into a
// shared-memory direct map  column id -> slot -> (32-bit query mask, idf32)  and then streams truth
// rows, one row per thread, with 32 register accumulators.  Adding the row's columns in ascending
// order reproduces the reference's float32 accumulation order for every (query, row) pair at once;
// the score matrix never exists.  Selection is a conservative float32 threshold test per pair
// (no false negatives, see `filter_from_threshold`); survivors are re-scored in float64 exactly as
// the reference does and merged into a per-query retained list that drives the threshold.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <new>
#include <type_traits>

#include "ds_common.cuh"

namespace ds {

thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_kernel_launches{0};

// optional timing of the dominant kernel (k_scan) for bench.py's roofline: CUDA events on the launching stream
struct ScanProfile {
    bool enabled = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events[2];   // [0] k_scan launches, [1] k_post launches
    double pairs[2] = {0.0, 0.0};
};
static ScanProfile g_profile;
static std::mutex g_profile_lock;   // ds_profile_* and the launch helpers may run on different host threads

constexpr int TQ = 32;              // queries per tile (one bit each in the slot mask)
constexpr int CHUNK_COLS = 4;       // u16 column ids per chunk: 4 -> one LDG.64 per chunk (7.7 % sentinel padding at the
                                    // example length distribution; 8 -> LDG.128, 17 % padding)
typedef uint2 chunk_t;
static_assert(sizeof(chunk_t) == CHUNK_COLS * 2, "chunk type and CHUNK_COLS disagree");
__device__ __forceinline__ chunk_t zero_chunk() { return make_uint2(0, 0); }
constexpr int MAX_SLOTS = 2048;     // slot 0 = "column not in this tile"
constexpr int CAND_CAP_MAX = 1024;  // per query candidate buffer (per scan launch): clamp(8 * k, 256, 1024)
constexpr int CAND_CAP_WIDE = 4096; // second try for the queries whose buffer overflowed (massive ties), before the dense fallback
constexpr int DENSE_ROWS = 256;     // rows of the first (dense, threshold seeding) chunk
constexpr int FIXED_ROWS = 1024;    // chunk rows of the bounded-memory fallback / rescan passes
constexpr int QUERY_BATCH = 65536;  // queries per workspace batch
constexpr int STATE_OVERFLOW = 1;
constexpr int SORT_BLOCK = 4096;    // rows are re-ordered (by their weight sum) inside blocks of this many consecutive rows
constexpr int POST_ROWS = 4096;     // row positions per posting block = u16 fixed-point accumulators per warp in k_post (8 KB)
constexpr int POST_WARPS = 13;      // warps per CTA; two CTAs (26 warps, 213 KB of accumulators) per SM
constexpr int POST_LIST = 32;       // per warp: rows of the swept block waiting for the full filter
constexpr int POST_DESC = 64;       // per warp: listed pieces (start inside the block's postings | length << 24, fixed-point weight) of the block being walked
constexpr int DENSE_MAX = 32;       // columns kept out of the postings: their presence is a 32-bit pattern per row
constexpr int DENSE_MIN_SHARE = 128;   // a column is dense only if it sits in at least 1 / 128 of the rows
// tuning knobs of k_post, overridable at build time (-DDS_POST_...=n) for A/B builds
#ifndef DS_POST_RUN
#define DS_POST_RUN 4
#endif
#ifndef DS_POST_DEPTH
#define DS_POST_DEPTH 4
#endif
#ifndef DS_POST_TASKS
#define DS_POST_TASKS 8
#endif
#ifndef DS_POST_GROWTH
#define DS_POST_GROWTH 3
#endif
#ifndef DS_SMALL_BATCH
#define DS_SMALL_BATCH 32768   // batches up to this many queries sweep with x4 growing ranges
#endif
constexpr int POST_RUN = DS_POST_RUN;      // consecutive posting blocks per warp task (the query's columns are loaded once)
constexpr int POST_DEPTH = DS_POST_DEPTH;  // posting pieces (<= 64 postings each) in flight per warp
constexpr int POST_GROUPS = POST_ROWS / 128;   // bars are kept per 128 rows: 32 groups, one lane each
constexpr int POST_CTAS = 2;
typedef uint32_t post_off_t;                   // segment offsets inside a block
typedef uint16_t post_acc_t;                   // fixed-point accumulator
static_assert(SORT_BLOCK % POST_ROWS == 0 && POST_GROUPS == 32 && POST_ROWS <= 65536, "posting blocks must tile the sort blocks");
constexpr int MODE_SCORE = 0;       // retained list = best m by (score, row); drives the running threshold
constexpr int MODE_ROW = 1;         // retained list = the k highest rows with s64 >= a fixed threshold (rescan)

struct Index {
    int device = 0;
    cudaStream_t last_stream = nullptr;   // stream of the latest call that used the index: ds_index_destroy frees on it
    int64_t n_truth = 0;
    int32_t n_vocab = 0;
    int64_t row_offset = 0;
    int64_t n_total = 0;
    uint64_t n_chunks = 0;
    // Rows are stored in a permuted order: inside every block of SORT_BLOCK consecutive rows they are sorted
    // by their weight sum (which also sorts them by length, nearly: the 32 rows of a k_scan warp stream about the
    // same number of chunks).  Everything indexed by "position" below uses that order; perm[position] is the
    // shard-local original row.
    int32_t *perm = nullptr;        // [n_truth] position -> original row
    chunk_t *chunks = nullptr;      // [n_chunks] CHUNK_COLS ascending u16 column ids, sentinel = n_vocab
    uint32_t *chunk_ptr = nullptr;  // [n_truth + 1] by position
    float *sums_pos = nullptr;      // [n_truth] sums_matrix_truth by position
    float *sums = nullptr;          // [n_truth] sums_matrix_truth by original row
    float *w32 = nullptr;           // [n_vocab + 1], w32[n_vocab] = 0 (sentinel)
    double *w64 = nullptr;          // [n_vocab]
    // The same rows inverted (k_post): for every block of POST_ROWS positions and every column id, the
    // block-local positions of the rows holding that column.  Segment (block s, column c) is
    // post[seg_base[s] + seg_off[s * (n_vocab + 1) + c] .. seg_base[s] + seg_off[s * (n_vocab + 1) + c + 1]).
    // Null when the index is small, a weight or a row sum is negative / NaN (partial sums must grow and the
    // filter's right-hand side must not fall below b), a row repeats a column or a block holds > 65,535 postings.
    uint16_t *post = nullptr;
    post_off_t *seg_off = nullptr;
    uint32_t *seg_base = nullptr;
    float *sums_floor = nullptr;    // [n_sub * POST_GROUPS] smallest sums_pos of every group of 128 positions (+inf padded)
    int n_sub = 0;
    // Dense columns: the up to 32 columns with the highest document frequency (each in >= 1 / 128 of the rows) have no
    // postings; which of them a row holds is a 32-bit pattern (bit 31 = the commonest column).
    uint8_t *dense_bit = nullptr;   // [n_vocab + 1] bit of a dense column, 255 = not dense
    float *dense_w = nullptr;       // [32] w32 of the dense column of every bit (0 where unused)
    uint32_t *pat_pos = nullptr;    // [n_truth] dense pattern by position
    uint32_t *group_pat = nullptr;  // [n_sub * POST_GROUPS] OR of the patterns of every group of 128 positions
    int n_dense = 0;
};

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
// s64 exactly as numba evaluates `scores / (sums + (mx - scores))` (match_maker.py:50): float64, this
// association, one IEEE rounding per operation, no FMA contraction.
__device__ __forceinline__ double exact_score(float sc, float sum_truth, double mx) {
    double sc64 = (double)sc;
    double den = __dadd_rn((double)sum_truth, __dsub_rn(mx, sc64));
    return __ddiv_rn(sc64, den);
}

// Conservative float32 pre-filter for "s64 >= theta":  sc > fmaf(a, sums, b).
// For theta > 0:  s64 >= theta  <=>  sc >= c (sums + mx), c = theta / (1 + theta).  a and b are rounded
// DOWN with a 4e-6 relative margin, two orders of magnitude above the float32 rounding of the test
// itself and the float64 rounding of s64, so a true qualifier can never be rejected.  theta <= 0 gives
// a = b = 0, i.e. every positive score passes.
__device__ __forceinline__ float2 filter_from_threshold(double theta, double mx) {
    if (!(theta > 0.0)) return make_float2(0.0f, 0.0f);
    double c = theta / (1.0 + theta);
    float a = __double2float_rd(c * (1.0 - 4e-6));
    float b = __double2float_rd((double)a * mx * (1.0 - 1e-6));
    if (!(b >= 0.0f)) b = 0.0f;
    return make_float2(a, b);
}

__device__ __forceinline__ double threshold_from_key(float kth_key) {
    const float buffer = 1e-6f;  // settings.py:72 np.finfo(np.float32).resolution
    return __dsub_rn((double)kth_key, (double)buffer);
}

// strict total order "better": higher score first, ties: higher row first
__device__ __forceinline__ bool better(double sa, int64_t ra, double sb, int64_t rb) {
    return sa > sb || (sa == sb && ra > rb);
}

// ---------------------------------------------------------------------------------------------------
// index build kernels
// ---------------------------------------------------------------------------------------------------
__global__ void k_weights(const double *__restrict__ w64, float *__restrict__ w32, int n_vocab, int *__restrict__ not_monotone) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vocab) {
        const float w = (float)w64[i];  // numpy astype(float32): round to nearest even
        w32[i] = w;
        if (!(w >= 0.0f)) atomicOr(not_monotone, 1);   // negative / NaN idf: partial sums may decrease (k_post needs growth)
    }
    if (i == n_vocab) w32[i] = 0.0f;
}

// Postings of the packed rows, one thread per row position.  PASS 0 counts the (block, column) segment sizes
// into `seg`; PASS 1 appends the block-local position to each of its segments (`seg` = running cursors).
template <int PASS>
__global__ void k_post_build(const uint16_t *__restrict__ packed, const uint32_t *__restrict__ chunk_ptr, int64_t n_rows,
                             int n_vocab, const uint8_t *__restrict__ dense_bit, uint32_t *__restrict__ seg, uint16_t *__restrict__ post,
                             int *__restrict__ repeated) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n_rows) return;
    const uint16_t *cols = packed + (size_t)chunk_ptr[pos] * CHUNK_COLS;
    const int n = (int)(chunk_ptr[pos + 1] - chunk_ptr[pos]) * CHUNK_COLS;
    uint32_t *block_seg = seg + (size_t)(pos / POST_ROWS) * n_vocab;
    int prev = -1;
    for (int i = 0; i < n; ++i) {
        const int col = cols[i];
        if (col >= n_vocab) break;   // sentinel padding: ascending, so nothing follows
        if (PASS == 0 && col == prev) atomicOr(repeated, 1);
        prev = col;
        if (dense_bit[col] != 255) continue;   // dense columns live in the row patterns, not in the postings
        if (PASS == 0) atomicAdd(block_seg + col, 1u);
        else post[atomicAdd(block_seg + col, 1u)] = (uint16_t)(pos % POST_ROWS);
    }
}

// one thread per row position: chunk count, row sum and dense pattern of the row that sits there
__global__ void k_row_prepare(const int64_t *__restrict__ row_ptr, int64_t n_rows, const int32_t *__restrict__ perm,
                              const uint32_t *__restrict__ pat_row, uint32_t *__restrict__ pat_pos, uint32_t *__restrict__ n_chunks,
                              const float *__restrict__ sums, float *__restrict__ sums_pos) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n_rows) return;
    const int64_t r = perm[pos];
    pat_pos[pos] = pat_row[r];
    n_chunks[pos] = (uint32_t)((row_ptr[r + 1] - row_ptr[r] + CHUNK_COLS - 1) / CHUNK_COLS);
    sums_pos[pos] = sums[r];
}

// one warp per row: rank-sort the row's column ids ascending into its sentinel padded chunks
__global__ void k_row_pack(const int64_t *__restrict__ row_ptr, const uint16_t *__restrict__ cols,
                           const uint32_t *__restrict__ chunk_ptr, int64_t n_rows, const int32_t *__restrict__ perm,
                           uint16_t sentinel, uint16_t *__restrict__ out) {
    int64_t pos = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int lane = threadIdx.x & 31;
    if (pos >= n_rows) return;
    const int64_t r = perm[pos];
    int64_t p0 = row_ptr[r];
    int g = (int)(row_ptr[r + 1] - p0);
    uint16_t *dst = out + (size_t)chunk_ptr[pos] * CHUNK_COLS;
    int padded = (int)(chunk_ptr[pos + 1] - chunk_ptr[pos]) * CHUNK_COLS;
    for (int i = lane; i < padded; i += 32) {
        if (i >= g) dst[i] = sentinel;
    }
    for (int i = lane; i < g; i += 32) {
        uint16_t x = min(cols[p0 + i], sentinel);
        int rank = 0;
        for (int j = 0; j < g; ++j) {
            uint16_t y = min(cols[p0 + j], sentinel);
            rank += (y < x) || (y == x && j < i);
        }
        dst[rank] = x;
    }
}

// smallest row sum and OR of the dense patterns of every group of 128 consecutive positions (one warp per group)
__global__ void k_sums_floor(const float *__restrict__ sums_pos, const uint32_t *__restrict__ pat_pos, int64_t n_rows, int64_t n_groups,
                             float *__restrict__ out, uint32_t *__restrict__ out_pat) {
    int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int lane = threadIdx.x & 31;
    if (group >= n_groups) return;
    float low = __int_as_float(0x7f800000);
    uint32_t pat = 0;
    for (int i = lane; i < 128; i += 32) {
        int64_t pos = group * 128 + i;
        if (pos < n_rows) {
            low = fminf(low, sums_pos[pos]);
            pat |= pat_pos[pos];
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) low = fminf(low, __shfl_xor_sync(0xffffffffu, low, d));
    pat = __reduce_or_sync(0xffffffffu, pat);
    if (lane == 0) {
        out[group] = low;
        out_pat[group] = pat;
    }
}

// document frequency of every column over the rows of this index (rows hold a column once: checked by k_post_build)
__global__ void k_col_df(const uint16_t *__restrict__ cols, int64_t nnz, int n_vocab, uint32_t *__restrict__ df) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz && cols[i] < n_vocab) atomicAdd(df + cols[i], 1u);
}

// Per original row: sums_matrix_truth = sequential f32 sum in the caller's column order (match_maker.py:172-174; taken
// as given when the caller provides it), its dense pattern, and the key its position inside the sort block is sorted
// by: (sort block, row sum) - the top field keeps every block of SORT_BLOCK rows in place.  Rows of (nearly) equal sum
// sit together, so the smallest sum of a group of 128 positions is close to every row's own and the group bars of
// k_post are nearly as tight as the row tests.  (Sorting by pattern first was measured and dropped: a block of 4,096
// rows holds ~1,300 distinct patterns, the runs are too short to help and they scatter the sums - 1,060 instead of
// 180 rows per query end up above their group's bar at N = 200k.)
__global__ void k_row_keys(const int64_t *__restrict__ row_ptr, const uint16_t *__restrict__ cols, const float *__restrict__ w32, int n_vocab,
                           int64_t n_rows, const uint8_t *__restrict__ dense_bit, int compute_sums, float *__restrict__ sums,
                           uint32_t *__restrict__ pat_row, unsigned long long *__restrict__ keys, int32_t *__restrict__ ids,
                           int *__restrict__ not_monotone) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int64_t p0 = row_ptr[r], p1 = row_ptr[r + 1];
    uint32_t pat = 0;
    float acc = 0.0f;
    for (int64_t p = p0; p < p1; ++p) {
        const int col = min((int)cols[p], n_vocab);
        const int bit = dense_bit[col];
        if (bit != 255) pat |= 1u << bit;
        acc = __fadd_rn(acc, w32[col]);
    }
    if (compute_sums) sums[r] = acc;
    else acc = sums[r];
    if (!(acc >= 0.0f)) atomicOr(not_monotone, 1);
    pat_row[r] = pat;
    // 2^-10 units, saturating at 2^21 (sums are a few hundred at most)
    const unsigned long long sum_key = (unsigned long long)(fminf(fmaxf(acc, 0.0f), 2097151.0f) * 1024.0f);
    keys[r] = ((unsigned long long)(r / SORT_BLOCK) << 42) | sum_key;
    ids[r] = (int32_t)r;
}

// Re-orders the postings of every long segment so that 32 consecutive entries fall into 32 different shared-memory
// banks of k_post's u16 accumulators (bank = (row / 2) & 31): entries are ranked by (index within their bank, bank).
// The rows of a segment are distinct and their additions commute, so the order inside a segment is free.  One warp
// per segment.
constexpr int BALANCE_WARPS = 2;
__device__ __forceinline__ int post_bank(int row) { return (row >> 1) & 31; }
__global__ void __launch_bounds__(BALANCE_WARPS * 32) k_post_balance(uint16_t *__restrict__ post, const post_off_t *__restrict__ seg_off,
                                                                     const uint32_t *__restrict__ seg_base, int64_t n_segments, int n_vocab) {
    __shared__ uint16_t s_rows[BALANCE_WARPS][POST_ROWS];
    __shared__ uint16_t s_rank[BALANCE_WARPS][POST_ROWS];
    __shared__ int s_count[BALANCE_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * BALANCE_WARPS + warp;
    if (seg >= n_segments) return;
    const int64_t s = seg / n_vocab;
    const int c = (int)(seg % n_vocab);
    const post_off_t *o = seg_off + s * (n_vocab + 1) + c;
    const int n = (int)(o[1] - o[0]);
    if (n <= 32 || n > POST_ROWS) return;
    uint16_t *entries = post + seg_base[s] + o[0];
    s_count[warp][lane] = 0;
    for (int i = lane; i < n; i += 32) s_rows[warp][i] = entries[i];
    __syncwarp();
    for (int i = lane; i < n; i += 32) s_rank[warp][i] = (uint16_t)atomicAdd(&s_count[warp][post_bank(s_rows[warp][i])], 1);
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const int row = s_rows[warp][i], bank = post_bank(row), k = s_rank[warp][i];
        int at = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int cnt = s_count[warp][b];
            at += min(cnt, k) + ((b < bank && cnt > k) ? 1 : 0);
        }
        entries[at] = (uint16_t)row;
    }
}

// global segment starts -> per block base + offsets inside the block
__global__ void k_post_offsets(const uint32_t *__restrict__ seg_start, int n_sub, int n_vocab, post_off_t *__restrict__ seg_off,
                               uint32_t *__restrict__ seg_base, int *__restrict__ too_long) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_sub * (n_vocab + 1)) return;
    const int s = (int)(i / (n_vocab + 1)), c = (int)(i % (n_vocab + 1));
    const uint32_t base = seg_start[(size_t)s * n_vocab];
    const uint32_t off = seg_start[(size_t)s * n_vocab + c] - base;   // c == n_vocab: the start of the next block
    if (off >= (1u << 24)) atomicOr(too_long, 1);   // piece descriptors keep starts in 24 bits
    seg_off[i] = (post_off_t)off;
    if (c == 0) seg_base[s] = base;
}

// ---------------------------------------------------------------------------------------------------
// query preparation: ascending column ids (match_maker.py:111-120) and mx (match_maker.py:197)
// ---------------------------------------------------------------------------------------------------
__global__ void k_query_prepare(const int64_t *__restrict__ q_ptr, const uint16_t *__restrict__ q_cols,
                                const double *__restrict__ w64, int n_vocab, const double *__restrict__ mx_in, int mx_mode,
                                int64_t n_q, uint16_t *__restrict__ sorted, double *__restrict__ mx_out) {
    int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int lane = threadIdx.x & 31;
    if (q >= n_q) return;
    int64_t p0 = q_ptr[q];
    int g = (int)(q_ptr[q + 1] - p0);
    for (int i = lane; i < g; i += 32) {
        uint16_t x = min(q_cols[p0 + i], (uint16_t)n_vocab);  // out-of-range ids become the weightless sentinel
        int rank = 0;
        for (int j = 0; j < g; ++j) {
            uint16_t y = min(q_cols[p0 + j], (uint16_t)n_vocab);
            rank += (y < x) || (y == x && j < i);
        }
        sorted[p0 + rank] = x;
    }
    __syncwarp();
    if (lane == 0) {
        if (mx_in != nullptr) {
            mx_out[q] = mx_in[q];
        } else {
            // CPython's builtin sum() over python floats (Python/bltinmodule.c): the first item is added to
            // int 0 exactly; the remaining ones go through Neumaier's compensated loop on >= 3.12.
            double s = 0.0, c = 0.0;
            for (int i = 0; i < g; ++i) {
                int col = sorted[p0 + i];
                double x = col < n_vocab ? w64[col] : 0.0;
                if (i == 0) {
                    s = x;
                    continue;
                }
                double t = __dadd_rn(s, x);
                if (mx_mode == DS_MX_PY312_COMPENSATED) {
                    if (fabs(s) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(s, t), x));
                    else c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, t), s));
                }
                s = t;
            }
            if (c != 0.0 && isfinite(c)) s = __dadd_rn(s, c);
            mx_out[q] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K1 scan kernel
// ---------------------------------------------------------------------------------------------------
struct ScanParams {
    const chunk_t *chunks;
    const uint32_t *chunk_ptr;
    const float *sums;          // by position
    const float *w32;
    int n_vocab;
    const uint16_t *q_sorted;   // call-level CSR, ascending per query
    const int64_t *q_ptr;
    const int32_t *batch_q;     // [n_batch] batch-local -> call-level query id
    const int32_t *tile_q;      // [n_tiles * TQ] batch-local query ids, -1 = empty slot
    const float2 *ab;           // [n_batch] filter constants
    int r0, r1, rows_per_cta;
    uint2 *cand;                // [n_batch * cap] (row position, f32 bits of sc)
    int *cand_count;            // [n_batch]
    int cap;
    float *dense;               // non-null: write every sc to dense[b * dense_stride + (row - r0)]
    int dense_stride;
};

__device__ __forceinline__ void accumulate_column(float (&sc)[TQ], uint32_t mask, float w) {
#pragma unroll
    for (int j = 0; j < TQ; ++j) {
        if (mask & (1u << j)) sc[j] = __fadd_rn(sc[j], w);
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) k_scan(ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int slot_bytes = ((p.n_vocab + 1) * 2 + 15) & ~15;
    uint16_t *slot16 = reinterpret_cast<uint16_t *>(smem);
    uint2 *entries = reinterpret_cast<uint2 *>(smem + slot_bytes);
    float2 *s_ab = reinterpret_cast<float2 *>(smem + slot_bytes + MAX_SLOTS * 8);
    int *s_qid = reinterpret_cast<int *>(smem + slot_bytes + MAX_SLOTS * 8 + TQ * 8);
    int *s_base = s_qid + TQ;

    const int tid = threadIdx.x;
    const int tile = blockIdx.y;
    const int cta_r0 = p.r0 + blockIdx.x * p.rows_per_cta;
    const int cta_r1 = min(p.r1, cta_r0 + p.rows_per_cta);

    // ---- build the tile table ----
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        const int n16 = (slot_bytes + MAX_SLOTS * 8) / 16;
        for (int i = tid; i < n16; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
        if (tid < TQ) {
            int b = p.tile_q[tile * TQ + tid];
            s_qid[tid] = b;
            int g = 0;
            float2 ab = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));  // +inf: never passes
            if (b >= 0) {
                int q = p.batch_q[b];
                g = (int)(p.q_ptr[q + 1] - p.q_ptr[q]);
                ab = p.ab[b];
            }
            s_ab[tid] = ab;
            int incl = g;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (tid >= d) incl += v;
            }
            s_base[tid] = incl - g;
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < TQ; j += THREADS / 32) {
        int b = s_qid[j];
        if (b < 0) continue;
        int q = p.batch_q[b];
        int64_t q0 = p.q_ptr[q];
        int g = (int)(p.q_ptr[q + 1] - q0);
        for (int i = lane; i < g; i += 32) slot16[p.q_sorted[q0 + i]] = (uint16_t)(s_base[j] + i + 1);
    }
    __syncthreads();
    for (int j = warp; j < TQ; j += THREADS / 32) {
        int b = s_qid[j];
        if (b < 0) continue;
        int q = p.batch_q[b];
        int64_t q0 = p.q_ptr[q];
        int g = (int)(p.q_ptr[q + 1] - q0);
        for (int i = lane; i < g; i += 32) {
            uint16_t col = p.q_sorted[q0 + i];
            int s = slot16[col];
            atomicOr(&entries[s].x, 1u << j);
            entries[s].y = __float_as_uint(p.w32[col]);
        }
    }
    __syncthreads();

    // ---- stream the truth rows: one row per thread, all TQ queries at once ----
    // Software pipelined: the next row's metadata and the next 16-byte chunk are requested before the
    // current chunk is consumed (only 16 warps fit an SM at 128 registers, so latency is hidden by ILP).
    int row = cta_r0 + tid;
    uint32_t c0 = 0, c1 = 0;
    float sum_t = 0.0f;
    chunk_t ch = zero_chunk();
    if (row < cta_r1) {
        c0 = p.chunk_ptr[row];
        c1 = p.chunk_ptr[row + 1];
        sum_t = p.sums[row];
        if (c0 < c1) ch = __ldg(p.chunks + c0);
    }
    while (row < cta_r1) {
        const int next_row = row + THREADS;
        uint32_t n0 = 0, n1 = 0;
        float next_sum = 0.0f;
        if (next_row < cta_r1) {
            n0 = p.chunk_ptr[next_row];
            n1 = p.chunk_ptr[next_row + 1];
            next_sum = p.sums[next_row];
        }
        float sc[TQ];
#pragma unroll
        for (int j = 0; j < TQ; ++j) sc[j] = 0.0f;
#pragma unroll 1
        for (uint32_t c = c0; c < c1; ++c) {
            chunk_t next_ch = zero_chunk();
            if (c + 1 < c1) next_ch = __ldg(p.chunks + c + 1);
            else if (n0 < n1) next_ch = __ldg(p.chunks + n0);
            const uint32_t words[CHUNK_COLS / 2] = {ch.x, ch.y};
            uint2 ent[CHUNK_COLS];
#pragma unroll
            for (int i = 0; i < CHUNK_COLS; ++i) {
                uint32_t col = (words[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                ent[i] = entries[slot16[col]];
            }
#pragma unroll
            for (int i = 0; i < CHUNK_COLS; ++i) accumulate_column(sc, ent[i].x, __uint_as_float(ent[i].y));
            ch = next_ch;
        }
        if (c0 == c1 && n0 < n1) ch = __ldg(p.chunks + n0);   // empty row: nothing was prefetched above
        if (p.dense != nullptr) {
#pragma unroll
            for (int j = 0; j < TQ; ++j) {
                int b = s_qid[j];
                if (b >= 0) p.dense[(size_t)b * p.dense_stride + (row - p.r0)] = sc[j];
            }
        } else {
            uint32_t pass = 0;
#pragma unroll
            for (int j = 0; j < TQ; ++j) {
                const float2 ab = s_ab[j];
                pass |= (sc[j] > fmaf(ab.x, sum_t, ab.y)) ? (1u << j) : 0u;
            }
            if (pass != 0) {   // rare: spill the accumulators once and emit the survivors
                float spilled[TQ];
#pragma unroll
                for (int j = 0; j < TQ; ++j) spilled[j] = sc[j];
                while (pass != 0) {
                    const int j = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const int b = s_qid[j];
                    const int pos = atomicAdd(p.cand_count + b, 1);
                    if (pos < p.cap) p.cand[(size_t)b * p.cap + pos] = make_uint2((uint32_t)row, __float_as_uint(spilled[j]));
                }
            }
        }
        row = next_row;
        c0 = n0;
        c1 = n1;
        sum_t = next_sum;
    }
}

// ---------------------------------------------------------------------------------------------------
// K1 posting kernel: the candidates of k_scan's filter pass (a superset of them, never fewer) from the inverted
// rows.  One warp task = one query x a run of consecutive blocks of POST_ROWS row positions.
//
// A query's columns are split in two.  DENSE columns (the <= 32 commonest trigrams of the index: 3/4 of all
// (row, column) incidences a query touches) are never walked: which of them a row holds is a 32-bit pattern, rows
// are sorted by pattern, and the weight a group of 128 rows can gain from them is the sum over (group's OR-ed
// pattern & query mask).  SPARSE columns are walked through their posting segments: every posting adds the
// column's weight to its row's accumulator in shared memory.  A row appears once per segment, so the lanes of a
// piece never collide; __syncwarp() orders consecutive pieces.  Postings are requested POST_DEPTH pieces ahead of
// their use, the next block's segment bounds one block ahead.
//
// Nothing here has to be exact, only conservative: every row that survives is re-scored by k_select in the
// reference's own float32 order (ascending column ids over the row's resident chunks, match_maker.py:33-47).
// So the accumulators are 16-bit fixed point (weights rounded UP to 1/S units, S chosen per query so that the
// largest possible sum fits 65,535): 8 KB per warp hold 4,096 rows, and the sweep that finds the rows above their
// bar and clears the block reads 8 rows per lane and 128-bit load with two packed-u16 maxima.
// `post_test` is the one conservative test (monotone in every argument); a group's bar is the largest accumulator
// value that fails it with the group's smallest row sum and largest dense gain.
// ---------------------------------------------------------------------------------------------------
struct PostParams {
    const uint16_t *post;
    const post_off_t *seg_off;
    const uint32_t *seg_base;
    int n_vocab;
    int64_t n_truth;
    const float *sums;          // by position
    const float *sums_floor;    // [blocks * POST_GROUPS] smallest row sum of every group of 128 positions
    const uint32_t *group_pat;  // [blocks * POST_GROUPS] OR of the dense patterns of the group
    const uint32_t *pat_pos;    // [n_truth] dense pattern by position
    const uint8_t *dense_bit;   // [n_vocab + 1]
    const float *dense_w;       // [32]
    const float *w32;
    const uint16_t *q_sorted;
    const int64_t *q_ptr;
    const int32_t *batch_q;
    const float2 *ab;
    int n_batch;
    int s0, s1;                 // posting blocks of this launch
    int run_len;                // blocks per task
    long long n_tasks;          // runs x n_batch, the query index fastest
    int tasks_per_cta;
    uint2 *cand;
    int *cand_count;
    int cap;
};

struct PostPiece {
    uint32_t r0, r1;   // block-local rows of this lane's postings (entries lane and lane + 32 of the piece)
    uint32_t w;        // fixed-point weight
    int n;             // postings in the piece (warp uniform); 0 = the stream has ended
};

// per query constants of the conservative test
struct PostQuery {
    float2 ab;         // k_scan's filter constants
    float inv_scale;   // >= 1 / S (0 when the query has no sparse column)
    float grow;        // 1 + e: covers every float32 rounding between the reference's sum and this bound
    uint32_t mask;     // the query's dense columns
};

// float32 sum of the dense weights selected by `bits`, lowest bit first: monotone in `bits` (a superset never sums lower)
__device__ __forceinline__ float dense_gain(uint32_t bits, const float *s_dense_w) {
    float v = 0.0f;
    while (bits != 0) {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        v = __fadd_rn(v, s_dense_w[bit]);
    }
    return v;
}

// Can a row with fixed-point sparse sum `acc`, dense gain `gain` and filter bar `bar` = fmaf(a, sums, b) reach the
// threshold?  acc * inv_scale >= the real sparse sum, gain >= the real dense sum up to float32 rounding, and `grow`
// covers those roundings and the reference's own: the reference's float32 score sum is <= the left-hand side, so
// k_scan's filter `sc > bar` implies this test.  Monotone: non-decreasing in acc and gain, non-increasing in bar.
__device__ __forceinline__ bool post_test(int acc, float gain, float bar, const PostQuery &q) {
    const float upper = __fmul_ru(__fadd_ru(__fmul_ru((float)acc, q.inv_scale), gain), q.grow);
    return upper > bar;
}

// what the rare paths of k_post need from the kernel parameters, kept in shared memory so that the hot loops do not
// re-load them from the constant bank around every (inlined or not) call
struct PostShared {
    int64_t n_truth;
    const uint32_t *pat_pos;
    const float *sums;
    int *cand_count;
    uint2 *cand;
    int cap;
    float dense_w[DENSE_MAX];
};

// full test for the listed rows of a swept block (kept out of line: the sweep loop stays small)
__device__ __noinline__ void post_flush(post_acc_t *acc, const uint16_t *list, int n_list, int base_block, float2 ab, float inv_scale,
                                        float grow, uint32_t mask, const PostShared *sh, int b, int lane) {
    PostQuery q;
    q.ab = ab;
    q.inv_scale = inv_scale;
    q.grow = grow;
    q.mask = mask;
    const int64_t base_pos = (int64_t)base_block * POST_ROWS;
    const int64_t n_truth = sh->n_truth;
    const uint32_t *__restrict__ pat_pos = sh->pat_pos;
    const float *__restrict__ sums = sh->sums;
    int *cand_count = sh->cand_count;
    uint2 *cand = sh->cand;
    const int cap = sh->cap;
    const float *s_dense_w = sh->dense_w;
    __syncwarp();
    for (int i = lane; i < n_list; i += 32) {
        const int r = list[i];
        const int value = acc[r];
        acc[r] = 0;
        const int64_t pos = base_pos + r;
        if (pos >= n_truth) continue;   // padding of the last block (listed only when a bar is negative)
        const float gain = dense_gain(__ldg(pat_pos + pos) & q.mask, s_dense_w);
        if (post_test(value, gain, fmaf(q.ab.x, __ldg(sums + pos), q.ab.y), q)) {
            const int at = atomicAdd(cand_count + b, 1);
            if (at < cap) cand[(size_t)b * cap + at] = make_uint2((uint32_t)pos, (uint32_t)value);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(POST_WARPS * 32, POST_CTAS) k_post(PostParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_next;
    __shared__ PostShared s_shared;
    const float *s_dense_w = s_shared.dense_w;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    post_acc_t *acc = reinterpret_cast<post_acc_t *>(smem) + (size_t)warp * POST_ROWS;
    uint4 *acc4 = reinterpret_cast<uint4 *>(acc);
    uint16_t *list = reinterpret_cast<uint16_t *>(smem + (size_t)POST_WARPS * POST_ROWS * sizeof(post_acc_t)) + warp * POST_LIST;
    uint2 *desc = reinterpret_cast<uint2 *>(smem + (size_t)POST_WARPS * (POST_ROWS * sizeof(post_acc_t) + POST_LIST * 2)) + warp * POST_DESC;
    if (threadIdx.x == 0) {
        s_next = 0;
        s_shared.n_truth = p.n_truth;
        s_shared.pat_pos = p.pat_pos;
        s_shared.sums = p.sums;
        s_shared.cand_count = p.cand_count;
        s_shared.cand = p.cand;
        s_shared.cap = p.cap;
    }
    if (threadIdx.x < DENSE_MAX) s_shared.dense_w[threadIdx.x] = p.dense_w[threadIdx.x];
    __syncthreads();
    const long long cta_first = (long long)blockIdx.x * p.tasks_per_cta;
    const int cta_tasks = (int)min((long long)p.tasks_per_cta, p.n_tasks - cta_first);
    const int stride = p.n_vocab + 1;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    constexpr int SWEEPS = POST_ROWS / 256;   // the sweep takes 256 rows (8 per lane) at a time
#pragma unroll
    for (int i = 0; i < SWEEPS; ++i) acc4[i * 32 + lane] = zero4;   // every sweep leaves the block zeroed again
    __syncwarp();

    for (;;) {
        int local = 0;
        if (lane == 0) local = atomicAdd(&s_next, 1);
        local = __shfl_sync(0xffffffffu, local, 0);
        if (local >= cta_tasks) break;
        const long long task = cta_first + local;
        const int run = (int)(task / p.n_batch);
        const int b = (int)(task % p.n_batch);
        PostQuery pq;
        pq.ab = p.ab[b];
        if (!(pq.ab.y < __int_as_float(0x7f800000))) continue;   // overflowed / dense-only query: nothing can pass
        const int s_begin = p.s0 + run * p.run_len;
        const int s_end = min(p.s1, s_begin + p.run_len);
        const int q = p.batch_q[b];
        const int64_t q0 = p.q_ptr[q];
        const int g = (int)(p.q_ptr[q + 1] - q0);

        // one pass over the query's columns: dense mask, the sum of the sparse weights (rounded up) -> the scale.
        // Up to 32 columns stay in registers for the whole run: lane = column.
        const bool cached = g <= 32;
        int col = p.n_vocab;
        float col_w = 0.0f;
        bool col_sparse = false;
        float sparse_sum = 0.0f;
        pq.mask = 0;
        for (int g0 = 0; g0 < g; g0 += 32) {
            int c = p.n_vocab;
            if (g0 + lane < g) c = p.q_sorted[q0 + g0 + lane];
            float w = 0.0f;
            bool sparse = false;
            if (c < p.n_vocab) {
                w = __ldg(p.w32 + c);
                const int bit = __ldg(p.dense_bit + c);
                if (bit != 255) pq.mask |= 1u << bit;
                else sparse = true;
            }
            if (sparse) sparse_sum = __fadd_ru(sparse_sum, w);
            if (cached) {
                col = c;
                col_w = w;
                col_sparse = sparse;
            }
        }
        pq.mask = __reduce_or_sync(0xffffffffu, pq.mask);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sparse_sum = __fadd_ru(sparse_sum, __shfl_xor_sync(0xffffffffu, sparse_sum, d));
        // S * (sum of the real sparse weights) + 2 g <= 65,531: no accumulator can overflow (a row holds a column once,
        // every weight is rounded up by less than two units)
        float scale = 0.0f;
        pq.inv_scale = 0.0f;
        if (sparse_sum > 0.0f && sparse_sum < __int_as_float(0x7f800000)) {
            scale = __fdiv_rd((float)(65535 - 4 - 2 * min(g, 16000)), sparse_sum);
            pq.inv_scale = __frcp_ru(scale);
        }
        pq.grow = __fadd_ru(1.0f, __fmul_ru((float)(2 * g + 64), 5.9604645e-8f));   // 1 + (2 g + 64) * 2^-24
        uint32_t col_fix = 0;
        post_off_t next_beg = 0, next_end = 0;
        if (col_sparse) {
            col_fix = min(65535u, __float2uint_ru(__fmul_ru(col_w, scale)));
            const post_off_t *o = p.seg_off + (size_t)s_begin * stride + col;
            next_beg = __ldg(o);
            next_end = __ldg(o + 1);
        }
        uint32_t next_base = __ldg(p.seg_base + s_begin);
        // lane i: the smallest row sum and the OR-ed dense pattern of the block's i-th group of 128 rows
        float next_floor = __ldg(p.sums_floor + (size_t)s_begin * POST_GROUPS + lane);
        uint32_t next_pat = __ldg(p.group_pat + (size_t)s_begin * POST_GROUPS + lane);
        int n_list = 0;

        for (int s = s_begin; s < s_end; ++s) {
            const uint32_t base = next_base;
            // This lane's group bar: the largest accumulator value that still fails the test with the group's smallest
            // row sum (fmaf is monotone in sums, a >= 0) and its largest dense gain (the OR of its rows' patterns).
            // A first guess from the real-number form, then corrected against post_test itself (monotone in acc).
            int my_bar;
            {
                const float bar_f = fmaf(pq.ab.x, next_floor, pq.ab.y);
                const float gain = dense_gain(next_pat & pq.mask, s_dense_w);
                // acc * inv_scale + gain > bar / grow, with 1 / grow ~ 2 - grow; NaN (padded group: 0 * inf) compares false
                const float room = __fmul_rd(__fsub_rd(__fmul_rd(bar_f, __fsub_rd(2.0f, pq.grow)), gain), scale);
                int guess = room >= 65534.0f ? 65534 : (room >= 1.0f ? (int)room - 1 : -1);
                if (!(bar_f == bar_f)) guess = 65535;
                if (pq.inv_scale == 0.0f) guess = post_test(0, gain, bar_f, pq) ? -1 : 65535;   // no sparse column: acc stays 0
                if (guess < 0 && !post_test(0, gain, bar_f, pq)) guess = 0;
                if (guess >= 0 && guess < 65535) {
                    for (int it = 0; it < 64 && guess < 65535 && !post_test(guess + 1, gain, bar_f, pq); ++it) ++guess;
                    for (int it = 0; it < 64 && guess >= 0 && post_test(guess, gain, bar_f, pq); ++it) --guess;
                    if (guess >= 0 && post_test(guess, gain, bar_f, pq)) guess = -1;   // still passing: list the whole group
                }
                my_bar = guess;
            }
            if (s + 1 < s_end) {
                next_base = __ldg(p.seg_base + s + 1);
                next_floor = __ldg(p.sums_floor + (size_t)(s + 1) * POST_GROUPS + lane);
                next_pat = __ldg(p.group_pat + (size_t)(s + 1) * POST_GROUPS + lane);
            }
            for (int g0 = 0; g0 < g; g0 += 32) {
                uint32_t my_beg = 0, my_end = 0, my_w = 0;
                if (cached) {
                    my_beg = next_beg;
                    my_end = next_end;
                    my_w = col_fix;
                    if (s + 1 < s_end && col_sparse) {
                        const post_off_t *o = p.seg_off + (size_t)(s + 1) * stride + col;
                        next_beg = __ldg(o);
                        next_end = __ldg(o + 1);
                    }
                } else {
                    int c = p.n_vocab;
                    if (g0 + lane < g) c = p.q_sorted[q0 + g0 + lane];
                    if (c < p.n_vocab && __ldg(p.dense_bit + c) == 255) {
                        const post_off_t *o = p.seg_off + (size_t)s * stride + c;
                        my_beg = __ldg(o);
                        my_end = __ldg(o + 1);
                        my_w = min(65535u, __float2uint_ru(__fmul_ru(__ldg(p.w32 + c), scale)));
                    }
                }
                // The non-empty segments are cut into pieces of <= 64 postings, listed in shared memory (lane = column:
                // a prefix sum of the piece counts gives every lane its slots); the walk then is a plain counted loop
                // over the list.  Lists longer than POST_DESC pieces go in rounds.
                const uint32_t n_mine = my_end > my_beg ? my_end - my_beg : 0u;
                unsigned remaining = __ballot_sync(0xffffffffu, n_mine != 0);
                const uint16_t *lane_post = p.post + base + lane;   // piece starts are kept relative to the block
                uint32_t done = 0;                                   // postings of this lane's segment already listed
                while (remaining != 0) {
                    const bool mine = (remaining >> lane) & 1u;
                    const uint32_t left_mine = n_mine - done;
                    const uint32_t pieces = mine ? (left_mine + 63u) >> 6 : 0u;
                    uint32_t incl = pieces;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    // lanes whose pieces all fit go in this round; the first lane that does not fit lists what fits of
                    // its segment (a segment of a 4,096-row block can alone hold more than POST_DESC pieces)
                    const uint32_t before = incl - pieces;
                    const bool fits = mine && incl <= (uint32_t)POST_DESC;
                    const bool partial = mine && !fits && before < (uint32_t)POST_DESC;
                    const uint32_t take = fits ? pieces : (partial ? (uint32_t)POST_DESC - before : 0u);
                    const unsigned fit_mask = __ballot_sync(0xffffffffu, fits);
                    const unsigned part_mask = __ballot_sync(0xffffffffu, partial);
                    const int last_lane = 31 - __clz(fit_mask | part_mask);   // the masks are not both empty: the first remaining lane takes >= 1
                    const int n_pieces = (int)__shfl_sync(0xffffffffu, before + take, last_lane);
                    if (take != 0) {
                        uint32_t at = before, start = my_beg + done, left = left_mine;
                        for (uint32_t t = 0; t < take; ++t) {
                            desc[at] = make_uint2(start | (min(left, 64u) << 24), my_w);   // start < 2^24: checked at build time
                            ++at;
                            start += 64u;
                            left -= min(left, 64u);
                        }
                        if (partial) done += take * 64u;
                    }
                    remaining &= ~fit_mask;
                    __syncwarp();

                    // ring of POST_DEPTH pieces in flight: read the accumulators of the oldest piece, request the piece
                    // POST_DEPTH ahead (its address arithmetic covers the shared-memory latency), then add and store
                    auto fetch = [&](int index, PostPiece &piece) {
                        const uint2 d = desc[index];
                        const int n = (int)(d.x >> 24);
                        piece.n = n;
                        piece.w = d.y;
                        const uint16_t *src = lane_post + (d.x & 0xffffffu);
                        if (lane < n) piece.r0 = __ldg(src);
                        if (lane + 32 < n) piece.r1 = __ldg(src + 32);
                    };
                    PostPiece ring[POST_DEPTH];
#pragma unroll
                    for (int d = 0; d < POST_DEPTH; ++d) {
                        ring[d].r0 = 0;
                        ring[d].r1 = 0;
                        ring[d].n = 0;
                        ring[d].w = 0;
                        if (d < n_pieces) fetch(d, ring[d]);
                    }
                    for (int first_piece = 0; first_piece < n_pieces; first_piece += POST_DEPTH) {
#pragma unroll
                        for (int d = 0; d < POST_DEPTH; ++d) {
                            const int index = first_piece + d;
                            if (index >= n_pieces) break;
                            const int n = ring[d].n;
                            const uint32_t add = ring[d].w;
                            post_acc_t *slot0 = acc + ring[d].r0, *slot1 = acc + ring[d].r1;
                            const bool first = lane < n, second = lane + 32 < n;
                            uint32_t a0 = 0, a1 = 0;
                            if (first) a0 = *slot0;      // predicated: idle lanes would only add bank conflicts
                            if (second) a1 = *slot1;
                            if (index + POST_DEPTH < n_pieces) fetch(index + POST_DEPTH, ring[d]);
                            if (first) *slot0 = (post_acc_t)(a0 + add);
                            if (second) *slot1 = (post_acc_t)(a1 + add);
                            __syncwarp();   // the next piece may belong to another column and touch the same rows
                        }
                    }
                    __syncwarp();   // the list is rewritten by the next round
                }
            }

            // sweep and zero the block, 256 rows (8 per lane, one 128-bit load) at a time; rows above their group's bar
            // wait in `list` (their accumulators stay) for the full test
#pragma unroll 4
            for (int i = 0; i < SWEEPS; ++i) {
                const int idx = i * 32 + lane;
                const int bar = __shfl_sync(0xffffffffu, my_bar, 2 * i + (lane >> 4));   // lanes 0-15 / 16-31: two groups
                const uint4 v = acc4[idx];
                unsigned m = __vimax3_u16x2(v.x, v.y, v.z);
                m = __vmaxu2(m, v.w);
                const bool hit = (int)max(m & 0xffffu, m >> 16) > bar;
                if (__ballot_sync(0xffffffffu, hit) == 0) {
                    acc4[idx] = zero4;
                    continue;
                }
                // some row of these 256 is above its bar: every lane marks which of its 8 values are (they stay in the
                // block for the full test, the others are cleared) and the lanes with marks take list slots in lane order
                const uint32_t words[4] = {v.x, v.y, v.z, v.w};
                uint32_t kept[4];
                unsigned marks = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t lo = words[c] & 0xffffu, hi = words[c] >> 16;
                    const bool lo_above = (int)lo > bar, hi_above = (int)hi > bar;
                    kept[c] = (lo_above ? lo : 0u) | (hi_above ? (hi << 16) : 0u);
                    marks |= (lo_above ? 1u : 0u) << (2 * c) | (hi_above ? 1u : 0u) << (2 * c + 1);
                }
                acc4[idx] = make_uint4(kept[0], kept[1], kept[2], kept[3]);
                const int mine = __popc(marks);
                int incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += up;
                }
                const int before = incl - mine;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                // the list holds POST_LIST rows; 256 candidates at once only happen with a negative bar
                for (int first = 0; first < total; first += POST_LIST) {
                    if (n_list > 0 && n_list + min(total - first, POST_LIST) > POST_LIST) {
                        post_flush(acc, list, n_list, s, pq.ab, pq.inv_scale, pq.grow, pq.mask, &s_shared, b, lane);
                        n_list = 0;
                    }
                    unsigned left = marks;
                    for (int j = before; left != 0; ++j) {
                        const int c = __ffs(left) - 1;
                        left &= left - 1;
                        if (j >= first && j < first + POST_LIST) list[n_list + j - first] = (uint16_t)(idx * 8 + c);
                    }
                    n_list += min(total - first, POST_LIST);
                    if (first + POST_LIST < total) {
                        post_flush(acc, list, n_list, s, pq.ab, pq.inv_scale, pq.grow, pq.mask, &s_shared, b, lane);
                        n_list = 0;
                    }
                }
            }
            if (n_list > 0) {
                post_flush(acc, list, n_list, s, pq.ab, pq.inv_scale, pq.grow, pq.mask, &s_shared, b, lane);
                n_list = 0;
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// selection: merge this launch's candidates into the per-query retained list (best m by score) and
// refresh the threshold / filter constants.  One warp per batch query.
// ---------------------------------------------------------------------------------------------------
struct SelectParams {
    int mode;                 // MODE_SCORE / MODE_ROW
    const int32_t *batch_q;
    const double *q_mx;
    const double *threshold;  // MODE_ROW: call-level fixed thresholds
    const float *sums;        // by position
    const int32_t *perm;      // position -> shard-local original row
    int n_batch;
    int k, m;                 // m = retained list length (MODE_ROW: m == k)
    const uint2 *cand;
    int *cand_count;
    int cap;
    const float *dense;
    int dense_stride, dense_rows, dense_r0;
    double *ret_score;  // [n_batch * m]
    int32_t *ret_row;   // [n_batch * m] original rows
    int *ret_n;
    double *theta;      // [n_batch] current exact threshold (<= 0: none yet)
    float2 *ab;
    int *state;
    int *overflow_count;
    int max_items;      // smem capacity per warp
    // Thresholds shared between the shards of one truth DB (MODE_SCORE): every shard publishes, per query, the best lower
    // bound of the GLOBAL k-th best score it knows (its own k-th best so far, or a peer's) and reads its peers' over
    // NVLink peer memory.  The k-th best of any subset of the rows is a valid bound, stale values only prune less.
    double *theta_own;                       // [n_q] call level, nullable
    const double *theta_peers[DS_MAX_PEERS]; // the same array of the other shards (peer-mapped device pointers)
    int n_peers;
    // candidates of k_post carry no score: it is computed here, in the reference's own float32 order
    int rescore;
    const chunk_t *chunks;
    const uint32_t *chunk_ptr;
    const float *w32;
    int n_vocab;
    const uint16_t *q_sorted;
    const int64_t *q_ptr;
};

// fast_jaccard's float32 sum for one (query, row) pair (match_maker.py:33-47): idf32 of the shared columns added in
// ascending column id order.  The row's chunks are ascending, so walking them in order and looking every column up
// in the query gives exactly that order.  Two forms of the lookup: a per-warp hash table of the query's columns in
// shared memory (queries of <= SELECT_HASH_MAX columns: one or two probes per row column, no dependent global loads)
// and a merge against the query's ascending column list in global memory.
constexpr int SELECT_HASH = 256;       // slots per warp
constexpr int SELECT_HASH_MAX = 128;   // most columns a hashed query may have (load factor 1/2)
constexpr uint32_t SELECT_EMPTY = 0xffffffffu;
__device__ __forceinline__ uint32_t select_hash(uint32_t col) { return (col * 0x9E3779B1u) >> 24; }

__device__ __forceinline__ float exact_intersection_hashed(const uint32_t *h_key, const float *h_w, const chunk_t *__restrict__ chunks,
                                                           uint32_t c0, uint32_t c1, int n_vocab) {
    float sc = 0.0f;
    chunk_t next = c0 < c1 ? __ldg(chunks + c0) : zero_chunk();
    for (uint32_t c = c0; c < c1; ++c) {
        const chunk_t ch = next;
        if (c + 1 < c1) next = __ldg(chunks + c + 1);
        const uint32_t words[CHUNK_COLS / 2] = {ch.x, ch.y};
#pragma unroll
        for (int k = 0; k < CHUNK_COLS; ++k) {
            const uint32_t rc = (words[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
            if ((int)rc >= n_vocab) break;   // sentinel padding
            uint32_t h = select_hash(rc);
            for (;;) {
                const uint32_t key = h_key[h];
                if (key == rc) {
                    sc = __fadd_rn(sc, h_w[h]);
                    break;
                }
                if (key == SELECT_EMPTY) break;
                h = (h + 1) & (SELECT_HASH - 1);
            }
        }
    }
    return sc;
}

__device__ __forceinline__ float exact_intersection(const uint16_t *__restrict__ q_cols, int g, const chunk_t *__restrict__ chunks, uint32_t c0,
                                                    uint32_t c1, const float *__restrict__ w32, int n_vocab) {
    float sc = 0.0f;
    int i = 0;
    for (uint32_t c = c0; c < c1 && i < g; ++c) {
        const chunk_t ch = __ldg(chunks + c);
        const uint32_t words[CHUNK_COLS / 2] = {ch.x, ch.y};
#pragma unroll
        for (int k = 0; k < CHUNK_COLS; ++k) {
            const int rc = (int)((words[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
            if (rc >= n_vocab) break;   // sentinel padding
            while (i < g && (int)q_cols[i] < rc) ++i;
            if (i < g && (int)q_cols[i] == rc) sc = __fadd_rn(sc, __ldg(w32 + rc));
        }
    }
    return sc;
}

__global__ void __launch_bounds__(128) k_select(SelectParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= p.n_batch) return;
    double *s_score = reinterpret_cast<double *>(smem) + (size_t)warp * p.max_items;
    int32_t *s_row = reinterpret_cast<int32_t *>(smem + (size_t)4 * p.max_items * 8) + (size_t)warp * p.max_items;
    uint32_t *h_key = reinterpret_cast<uint32_t *>(smem + (size_t)4 * p.max_items * 12) + (size_t)warp * SELECT_HASH;   // rescore only
    float *h_w = reinterpret_cast<float *>(smem + (size_t)4 * p.max_items * 12 + (size_t)4 * SELECT_HASH * 4) + (size_t)warp * SELECT_HASH;
    __shared__ double s_kth[4];

    if (p.state[b] & STATE_OVERFLOW) return;
    const bool by_row = p.mode == MODE_ROW;
    const int64_t q = p.batch_q[b];
    const double mx = p.q_mx[q];
    double theta = by_row ? p.threshold[q] : p.theta[b];
    if (!by_row && p.n_peers > 0) {
        double ext = 0.0;
        if (lane < p.n_peers) ext = *reinterpret_cast<const volatile double *>(p.theta_peers[lane] + q);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ext = fmax(ext, __shfl_xor_sync(0xffffffffu, ext, d));
        if (ext > theta) {   // a peer knows a better bound: prune with it from now on
            theta = ext;
            if (lane == 0) {
                p.theta[b] = ext;
                p.ab[b] = filter_from_threshold(ext, mx);
                p.theta_own[q] = ext;
            }
        }
    }
    int n = 0;
    if (p.dense != nullptr) {
        for (int i0 = 0; i0 < p.dense_rows; i0 += 32) {
            int i = i0 + lane;
            bool pass = false;
            double s = 0.0;
            int pos = p.dense_r0 + i;
            if (i < p.dense_rows) {
                float sc = p.dense[(size_t)b * p.dense_stride + i];
                if (by_row) {
                    // every row is a candidate: zero scores qualify for thr <= 0 and 0/0 = NaN never does,
                    // exactly like numpy's `array >= threshold` (match_maker.py:71)
                    s = exact_score(sc, p.sums[pos], mx);
                    pass = s >= theta;
                } else if (sc > 0.0f) {
                    s = exact_score(sc, p.sums[pos], mx);
                    pass = s > 0.0 && s >= theta;
                }
            }
            unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                int at = n + __popc(ballot & ((1u << lane) - 1));
                s_score[at] = s;
                s_row[at] = p.perm[pos];
            }
            n += __popc(ballot);
        }
    } else {
        int cnt = p.cand_count[b];
        if (cnt == 0) return;
        if (cnt > p.cap) {
            if (lane == 0) {
                p.state[b] |= STATE_OVERFLOW;
                p.ab[b] = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
                p.cand_count[b] = 0;
                atomicAdd(p.overflow_count, 1);
            }
            return;
        }
        const int64_t q0 = p.q_ptr[q];
        const int g = (int)(p.q_ptr[q + 1] - q0);
        const bool hashed = p.rescore && g <= SELECT_HASH_MAX;
        if (hashed) {   // the query's columns -> (column, idf32) hash table
            for (int i = lane; i < SELECT_HASH; i += 32) h_key[i] = SELECT_EMPTY;
            __syncwarp();
            for (int i = lane; i < g; i += 32) {
                const uint32_t col = p.q_sorted[q0 + i];
                if ((int)col >= p.n_vocab) continue;
                uint32_t h = select_hash(col);
                for (;;) {
                    const uint32_t old = atomicCAS(h_key + h, SELECT_EMPTY, col);
                    if (old == SELECT_EMPTY || old == col) {
                        h_w[h] = __ldg(p.w32 + col);
                        break;
                    }
                    h = (h + 1) & (SELECT_HASH - 1);
                }
            }
            __syncwarp();
        }
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            int i = i0 + lane;
            bool pass = false;
            double s = 0.0;
            int row = 0;
            if (i < cnt) {
                uint2 c = p.cand[(size_t)b * p.cap + i];
                float sc = __uint_as_float(c.y);
                if (hashed) sc = exact_intersection_hashed(h_key, h_w, p.chunks, p.chunk_ptr[c.x], p.chunk_ptr[c.x + 1], p.n_vocab);
                else if (p.rescore) sc = exact_intersection(p.q_sorted + q0, g, p.chunks, p.chunk_ptr[c.x], p.chunk_ptr[c.x + 1], p.w32, p.n_vocab);
                s = exact_score(sc, p.sums[c.x], mx);
                row = p.perm[c.x];
                pass = s > 0.0 && s >= theta;
            }
            unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                int at = n + __popc(ballot & ((1u << lane) - 1));
                s_score[at] = s;
                s_row[at] = row;
            }
            n += __popc(ballot);
        }
    }
    if (n == 0) {
        if (lane == 0 && p.dense == nullptr) p.cand_count[b] = 0;
        return;
    }
    const int n_old = p.ret_n[b];
    // Candidates [0, n) are padded to a power of two with entries that lose every comparison and sorted in place
    // (bitonic, best first in this mode's order); the old list, already sorted, sits behind them.
    int padded = 1;
    while (padded < n) padded <<= 1;
    for (int i = n + lane; i < padded; i += 32) {
        s_score[i] = -INFINITY;
        s_row[i] = -1;
    }
    double *o_score = s_score + padded;
    int32_t *o_row = s_row + padded;
    for (int i = lane; i < n_old; i += 32) {
        o_score[i] = p.ret_score[(size_t)b * p.m + i];
        o_row[i] = p.ret_row[(size_t)b * p.m + i];
    }
    __syncwarp();
    for (int span = 2; span <= padded; span <<= 1) {
        for (int j = span >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (padded >> 1); t += 32) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // bit j clear
                const int hi = lo | j;
                const double s_lo = s_score[lo], s_hi = s_score[hi];
                const int r_lo = s_row[lo], r_hi = s_row[hi];
                const bool hi_ahead = by_row ? (r_hi > r_lo) : better(s_hi, r_hi, s_lo, r_lo);
                const bool best_first = (lo & span) == 0;
                if (hi_ahead == best_first) {
                    s_score[lo] = s_hi;
                    s_row[lo] = r_hi;
                    s_score[hi] = s_lo;
                    s_row[hi] = r_lo;
                }
            }
            __syncwarp();
        }
    }
    // Final rank in the union of two sorted lists: own position plus a binary search into the other list.
    const int total = n + n_old;
    for (int i = lane; i < total; i += 32) {
        const bool is_new = i < n;
        const int own = is_new ? i : i - n;
        const double si = is_new ? s_score[own] : o_score[own];
        const int ri = is_new ? s_row[own] : o_row[own];
        const double *other_score = is_new ? o_score : s_score;
        const int32_t *other_row = is_new ? o_row : s_row;
        int lo = 0, hi = is_new ? n_old : n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const bool ahead = by_row ? (other_row[mid] > ri) : better(other_score[mid], other_row[mid], si, ri);
            if (ahead) lo = mid + 1;
            else hi = mid;
        }
        const int rank = own + lo;
        if (rank < p.m) {
            p.ret_score[(size_t)b * p.m + rank] = si;
            p.ret_row[(size_t)b * p.m + rank] = ri;
        }
        if (rank == p.k - 1) s_kth[warp] = si;
    }
    __syncwarp();
    if (lane == 0) {
        p.ret_n[b] = min(total, p.m);
        if (p.dense == nullptr) p.cand_count[b] = 0;
        if (!by_row && total >= p.k) {
            double th = threshold_from_key(__double2float_rn(s_kth[warp]));
            if (!(th > theta)) th = theta;   // never below a bound already in use (a peer's)
            p.theta[b] = th;
            p.ab[b] = filter_from_threshold(th, mx);
            if (p.theta_own != nullptr) p.theta_own[q] = th;
        }
    }
}

// MODE_ROW set-up: filter constants from the fixed thresholds; queries whose threshold is <= 0 (every
// non-NaN row qualifies) cannot use the positive-score filter and are marked for the dense pass
__global__ void k_rescan_init(const int32_t *__restrict__ batch_q, int n_batch, const double *__restrict__ threshold,
                              const double *__restrict__ q_mx, float2 *__restrict__ ab, int *__restrict__ state,
                              int *__restrict__ overflow_count) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    const int64_t q = batch_q[b];
    const double thr = threshold[q];
    if (thr > 0.0) {
        ab[b] = filter_from_threshold(thr, q_mx[q]);
        state[b] = 0;
    } else {
        ab[b] = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
        state[b] = STATE_OVERFLOW;
        atomicAdd(overflow_count, 1);
    }
}

// MODE_ROW result: retained rows (descending) -> call-level out_rows / out_count as global rows
__global__ void k_export_rows(const int32_t *__restrict__ batch_q, int n_batch, int k, const int32_t *__restrict__ ret_row,
                              const int *__restrict__ ret_n, int64_t row_offset, int64_t *__restrict__ out_rows,
                              int32_t *__restrict__ out_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_batch * k) return;
    int b = (int)(i / k), slot = (int)(i % k);
    int64_t q = batch_q[b];
    out_rows[q * k + slot] = slot < ret_n[b] ? (int64_t)ret_row[i] + row_offset : -1;
    if (slot == 0) out_count[q] = ret_n[b];
}

// copy the retained lists to the phase-1 output layout: [n_q, m] score desc / global row, -1 padded
__global__ void k_export(const int32_t *__restrict__ batch_q, int n_batch, int m, const double *__restrict__ ret_score,
                         const int32_t *__restrict__ ret_row, const int *__restrict__ ret_n, int64_t row_offset,
                         double *__restrict__ out_score, int64_t *__restrict__ out_row) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_batch * m) return;
    int b = (int)(i / m), slot = (int)(i % m);
    int64_t q = batch_q[b];
    bool valid = slot < ret_n[b];
    out_score[q * m + slot] = valid ? ret_score[i] : -1.0;
    out_row[q * m + slot] = valid ? (int64_t)ret_row[i] + row_offset : -1;
}

// ---------------------------------------------------------------------------------------------------
// merge (phase 2): one CTA per query over the gathered [n_shards, n_q, m] candidates
// ---------------------------------------------------------------------------------------------------
struct MergeParams {
    int n_shards;
    int64_t n_q;
    int k, m;
    int64_t n_total;
    const double *all_score;
    const int64_t *all_row;
    const double *q_mx;  // nullable
    int64_t *out_rows;
    int32_t *out_count;
    float *out_kth;
    double *out_threshold;
    int32_t *out_flags;
    int *flagged_count;
};

// number of entries of the (score desc, row desc) sorted list [0, n) that are strictly better than (sc, rw)
__device__ __forceinline__ int count_better(const double *__restrict__ score, const int64_t *__restrict__ row, int n, double sc,
                                            int64_t rw) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (better(score[mid], row[mid], sc, rw)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// One warp per query.  Every shard's list is sorted by (score desc, row desc) and shard s holds the rows of
// the s-th contiguous ascending range, so: the global k-th best lies among the first k entries of each list and
// its rank is a sum of binary searches; the qualifiers of a list are a prefix; and the k highest qualifying
// rows are taken shard by shard from the highest range down.
__global__ void __launch_bounds__(128) k_merge(MergeParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + warp;
    if (q >= p.n_q) return;
    __shared__ int s_count[4][128];      // valid entries per shard (n_shards <= 128)
    __shared__ int s_qual[4][128];       // qualifying entries per shard
    __shared__ double s_kth[4];
    int *n_valid = s_count[warp], *n_qual = s_qual[warp];
    const size_t list_stride = (size_t)p.n_q * p.m;
    const double *score_q = p.all_score + (size_t)q * p.m;
    const int64_t *row_q = p.all_row + (size_t)q * p.m;

    int positives = 0;
    for (int s = lane; s < p.n_shards; s += 32) {
        const int64_t *rows = row_q + s * list_stride;
        int lo = 0, hi = p.m;                       // first unused slot (row < 0): slots are front packed
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rows[mid] >= 0) lo = mid + 1;
            else hi = mid;
        }
        n_valid[s] = lo;
        positives += lo;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) positives += __shfl_xor_sync(0xffffffffu, positives, d);
    __syncwarp();

    if (lane == 0) s_kth[warp] = 0.0;
    __syncwarp();
    if (positives >= p.k) {
        // rank of every head element (slot < k of its list) in the union; exactly one has rank k - 1
        const int heads = p.n_shards * p.k;
        for (int e = lane; e < heads; e += 32) {
            const int s = e / p.k, i = e % p.k;
            if (i >= n_valid[s]) continue;
            const double sc = score_q[s * list_stride + i];
            const int64_t rw = row_q[s * list_stride + i];
            int rank = i;
            for (int t = 0; t < p.n_shards && rank < p.k; ++t) {
                if (t == s) continue;
                rank += count_better(score_q + t * list_stride, row_q + t * list_stride, n_valid[t], sc, rw);
            }
            if (rank == p.k - 1) s_kth[warp] = sc;
        }
    }
    __syncwarp();
    const float kth_key = (positives >= p.k) ? __double2float_rn(s_kth[warp]) : 0.0f;
    const double thr = threshold_from_key(kth_key);
    int64_t *rows_out = p.out_rows + q * p.k;
    int flags = 0;
    if (!(thr > 0.0)) {
        // every row with a non-NaN score qualifies (scores are >= 0): the last k rows of the DB, unless the
        // query has mx == 0, where rows with sums == 0 score 0/0 = NaN and must be skipped by the rescan.
        flags = DS_FLAG_FEW_POSITIVE;
        if ((p.q_mx != nullptr) && (p.q_mx[q] == 0.0)) flags |= DS_FLAG_RESCAN;
        const int64_t count = min((int64_t)p.k, p.n_total);
        for (int i = lane; i < p.k; i += 32) rows_out[i] = (i < count) ? (p.n_total - 1 - i) : -1;
        if (lane == 0) p.out_count[q] = (int32_t)count;
    } else {
        int incomplete = 0, total_qual = 0;
        for (int s = lane; s < p.n_shards; s += 32) {
            const double *scores = score_q + s * list_stride;
            int lo = 0, hi = n_valid[s];            // qualifiers = prefix with score >= thr
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (scores[mid] >= thr) lo = mid + 1;
                else hi = mid;
            }
            n_qual[s] = lo;
            total_qual += lo;
            // a full list whose lowest entry still qualifies may have dropped further qualifiers
            incomplete |= (n_valid[s] == p.m && lo == p.m);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            total_qual += __shfl_xor_sync(0xffffffffu, total_qual, d);
            incomplete |= __shfl_xor_sync(0xffffffffu, incomplete, d);
        }
        __syncwarp();
        for (int i = lane; i < p.k; i += 32) rows_out[i] = -1;
        __syncwarp();
        if (incomplete) {
            flags = DS_FLAG_RESCAN;
            if (lane == 0) p.out_count[q] = 0;
        } else {
            int filled = 0;
            for (int s = p.n_shards - 1; s >= 0 && filled < p.k; --s) {
                const int c = n_qual[s];
                const int64_t *rows = row_q + s * list_stride;
                const int room = p.k - filled;
                for (int i = lane; i < c; i += 32) {
                    const int64_t ri = rows[i];
                    int rank = 0;
                    for (int j = 0; j < c; ++j) rank += rows[j] > ri;
                    if (rank < room) rows_out[filled + rank] = ri;
                }
                filled += min(c, room);
            }
            if (lane == 0) p.out_count[q] = min(total_qual, p.k);
        }
    }
    if (lane == 0) {
        if (p.out_kth) p.out_kth[q] = kth_key;
        if (p.out_threshold) p.out_threshold[q] = thr;
        if (p.out_flags) p.out_flags[q] = flags;
        if ((flags & DS_FLAG_RESCAN) && p.flagged_count) atomicAdd(p.flagged_count, 1);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct QuerySet {
    int64_t n_q = 0;
    std::vector<int64_t> h_ptr;
    const int64_t *d_ptr = nullptr;
    uint16_t *d_sorted = nullptr;
    double *d_mx = nullptr;
};

static size_t scan_smem_bytes(int n_vocab) {
    size_t slot_bytes = (((size_t)n_vocab + 1) * 2 + 15) & ~(size_t)15;
    return slot_bytes + MAX_SLOTS * 8 + TQ * 8 + TQ * 4 * 2;
}

static int prepare_queries(Workspace &ws, const Index &ix, int64_t n_q, const int64_t *q_row_ptr,
                           const uint16_t *q_col_ids, const double *q_mx, int32_t mx_mode, QuerySet *qs) {
    cudaStream_t stream = ws.stream();
    qs->n_q = n_q;
    qs->h_ptr.resize((size_t)n_q + 1);
    if (is_device_pointer(q_row_ptr)) {
        DS_CUDA(cudaMemcpyAsync(qs->h_ptr.data(), q_row_ptr, (size_t)(n_q + 1) * 8, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
    } else {
        memcpy(qs->h_ptr.data(), q_row_ptr, (size_t)(n_q + 1) * 8);
    }
    const int64_t nnz = qs->h_ptr[n_q];
    if (qs->h_ptr[0] != 0 || nnz < 0) return fail(DS_ERR_BAD_ARG, "q_row_ptr must start at 0 and be non-decreasing");
    for (int64_t q = 0; q < n_q; ++q) {
        int64_t g = qs->h_ptr[q + 1] - qs->h_ptr[q];
        if (g < 0) return fail(DS_ERR_BAD_ARG, "q_row_ptr is decreasing at query %lld", (long long)q);
        if (g > MAX_SLOTS - 1)
            return fail(DS_ERR_UNSUPPORTED, "query %lld has %lld columns (max %d)", (long long)q, (long long)g, MAX_SLOTS - 1);
    }
    DS_CHECK(ws.stage_in(&qs->d_ptr, q_row_ptr, (size_t)n_q + 1));
    const uint16_t *d_cols = nullptr;
    DS_CHECK(ws.stage_in(&d_cols, q_col_ids, (size_t)nnz));
    const double *d_mx_in = nullptr;
    DS_CHECK(ws.stage_in(&d_mx_in, q_mx, (size_t)n_q));
    DS_CHECK(ws.alloc(&qs->d_sorted, (size_t)nnz));
    DS_CHECK(ws.alloc(&qs->d_mx, (size_t)n_q));
    if (n_q > 0) {
        int64_t threads = n_q * 32;
        k_query_prepare<<<(unsigned)ceil_div(threads, 256), 256, 0, stream>>>(qs->d_ptr, d_cols, ix.w64, ix.n_vocab, d_mx_in, mx_mode, n_q,
                                                                              qs->d_sorted, qs->d_mx);
        DS_LAUNCHED("k_query_prepare");
    }
    return DS_OK;
}

// tiles: queries grouped by column count so that a tile's columns fit the slot table
static void plan_tiles(const QuerySet &qs, const std::vector<int32_t> &batch, std::vector<int32_t> *tile_q) {
    // class c holds queries with at most limit[c] columns, tiles of size[c] queries
    const int limits[6] = {63, 127, 255, 511, 1023, MAX_SLOTS - 1};
    const int sizes[6] = {32, 16, 8, 4, 2, 1};
    std::vector<int32_t> classes[6];
    for (int32_t b = 0; b < (int32_t)batch.size(); ++b) {
        int64_t q = batch[b];
        int64_t g = qs.h_ptr[q + 1] - qs.h_ptr[q];
        int c = 0;
        while (g > limits[c]) ++c;
        classes[c].push_back(b);
    }
    tile_q->clear();
    for (int c = 0; c < 6; ++c) {
        const std::vector<int32_t> &members = classes[c];
        for (size_t i = 0; i < members.size(); i += sizes[c]) {
            size_t base = tile_q->size();
            tile_q->resize(base + TQ, -1);
            for (int j = 0; j < sizes[c] && i + j < members.size(); ++j) (*tile_q)[base + j] = members[i + j];
        }
    }
}

static int launch_scan(const Index &ix, cudaStream_t stream, ScanParams sp, int n_tiles, int n_queries) {
    const size_t smem = scan_smem_bytes(ix.n_vocab);
    const int64_t rows = sp.r1 - sp.r0;
    if (rows <= 0 || n_tiles <= 0) return DS_OK;
    const bool big = smem > 110 * 1024;
    const int threads = big ? 512 : 256;
    // enough CTAs to fill 148 SMs a few times, but at least 4 rows per thread to amortise the table build
    int64_t rows_per_cta = std::max<int64_t>(threads * 4, ceil_div(rows, std::max<int64_t>(1, ceil_div(148 * 32, n_tiles))));
    rows_per_cta = ceil_div(rows_per_cta, threads) * threads;
    sp.rows_per_cta = (int)std::min<int64_t>(rows_per_cta, 1 << 30);
    int64_t row_ctas = ceil_div(rows, sp.rows_per_cta);
    for (int t0 = 0; t0 < n_tiles; t0 += 65535) {
        int nt = std::min(65535, n_tiles - t0);
        ScanParams part = sp;
        part.tile_q = sp.tile_q + (size_t)t0 * TQ;
        dim3 grid((unsigned)row_ctas, (unsigned)nt);
        cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
        if (g_profile.enabled) {
            DS_CUDA(cudaEventCreate(&ev_start));
            DS_CUDA(cudaEventCreate(&ev_stop));
            DS_CUDA(cudaEventRecord(ev_start, stream));
        }
        if (big) {
            DS_CHECK(ensure_dynamic_smem(reinterpret_cast<const void *>(&k_scan<512>), smem));
            k_scan<512><<<grid, 512, smem, stream>>>(part);
        } else {
            DS_CHECK(ensure_dynamic_smem(reinterpret_cast<const void *>(&k_scan<256>), smem));
            k_scan<256><<<grid, 256, smem, stream>>>(part);
        }
        DS_LAUNCHED("k_scan");
        if (ev_start != nullptr) {
            DS_CUDA(cudaEventRecord(ev_stop, stream));
            std::lock_guard<std::mutex> guard(g_profile_lock);
            g_profile.events[0].emplace_back(ev_start, ev_stop);
            g_profile.pairs[0] += (double)rows * (double)n_queries * ((double)nt / (double)n_tiles);
        }
    }
    return DS_OK;
}

// posting blocks [s0, s1) for every batch query; counted with k_scan in the profile (same pairs, same role)
static int launch_post(const Index &ix, cudaStream_t stream, PostParams pp, int s0, int s1) {
    if (s1 <= s0 || pp.n_batch <= 0) return DS_OK;
    pp.post = ix.post;
    pp.seg_off = ix.seg_off;
    pp.seg_base = ix.seg_base;
    pp.n_vocab = ix.n_vocab;
    pp.n_truth = ix.n_truth;
    pp.sums = ix.sums_pos;
    pp.sums_floor = ix.sums_floor;
    pp.group_pat = ix.group_pat;
    pp.pat_pos = ix.pat_pos;
    pp.dense_bit = ix.dense_bit;
    pp.dense_w = ix.dense_w;
    pp.w32 = ix.w32;
    pp.s0 = s0;
    pp.s1 = s1;
    // a task = one query x a run of blocks: long runs amortise the query set-up, short ones keep small batches parallel
    const long long wanted_tasks = (long long)148 * 2 * POST_WARPS * 2;
    pp.run_len = (int)std::min<long long>(POST_RUN, std::max<long long>(1, (long long)(s1 - s0) * pp.n_batch / wanted_tasks));
    pp.n_tasks = ceil_div(s1 - s0, pp.run_len) * (long long)pp.n_batch;
    // Tasks are handed out warp by warp inside a CTA (up to DS_POST_TASKS per warp level the very uneven task lengths);
    // the grid is a whole number of waves of 148 SMs x 2 resident CTAs, so that a small launch (a few thousand
    // queries per rank at 8 GPUs) is one full wave instead of one full and one nearly empty one.
    const long long resident = (long long)148 * POST_CTAS;
    const long long most_per_cta = (long long)POST_WARPS * DS_POST_TASKS;
    const long long waves = std::max<long long>(1, ceil_div(pp.n_tasks, resident * most_per_cta));
    pp.tasks_per_cta = (int)std::max<long long>(1, ceil_div(pp.n_tasks, resident * waves));
    const long long ctas = ceil_div(pp.n_tasks, (long long)pp.tasks_per_cta);
    if (ctas > INT32_MAX) return fail(DS_ERR_UNSUPPORTED, "too many posting tasks in one launch");
    const size_t smem = (size_t)POST_WARPS * (POST_ROWS * sizeof(post_acc_t) + POST_LIST * 2 + POST_DESC * 8);
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    if (g_profile.enabled) {
        DS_CUDA(cudaEventCreate(&ev_start));
        DS_CUDA(cudaEventCreate(&ev_stop));
        DS_CUDA(cudaEventRecord(ev_start, stream));
    }
    DS_CHECK(ensure_dynamic_smem(reinterpret_cast<const void *>(&k_post), smem));
    k_post<<<(unsigned)ctas, POST_WARPS * 32, smem, stream>>>(pp);
    DS_LAUNCHED("k_post");
    if (ev_start != nullptr) {
        DS_CUDA(cudaEventRecord(ev_stop, stream));
        std::lock_guard<std::mutex> guard(g_profile_lock);
        g_profile.events[1].emplace_back(ev_start, ev_stop);
        const int64_t rows = std::min<int64_t>(ix.n_truth, (int64_t)s1 * POST_ROWS) - (int64_t)s0 * POST_ROWS;
        g_profile.pairs[1] += (double)rows * (double)pp.n_batch;
    }
    return DS_OK;
}

static int launch_select(cudaStream_t stream, SelectParams sp) {
    if (sp.n_batch <= 0) return DS_OK;
    int widest = 1;   // candidates are padded to a power of two for the in-place sort
    while (widest < std::max(sp.cap, sp.dense_rows)) widest <<= 1;
    int items = widest + sp.m;
    sp.max_items = items;
    size_t smem = (size_t)4 * items * 12 + (size_t)4 * SELECT_HASH * 8;
    DS_CHECK(ensure_dynamic_smem(reinterpret_cast<const void *>(&k_select), smem));
    k_select<<<(unsigned)ceil_div(sp.n_batch, 4), 128, smem, stream>>>(sp);
    DS_LAUNCHED("k_select");
    return DS_OK;
}

// Runs the scan + select pipeline for one batch of queries (ids into the call's query set).
//   MODE_SCORE  phase 1: retained list = best m by score, exported as (score64, global row)
//   MODE_ROW    phase 3: retained list = the k highest rows reaching the fixed threshold, exported as rows
// `fixed` = bounded-memory form: dense chunks of FIXED_ROWS rows only (cannot overflow).  Otherwise the
// first chunk (MODE_SCORE) is dense to seed the thresholds and the rest goes through the candidate
// buffers; queries whose buffer overflows are returned in `overflowed` to be redone with `fixed`.
struct SharedTheta {
    double *own = nullptr;
    const double *peers[DS_MAX_PEERS] = {nullptr};
    int n_peers = 0;
};

struct PipelineOut {
    double *score = nullptr;      // MODE_SCORE [n_q * m]
    int64_t *row = nullptr;       // MODE_SCORE [n_q * m]
    int64_t *rows_k = nullptr;    // MODE_ROW   [n_q * k]
    int32_t *count = nullptr;     // MODE_ROW   [n_q]
};

static int run_pipeline(Workspace &call_ws, const Index &ix, const QuerySet &qs, const std::vector<int32_t> &batch, int mode,
                        int k, int m, bool fixed, bool wide, const double *d_threshold, const PipelineOut &out,
                        std::vector<int32_t> *overflowed, const SharedTheta &shared = SharedTheta()) {
    cudaStream_t stream = call_ws.stream();
    const int n_batch = (int)batch.size();
    if (n_batch == 0) return DS_OK;
    Workspace ws(stream);
    std::vector<int32_t> tile_q;
    plan_tiles(qs, batch, &tile_q);
    const int n_tiles = (int)(tile_q.size() / TQ);

    int32_t *d_batch = nullptr, *d_tile = nullptr, *d_ret_row = nullptr;
    int *d_cand_count = nullptr, *d_ret_n = nullptr, *d_state = nullptr, *d_overflow = nullptr;
    float2 *d_ab = nullptr;
    uint2 *d_cand = nullptr;
    double *d_ret_score = nullptr, *d_theta = nullptr;
    float *d_dense = nullptr;
    const int cand_cap = wide ? CAND_CAP_WIDE : (mode == MODE_ROW ? CAND_CAP_MAX : std::min(CAND_CAP_MAX, std::max(256, 8 * k)));
    const bool dense_first = mode == MODE_SCORE;
    const int dense_rows = fixed ? FIXED_ROWS : std::max(DENSE_ROWS, ((2 * k + 255) / 256) * 256);
    DS_CHECK(ws.alloc(&d_batch, n_batch));
    DS_CHECK(ws.alloc(&d_tile, tile_q.size()));
    DS_CHECK(ws.alloc(&d_ret_row, (size_t)n_batch * m));
    DS_CHECK(ws.alloc(&d_ret_score, (size_t)n_batch * m));
    DS_CHECK(ws.alloc(&d_cand_count, n_batch));
    DS_CHECK(ws.alloc(&d_ret_n, n_batch));
    DS_CHECK(ws.alloc(&d_state, n_batch));
    DS_CHECK(ws.alloc(&d_overflow, 1));
    DS_CHECK(ws.alloc(&d_ab, n_batch));
    DS_CHECK(ws.alloc(&d_theta, n_batch));
    if (fixed || dense_first) DS_CHECK(ws.alloc(&d_dense, (size_t)n_batch * dense_rows));
    if (!fixed) DS_CHECK(ws.alloc(&d_cand, (size_t)n_batch * cand_cap));
    DS_CUDA(cudaMemcpyAsync(d_batch, batch.data(), (size_t)n_batch * 4, cudaMemcpyHostToDevice, stream));
    DS_CUDA(cudaMemcpyAsync(d_tile, tile_q.data(), tile_q.size() * 4, cudaMemcpyHostToDevice, stream));
    DS_CUDA(cudaMemsetAsync(d_cand_count, 0, (size_t)n_batch * 4, stream));
    DS_CUDA(cudaMemsetAsync(d_ret_n, 0, (size_t)n_batch * 4, stream));
    DS_CUDA(cudaMemsetAsync(d_state, 0, (size_t)n_batch * 4, stream));
    DS_CUDA(cudaMemsetAsync(d_overflow, 0, 4, stream));
    DS_CUDA(cudaMemsetAsync(d_ab, 0, (size_t)n_batch * 8, stream));
    DS_CUDA(cudaMemsetAsync(d_theta, 0, (size_t)n_batch * 8, stream));
    if (mode == MODE_ROW && !fixed) {
        k_rescan_init<<<(unsigned)ceil_div(n_batch, 256), 256, 0, stream>>>(d_batch, n_batch, d_threshold, qs.d_mx, d_ab, d_state,
                                                                             d_overflow);
        DS_LAUNCHED("k_rescan_init");
    }

    ScanParams sp{};
    sp.chunks = ix.chunks;
    sp.chunk_ptr = ix.chunk_ptr;
    sp.sums = ix.sums_pos;
    sp.w32 = ix.w32;
    sp.n_vocab = ix.n_vocab;
    sp.q_sorted = qs.d_sorted;
    sp.q_ptr = qs.d_ptr;
    sp.batch_q = d_batch;
    sp.tile_q = d_tile;
    sp.ab = d_ab;
    sp.cand = d_cand;
    sp.cand_count = d_cand_count;
    sp.cap = cand_cap;
    sp.dense_stride = dense_rows;

    PostParams pp{};
    pp.q_sorted = qs.d_sorted;
    pp.q_ptr = qs.d_ptr;
    pp.batch_q = d_batch;
    pp.ab = d_ab;
    pp.n_batch = n_batch;
    pp.cand = d_cand;
    pp.cand_count = d_cand_count;
    pp.cap = cand_cap;

    SelectParams sel{};
    sel.mode = mode;
    sel.batch_q = d_batch;
    sel.q_mx = qs.d_mx;
    sel.threshold = d_threshold;
    sel.sums = ix.sums_pos;
    sel.perm = ix.perm;
    sel.n_batch = n_batch;
    sel.k = k;
    sel.m = m;
    sel.cand = d_cand;
    sel.cand_count = d_cand_count;
    sel.cap = fixed ? 0 : cand_cap;
    sel.dense_stride = dense_rows;
    sel.ret_score = d_ret_score;
    sel.ret_row = d_ret_row;
    sel.ret_n = d_ret_n;
    sel.theta = d_theta;
    sel.ab = d_ab;
    sel.state = d_state;
    sel.overflow_count = d_overflow;
    sel.chunks = ix.chunks;
    sel.chunk_ptr = ix.chunk_ptr;
    sel.w32 = ix.w32;
    sel.n_vocab = ix.n_vocab;
    sel.q_sorted = qs.d_sorted;
    sel.q_ptr = qs.d_ptr;
    sel.theta_own = mode == MODE_SCORE ? shared.own : nullptr;
    sel.n_peers = mode == MODE_SCORE && shared.own != nullptr ? shared.n_peers : 0;
    for (int i = 0; i < DS_MAX_PEERS; ++i) sel.theta_peers[i] = shared.peers[i];

    const int64_t n = ix.n_truth;
    int64_t r0 = 0;
    bool first = true;
    while (r0 < n) {
        const bool dense = fixed || (first && dense_first);
        int64_t r1;
        if (dense) r1 = std::min<int64_t>(n, r0 + dense_rows);
        else if (mode == MODE_ROW) r1 = n;  // the threshold is fixed: one pass over all rows
        else {
            // Growing sweep: thresholds tighten early.  After R rows the threshold is the k-th best of R rows, so about
            // k * M / R of the next M rows pass it and a range may grow by more than x2 without overflowing the candidate
            // buffers; but the blocks swept with a stale threshold list more rows for the full test.  Measured per C3 step
            // (100k queries): x2 37.6 ms, x3 36.6, x4 37.4, x8 39.6.  Small batches (a rank of an 8-GPU run holds 12,500
            // queries) are bound by the latency of their short launches, not by instructions, and fewer, longer launches
            // win: 12,500 queries x2 5.95 ms, x3 5.70, x4 5.69, x8 5.83; 25,000 queries x2 11.6, x4 10.9.
            // The growth is capped at 1 + cap / (5 k) (top_n = 100: x2 - 73.2 ms per C3 step against 76.5 with x3).
            // Ranges end on posting-block boundaries.  DS_POST_GROWTH in the environment overrides.
            static const int64_t growth_env = []() {
                const char *env = getenv("DS_POST_GROWTH");
                return env ? std::max<int64_t>(2, strtoll(env, nullptr, 10)) : (int64_t)0;
            }();
            const int64_t wanted = growth_env > 0 ? growth_env : (n_batch <= DS_SMALL_BATCH ? 4 : DS_POST_GROWTH);
            const int64_t growth = r0 >= POST_ROWS ? std::min<int64_t>(wanted, std::max<int64_t>(2, 1 + (int64_t)(cand_cap / (5.0 * k)))) : 2;
            r1 = std::max<int64_t>(growth * r0, r0 + dense_rows);
            if (r1 > POST_ROWS) r1 = ceil_div(r1, (int64_t)POST_ROWS) * POST_ROWS;
            r1 = std::min<int64_t>(n, r1);
        }
        // The first `scan_rows` rows of a MODE_SCORE sweep go through the dense k_scan, whose cost (1.33 ms per 4,096-row
        // block and 100k queries) does not depend on the thresholds; k_post costs 0.82 ms on the first block after the seed
        // and 0.29 ms on the last at top_n = 10, but lists far more rows per block while the thresholds settle when top_n
        // is large.  Measured on C3: top_n = 10: 38.2 ms with 4,096 or 16,384 scanned rows, 42.3 with 65,536; top_n = 100:
        // 100.1 ms with 4,096, 74.1 ms with 65,536.  DS_SCAN_ROWS (environment) overrides.
        static const int64_t scan_rows_env = []() {
            const char *env = getenv("DS_SCAN_ROWS");
            return env ? std::max<int64_t>(POST_ROWS, ceil_div(strtoll(env, nullptr, 10), (int64_t)POST_ROWS) * POST_ROWS) : (int64_t)0;
        }();
        const int64_t scan_rows = scan_rows_env > 0 ? scan_rows_env : (int64_t)POST_ROWS * std::min(16, std::max(1, k / 6));
        const bool inverted = !dense && ix.post != nullptr && r0 % POST_ROWS == 0 && (r0 >= scan_rows || mode == MODE_ROW);
        if (inverted) {
            DS_CHECK(launch_post(ix, stream, pp, (int)(r0 / POST_ROWS), (int)ceil_div(r1, (int64_t)POST_ROWS)));
        } else {
            sp.r0 = (int)r0;
            sp.r1 = (int)r1;
            sp.dense = dense ? d_dense : nullptr;
            DS_CHECK(launch_scan(ix, stream, sp, n_tiles, n_batch));
        }
        sel.dense = dense ? d_dense : nullptr;
        sel.dense_rows = dense ? (int)(r1 - r0) : 0;
        sel.dense_r0 = (int)r0;
        sel.rescore = inverted ? 1 : 0;
        DS_CHECK(launch_select(stream, sel));
        r0 = r1;
        first = false;
    }
    if (mode == MODE_SCORE) {
        k_export<<<(unsigned)ceil_div((int64_t)n_batch * m, 256), 256, 0, stream>>>(d_batch, n_batch, m, d_ret_score, d_ret_row,
                                                                                    d_ret_n, ix.row_offset, out.score, out.row);
        DS_LAUNCHED("k_export");
    } else {
        k_export_rows<<<(unsigned)ceil_div((int64_t)n_batch * k, 256), 256, 0, stream>>>(d_batch, n_batch, k, d_ret_row, d_ret_n,
                                                                                         ix.row_offset, out.rows_k, out.count);
        DS_LAUNCHED("k_export_rows");
    }
    if (!fixed && overflowed != nullptr) {
        int h_overflow = 0;
        DS_CUDA(cudaMemcpyAsync(&h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
        if (h_overflow > 0) {
            std::vector<int> h_state(n_batch);
            DS_CUDA(cudaMemcpyAsync(h_state.data(), d_state, (size_t)n_batch * 4, cudaMemcpyDeviceToHost, stream));
            DS_CUDA(cudaStreamSynchronize(stream));
            for (int b = 0; b < n_batch; ++b)
                if (h_state[b] & STATE_OVERFLOW) overflowed->push_back(batch[b]);
        }
    }
    return DS_OK;
}

// runs `ids` through the pipeline in workspace-sized batches, then redoes the overflowed ones with wide buffers and
// what still overflows in the bounded-memory form (adversarial row order, massive ties, thresholds <= 0)
static int run_batches(Workspace &ws, const Index &ix, const QuerySet &qs, const std::vector<int32_t> &ids, int mode, int k, int m,
                       const double *d_threshold, const PipelineOut &out, const SharedTheta &shared = SharedTheta()) {
    std::vector<int32_t> overflowed, still_overflowed;
    for (size_t i0 = 0; i0 < ids.size(); i0 += QUERY_BATCH) {
        size_t i1 = std::min(ids.size(), i0 + QUERY_BATCH);
        std::vector<int32_t> batch(ids.begin() + i0, ids.begin() + i1);
        DS_CHECK(run_pipeline(ws, ix, qs, batch, mode, k, m, false, false, d_threshold, out, &overflowed, shared));
    }
    // second try with 4,096-entry buffers (thousands of rows tied at the threshold, e.g. a title the truth DB repeats)
    const bool wide_fits = (size_t)4 * (CAND_CAP_WIDE + m) * 12 + (size_t)4 * SELECT_HASH * 8 <= 227 * 1024;   // k_select's shared memory (4 warps per CTA)
    if (!wide_fits) still_overflowed.swap(overflowed);
    for (size_t i0 = 0; i0 < overflowed.size(); i0 += QUERY_BATCH / 8) {
        size_t i1 = std::min(overflowed.size(), i0 + QUERY_BATCH / 8);
        std::vector<int32_t> batch(overflowed.begin() + i0, overflowed.begin() + i1);
        DS_CHECK(run_pipeline(ws, ix, qs, batch, mode, k, m, false, true, d_threshold, out, &still_overflowed, shared));
    }
    for (size_t i0 = 0; i0 < still_overflowed.size(); i0 += QUERY_BATCH / 8) {
        size_t i1 = std::min(still_overflowed.size(), i0 + QUERY_BATCH / 8);
        std::vector<int32_t> batch(still_overflowed.begin() + i0, still_overflowed.begin() + i1);
        DS_CHECK(run_pipeline(ws, ix, qs, batch, mode, k, m, true, false, d_threshold, out, nullptr, shared));
    }
    return DS_OK;
}

static int local_topn(Workspace &ws, const Index &ix, const QuerySet &qs, int k, int m, double *d_out_score,
                      int64_t *d_out_row, const SharedTheta &shared = SharedTheta()) {
    std::vector<int32_t> ids((size_t)qs.n_q);
    for (int64_t q = 0; q < qs.n_q; ++q) ids[(size_t)q] = (int32_t)q;
    PipelineOut out;
    out.score = d_out_score;
    out.row = d_out_row;
    return run_batches(ws, ix, qs, ids, MODE_SCORE, k, m, nullptr, out, shared);
}

static int merge_topn(cudaStream_t stream, int n_shards, int64_t n_q, int k, int m, int64_t n_total, const double *d_score,
                      const int64_t *d_row, const double *d_mx, int64_t *d_out_rows, int32_t *d_out_count, float *d_out_kth,
                      double *d_out_threshold, int32_t *d_out_flags, int *d_flagged_count) {
    if (n_q <= 0) return DS_OK;
    MergeParams mp{};
    mp.n_shards = n_shards;
    mp.n_q = n_q;
    mp.k = k;
    mp.m = m;
    mp.n_total = n_total;
    mp.all_score = d_score;
    mp.all_row = d_row;
    mp.q_mx = d_mx;
    mp.out_rows = d_out_rows;
    mp.out_count = d_out_count;
    mp.out_kth = d_out_kth;
    mp.out_threshold = d_out_threshold;
    mp.out_flags = d_out_flags;
    mp.flagged_count = d_flagged_count;
    k_merge<<<(unsigned)ceil_div(n_q, 4), 128, 0, stream>>>(mp);
    DS_LAUNCHED("k_merge");
    return DS_OK;
}

// exact re-scan of the local shard for `flagged` queries with known thresholds: the k highest local rows
// whose float64 score reaches the threshold (phase 3)
static int rescan_topn(Workspace &ws, const Index &ix, const QuerySet &qs, const double *d_threshold,
                       const std::vector<int32_t> &flagged, int k, int64_t *d_out_rows, int32_t *d_out_count) {
    PipelineOut out;
    out.rows_k = d_out_rows;
    out.count = d_out_count;
    return run_batches(ws, ix, qs, flagged, MODE_ROW, k, k, d_threshold, out);
}

static int check_topn_args(const Index *ix, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids, int32_t k) {
    if (ix == nullptr) return fail(DS_ERR_BAD_ARG, "index is NULL");
    if (n_q < 0) return fail(DS_ERR_BAD_ARG, "n_q < 0");
    if (n_q > 0 && q_row_ptr == nullptr) return fail(DS_ERR_BAD_ARG, "q_row_ptr is NULL");
    if (k < 1 || k > DS_MAX_TOP_N) return fail(DS_ERR_UNSUPPORTED, "top_n %d outside 1..%d", k, DS_MAX_TOP_N);
    if (n_q > INT32_MAX) return fail(DS_ERR_UNSUPPORTED, "n_q exceeds 2^31-1");
    (void)q_col_ids;
    return DS_OK;
}

}  // namespace ds

using namespace ds;

struct ds_index {
    Index ix;
};

extern "C" {

int ds_version(void) { return DS_VERSION; }

int ds_profile_begin(void) {
    std::lock_guard<std::mutex> guard(g_profile_lock);
    for (auto &events : g_profile.events) {
        for (auto &e : events) {
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        events.clear();
    }
    g_profile.pairs[0] = g_profile.pairs[1] = 0.0;
    g_profile.enabled = true;
    return DS_OK;
}

int ds_profile_end_split(double *ms, int64_t *launches, double *pairs) {
    std::lock_guard<std::mutex> guard(g_profile_lock);
    g_profile.enabled = false;
    for (int kernel = 0; kernel < 2; ++kernel) {
        double total = 0.0;
        for (auto &e : g_profile.events[kernel]) {
            float one = 0.0f;
            DS_CUDA(cudaEventSynchronize(e.second));
            DS_CUDA(cudaEventElapsedTime(&one, e.first, e.second));
            total += one;
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        if (ms) ms[kernel] = total;
        if (launches) launches[kernel] = (int64_t)g_profile.events[kernel].size();
        if (pairs) pairs[kernel] = g_profile.pairs[kernel];
        g_profile.events[kernel].clear();
    }
    return DS_OK;
}

int ds_profile_end(double *scan_ms, int64_t *scan_launches, double *scan_pairs) {
    double ms[2];
    int64_t launches[2];
    double pairs[2];
    DS_CHECK(ds_profile_end_split(ms, launches, pairs));
    if (scan_ms) *scan_ms = ms[0] + ms[1];
    if (scan_launches) *scan_launches = launches[0] + launches[1];
    if (scan_pairs) *scan_pairs = pairs[0] + pairs[1];
    return DS_OK;
}
const char *ds_last_error(void) { return g_last_error; }
int64_t ds_kernel_launches(void) { return g_kernel_launches.load(); }

int ds_trim(int device) {
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || device < 0 || device >= n_devices) {
        cudaGetLastError();
        return fail(DS_ERR_BAD_ARG, "device %d out of range", device);
    }
    cudaMemPool_t pool;
    DS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    DS_CUDA(cudaMemPoolTrimTo(pool, 0));
    return DS_OK;
}

int32_t ds_topn_retained(int32_t k) {
    int extra = std::max(32, k / 4);
    return ((k + extra + 31) / 32) * 32;
}

int ds_index_create(ds_index **out, int device, int64_t n_truth, int32_t n_vocab, const int64_t *t_row_ptr,
                    const uint16_t *t_col_ids, const double *idf64_by_col, const float *sums_truth_f32,
                    int64_t global_row_offset, int64_t n_truth_total, void *stream_) {
    if (out == nullptr) return fail(DS_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (n_truth < 0 || n_truth >= (int64_t)1 << 31) return fail(DS_ERR_UNSUPPORTED, "n_truth %lld outside 0..2^31-1", (long long)n_truth);
    if (n_vocab < 1 || n_vocab > 65535) return fail(DS_ERR_UNSUPPORTED, "n_vocab %d outside 1..65535 (u16 column ids)", n_vocab);
    if (t_row_ptr == nullptr || idf64_by_col == nullptr) return fail(DS_ERR_BAD_ARG, "t_row_ptr / idf64_by_col is NULL");
    if (scan_smem_bytes(n_vocab) > 227 * 1024) return fail(DS_ERR_UNSUPPORTED, "n_vocab %d needs more than 227 KB of shared memory", n_vocab);
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(DS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n_devices) return fail(DS_ERR_BAD_ARG, "device %d out of range (%d devices)", device, n_devices);
    DeviceGuard guard(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    std::vector<int64_t> h_ptr((size_t)n_truth + 1);
    if (is_device_pointer(t_row_ptr)) {
        DS_CUDA(cudaMemcpyAsync(h_ptr.data(), t_row_ptr, ((size_t)n_truth + 1) * 8, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
    } else {
        memcpy(h_ptr.data(), t_row_ptr, ((size_t)n_truth + 1) * 8);
    }
    const int64_t nnz = h_ptr[(size_t)n_truth];
    if (h_ptr[0] != 0 || nnz < 0) return fail(DS_ERR_BAD_ARG, "t_row_ptr must start at 0 and be non-decreasing");
    for (int64_t r = 0; r < n_truth; ++r)
        if (h_ptr[(size_t)r + 1] < h_ptr[(size_t)r]) return fail(DS_ERR_BAD_ARG, "t_row_ptr is decreasing at row %lld", (long long)r);
    if (nnz > 0 && t_col_ids == nullptr) return fail(DS_ERR_BAD_ARG, "t_col_ids is NULL");
    if (ceil_div(nnz, CHUNK_COLS) + n_truth >= ((int64_t)1 << 32)) return fail(DS_ERR_UNSUPPORTED, "index too large for 32-bit chunk offsets");

    ds_index *handle = new (std::nothrow) ds_index();
    if (handle == nullptr) return fail(DS_ERR_NO_MEMORY, "out of host memory");
    Index &ix = handle->ix;
    ix.device = device;
    ix.n_truth = n_truth;
    ix.n_vocab = n_vocab;
    ix.row_offset = global_row_offset;
    ix.n_total = n_truth_total > 0 ? n_truth_total : n_truth;
    ix.last_stream = stream;
    int status = [&]() -> int {
        Workspace ws(stream);
        const int64_t *d_ptr = nullptr;
        const uint16_t *d_cols = nullptr;
        const double *d_w64_in = nullptr;
        const float *d_sums_in = nullptr;
        DS_CHECK(ws.stage_in(&d_ptr, t_row_ptr, (size_t)n_truth + 1));
        DS_CHECK(ws.stage_in(&d_cols, t_col_ids, (size_t)nnz));
        DS_CHECK(ws.stage_in(&d_w64_in, idf64_by_col, (size_t)n_vocab));
        DS_CHECK(ws.stage_in(&d_sums_in, sums_truth_f32, (size_t)n_truth));
        const size_t rows_alloc = (size_t)std::max<int64_t>(1, n_truth);
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.w64), (size_t)n_vocab * 8, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.w32), ((size_t)n_vocab + 1) * 4, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.sums), rows_alloc * 4, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.sums_pos), rows_alloc * 4, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.perm), rows_alloc * 4, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.pat_pos), rows_alloc * 4, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.chunk_ptr), ((size_t)n_truth + 1) * 4, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.dense_bit), (size_t)n_vocab + 1, stream));
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.dense_w), DENSE_MAX * 4, stream));
        DS_CUDA(cudaMemcpyAsync(ix.w64, d_w64_in, (size_t)n_vocab * 8, cudaMemcpyDeviceToDevice, stream));
        int *d_post_flags = nullptr;   // [0]: a weight or row sum is negative / NaN, [1]: a row repeats a column, [2]: a block is too long
        DS_CHECK(ws.alloc(&d_post_flags, 3));
        DS_CUDA(cudaMemsetAsync(d_post_flags, 0, 12, stream));
        k_weights<<<(unsigned)ceil_div(n_vocab + 1, 256), 256, 0, stream>>>(ix.w64, ix.w32, n_vocab, d_post_flags);
        DS_LAUNCHED("k_weights");

        // inverted form for k_post wanted?  (small indexes: the dense sweep of the first rows covers them)
        const int64_t n_sub = ceil_div(n_truth, (int64_t)POST_ROWS);
        const int64_t n_seg = n_sub * n_vocab;
        const bool want_post = n_truth > 2 * POST_ROWS && n_seg < ((int64_t)1 << 30) && (uint64_t)(nnz + 4 * n_truth) < ((uint64_t)1 << 32) - 256;

        // dense columns: the (at most 32) commonest columns that sit in >= 1 / 128 of the rows; bit 31 = the commonest
        std::vector<uint8_t> h_dense_bit((size_t)n_vocab + 1, 255);
        float h_dense_w[DENSE_MAX] = {0.0f};
        ix.n_dense = 0;
        if (want_post && nnz > 0) {
            uint32_t *d_df = nullptr;
            DS_CHECK(ws.alloc(&d_df, (size_t)n_vocab));
            DS_CUDA(cudaMemsetAsync(d_df, 0, (size_t)n_vocab * 4, stream));
            k_col_df<<<(unsigned)ceil_div(nnz, 256), 256, 0, stream>>>(d_cols, nnz, n_vocab, d_df);
            DS_LAUNCHED("k_col_df");
            std::vector<uint32_t> h_df((size_t)n_vocab);
            std::vector<float> h_w32((size_t)n_vocab);
            DS_CUDA(cudaMemcpyAsync(h_df.data(), d_df, (size_t)n_vocab * 4, cudaMemcpyDeviceToHost, stream));
            DS_CUDA(cudaMemcpyAsync(h_w32.data(), ix.w32, (size_t)n_vocab * 4, cudaMemcpyDeviceToHost, stream));
            DS_CUDA(cudaStreamSynchronize(stream));
            std::vector<int32_t> order;
            const uint32_t least = (uint32_t)std::max<int64_t>(1, n_truth / DENSE_MIN_SHARE);
            for (int32_t c = 0; c < n_vocab; ++c)
                if (h_df[(size_t)c] >= least && h_w32[(size_t)c] >= 0.0f && std::isfinite(h_w32[(size_t)c])) order.push_back(c);
            std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return h_df[(size_t)x] != h_df[(size_t)y] ? h_df[(size_t)x] > h_df[(size_t)y] : x < y; });
            if (order.size() > (size_t)DENSE_MAX) order.resize(DENSE_MAX);
            ix.n_dense = (int)order.size();
            for (int rank = 0; rank < ix.n_dense; ++rank) {
                const int bit = DENSE_MAX - 1 - rank;
                h_dense_bit[(size_t)order[(size_t)rank]] = (uint8_t)bit;
                h_dense_w[bit] = h_w32[(size_t)order[(size_t)rank]];
            }
        }
        DS_CUDA(cudaMemcpyAsync(ix.dense_bit, h_dense_bit.data(), (size_t)n_vocab + 1, cudaMemcpyHostToDevice, stream));
        DS_CUDA(cudaMemcpyAsync(ix.dense_w, h_dense_w, DENSE_MAX * 4, cudaMemcpyHostToDevice, stream));

        // position -> original row: inside every block of SORT_BLOCK rows sorted by the row sum
        uint32_t *d_pat_row = nullptr;
        unsigned long long *d_keys = nullptr, *d_keys_sorted = nullptr;
        int32_t *d_ids = nullptr;
        DS_CHECK(ws.alloc(&d_pat_row, rows_alloc));
        DS_CHECK(ws.alloc(&d_keys, rows_alloc));
        DS_CHECK(ws.alloc(&d_keys_sorted, rows_alloc));
        DS_CHECK(ws.alloc(&d_ids, rows_alloc));
        uint32_t *d_counts = nullptr;
        DS_CHECK(ws.alloc(&d_counts, (size_t)n_truth + 1));
        DS_CUDA(cudaMemsetAsync(d_counts, 0, ((size_t)n_truth + 1) * 4, stream));
        if (n_truth > 0) {
            if (d_sums_in != nullptr)
                DS_CUDA(cudaMemcpyAsync(ix.sums, d_sums_in, (size_t)n_truth * 4, cudaMemcpyDeviceToDevice, stream));
            k_row_keys<<<(unsigned)ceil_div(n_truth, 256), 256, 0, stream>>>(d_ptr, d_cols, ix.w32, n_vocab, n_truth, ix.dense_bit,
                                                                            d_sums_in == nullptr ? 1 : 0, ix.sums, d_pat_row, d_keys, d_ids, d_post_flags);
            DS_LAUNCHED("k_row_keys");
            int end_bit = 42;
            while (end_bit < 64 && ((int64_t)1 << (end_bit - 42)) <= ceil_div(n_truth, (int64_t)SORT_BLOCK)) ++end_bit;
            size_t sort_bytes = 0;
            DS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys_sorted, d_ids, ix.perm, (int)n_truth, 0, end_bit, stream));
            unsigned char *d_sort_temp = nullptr;
            DS_CHECK(ws.alloc(&d_sort_temp, sort_bytes));
            DS_CUDA(cub::DeviceRadixSort::SortPairs(d_sort_temp, sort_bytes, d_keys, d_keys_sorted, d_ids, ix.perm, (int)n_truth, 0, end_bit, stream));
            g_kernel_launches.fetch_add(1);
            k_row_prepare<<<(unsigned)ceil_div(n_truth, 256), 256, 0, stream>>>(d_ptr, n_truth, ix.perm, d_pat_row, ix.pat_pos, d_counts, ix.sums,
                                                                               ix.sums_pos);
            DS_LAUNCHED("k_row_prepare");
        }
        size_t temp_bytes = 0;
        DS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_counts, ix.chunk_ptr, (int)(n_truth + 1), stream));
        unsigned char *d_temp = nullptr;
        DS_CHECK(ws.alloc(&d_temp, temp_bytes));
        DS_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_counts, ix.chunk_ptr, (int)(n_truth + 1), stream));
        g_kernel_launches.fetch_add(1);
        uint32_t total_chunks = 0;
        DS_CUDA(cudaMemcpyAsync(&total_chunks, ix.chunk_ptr + n_truth, 4, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
        ix.n_chunks = total_chunks;
        DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.chunks), std::max<size_t>(1, (size_t)total_chunks) * sizeof(chunk_t), stream));
        if (n_truth > 0) {
            k_row_pack<<<(unsigned)ceil_div(n_truth * 32, 256), 256, 0, stream>>>(d_ptr, d_cols, ix.chunk_ptr, n_truth, ix.perm,
                                                                                 (uint16_t)n_vocab, reinterpret_cast<uint16_t *>(ix.chunks));
            DS_LAUNCHED("k_row_pack");
        }
        if (want_post) {
            ix.n_sub = (int)n_sub;
            uint32_t *d_seg_start = nullptr;
            unsigned char *d_seg_temp = nullptr;
            DS_CHECK(ws.alloc(&d_seg_start, (size_t)n_seg + 1));
            DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.seg_off), (size_t)n_sub * (n_vocab + 1) * sizeof(post_off_t), stream));
            DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.seg_base), (size_t)n_sub * 4, stream));
            DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.sums_floor), (size_t)n_sub * POST_GROUPS * 4, stream));
            DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.group_pat), (size_t)n_sub * POST_GROUPS * 4, stream));
            k_sums_floor<<<(unsigned)ceil_div(n_sub * POST_GROUPS * 32, 256), 256, 0, stream>>>(ix.sums_pos, ix.pat_pos, n_truth, n_sub * POST_GROUPS,
                                                                                               ix.sums_floor, ix.group_pat);
            DS_LAUNCHED("k_sums_floor");
            DS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ix.post), std::max<size_t>(1, (size_t)total_chunks * CHUNK_COLS) * 2, stream));
            DS_CUDA(cudaMemsetAsync(d_seg_start, 0, ((size_t)n_seg + 1) * 4, stream));
            const uint16_t *packed = reinterpret_cast<const uint16_t *>(ix.chunks);
            k_post_build<0><<<(unsigned)ceil_div(n_truth, 256), 256, 0, stream>>>(packed, ix.chunk_ptr, n_truth, n_vocab, ix.dense_bit, d_seg_start,
                                                                                  nullptr, d_post_flags + 1);
            DS_LAUNCHED("k_post_build");
            size_t seg_temp_bytes = 0;
            DS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, seg_temp_bytes, d_seg_start, d_seg_start, (int)(n_seg + 1), stream));
            DS_CHECK(ws.alloc(&d_seg_temp, seg_temp_bytes));
            DS_CUDA(cub::DeviceScan::ExclusiveSum(d_seg_temp, seg_temp_bytes, d_seg_start, d_seg_start, (int)(n_seg + 1), stream));
            g_kernel_launches.fetch_add(1);
            k_post_offsets<<<(unsigned)ceil_div(n_sub * (n_vocab + 1), 256), 256, 0, stream>>>(d_seg_start, (int)n_sub, n_vocab, ix.seg_off,
                                                                                              ix.seg_base, d_post_flags + 2);
            DS_LAUNCHED("k_post_offsets");
            k_post_build<1><<<(unsigned)ceil_div(n_truth, 256), 256, 0, stream>>>(packed, ix.chunk_ptr, n_truth, n_vocab, ix.dense_bit, d_seg_start,
                                                                                  ix.post, d_post_flags + 1);   // d_seg_start = running cursors from here on
            DS_LAUNCHED("k_post_build");
            k_post_balance<<<(unsigned)ceil_div(n_seg, BALANCE_WARPS), BALANCE_WARPS * 32, 0, stream>>>(ix.post, ix.seg_off, ix.seg_base, n_seg,
                                                                                                 n_vocab);
            DS_LAUNCHED("k_post_balance");
        }
        int h_post_flags[3] = {0, 0, 0};
        DS_CUDA(cudaMemcpyAsync(h_post_flags, d_post_flags, 12, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
        if ((h_post_flags[0] != 0 || h_post_flags[1] != 0 || h_post_flags[2] != 0) && ix.post != nullptr) {
            cudaFreeAsync(ix.post, stream);
            cudaFreeAsync(ix.seg_off, stream);
            cudaFreeAsync(ix.seg_base, stream);
            cudaFreeAsync(ix.sums_floor, stream);
            cudaFreeAsync(ix.group_pat, stream);
            ix.sums_floor = nullptr;
            ix.group_pat = nullptr;
            ix.post = nullptr;
            ix.seg_off = nullptr;
            ix.seg_base = nullptr;
        }
        return DS_OK;
    }();
    if (status != DS_OK) {
        ds_index_destroy(handle);
        return status;
    }
    *out = handle;
    return DS_OK;
}

int ds_index_destroy(ds_index *index) {
    if (index == nullptr) return DS_OK;
    DeviceGuard guard(index->ix.device);
    // stream-ordered frees (the buffers come from the same pool as the per-call workspace): no device-wide
    // synchronisation, unlike cudaFree.  Freed on the stream of the latest call that used the index, so the
    // pool cannot hand the memory out again while kernels of that call are still in flight.
    void *buffers[] = {index->ix.chunks, index->ix.chunk_ptr, index->ix.sums, index->ix.sums_pos, index->ix.perm, index->ix.w32, index->ix.w64,
                       index->ix.post, index->ix.seg_off, index->ix.seg_base, index->ix.sums_floor, index->ix.group_pat, index->ix.pat_pos,
                       index->ix.dense_bit, index->ix.dense_w};
    for (void *b : buffers)
        if (b != nullptr) cudaFreeAsync(b, index->ix.last_stream);
    delete index;
    return DS_OK;
}

int ds_index_get_sums(const ds_index *index, float *out, void *stream_) {
    if (index == nullptr || out == nullptr) return fail(DS_ERR_BAD_ARG, "index / out is NULL");
    DeviceGuard guard(index->ix.device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const_cast<ds_index *>(index)->ix.last_stream = stream;
    DS_CUDA(cudaMemcpyAsync(out, index->ix.sums, (size_t)index->ix.n_truth * 4, cudaMemcpyDefault, stream));
    if (!is_device_pointer(out)) DS_CUDA(cudaStreamSynchronize(stream));
    return DS_OK;
}

int ds_topn_local_shared(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids, const double *q_mx,
                         int32_t mx_mode, int32_t k, double *out_score, int64_t *out_row, double *out_mx, double *theta_own,
                         const double *const *theta_peers, int32_t n_peers, void *stream_) {
    if (n_peers < 0 || n_peers > DS_MAX_PEERS) return fail(DS_ERR_BAD_ARG, "n_peers %d outside 0..%d", n_peers, DS_MAX_PEERS);
    if (n_peers > 0 && (theta_own == nullptr || theta_peers == nullptr)) return fail(DS_ERR_BAD_ARG, "theta_own / theta_peers is NULL");
    if (theta_own != nullptr && !is_device_pointer(theta_own)) return fail(DS_ERR_BAD_ARG, "theta_own must be device memory");
    SharedTheta shared;
    shared.own = theta_own;
    shared.n_peers = n_peers;
    for (int i = 0; i < n_peers; ++i) shared.peers[i] = theta_peers[i];
    DS_CHECK(check_topn_args(index ? &index->ix : nullptr, n_q, q_row_ptr, q_col_ids, k));
    if (out_score == nullptr || out_row == nullptr) return fail(DS_ERR_BAD_ARG, "out_score / out_row is NULL");
    const Index &ix = index->ix;
    DeviceGuard guard(ix.device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    index->ix.last_stream = stream;
    if (n_q == 0) return DS_OK;
    Workspace ws(stream);
    QuerySet qs;
    DS_CHECK(prepare_queries(ws, ix, n_q, q_row_ptr, q_col_ids, q_mx, mx_mode, &qs));
    const int m = ds_topn_retained(k);
    double *d_score = nullptr, *d_mx_out = nullptr;
    int64_t *d_row = nullptr;
    DS_CHECK(ws.stage_out(&d_score, out_score, (size_t)n_q * m));
    DS_CHECK(ws.stage_out(&d_row, out_row, (size_t)n_q * m));
    DS_CHECK(ws.stage_out(&d_mx_out, out_mx, (size_t)n_q));
    DS_CHECK(local_topn(ws, ix, qs, k, m, d_score, d_row, shared));
    if (d_mx_out != nullptr) DS_CUDA(cudaMemcpyAsync(d_mx_out, qs.d_mx, (size_t)n_q * 8, cudaMemcpyDeviceToDevice, stream));
    return ws.finish_outputs();
}

int ds_topn_local(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids, const double *q_mx,
                  int32_t mx_mode, int32_t k, double *out_score, int64_t *out_row, double *out_mx, void *stream_) {
    return ds_topn_local_shared(index, n_q, q_row_ptr, q_col_ids, q_mx, mx_mode, k, out_score, out_row, out_mx, nullptr, nullptr, 0, stream_);
}

int ds_topn_merge(int32_t n_shards, int64_t n_q, int32_t k, int64_t n_truth_total, const double *all_score,
                  const int64_t *all_row, const double *q_mx, int64_t *out_rows, int32_t *out_count, float *out_kth_f32,
                  double *out_threshold, int32_t *out_flags, int device, void *stream_) {
    if (n_shards < 1 || n_shards > 128) return fail(DS_ERR_BAD_ARG, "n_shards %d outside 1..128", n_shards);
    if (n_q < 0 || k < 1 || k > DS_MAX_TOP_N) return fail(DS_ERR_BAD_ARG, "bad n_q / k");
    if (n_q == 0) return DS_OK;
    if (all_score == nullptr || all_row == nullptr || out_rows == nullptr || out_count == nullptr)
        return fail(DS_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    const int m = ds_topn_retained(k);
    const double *d_score = nullptr, *d_mx = nullptr;
    const int64_t *d_row = nullptr;
    DS_CHECK(ws.stage_in(&d_score, all_score, (size_t)n_shards * n_q * m));
    DS_CHECK(ws.stage_in(&d_row, all_row, (size_t)n_shards * n_q * m));
    DS_CHECK(ws.stage_in(&d_mx, q_mx, (size_t)n_q));
    int64_t *d_rows = nullptr;
    int32_t *d_count = nullptr, *d_flags = nullptr;
    float *d_kth = nullptr;
    double *d_thr = nullptr;
    DS_CHECK(ws.stage_out(&d_rows, out_rows, (size_t)n_q * k));
    DS_CHECK(ws.stage_out(&d_count, out_count, (size_t)n_q));
    DS_CHECK(ws.stage_out(&d_kth, out_kth_f32, (size_t)n_q));
    DS_CHECK(ws.stage_out(&d_thr, out_threshold, (size_t)n_q));
    DS_CHECK(ws.stage_out(&d_flags, out_flags, (size_t)n_q));
    DS_CHECK(merge_topn(stream, n_shards, n_q, k, m, n_truth_total, d_score, d_row, d_mx, d_rows, d_count, d_kth, d_thr, d_flags, nullptr));
    return ws.finish_outputs();
}

static int collect_flagged(cudaStream_t stream, const int32_t *d_flags, int64_t n_q, std::vector<int32_t> *flagged) {
    std::vector<int32_t> h_flags((size_t)n_q);
    DS_CUDA(cudaMemcpyAsync(h_flags.data(), d_flags, (size_t)n_q * 4, cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    for (int64_t q = 0; q < n_q; ++q)
        if (h_flags[(size_t)q] & DS_FLAG_RESCAN) flagged->push_back((int32_t)q);
    return DS_OK;
}

int ds_topn_rescan(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids, const double *q_mx,
                   const double *threshold, const int32_t *flags, int32_t k, int64_t *out_rows, int32_t *out_count,
                   void *stream_) {
    DS_CHECK(check_topn_args(index ? &index->ix : nullptr, n_q, q_row_ptr, q_col_ids, k));
    if (threshold == nullptr || flags == nullptr || out_rows == nullptr || out_count == nullptr)
        return fail(DS_ERR_BAD_ARG, "NULL argument");
    const Index &ix = index->ix;
    DeviceGuard guard(ix.device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    index->ix.last_stream = stream;
    if (n_q == 0) return DS_OK;
    Workspace ws(stream);
    const int32_t *d_flags = nullptr;
    DS_CHECK(ws.stage_in(&d_flags, flags, (size_t)n_q));
    std::vector<int32_t> flagged;
    if (is_device_pointer(flags)) {
        DS_CHECK(collect_flagged(stream, d_flags, n_q, &flagged));
    } else {
        for (int64_t q = 0; q < n_q; ++q)
            if (flags[q] & DS_FLAG_RESCAN) flagged.push_back((int32_t)q);
    }
    if (flagged.empty()) return DS_OK;
    if (q_mx == nullptr) return fail(DS_ERR_BAD_ARG, "ds_topn_rescan needs the q_mx returned by ds_topn_local");
    QuerySet qs;
    DS_CHECK(prepare_queries(ws, ix, n_q, q_row_ptr, q_col_ids, q_mx, 0, &qs));
    const double *d_thr = nullptr;
    DS_CHECK(ws.stage_in(&d_thr, threshold, (size_t)n_q));
    // outputs are patched in place: host destinations are staged in full first
    int64_t *d_rows = nullptr;
    int32_t *d_count = nullptr;
    if (is_device_pointer(out_rows)) {
        d_rows = out_rows;
    } else {
        DS_CHECK(ws.stage_out(&d_rows, out_rows, (size_t)n_q * k));
        DS_CUDA(cudaMemcpyAsync(d_rows, out_rows, (size_t)n_q * k * 8, cudaMemcpyHostToDevice, stream));
    }
    if (is_device_pointer(out_count)) {
        d_count = out_count;
    } else {
        DS_CHECK(ws.stage_out(&d_count, out_count, (size_t)n_q));
        DS_CUDA(cudaMemcpyAsync(d_count, out_count, (size_t)n_q * 4, cudaMemcpyHostToDevice, stream));
    }
    DS_CHECK(rescan_topn(ws, ix, qs, d_thr, flagged, k, d_rows, d_count));
    return ws.finish_outputs();
}

int ds_topn(ds_index *index, int64_t n_q, const int64_t *q_row_ptr, const uint16_t *q_col_ids, const double *q_mx,
            int32_t mx_mode, int32_t k, int64_t *out_rows, int32_t *out_count, float *out_kth_f32, int32_t *out_flags,
            void *stream_) {
    DS_CHECK(check_topn_args(index ? &index->ix : nullptr, n_q, q_row_ptr, q_col_ids, k));
    if (out_rows == nullptr || out_count == nullptr) return fail(DS_ERR_BAD_ARG, "out_rows / out_count is NULL");
    const Index &ix = index->ix;
    if (ix.row_offset != 0 || ix.n_total != ix.n_truth)
        return fail(DS_ERR_BAD_ARG, "ds_topn needs an unsharded index; use ds_topn_local / _merge / _rescan for shards");
    DeviceGuard guard(ix.device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    index->ix.last_stream = stream;
    if (n_q == 0) return DS_OK;
    Workspace ws(stream);
    QuerySet qs;
    DS_CHECK(prepare_queries(ws, ix, n_q, q_row_ptr, q_col_ids, q_mx, mx_mode, &qs));
    const int m = ds_topn_retained(k);
    double *d_score = nullptr, *d_thr = nullptr;
    int64_t *d_row = nullptr, *d_rows = nullptr;
    int32_t *d_count = nullptr, *d_flags = nullptr;
    float *d_kth = nullptr;
    int *d_flagged_count = nullptr;
    DS_CHECK(ws.alloc(&d_score, (size_t)n_q * m));
    DS_CHECK(ws.alloc(&d_row, (size_t)n_q * m));
    DS_CHECK(ws.alloc(&d_thr, (size_t)n_q));
    DS_CHECK(ws.alloc(&d_flagged_count, 1));
    DS_CHECK(ws.stage_out(&d_rows, out_rows, (size_t)n_q * k));
    DS_CHECK(ws.stage_out(&d_count, out_count, (size_t)n_q));
    DS_CHECK(ws.stage_out(&d_kth, out_kth_f32, (size_t)n_q));
    if (out_flags != nullptr) DS_CHECK(ws.stage_out(&d_flags, out_flags, (size_t)n_q));
    else DS_CHECK(ws.alloc(&d_flags, (size_t)n_q));
    DS_CUDA(cudaMemsetAsync(d_flagged_count, 0, 4, stream));
    DS_CHECK(local_topn(ws, ix, qs, k, m, d_score, d_row));
    DS_CHECK(merge_topn(stream, 1, n_q, k, m, ix.n_total, d_score, d_row, qs.d_mx, d_rows, d_count, d_kth, d_thr, d_flags, d_flagged_count));
    int h_flagged = 0;
    DS_CUDA(cudaMemcpyAsync(&h_flagged, d_flagged_count, 4, cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    if (h_flagged > 0) {
        std::vector<int32_t> flagged;
        DS_CHECK(collect_flagged(stream, d_flags, n_q, &flagged));
        DS_CHECK(rescan_topn(ws, ix, qs, d_thr, flagged, k, d_rows, d_count));
    }
    return ws.finish_outputs();
}

}  // extern "C"
