// ds_pairs.cu - K2 (batched InDel ratio) and K3 (construct_features) for candidate pairs.
//
// Reference semantics reproduced (paths relative to /root/reference):
//   fast_levenshtein_ratio   doppelspeller/feature_engineering.py:25-63
//       InDel DP (match 0, mismatch 2, gap 1) whose cells are stored as uint8 (wrap on store), result
//       uint8((L - d) * 100 / L) = (100 * (L - d)) / L as executed under numba fastmath.
//   construct_features       doppelspeller/feature_engineering.py:75-169
//   levenshtein_ratio        doppelspeller/common.py:161-162 (python-levenshtein ratio, true InDel
//       distance, int(round(ratio * 100)) with round-half-even on the float64)
//
// Design (B200): no DP matrices on the common path.  When la + lb <= 255 no uint8 cell can wrap and
// d = la + lb - 2 * LCS(a, b); the LCS comes from the Hyyro / Allison-Dix bit-vector recurrence
//     U = V & M[c];  V = (V + U) | (V & ~M[c])          (one machine word per 32 / 64 pattern characters)
// with the match masks M in shared memory.  Pairs are radix-sorted by (length class, la + lb) so the
// lanes of a warp run the same path for about the same number of steps; strings are fetched with aligned
// 32-bit loads, re-aligned with funnel shifts into a per-lane shared-memory slot and consumed four
// characters per LDS.  Only the uint8 wrap region (la + lb > 255) runs a DP - one pair per warp, swept by
// anti-diagonals.  construct_features = one warp per pair for the sliding word windows (lanes = window
// starts) + two passes of the batched ratio kernel for the whole-title ratios.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>

#include "ds_common.cuh"

namespace ds {

constexpr int PM_CODES = 40;       // alphabet '- a..z0..9' = codes 0..37 (feature_engineering.py:200)
constexpr int N_WORDS = DS_N_WORDS;
constexpr int RECON_STRIDE = 288;  // reconstruction <= 255 (windows) + 15 (spaces) + 15 (unmatched words) = 285
constexpr int MAX_LONG = 320;      // longest string any kernel sees (the reconstruction)

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------------
// pair addressing: each side is either padded rows [P, stride] with a length array, or a compact
// (bytes, offsets, index) table
// ---------------------------------------------------------------------------------------------------
struct Side {
    const uint8_t *base;
    int64_t stride;          // > 0: row p at base + p * stride, length len8[p] / len16[p]
    const uint8_t *len8;
    const uint16_t *len16;
    const int64_t *off;      // stride == 0: string idx[p] = [base + off[i], base + off[i + 1])
    const int32_t *idx;
};

struct PairSource {
    Side a, b;
};

__device__ __forceinline__ void load_side(const Side &s, int64_t p, const uint8_t **ptr, int *len, int64_t *id) {
    if (s.stride > 0) {
        *ptr = s.base + p * s.stride;
        *len = s.len16 ? (int)s.len16[p] : (int)s.len8[p];
        *id = p;
    } else {
        const int64_t i = s.idx[p];
        const int64_t o = s.off[i];
        *ptr = s.base + o;
        *len = (int)min((int64_t)DS_MAX_TITLE, s.off[i + 1] - o);
        *id = i;
    }
}

__device__ __forceinline__ int side_length(const Side &s, int64_t p) {
    if (s.stride > 0) return s.len16 ? (int)s.len16[p] : (int)s.len8[p];
    const int64_t i = s.idx[p];
    return (int)min((int64_t)DS_MAX_TITLE, s.off[i + 1] - s.off[i]);
}

// ASCII mode maps raw title bytes onto the 38 codes; anything else is "outside" (>= PM_CODES)
__device__ __forceinline__ int map_code(uint8_t c, int ascii) {
    if (!ascii) return c;
    if (c >= 'a' && c <= 'z') return c - 'a' + 2;
    if (c >= '0' && c <= '9') return c - '0' + 28;
    if (c == ' ') return 1;
    if (c == '-') return 0;
    return 255;
}

template <int MODE>
__device__ __forceinline__ int table_code(uint32_t byte, bool &good) {
    int c = MODE == 0 ? (int)byte : map_code((uint8_t)byte, 1);
    good &= c < PM_CODES;
    return min(c, PM_CODES - 1);
}

__device__ __forceinline__ int ratio_u8(int total, int d) {
    // uint8(((L - d) * 100) / L): integer exact (feature_engineering.py:63 under fastmath)
    return total == 0 ? 0 : ((100 * (total - d)) / total) & 0xff;
}

// Same value without an integer division: num = 100 * (L - d) <= 51,000 and L <= 510, so the float product
// num * (1 / L) is within 1.3e-5 of the true quotient, whose fractional part is 0 or >= 1 / 510; the 1e-4 bias
// therefore makes truncation exact.  Used where the ratio is evaluated per lane per round.
__device__ __forceinline__ int ratio_u8_fast(int total, int d) {
    const float q = fmaf((float)(100 * (total - d)), __frcp_rn((float)total), 1e-4f);
    return total == 0 ? 0 : (__float2int_rz(q) & 0xff);
}

// literal restatement of the uint8 DP (feature_engineering.py:42-61), one thread; `y` (length ly <= 255)
// indexes the row buffer, `x` the outer loop.  The recurrence is symmetric, so which string plays which
// role does not change any cell value (the reference puts the shorter one on the rows, :35-37).
// Only used for strings carrying bytes outside the 40-symbol table.
__device__ int indel_u8_dp(const uint8_t *x, int lx, const uint8_t *y, int ly) {
    uint8_t row[256];
    for (int j = 0; j <= ly; ++j) row[j] = (uint8_t)j;
    for (int i = 1; i <= lx; ++i) {
        int diag = row[0];
        row[0] = (uint8_t)i;
        const uint8_t xi = x[i - 1];
        for (int j = 1; j <= ly; ++j) {
            int up = row[j] + 1, left = row[j - 1] + 1;
            int d = diag + (xi == y[j - 1] ? 0 : 2);
            int v = min(min(up, left), d);
            diag = row[j];
            row[j] = (uint8_t)v;
        }
    }
    return row[ly];
}

// true (non-wrapping) InDel distance, same orientation rules
__device__ int indel_true_dp(const uint8_t *x, int lx, const uint8_t *y, int ly) {
    uint16_t row[256];
    for (int j = 0; j <= ly; ++j) row[j] = (uint16_t)j;
    for (int i = 1; i <= lx; ++i) {
        int diag = row[0];
        row[0] = (uint16_t)i;
        const uint8_t xi = x[i - 1];
        for (int j = 1; j <= ly; ++j) {
            int up = row[j] + 1, left = row[j - 1] + 1;
            int d = diag + (xi == y[j - 1] ? 0 : 2);
            int v = min(min(up, left), d);
            diag = row[j];
            row[j] = (uint16_t)v;
        }
    }
    return row[ly];
}

// ---------------------------------------------------------------------------------------------------
// K2
//   MODE 0: fast_levenshtein_ratio (uint8 wrap semantics)   MODE 1: common.levenshtein_ratio on raw bytes
//   class 0  shorter string <= 32, longer <= 64   one pair per lane, one 32-bit word
//   class 1  both <= 64                           one pair per lane, one 64-bit word
//   class 2  bit-vector path for longer strings (MODE 0: la + lb <= 255, MODE 1: all <= 255): pattern in
//            64-character blocks, the carry out of every text step replayed into the next block
//   class 3  MODE 0, la + lb > 255 (uint8 cells can wrap): one pair per WARP, literal uint8 DP swept by
//            anti-diagonals over 32-column strips
// ---------------------------------------------------------------------------------------------------
struct K2Out {
    uint8_t *u8;        // MODE 0: ratio
    uint16_t *dist;     // MODE 0: wrapped distance (nullable)
    int32_t *i32;       // MODE 1: levenshtein_ratio
    float *feat;        // MODE 0: ratio as float at feat[p * feat_stride + feat_col] (construct_features)
    int feat_stride, feat_col;
};

template <int MODE>
__device__ __forceinline__ void store_result(const K2Out &out, int total, int d, int64_t p) {
    if (MODE == 0) {
        const int r = ratio_u8(total, d);
        if (out.u8) out.u8[p] = (uint8_t)r;
        if (out.dist) out.dist[p] = (uint16_t)d;
        if (out.feat) out.feat[p * out.feat_stride + out.feat_col] = (float)r;
    } else {
        int result = 100;  // ratio 1.0 for two empty strings
        if (total > 0) {
            // int(round(ratio * 100)): float64 divide, float64 multiply, round half to even
            const double ratio = __ddiv_rn((double)(total - d), (double)total);
            result = (int)rint(__dmul_rn(ratio, 100.0));
        }
        out.i32[p] = result;
    }
}

template <typename W, int MAXLEN, int BLOCK>
struct K2Smem {
    static constexpr int STAGE_WORDS = ((MAXLEN + 3) / 4) | 1;   // words per staged string, odd lane stride
    W pm[PM_CODES * BLOCK];
    uint32_t stage_a[STAGE_WORDS * BLOCK];
    uint32_t stage_b[STAGE_WORDS * BLOCK];
};

// Copies [g, g + len) into slot[0 .. ceil(len / 4)) so that byte i of the string is byte i of the slot.
// Only aligned words holding at least one valid byte are read (never leaves the page of a valid byte).
__device__ __forceinline__ void stage_string(const uint8_t *g, int len, uint32_t *slot) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(g);
    const int shift = (int)(addr & 3);
    const uint32_t *g32 = reinterpret_cast<const uint32_t *>(addr - shift);
    const int words_out = (len + 3) >> 2;
    const int words_in = (shift + len + 3) >> 2;
    if (words_out == 0) return;
    uint32_t cur = __ldg(g32);
    for (int k = 0; k < words_out; ++k) {
        uint32_t next = (k + 1 < words_in) ? __ldg(g32 + k + 1) : 0u;
        slot[k] = __funnelshift_r(cur, next, shift * 8);
        cur = next;
    }
}

// one-word LCS: pattern m <= bits(W) staged in `pat`, text n staged in `txt` (word-aligned slots)
template <typename W, int BLOCK, int MODE>
__device__ __forceinline__ int lcs_one_word(W *pm, const uint32_t *pat, int m, const uint32_t *txt, int n, bool *ok) {
    bool good = true;
    for (int i0 = 0; i0 < m; i0 += 4) {
        const uint32_t w = pat[i0 >> 2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < m) {
                const int c = table_code<MODE>((w >> (8 * k)) & 0xffu, good);
                pm[c * BLOCK] |= (W)1 << (i0 + k);
            }
        }
    }
    W v = ~(W)0;
    for (int j0 = 0; j0 < n; j0 += 4) {
        const uint32_t w = txt[j0 >> 2];
        W mm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool in_table = true;   // bytes past the end of the text are whatever follows it in memory: ignored
            mm[k] = pm[table_code<MODE>((w >> (8 * k)) & 0xffu, in_table) * BLOCK];
            good &= in_table | (j0 + k >= n);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (j0 + k < n) {
                const W u = v & mm[k];
                v = (v + u) | (v & ~mm[k]);
            }
        }
    }
    *ok = good;
    const W valid = (m == (int)(8 * sizeof(W))) ? ~(W)0 : (((W)1 << m) - 1);
    return sizeof(W) == 8 ? __popcll((u64)(~v & valid)) : __popc((uint32_t)(~v & valid));
}

// general LCS: pattern m <= 255 in up to four 64-character blocks, text n <= 255
template <int BLOCK, int MODE>
__device__ int lcs_blocks(u64 *pm, const uint32_t *pat, int m, const uint32_t *txt, int n, bool *ok) {
    bool good = true;
    int lcs = 0;
    u64 carry[4] = {0, 0, 0, 0};
    for (int w0 = 0; w0 < m; w0 += 64) {
        const int mw = min(64, m - w0);
        for (int i0 = 0; i0 < mw; i0 += 4) {
            const uint32_t w = pat[(w0 + i0) >> 2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + k < mw) {
                    const int c = table_code<MODE>((w >> (8 * k)) & 0xffu, good);
                    pm[c * BLOCK] |= 1ull << (i0 + k);
                }
            }
        }
        u64 v = ~0ull;
        u64 out[4] = {0, 0, 0, 0};
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) {
            const u64 cw = carry[seg];
            u64 ow = 0;
            const int j_end = min(n, seg * 64 + 64);
            for (int j0 = seg * 64; j0 < j_end; j0 += 4) {
                const uint32_t w = txt[j0 >> 2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (j0 + k < j_end) {
                        const u64 mm = pm[table_code<MODE>((w >> (8 * k)) & 0xffu, good) * BLOCK];
                        const u64 u = v & mm;
                        const u64 cin = (cw >> ((j0 + k) & 63)) & 1ull;
                        const u64 s1 = v + u;
                        const u64 s2 = s1 + cin;
                        ow |= (u64)((s1 < v) | (s2 < s1)) << ((j0 + k) & 63);
                        v = s2 | (v & ~mm);
                    }
                }
            }
            out[seg] = ow;
        }
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) carry[seg] = out[seg];
        const u64 valid = (mw == 64) ? ~0ull : ((1ull << mw) - 1);
        lcs += __popcll(~v & valid);
        if (w0 + 64 < m) {   // the next block reuses the column
            for (int i0 = 0; i0 < mw; i0 += 4) {
                const uint32_t w = pat[(w0 + i0) >> 2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    bool ignored = true;
                    if (i0 + k < mw) pm[table_code<MODE>((w >> (8 * k)) & 0xffu, ignored) * BLOCK] = 0;
                }
            }
        }
    }
    *ok = good;
    return lcs;
}

__device__ __forceinline__ int pair_class(int la, int lb, int mode) {
    const int lo = min(la, lb), hi = max(la, lb);
    if (lo <= 32 && hi <= 64) return 0;
    if (hi <= 64) return 1;
    if (hi <= 255 && (mode == 1 || la + lb <= 255)) return 2;
    return 3;
}

// sort key = (class << 10) | (la + lb); values = pair ids; counts per class.  `subset` (nullable): only these pairs.
__global__ void k_pair_keys(PairSource src, const int32_t *__restrict__ subset, int64_t n, int mode, uint16_t *__restrict__ keys,
                            int32_t *__restrict__ ids, int *__restrict__ counts) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1;
    if (slot < n) {
        const int64_t p = subset ? (int64_t)subset[slot] : slot;
        const int la = side_length(src.a, p), lb = side_length(src.b, p);
        cls = pair_class(la, lb, mode);
        keys[slot] = (uint16_t)((cls << 10) | min(la + lb, 1023));
        ids[slot] = (int32_t)p;
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned ballot = __ballot_sync(0xffffffffu, cls == c);
        if (ballot != 0 && lane == __ffs(ballot) - 1) atomicAdd(counts + c, __popc(ballot));
    }
}

// classes 0..2: one pair per lane.  W/MAXLEN select the path: <u32,64> class 0, <u64,64> class 1, <u64,255> class 2
template <typename W, int MAXLEN, int BLOCK, int MODE>
__global__ void __launch_bounds__(BLOCK) k_indel_pairs(PairSource src, const int32_t *__restrict__ pair_list, int64_t n, K2Out out,
                                                       const int *__restrict__ n_device = nullptr) {
    if (n_device != nullptr) n = min(n, (int64_t)*n_device);   // list filled on the device: its length never visits the host
    if ((int64_t)blockIdx.x * BLOCK >= n) return;
    extern __shared__ __align__(16) unsigned char k2_raw[];
    typedef K2Smem<W, MAXLEN, BLOCK> Smem;
    Smem &sm = *reinterpret_cast<Smem *>(k2_raw);
    {
        uint4 *z = reinterpret_cast<uint4 *>(sm.pm);
        for (int i = threadIdx.x; i < (int)(sizeof(sm.pm) / 16); i += BLOCK) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int64_t slot = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (slot >= n) return;
    const int64_t p = pair_list ? (int64_t)pair_list[slot] : slot;
    const uint8_t *a, *b;
    int la, lb;
    int64_t unused;
    load_side(src.a, p, &a, &la, &unused);
    load_side(src.b, p, &b, &lb, &unused);
    W *my_pm = sm.pm + threadIdx.x;
    const int total = la + lb;
    uint32_t *sa = sm.stage_a + threadIdx.x * Smem::STAGE_WORDS;
    uint32_t *sb = sm.stage_b + threadIdx.x * Smem::STAGE_WORDS;
    stage_string(a, la, sa);
    stage_string(b, lb, sb);
    bool ok;
    int lcs;
    const uint32_t *pat = (la <= lb) ? sa : sb, *txt = (la <= lb) ? sb : sa;
    const int m = min(la, lb), nn = max(la, lb);
    if (MAXLEN <= 64) lcs = lcs_one_word<W, BLOCK, MODE>(my_pm, pat, m, txt, nn, &ok);
    else lcs = lcs_blocks<BLOCK, MODE>(reinterpret_cast<u64 *>(my_pm), pat, m, txt, nn, &ok);
    int d = total - 2 * lcs;
    if (!ok) {   // bytes outside the 40-symbol table: literal DP on the original strings
        if (MODE == 0) d = (lb <= 255) ? indel_u8_dp(a, la, b, lb) : indel_u8_dp(b, lb, a, la);
        else d = indel_true_dp(a, la, b, lb);
    }
    store_result<MODE>(out, total, d, p);
}

// ---------------------------------------------------------------------------------------------------
// K2 for candidate lists: runs of `run` consecutive pairs that share their first title (the top_n candidates of one
// test title, predict.py:129-136; uniform run length: a [Q, top_n] list).  A group of 16 lanes (32 for runs longer
// than 16) takes one run: the match-mask table of the shared title is built ONCE per run in shared memory by the
// group (pattern = that title; the LCS is symmetric), every lane then streams its own second title through registers
// as aligned 32-bit words.  Nothing is sorted, no pair ids are gathered.  Pairs this path does not take (a run whose
// pairs do not all share the title, a first title longer than 64 characters, bytes outside the 40-symbol table, MODE 0
// with la + lb > 255: the uint8 wrap region) go to `rest` and from there through the sorted class kernels.
// ---------------------------------------------------------------------------------------------------
constexpr int GROUP_BLOCK = 256;

// counts the positions whose first title differs from the previous pair's (run structure of a pair list)
__global__ void k_pair_runs(Side a, int64_t n, unsigned long long *__restrict__ changes) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool change = false;
    if (p < n) {
        if (a.stride > 0) change = true;   // padded rows: every pair has its own copy of the title
        else change = p == 0 || a.idx[p] != a.idx[p - 1];
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, change);
    if ((threadIdx.x & 31) == 0 && ballot != 0) atomicAdd(changes, (unsigned long long)__popc(ballot));
}

template <int MODE>
__global__ void __launch_bounds__(GROUP_BLOCK) k_indel_groups(PairSource src, int64_t n, int run, int group_lanes, K2Out out,
                                                               int32_t *__restrict__ rest, int32_t *__restrict__ rest_wrap,
                                                               int *__restrict__ rest_count) {
    // group_lanes = min(run, 32) consecutive lanes take one run; a warp holds 32 / group_lanes runs (30 of 32 lanes busy at
    // top_n = 10); runs longer than 32 pairs go through their group in rounds
    constexpr int WARPS = GROUP_BLOCK / 32;
    __shared__ __align__(8) uint32_t pm[WARPS * 8][PM_CODES][2];   // per group (>= 4 lanes: <= 8 per warp): [code][low / high 32 positions]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int GROUP = group_lanes;
    const int per_warp = 32 / GROUP;
    const int group_in_warp = lane / GROUP;
    const int sub = lane - group_in_warp * GROUP;   // lane inside the group
    if (group_in_warp >= per_warp) return;          // the warp's spare lanes
    const unsigned group_mask = (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u)) << (group_in_warp * GROUP);
    const int64_t r = ((int64_t)blockIdx.x * WARPS + warp) * per_warp + group_in_warp;
    const int64_t first = r * run;
    if (first >= n) return;                         // whole groups leave together
    uint32_t(*my_pm)[2] = pm[warp * 8 + group_in_warp];
    for (int c = sub; c < PM_CODES; c += GROUP) {
        my_pm[c][0] = 0;
        my_pm[c][1] = 0;
    }
    const uint8_t *pa = nullptr;
    int la = 0;
    int64_t id_a = -1;
    load_side(src.a, first, &pa, &la, &id_a);
    __syncwarp(group_mask);
    bool good = la <= 64;
    for (int i = sub; i < la && i < 64; i += GROUP) {
        const int c = table_code<MODE>(pa[i], good);
        atomicOr(&my_pm[c][i >> 5], 1u << (i & 31));
    }
    __syncwarp(group_mask);
    const bool table_ok = __all_sync(group_mask, good);
    const bool narrow = la <= 32;                   // uniform inside the group
    for (int j0 = 0; j0 < run; j0 += GROUP) {
        const int64_t p = first + j0 + sub;
        const bool active = j0 + sub < run && p < n;
        const uint8_t *pb = nullptr;
        int lb = 0;
        bool here = false;
        if (active) {
            int64_t id_b = -1;
            load_side(src.b, p, &pb, &lb, &id_b);
            const bool same_title = src.a.stride > 0 ? false : src.a.idx[p] == (int32_t)id_a;
            here = table_ok && same_title && lb <= 255 && (MODE == 1 || la + lb <= 255);
        }
        bool ok = here;
        int lcs = 0;
        if (here) {
            const uintptr_t addr = reinterpret_cast<uintptr_t>(pb);
            const int shift = (int)(addr & 3);
            const uint32_t *g32 = reinterpret_cast<const uint32_t *>(addr - shift);
            const int words_in = (shift + lb + 3) >> 2;
            uint32_t cur = words_in > 0 ? __ldg(g32) : 0u;
            uint32_t next = words_in > 1 ? __ldg(g32 + 1) : 0u;
            if (narrow) {
                uint32_t v = ~0u;
                for (int t0 = 0, k = 0; t0 < lb; t0 += 4, ++k) {
                    const uint32_t after = (k + 2 < words_in) ? __ldg(g32 + k + 2) : 0u;
                    const uint32_t w = __funnelshift_r(cur, next, shift * 8);
                    cur = next;
                    next = after;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (t0 + t < lb) {
                            bool in_table = true;
                            const uint32_t mm = my_pm[table_code<MODE>((w >> (8 * t)) & 0xffu, in_table)][0];
                            ok &= in_table;
                            const uint32_t u = v & mm;
                            v = (v + u) | (v & ~mm);
                        }
                    }
                }
                const uint32_t valid = (la >= 32) ? ~0u : ((1u << la) - 1);
                lcs = __popc(~v & valid);
            } else {
                u64 v = ~0ull;
                for (int t0 = 0, k = 0; t0 < lb; t0 += 4, ++k) {
                    const uint32_t after = (k + 2 < words_in) ? __ldg(g32 + k + 2) : 0u;
                    const uint32_t w = __funnelshift_r(cur, next, shift * 8);
                    cur = next;
                    next = after;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (t0 + t < lb) {
                            bool in_table = true;
                            const uint2 m2 = *reinterpret_cast<const uint2 *>(my_pm[table_code<MODE>((w >> (8 * t)) & 0xffu, in_table)]);
                            ok &= in_table;
                            const u64 mm = ((u64)m2.y << 32) | m2.x;
                            const u64 u = v & mm;
                            v = (v + u) | (v & ~mm);
                        }
                    }
                }
                const u64 valid = (la >= 64) ? ~0ull : ((1ull << la) - 1);
                lcs = __popcll(~v & valid);
            }
        }
        if (ok) {
            store_result<MODE>(out, la + lb, la + lb - 2 * lcs, p);
        } else if (active) {
            // the pair's own first title decides where it goes on (it may differ from the group's)
            const int my_la = side_length(src.a, p);
            if (MODE == 0 && (my_la + lb > 255 || max(my_la, lb) > 255)) rest_wrap[atomicAdd(rest_count + 1, 1)] = (int32_t)p;
            else rest[atomicAdd(rest_count, 1)] = (int32_t)p;
        }
    }
}

// class 3 (MODE 0, la + lb > 255): the literal uint8 DP of feature_engineering.py:42-61, one pair per warp.
// Columns are processed in strips of 32 (lane = column); inside a strip the cells of an anti-diagonal are
// independent: at step s lane l owns row s - l + 1; `up` is its own previous cell, `left` / `diag` come from
// lane l - 1 (shuffle) or, for the strip's first column, from the previous strip's last column kept in
// shared memory.  Every cell is reduced mod 256 on store exactly like the uint8 matrix.
struct WrapSmem {
    uint8_t x[MAX_LONG];
    uint8_t y[256];
    uint8_t edge[2][MAX_LONG + 8];
};

__global__ void __launch_bounds__(128) k_indel_wrap(PairSource src, const int32_t *__restrict__ pair_list, int64_t n, K2Out out,
                                                    const int *__restrict__ n_device = nullptr) {
    if (n_device != nullptr) n = min(n, (int64_t)*n_device);
    __shared__ WrapSmem smem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WrapSmem &sm = smem[warp];
    for (int64_t slot = (int64_t)blockIdx.x * 4 + warp; slot < n; slot += (int64_t)gridDim.x * 4) {
    const int64_t p = pair_list ? (int64_t)pair_list[slot] : slot;
    const uint8_t *a, *b;
    int la, lb;
    int64_t unused;
    load_side(src.a, p, &a, &la, &unused);
    load_side(src.b, p, &b, &lb, &unused);
    // x = rows (outer sweep), y = columns (strips): the shorter string on the columns means fewer strips
    const uint8_t *gx = (la >= lb) ? a : b, *gy = (la >= lb) ? b : a;
    const int lx = min(max(la, lb), MAX_LONG), ly = min(min(la, lb), 255);
    for (int i = lane; i < lx; i += 32) sm.x[i] = gx[i];
    for (int i = lane; i < ly; i += 32) sm.y[i] = gy[i];
    for (int i = lane; i <= lx; i += 32) sm.edge[0][i] = (uint8_t)i;   // column 0: D[i][0] = i (uint8)
    __syncwarp();
    int result = lx & 0xff;   // ly == 0
    int cur_edge = 0;
    for (int c0 = 1; c0 <= ly; c0 += 32) {
        const int j = c0 + lane;                       // this lane's column (1-based)
        const bool active_col = j <= ly;
        const uint8_t yj = active_col ? sm.y[j - 1] : 0;
        int cur = j & 0xff;                            // D[0][j]
        int prev_left = (j - 1) & 0xff;                // D[0][j-1]: the diagonal of row 1
        const uint8_t *left_col = sm.edge[cur_edge];
        uint8_t *next_col = sm.edge[cur_edge ^ 1];
        const int last_lane = min(31, ly - c0);
        if (lane == last_lane) next_col[0] = (uint8_t)cur;
        for (int s = 0; s < lx + 31; ++s) {
            const int i = s - lane + 1;                // row of this lane at this step
            int left = __shfl_up_sync(0xffffffffu, cur, 1);
            if (lane == 0) left = (i >= 1 && i <= lx) ? left_col[i] : 0;
            if (i >= 1 && i <= lx && active_col) {
                const int diag = prev_left + (sm.x[i - 1] == yj ? 0 : 2);
                const int v = min(min(cur + 1, left + 1), diag) & 0xff;
                prev_left = left;
                cur = v;
                if (lane == last_lane) next_col[i] = (uint8_t)v;
            }
        }
        __syncwarp();
        result = __shfl_sync(0xffffffffu, cur, last_lane);
        cur_edge ^= 1;
    }
    if (lane == 0) store_result<0>(out, la + lb, result, p);
    __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// K3 word features: one warp per pair (grid-stride).  For every truth word (first 15) the sliding windows
// over the space-less title are spread over the lanes with one shared match-mask table; first strict
// maximum wins (feature_engineering.py:139-149).  Writes features 0..3 and 6..65 and the reconstructed
// title (feature 5's input) to scratch; features 4 and 5 come from two K2 passes.
// ---------------------------------------------------------------------------------------------------
constexpr int K3_WARPS = 8;

struct K3Smem {
    uint32_t pm_lo[64];       // 40 codes used; 64 entries so that stale bytes read past a window stay in range
    uint32_t pm_hi[64];
    uint8_t a_ns[256 + 32];   // + 32: the uniform window loop may read (not use) up to 31 bytes past the end
    uint8_t b_cur[256];
    uint8_t recon[RECON_STRIDE];
};

__global__ void __launch_bounds__(K3_WARPS * 32) k_feature_words(PairSource src, const uint32_t *__restrict__ counts, int space_code,
                                                                 uint32_t n_truth, int64_t n_pairs, float *__restrict__ out,
                                                                 uint8_t *__restrict__ recon_out, uint16_t *__restrict__ recon_len) {
    __shared__ K3Smem smem[K3_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K3Smem &sm = smem[warp];
    for (int i = lane; i < 64; i += 32) {
        sm.pm_lo[i] = 0;
        sm.pm_hi[i] = 0;
    }
    for (int i = lane; i < 256 + 32; i += 32) sm.a_ns[i] = 0;
    __syncwarp();
    const float nan_f = __int_as_float(0x7fc00000);
    const int64_t warps_total = (int64_t)gridDim.x * K3_WARPS;
    for (int64_t p = (int64_t)blockIdx.x * K3_WARPS + warp; p < n_pairs; p += warps_total) {
        const uint8_t *pa, *pb;
        int la, lb;
        int64_t unused, truth_id;
        load_side(src.a, p, &pa, &la, &unused);
        load_side(src.b, p, &pb, &lb, &truth_id);
        // stage the truth title and the space-less title; count words; check the code range
        int n_ns = 0, spaces_a = 0, spaces_b = 0;
        bool ok = space_code < PM_CODES;
        for (int i0 = 0; i0 < max(la, lb); i0 += 32) {
            const int i = i0 + lane;
            uint8_t ca = 0, cb = 0;
            const bool in_a = i < la, in_b = i < lb;
            if (in_a) ca = pa[i];
            if (in_b) {
                cb = pb[i];
                sm.b_cur[i] = cb;
            }
            const bool keep = in_a && ca != space_code;
            const unsigned keep_mask = __ballot_sync(0xffffffffu, keep);
            spaces_a += __popc(__ballot_sync(0xffffffffu, in_a && ca == space_code));
            spaces_b += __popc(__ballot_sync(0xffffffffu, in_b && cb == space_code));
            ok &= __all_sync(0xffffffffu, (!in_a || ca < PM_CODES) && (!in_b || cb < PM_CODES)) != 0;
            if (keep) sm.a_ns[n_ns + __popc(keep_mask & ((1u << lane) - 1))] = ca;
            n_ns += __popc(keep_mask);
        }
        __syncwarp();
        const int words_b = spaces_b + 1;

        float my_best = nan_f, my_wlen = nan_f, my_idf = nan_f;  // lane w < 15 owns truth word w
        int n_words = 0, last = 0, n_recon = 0;
        for (int base = 0; base <= lb && n_words < N_WORDS; base += 32) {
            const int pos_l = base + lane;
            const bool sep = pos_l <= lb && (pos_l == lb || sm.b_cur[pos_l] == space_code);
            unsigned sepmask = __ballot_sync(0xffffffffu, sep);
            while (sepmask != 0 && n_words < N_WORDS) {
                const int pos = base + __ffs(sepmask) - 1;
                sepmask &= sepmask - 1;
                const uint8_t *word = sm.b_cur + last;
                const int wl = pos - last;
                last = pos + 1;
                int best_ratio = 0, best_start = -1, best_len = 1;
                if (wl > 0 && n_ns > 0) {
                    const bool fast = ok && wl <= 64;
                    if (fast) {
                        for (int i = lane; i < wl; i += 32) {
                            if (i < 32) atomicOr(&sm.pm_lo[word[i]], 1u << i);
                            else atomicOr(&sm.pm_hi[word[i]], 1u << (i - 32));
                        }
                        __syncwarp();
                    }
                    for (int i0 = 0; i0 < n_ns; i0 += 32) {
                        const int i = i0 + lane;
                        int key = -1;
                        if (i < n_ns) {
                            const int pl = min(wl, n_ns - i);
                            const uint8_t *win = sm.a_ns + i;
                            int d;
                            if (fast && wl <= 32) {
                                // uniform trip count (wl): tail windows stop updating once their characters run out;
                                // reading a_ns a few bytes past n_ns stays inside the 256-byte buffer (i + wl <= 255 + 32)
                                uint32_t v = ~0u;
#pragma unroll 4
                                for (int j = 0; j < wl; ++j) {
                                    const uint32_t mm = sm.pm_lo[win[j] & 63];
                                    const uint32_t u = v & mm;
                                    const uint32_t nv = (v + u) | (v & ~mm);
                                    v = (j < pl) ? nv : v;
                                }
                                const uint32_t valid = (wl == 32) ? ~0u : ((1u << wl) - 1);
                                d = pl + wl - 2 * __popc(~v & valid);
                            } else if (fast) {
                                u64 v = ~0ull;
                                for (int j = 0; j < pl; ++j) {
                                    const uint8_t c = win[j];
                                    const u64 mm = ((u64)sm.pm_hi[c] << 32) | sm.pm_lo[c];
                                    const u64 u = v & mm;
                                    v = (v + u) | (v & ~mm);
                                }
                                const u64 valid = (wl == 64) ? ~0ull : ((1ull << wl) - 1);
                                d = pl + wl - 2 * __popcll(~v & valid);
                            } else {
                                d = indel_u8_dp(word, wl, win, pl);   // words > 64 characters / bytes outside the table
                            }
                            const int r = ratio_u8_fast(pl + wl, d);
                            key = (r << 16) | (0xffff - i);  // max key = highest ratio, then lowest start
                        }
                        const int round_best = __reduce_max_sync(0xffffffffu, key);
                        const int r_best = round_best >> 16;
                        if (round_best >= 0 && r_best > best_ratio) {  // strict: the first maximum wins
                            best_ratio = r_best;
                            best_start = 0xffff - (round_best & 0xffff);
                            best_len = min(wl, n_ns - best_start);
                        }
                    }
                    if (fast) {
                        __syncwarp();
                        for (int i = lane; i < wl; i += 32) {
                            if (i < 32) sm.pm_lo[word[i]] = 0;
                            else sm.pm_hi[word[i]] = 0;
                        }
                        __syncwarp();
                    }
                }
                if (lane == n_words) {
                    my_best = (float)best_ratio;
                    my_wlen = (float)wl;
                }
                // reconstructed title: best window (or a single space) followed by a space (:154-155)
                if (best_start < 0) {
                    if (lane == 0) sm.recon[n_recon] = (uint8_t)space_code;
                } else {
                    for (int i = lane; i < best_len; i += 32) sm.recon[n_recon + i] = sm.a_ns[best_start + i];
                }
                n_recon += best_len;
                if (lane == 0) sm.recon[n_recon] = (uint8_t)space_code;
                n_recon += 1;
                ++n_words;
            }
        }
        if (lane < n_words) {   // idf of every found word at once (:153): log(N / count), count 0 -> +inf
            const uint32_t cnt = counts[truth_id * N_WORDS + lane];
            my_idf = (float)log(__ddiv_rn((double)n_truth, (double)cnt));
        }
        const int recon_n = n_recon > 0 ? n_recon - 1 : 0;   // drop the trailing space (:161)
        // IDF ranks (:158): NaN for every slot unless all 15 word slots are filled
        float rank = nan_f;
        if (n_words == N_WORDS) {
            float mxv = (lane < N_WORDS) ? my_idf : -INFINITY;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mxv = fmaxf(mxv, __shfl_xor_sync(0xffffffffu, mxv, d));
            const float diff = __fsub_rn(mxv, my_idf);
            rank = (float)__dadd_rn(1.0, __ddiv_rn((double)diff, (double)words_b));
        }
        float *o = out + p * DS_N_FEATURES;
        if (lane < N_WORDS) {
            o[6 + lane] = my_best;
            o[6 + N_WORDS + lane] = my_wlen;
            o[6 + 2 * N_WORDS + lane] = my_idf;
            o[6 + 3 * N_WORDS + lane] = rank;
        }
        if (lane == 0) {
            o[0] = (float)la;
            o[1] = (float)lb;
            o[2] = (float)(spaces_a + 1);
            o[3] = (float)words_b;
            recon_len[p] = (uint16_t)recon_n;
        }
        __syncwarp();
        uint8_t *ro = recon_out + p * RECON_STRIDE;
        for (int i = lane; i < recon_n; i += 32) ro[i] = sm.recon[i];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// f2: the fuzzy pre-match of Prediction (predict.py:140-183) on the device.
//   k_prematch_filter   _get_levenshtein_deletion_ratio (:140-145) in float64, the written association: pairs below the
//                       threshold get 0 (:150-151), the others are listed for the ratio kernel
//   k_prematch_again    pairs whose levenshtein_ratio is <= threshold are listed for the token-sorted ratio (:154-155)
//   k_select_close      per test title: the candidate with ratio > threshold that alone attains the title's maximum
//                       (:172-176 keep `> 94`, groupby-max, drop titles whose maximum is attained twice)
// The list lengths stay on the device (the ratio kernels read them there): no host synchronisation in the cascade.
// ---------------------------------------------------------------------------------------------------
__global__ void k_prematch_filter(PairSource src, int64_t n, double threshold, int32_t *__restrict__ out, int32_t *__restrict__ list,
                                  int *__restrict__ count) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (p < n) {
        const int la = side_length(src.a, p), lb = side_length(src.b, p);
        const double total = (double)(la + lb), delta = (double)abs(la - lb);
        const double deletion = __dmul_rn(__ddiv_rn(__dsub_rn(total, delta), total), 100.0);   // 0 / 0 = NaN: not < threshold
        keep = !(deletion < threshold);
        if (!keep) out[p] = 0;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ballot != 0) base = atomicAdd(count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) list[base + __popc(ballot & ((1u << lane) - 1))] = (int32_t)p;
}

__global__ void k_prematch_again(const int32_t *__restrict__ list, const int *__restrict__ count, const int32_t *__restrict__ out,
                                 int32_t threshold, int32_t *__restrict__ list2, int *__restrict__ count2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool again = false;
    int32_t p = 0;
    if (i < *count) {
        p = list[i];
        again = out[p] <= threshold;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, again);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ballot != 0) base = atomicAdd(count2, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (again) list2[base + __popc(ballot & ((1u << lane) - 1))] = p;
}

// one thread per test title over its `run` consecutive pairs; invalid[p] != 0 (nullable) excludes a pair
__global__ void k_select_close(const int32_t *__restrict__ ratios, const uint8_t *__restrict__ invalid, int64_t n_titles, int run,
                               int32_t threshold, int64_t *__restrict__ out_pair) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_titles) return;
    int32_t best = threshold;
    int64_t at = -1;
    int times = 0;
    for (int j = 0; j < run; ++j) {
        const int64_t p = q * run + j;
        if (invalid != nullptr && invalid[p]) continue;
        const int32_t r = ratios[p];
        if (r > best) {
            best = r;
            at = p;
            times = 1;
        } else if (r == best && at >= 0) {
            ++times;
        }
    }
    out_pair[q] = times == 1 ? at : -1;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <typename W, int MAXLEN, int BLOCK, int MODE>
static int launch_indel_class(const PairSource &src, const int32_t *list, int64_t n, const K2Out &out, cudaStream_t stream) {
    if (n <= 0) return DS_OK;
    const size_t smem = sizeof(K2Smem<W, MAXLEN, BLOCK>);
    DS_CHECK((ensure_dynamic_smem(reinterpret_cast<const void *>(&k_indel_pairs<W, MAXLEN, BLOCK, MODE>), smem)));
    k_indel_pairs<W, MAXLEN, BLOCK, MODE><<<(unsigned)ceil_div(n, BLOCK), BLOCK, smem, stream>>>(src, list, n, out);
    DS_LAUNCHED("k_indel_pairs");
    return DS_OK;
}

// the sorted class pipeline over all n pairs (subset == nullptr) or over the listed pair ids
template <int MODE>
static int launch_indel_sorted(Workspace &ws, const PairSource &src, const int32_t *subset, int64_t n, const K2Out &out) {
    cudaStream_t stream = ws.stream();
    if (n <= 0) return DS_OK;
    if (n > INT32_MAX) return fail(DS_ERR_UNSUPPORTED, "more than 2^31-1 pairs per call");
    uint16_t *d_keys = nullptr, *d_keys_sorted = nullptr;
    int32_t *d_ids = nullptr, *d_ids_sorted = nullptr;
    int *d_counts = nullptr;
    DS_CHECK(ws.alloc(&d_keys, (size_t)n));
    DS_CHECK(ws.alloc(&d_keys_sorted, (size_t)n));
    DS_CHECK(ws.alloc(&d_ids, (size_t)n));
    DS_CHECK(ws.alloc(&d_ids_sorted, (size_t)n));
    DS_CHECK(ws.alloc(&d_counts, 4));
    DS_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int), stream));
    k_pair_keys<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(src, subset, n, MODE, d_keys, d_ids, d_counts);
    DS_LAUNCHED("k_pair_keys");
    size_t temp_bytes = 0;
    DS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys, d_keys_sorted, d_ids, d_ids_sorted, (int)n, 0, 12, stream));
    unsigned char *d_temp = nullptr;
    DS_CHECK(ws.alloc(&d_temp, temp_bytes));
    DS_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys_sorted, d_ids, d_ids_sorted, (int)n, 0, 12, stream));
    g_kernel_launches.fetch_add(3);
    int h[4] = {0, 0, 0, 0};
    DS_CUDA(cudaMemcpyAsync(h, d_counts, sizeof(h), cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    const int32_t *list = d_ids_sorted;
    DS_CHECK((launch_indel_class<uint32_t, 64, 128, MODE>(src, list, h[0], out, stream)));
    DS_CHECK((launch_indel_class<u64, 64, 128, MODE>(src, list + h[0], h[1], out, stream)));
    DS_CHECK((launch_indel_class<u64, 255, 64, MODE>(src, list + h[0] + h[1], h[2], out, stream)));
    if (h[3] > 0) {
        if (MODE != 0) return fail(DS_ERR_UNSUPPORTED, "strings longer than 255 bytes");
        k_indel_wrap<<<(unsigned)std::min<int64_t>(ceil_div(h[3], 4), 148 * 16), 128, 0, stream>>>(src, list + h[0] + h[1] + h[2], h[3], out);
        DS_LAUNCHED("k_indel_wrap");
    }
    return DS_OK;
}

// Candidate lists (uniform runs of pairs sharing their first title) take the grouped kernel; whatever it leaves and
// every other pair list takes the sorted class pipeline.
template <int MODE>
static int launch_indel_mode(Workspace &ws, const PairSource &src, int64_t n, const K2Out &out) {
    cudaStream_t stream = ws.stream();
    if (n <= 0) return DS_OK;
    if (n > INT32_MAX) return fail(DS_ERR_UNSUPPORTED, "more than 2^31-1 pairs per call");
    if (src.a.stride > 0 || n < 1024) return launch_indel_sorted<MODE>(ws, src, nullptr, n, out);
    unsigned long long *d_changes = nullptr;
    DS_CHECK(ws.alloc(&d_changes, 1));
    DS_CUDA(cudaMemsetAsync(d_changes, 0, 8, stream));
    k_pair_runs<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(src.a, n, d_changes);
    DS_LAUNCHED("k_pair_runs");
    unsigned long long h_changes = 0;
    DS_CUDA(cudaMemcpyAsync(&h_changes, d_changes, 8, cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    // uniform runs of at least 4 pairs (a [Q, top_n] candidate list) take the grouped kernel; the kernel itself sends every
    // run that turns out not to share its title to `rest`, so a wrong guess costs time, never correctness
    if (h_changes == 0 || (unsigned long long)n % h_changes != 0 || h_changes * 4 > (unsigned long long)n)
        return launch_indel_sorted<MODE>(ws, src, nullptr, n, out);
    const int64_t run = n / (int64_t)h_changes;
    if (run > 4096) return launch_indel_sorted<MODE>(ws, src, nullptr, n, out);
    int32_t *d_rest = nullptr, *d_rest_wrap = nullptr;
    int *d_rest_count = nullptr;   // [0] general, [1] uint8 wrap region
    DS_CHECK(ws.alloc(&d_rest, (size_t)n));
    DS_CHECK(ws.alloc(&d_rest_wrap, (size_t)n));
    DS_CHECK(ws.alloc(&d_rest_count, 2));
    DS_CUDA(cudaMemsetAsync(d_rest_count, 0, 8, stream));
    const int group_lanes = (int)std::max<int64_t>(4, std::min<int64_t>(run, 32));   // run >= 4 here
    const int runs_per_block = (GROUP_BLOCK / 32) * (32 / group_lanes);
    k_indel_groups<MODE><<<(unsigned)ceil_div((int64_t)h_changes, runs_per_block), GROUP_BLOCK, 0, stream>>>(src, n, (int)run, group_lanes, out,
                                                                                                        d_rest, d_rest_wrap, d_rest_count);
    DS_LAUNCHED("k_indel_groups");
    // The leftovers (a fraction of a percent: long first titles, foreign bytes, the wrap region) take the general
    // block kernel / the wrap kernel unsorted.  The list lengths stay on the device: the grids are sized for n and
    // the blocks past the end of a list leave at once.
    {
        const size_t smem = sizeof(K2Smem<u64, 255, 64>);
        DS_CHECK((ensure_dynamic_smem(reinterpret_cast<const void *>(&k_indel_pairs<u64, 255, 64, MODE>), smem)));
        k_indel_pairs<u64, 255, 64, MODE><<<(unsigned)ceil_div(n, 64), 64, smem, stream>>>(src, d_rest, n, out, d_rest_count);
        DS_LAUNCHED("k_indel_pairs");
        if (MODE == 0) {
            k_indel_wrap<<<(unsigned)std::min<int64_t>(ceil_div(n, 4), 148 * 8), 128, 0, stream>>>(src, d_rest_wrap, n, out, d_rest_count + 1);
            DS_LAUNCHED("k_indel_wrap");
        }
    }
    return DS_OK;
}

static int launch_indel(Workspace &ws, const PairSource &src, int64_t n, int mode, const K2Out &out) {
    return mode == 0 ? launch_indel_mode<0>(ws, src, n, out) : launch_indel_mode<1>(ws, src, n, out);
}

static Side advance_side(const Side &s, int64_t first) {
    Side r = s;
    if (s.stride > 0) {
        r.base = s.base + first * s.stride;
        if (s.len8) r.len8 = s.len8 + first;
        if (s.len16) r.len16 = s.len16 + first;
    } else {
        r.idx = s.idx + first;
    }
    return r;
}

// construct_features over n_pairs: word-feature kernel + two batched ratio passes, in chunks that bound the
// reconstruction scratch (288 B per pair)
static int launch_features(Workspace &ws, const PairSource &src, const uint32_t *counts, int counts_per_pair, uint8_t space_code,
                           uint32_t n_truth, int64_t n_pairs, float *out) {
    cudaStream_t stream = ws.stream();
    if (n_pairs <= 0) return DS_OK;
    const int64_t chunk = std::min<int64_t>(n_pairs, 8 << 20);
    uint8_t *d_recon = nullptr;
    uint16_t *d_recon_len = nullptr;
    DS_CHECK(ws.alloc(&d_recon, (size_t)chunk * RECON_STRIDE));
    DS_CHECK(ws.alloc(&d_recon_len, (size_t)chunk));
    for (int64_t p0 = 0; p0 < n_pairs; p0 += chunk) {
        const int64_t n = std::min(chunk, n_pairs - p0);
        PairSource part;
        part.a = advance_side(src.a, p0);
        part.b = advance_side(src.b, p0);
        // counts are indexed by the truth id of side b: per pair for padded rows, per truth title for tables
        const uint32_t *part_counts = counts + (counts_per_pair ? p0 * N_WORDS : 0);
        float *part_out = out + p0 * DS_N_FEATURES;
        const int64_t blocks = std::min<int64_t>(ceil_div(n, K3_WARPS), 148 * 16);
        k_feature_words<<<(unsigned)blocks, K3_WARPS * 32, 0, stream>>>(part, part_counts, space_code, n_truth, n, part_out, d_recon,
                                                                       d_recon_len);
        DS_LAUNCHED("k_feature_words");
        K2Out lev{};
        lev.feat = part_out;
        lev.feat_stride = DS_N_FEATURES;
        lev.feat_col = 4;                                   // lev_ratio(title, truth)            (:106)
        Workspace pass(stream);
        DS_CHECK(launch_indel(pass, part, n, 0, lev));
        PairSource recon_pair;
        recon_pair.a = Side{d_recon, RECON_STRIDE, nullptr, d_recon_len, nullptr, nullptr};
        recon_pair.b = part.b;
        lev.feat_col = 5;                                   // lev_ratio(reconstruction, truth)   (:161-162)
        DS_CHECK(launch_indel(pass, recon_pair, n, 0, lev));
    }
    return DS_OK;
}

static int require_device() {
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(DS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    return DS_OK;
}

// stages a compact title table; the offsets are needed on the host only to size the byte copy
static int stage_table(Workspace &ws, const uint8_t *bytes, const int64_t *offsets, int64_t n_titles, const int32_t *idx, int64_t n,
                       Side *side) {
    *side = Side{};
    DS_CHECK(ws.stage_in(&side->off, offsets, (size_t)n_titles + 1));
    DS_CHECK(ws.stage_in(&side->idx, idx, (size_t)n));
    if (is_device_pointer(bytes)) {
        side->base = bytes;
        return DS_OK;
    }
    if (is_device_pointer(offsets)) return fail(DS_ERR_BAD_ARG, "title bytes on the host need host offsets");
    return ws.stage_in(&side->base, bytes, (size_t)std::max<int64_t>(1, offsets[n_titles]));
}

static int stage_rows(Workspace &ws, const uint8_t *rows, const uint8_t *lens, int64_t stride, int64_t n, Side *side) {
    *side = Side{};
    side->stride = stride;
    DS_CHECK(ws.stage_in(&side->base, rows, (size_t)n * stride));
    DS_CHECK(ws.stage_in(&side->len8, lens, (size_t)n));
    return DS_OK;
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_indel_ratio_u8(const uint8_t *a, const uint8_t *b, int64_t stride, const uint8_t *la, const uint8_t *lb, int64_t n,
                      uint8_t *out_ratio, uint16_t *out_dist, void *stream_) {
    if (n < 0 || stride <= 0) return fail(DS_ERR_BAD_ARG, "n < 0 or stride <= 0");
    if (n == 0) return DS_OK;
    if (!a || !b || !la || !lb || !out_ratio) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({a, b, la, lb, out_ratio, out_dist}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src;
    DS_CHECK(stage_rows(ws, a, la, stride, n, &src.a));
    DS_CHECK(stage_rows(ws, b, lb, stride, n, &src.b));
    K2Out out{};
    DS_CHECK(ws.stage_out(&out.u8, out_ratio, (size_t)n));
    DS_CHECK(ws.stage_out(&out.dist, out_dist, (size_t)n));
    DS_CHECK(launch_indel(ws, src, n, 0, out));
    return ws.finish_outputs();
}

int ds_indel_ratio_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a, const uint8_t *bytes_b,
                         const int64_t *offsets_b, int64_t n_titles_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n,
                         uint8_t *out_ratio, uint16_t *out_dist, void *stream_) {
    if (n < 0) return fail(DS_ERR_BAD_ARG, "n < 0");
    if (n == 0) return DS_OK;
    if (!bytes_a || !offsets_a || !bytes_b || !offsets_b || !idx_a || !idx_b || !out_ratio) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({bytes_a, bytes_b, offsets_a, offsets_b, idx_a, idx_b, out_ratio, out_dist}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src;
    DS_CHECK(stage_table(ws, bytes_a, offsets_a, n_titles_a, idx_a, n, &src.a));
    DS_CHECK(stage_table(ws, bytes_b, offsets_b, n_titles_b, idx_b, n, &src.b));
    K2Out out{};
    DS_CHECK(ws.stage_out(&out.u8, out_ratio, (size_t)n));
    DS_CHECK(ws.stage_out(&out.dist, out_dist, (size_t)n));
    DS_CHECK(launch_indel(ws, src, n, 0, out));
    return ws.finish_outputs();
}

int ds_levenshtein_ratio_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a, const uint8_t *bytes_b,
                               const int64_t *offsets_b, int64_t n_titles_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n,
                               int32_t *out_ratio, void *stream_) {
    if (n < 0) return fail(DS_ERR_BAD_ARG, "n < 0");
    if (n == 0) return DS_OK;
    if (!bytes_a || !offsets_a || !bytes_b || !offsets_b || !idx_a || !idx_b || !out_ratio) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({bytes_a, bytes_b, offsets_a, offsets_b, idx_a, idx_b, out_ratio}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src;
    DS_CHECK(stage_table(ws, bytes_a, offsets_a, n_titles_a, idx_a, n, &src.a));
    DS_CHECK(stage_table(ws, bytes_b, offsets_b, n_titles_b, idx_b, n, &src.b));
    K2Out out{};
    DS_CHECK(ws.stage_out(&out.i32, out_ratio, (size_t)n));
    DS_CHECK(launch_indel(ws, src, n, 1, out));
    return ws.finish_outputs();
}

int ds_construct_features(const uint8_t *la, const uint8_t *lb, const uint8_t *a, const uint8_t *b, int64_t stride,
                          const uint32_t *counts, uint8_t space_code, uint32_t n_truth, int64_t n_pairs, float *out,
                          void *stream_) {
    if (n_pairs < 0 || stride <= 0) return fail(DS_ERR_BAD_ARG, "n_pairs < 0 or stride <= 0");
    if (n_pairs == 0) return DS_OK;
    if (!la || !lb || !a || !b || !counts || !out) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({a, b, la, lb, counts, out}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src;
    DS_CHECK(stage_rows(ws, a, la, stride, n_pairs, &src.a));
    DS_CHECK(stage_rows(ws, b, lb, stride, n_pairs, &src.b));
    const uint32_t *d_counts = nullptr;
    DS_CHECK(ws.stage_in(&d_counts, counts, (size_t)n_pairs * DS_N_WORDS));
    float *d_out = nullptr;
    DS_CHECK(ws.stage_out(&d_out, out, (size_t)n_pairs * DS_N_FEATURES));
    DS_CHECK(launch_features(ws, src, d_counts, 1, space_code, n_truth, n_pairs, d_out));
    return ws.finish_outputs();
}

int ds_construct_features_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a, const uint8_t *bytes_b,
                                const int64_t *offsets_b, int64_t n_titles_b, const uint32_t *counts_b, const int32_t *idx_a,
                                const int32_t *idx_b, uint8_t space_code, uint32_t n_truth, int64_t n_pairs, float *out,
                                void *stream_) {
    if (n_pairs < 0) return fail(DS_ERR_BAD_ARG, "n_pairs < 0");
    if (n_pairs == 0) return DS_OK;
    if (!bytes_a || !offsets_a || !bytes_b || !offsets_b || !counts_b || !idx_a || !idx_b || !out)
        return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({bytes_a, bytes_b, offsets_a, offsets_b, counts_b, idx_a, idx_b, out}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src;
    DS_CHECK(stage_table(ws, bytes_a, offsets_a, n_titles_a, idx_a, n_pairs, &src.a));
    DS_CHECK(stage_table(ws, bytes_b, offsets_b, n_titles_b, idx_b, n_pairs, &src.b));
    const uint32_t *d_counts = nullptr;
    DS_CHECK(ws.stage_in(&d_counts, counts_b, (size_t)n_titles_b * DS_N_WORDS));
    float *d_out = nullptr;
    DS_CHECK(ws.stage_out(&d_out, out, (size_t)n_pairs * DS_N_FEATURES));
    DS_CHECK(launch_features(ws, src, d_counts, 0, space_code, n_truth, n_pairs, d_out));
    return ws.finish_outputs();
}

int ds_prematch_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, const uint8_t *sorted_a, const int64_t *sorted_offsets_a,
                      int64_t n_titles_a, const uint8_t *bytes_b, const int64_t *offsets_b, const uint8_t *sorted_b,
                      const int64_t *sorted_offsets_b, int64_t n_titles_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n,
                      int32_t threshold, int32_t *out_ratio, void *stream_) {
    if (n < 0) return fail(DS_ERR_BAD_ARG, "n < 0");
    if (n == 0) return DS_OK;
    if (!bytes_a || !offsets_a || !sorted_a || !sorted_offsets_a || !bytes_b || !offsets_b || !sorted_b || !sorted_offsets_b || !idx_a ||
        !idx_b || !out_ratio)
        return fail(DS_ERR_BAD_ARG, "NULL argument");
    if (n > INT32_MAX) return fail(DS_ERR_UNSUPPORTED, "more than 2^31-1 pairs per call");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({bytes_a, bytes_b, sorted_a, sorted_b, offsets_a, offsets_b, idx_a, idx_b, out_ratio}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource raw, sorted;
    DS_CHECK(stage_table(ws, bytes_a, offsets_a, n_titles_a, idx_a, n, &raw.a));
    DS_CHECK(stage_table(ws, bytes_b, offsets_b, n_titles_b, idx_b, n, &raw.b));
    sorted.a = raw.a;
    sorted.b = raw.b;
    DS_CHECK(ws.stage_in(&sorted.a.off, sorted_offsets_a, (size_t)n_titles_a + 1));
    DS_CHECK(ws.stage_in(&sorted.b.off, sorted_offsets_b, (size_t)n_titles_b + 1));
    if (is_device_pointer(sorted_a)) sorted.a.base = sorted_a;
    else DS_CHECK(ws.stage_in(&sorted.a.base, sorted_a, (size_t)std::max<int64_t>(1, sorted_offsets_a[n_titles_a])));
    if (is_device_pointer(sorted_b)) sorted.b.base = sorted_b;
    else DS_CHECK(ws.stage_in(&sorted.b.base, sorted_b, (size_t)std::max<int64_t>(1, sorted_offsets_b[n_titles_b])));
    K2Out out{};
    DS_CHECK(ws.stage_out(&out.i32, out_ratio, (size_t)n));
    int32_t *d_list = nullptr, *d_list2 = nullptr;
    int *d_count = nullptr;   // [0] pairs past the length filter, [1] pairs that need the token-sorted ratio
    DS_CHECK(ws.alloc(&d_list, (size_t)n));
    DS_CHECK(ws.alloc(&d_list2, (size_t)n));
    DS_CHECK(ws.alloc(&d_count, 2));
    DS_CUDA(cudaMemsetAsync(d_count, 0, 8, stream));
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    k_prematch_filter<<<blocks, 256, 0, stream>>>(raw, n, (double)threshold, out.i32, d_list, d_count);
    DS_LAUNCHED("k_prematch_filter");
    const size_t smem = sizeof(K2Smem<u64, 255, 64>);
    DS_CHECK((ensure_dynamic_smem(reinterpret_cast<const void *>(&k_indel_pairs<u64, 255, 64, 1>), smem)));
    k_indel_pairs<u64, 255, 64, 1><<<(unsigned)ceil_div(n, 64), 64, smem, stream>>>(raw, d_list, n, out, d_count);
    DS_LAUNCHED("k_indel_pairs");
    k_prematch_again<<<blocks, 256, 0, stream>>>(d_list, d_count, out.i32, threshold, d_list2, d_count + 1);
    DS_LAUNCHED("k_prematch_again");
    k_indel_pairs<u64, 255, 64, 1><<<(unsigned)ceil_div(n, 64), 64, smem, stream>>>(sorted, d_list2, n, out, d_count + 1);
    DS_LAUNCHED("k_indel_pairs");
    return ws.finish_outputs();
}

int ds_select_close_matches(const int32_t *ratios, const uint8_t *invalid, int64_t n_titles, int32_t run, int32_t threshold,
                            int64_t *out_pair, void *stream_) {
    if (n_titles < 0 || run < 1) return fail(DS_ERR_BAD_ARG, "n_titles < 0 or run < 1");
    if (n_titles == 0) return DS_OK;
    if (!ratios || !out_pair) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    DeviceGuard guard(owning_device({ratios, invalid, out_pair}));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    const int32_t *d_ratios = nullptr;
    const uint8_t *d_invalid = nullptr;
    int64_t *d_out = nullptr;
    DS_CHECK(ws.stage_in(&d_ratios, ratios, (size_t)n_titles * run));
    DS_CHECK(ws.stage_in(&d_invalid, invalid, (size_t)n_titles * run));
    DS_CHECK(ws.stage_out(&d_out, out_pair, (size_t)n_titles));
    k_select_close<<<(unsigned)ceil_div(n_titles, 256), 256, 0, stream>>>(d_ratios, d_invalid, n_titles, run, threshold, d_out);
    DS_LAUNCHED("k_select_close");
    return ws.finish_outputs();
}

}  // extern "C"
