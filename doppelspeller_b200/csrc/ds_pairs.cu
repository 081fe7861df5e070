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
// Design (B200): no DP matrices.  When la + lb <= 255 no uint8 cell can wrap and
// d = la + lb - 2 * LCS(a, b); the LCS comes from the Hyyro / Allison-Dix bit-vector recurrence
//     U = V & M[c];  V = (V + U) | (V & ~M[c])          (one 64-bit word per 64 pattern characters)
// with the match masks M in shared memory (one 40-entry column per lane, or one table per warp for
// the sliding word windows of construct_features).  Patterns longer than 64 characters are processed
// 64 characters at a time, the carry out of each text step being replayed as the carry in of the next
// pass.  Only pairs that can wrap (la + lb > 255) or carry codes outside the 40-symbol table take the
// literal uint8 DP.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>

#include "ds_common.cuh"

namespace ds {

constexpr int PM_CODES = 40;       // alphabet '- a..z0..9' = codes 0..37 (feature_engineering.py:200)
constexpr int K2_BLOCK = 128;
constexpr int K3_WARPS = 4;
constexpr int RECON_MAX = 320;     // reconstruction <= 255 (windows) + 15 (spaces) + 15 (unmatched words)
constexpr int N_WORDS = DS_N_WORDS;

typedef unsigned long long u64;

// ASCII mode maps raw title bytes onto the 38 codes; anything else is "outside" (>= PM_CODES)
__device__ __forceinline__ int map_code(uint8_t c, int ascii) {
    if (!ascii) return c;
    if (c >= 'a' && c <= 'z') return c - 'a' + 2;
    if (c >= '0' && c <= '9') return c - '0' + 28;
    if (c == ' ') return 1;
    if (c == '-') return 0;
    return 255;
}

// literal restatement of the uint8 DP (feature_engineering.py:42-61); `y` (length ly <= 255) indexes the
// row buffer, `x` the outer loop.  The recurrence is symmetric, so which string plays which role does
// not change any cell value (the reference puts the shorter one on the rows, :35-37).
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

// Bit-vector LCS of pattern `pat` (m <= 255) against text `txt` (n <= 511) with this lane's match-mask
// column `pm` (entry of code c at pm[c * STRIDE], all zero on entry and on exit).  All codes < PM_CODES.
template <int STRIDE>
__device__ int lcs_bitvector(u64 *pm, const uint8_t *pat, int m, const uint8_t *txt, int n, int ascii) {
    if (m <= 64) {
        for (int i = 0; i < m; ++i) pm[map_code(pat[i], ascii) * STRIDE] |= 1ull << i;
        u64 v = ~0ull;
        for (int j = 0; j < n; ++j) {
            const u64 mm = pm[map_code(txt[j], ascii) * STRIDE];
            const u64 u = v & mm;
            v = (v + u) | (v & ~mm);
        }
        for (int i = 0; i < m; ++i) pm[map_code(pat[i], ascii) * STRIDE] = 0;
        const u64 valid = (m == 64) ? ~0ull : ((1ull << m) - 1);
        return __popcll(~v & valid);
    }
    uint32_t carry[16];  // carry into the current pattern block at text step j (n <= 512 bits)
    for (int i = 0; i < 16; ++i) carry[i] = 0;
    int lcs = 0;
    for (int w0 = 0; w0 < m; w0 += 64) {
        const int mw = min(64, m - w0);
        for (int i = 0; i < mw; ++i) pm[map_code(pat[w0 + i], ascii) * STRIDE] |= 1ull << i;
        u64 v = ~0ull;
        for (int j = 0; j < n; ++j) {
            const u64 mm = pm[map_code(txt[j], ascii) * STRIDE];
            const u64 u = v & mm;
            const u64 cin = (carry[j >> 5] >> (j & 31)) & 1u;
            const u64 s1 = v + u;
            const u64 s2 = s1 + cin;
            const uint32_t cout = (s1 < v) | (s2 < s1);
            v = s2 | (v & ~mm);
            carry[j >> 5] = (carry[j >> 5] & ~(1u << (j & 31))) | (cout << (j & 31));
        }
        for (int i = 0; i < mw; ++i) pm[map_code(pat[w0 + i], ascii) * STRIDE] = 0;
        const u64 valid = (mw == 64) ? ~0ull : ((1ull << mw) - 1);
        lcs += __popcll(~v & valid);
    }
    return lcs;
}

__device__ __forceinline__ bool codes_in_table(const uint8_t *s, int n, int ascii) {
    bool ok = true;
    for (int i = 0; i < n; ++i) ok &= map_code(s[i], ascii) < PM_CODES;
    return ok;
}

// uint8-wrapped distance of fast_levenshtein_ratio.  `codes_ok`: both strings only hold table codes.
template <int STRIDE>
__device__ int indel_distance_u8(u64 *pm, const uint8_t *a, int la, const uint8_t *b, int lb, bool codes_ok) {
    if (la + lb <= 255 && codes_ok) {
        // shorter string = pattern (fewer 64-character blocks)
        int lcs = (la <= lb) ? lcs_bitvector<STRIDE>(pm, a, la, b, lb, 0) : lcs_bitvector<STRIDE>(pm, b, lb, a, la, 0);
        return la + lb - 2 * lcs;
    }
    return (lb <= 255) ? indel_u8_dp(a, la, b, lb) : indel_u8_dp(b, lb, a, la);
}

__device__ __forceinline__ int ratio_u8(int total, int d) {
    // uint8(((L - d) * 100) / L): integer exact (feature_engineering.py:63 under fastmath)
    return total == 0 ? 0 : ((100 * (total - d)) / total) & 0xff;
}

// ---------------------------------------------------------------------------------------------------
// pair addressing: the reference's padded [P, stride] rows or a compact (bytes, offsets, index) table
// ---------------------------------------------------------------------------------------------------
struct PairSource {
    const uint8_t *a, *b;
    int64_t stride;             // > 0: padded rows + la/lb arrays
    const uint8_t *la, *lb;
    const int64_t *off_a, *off_b;  // compact tables
    const int32_t *idx_a, *idx_b;
};

__device__ __forceinline__ void load_pair(const PairSource &s, int64_t p, const uint8_t **pa, int *la, const uint8_t **pb,
                                          int *lb, int64_t *truth_id) {
    if (s.stride > 0) {
        *pa = s.a + p * s.stride;
        *pb = s.b + p * s.stride;
        *la = s.la[p];
        *lb = s.lb[p];
        *truth_id = p;
    } else {
        const int64_t ia = s.idx_a[p], ib = s.idx_b[p];
        const int64_t a0 = s.off_a[ia], b0 = s.off_b[ib];
        *pa = s.a + a0;
        *pb = s.b + b0;
        *la = (int)min((int64_t)DS_MAX_TITLE, s.off_a[ia + 1] - a0);
        *lb = (int)min((int64_t)DS_MAX_TITLE, s.off_b[ib + 1] - b0);
        *truth_id = ib;
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: batched InDel ratio.
//   MODE 0: fast_levenshtein_ratio (uint8 wrap semantics)  -> out_u8 / out_dist
//   MODE 1: common.levenshtein_ratio on raw bytes          -> out_i32
// Pairs are radix-sorted by (class, la + lb) so that the lanes of a warp run the same code path for about
// the same number of steps:
//   class 0  both strings <= 64 bytes: one pair per lane, one 64-bit word, one pass
//   class 1  bit-vector path for longer strings (MODE 0: la + lb <= 255, MODE 1: all): one pair per lane,
//            pattern in 64-character blocks, the carries of each text step replayed into the next block
//   class 2  MODE 0 only, la + lb > 255 (uint8 cells can wrap): one pair per WARP, literal uint8 DP swept
//            by anti-diagonals over 32-column strips
// Strings are fetched with aligned 32-bit loads, re-aligned with funnel shifts and kept in a per-lane
// shared-memory slot (odd word stride: conflict free), then consumed four characters per LDS.
// ---------------------------------------------------------------------------------------------------
template <int MAXLEN, int BLOCK>
struct K2Smem {
    static constexpr int STAGE_WORDS = ((MAXLEN + 3) / 4) | 1;   // words per staged string, odd lane stride
    u64 pm[PM_CODES * BLOCK];
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

template <int MODE>
__device__ __forceinline__ int table_code(uint32_t byte, bool &good) {
    int c = MODE == 0 ? (int)byte : map_code((uint8_t)byte, 1);
    good &= c < PM_CODES;
    return min(c, PM_CODES - 1);
}

// single-block LCS: pattern m <= 64 staged in `pat`, text n staged in `txt` (word-aligned slots)
template <int BLOCK, int MODE>
__device__ __forceinline__ int lcs_one_block(u64 *pm, const uint32_t *pat, int m, const uint32_t *txt, int n, bool *ok) {
    bool good = true;
    for (int i0 = 0; i0 < m; i0 += 4) {
        uint32_t w = pat[i0 >> 2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < m) {
                const int c = table_code<MODE>((w >> (8 * k)) & 0xffu, good);
                pm[c * BLOCK] |= 1ull << (i0 + k);
            }
        }
    }
    u64 v = ~0ull;
    for (int j0 = 0; j0 < n; j0 += 4) {
        const uint32_t w = txt[j0 >> 2];
        u64 mm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool in_table = true;   // bytes past the end of the text are whatever follows it in memory: ignored
            mm[k] = pm[table_code<MODE>((w >> (8 * k)) & 0xffu, in_table) * BLOCK];
            good &= in_table | (j0 + k >= n);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (j0 + k < n) {
                const u64 u = v & mm[k];
                v = (v + u) | (v & ~mm[k]);
            }
        }
    }
    *ok = good;
    const u64 valid = (m == 64) ? ~0ull : ((1ull << m) - 1);
    return __popcll(~v & valid);
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

__device__ __forceinline__ void pair_lengths(const PairSource &src, int64_t p, int *la, int *lb) {
    if (src.stride > 0) {
        *la = src.la[p];
        *lb = src.lb[p];
    } else {
        const int64_t ia = src.idx_a[p], ib = src.idx_b[p];
        *la = (int)min((int64_t)DS_MAX_TITLE, src.off_a[ia + 1] - src.off_a[ia]);
        *lb = (int)min((int64_t)DS_MAX_TITLE, src.off_b[ib + 1] - src.off_b[ib]);
    }
}

__device__ __forceinline__ int pair_class(int la, int lb, int mode) {
    if (la <= 64 && lb <= 64) return 0;
    if (mode == 1 || la + lb <= 255) return 1;
    return 2;
}

// sort key = (class << 9) | (la + lb); values = pair ids; counts per class
__global__ void k_pair_keys(PairSource src, int64_t n, int mode, uint16_t *__restrict__ keys, int32_t *__restrict__ ids,
                            int *__restrict__ counts) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1;
    if (p < n) {
        int la, lb;
        pair_lengths(src, p, &la, &lb);
        cls = pair_class(la, lb, mode);
        keys[p] = (uint16_t)((cls << 9) | (la + lb));
        ids[p] = (int32_t)p;
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const unsigned ballot = __ballot_sync(0xffffffffu, cls == c);
        if (ballot != 0 && lane == __ffs(ballot) - 1) atomicAdd(counts + c, __popc(ballot));
    }
}

template <int MODE>
__device__ __forceinline__ void store_result(int total, int d, int64_t p, uint8_t *out_u8, uint16_t *out_dist, int32_t *out_i32) {
    if (MODE == 0) {
        out_u8[p] = (uint8_t)ratio_u8(total, d);
        if (out_dist) out_dist[p] = (uint16_t)d;
    } else {
        int result = 100;  // ratio 1.0 for two empty strings
        if (total > 0) {
            // int(round(ratio * 100)): float64 divide, float64 multiply, round half to even
            const double ratio = __ddiv_rn((double)(total - d), (double)total);
            result = (int)rint(__dmul_rn(ratio, 100.0));
        }
        out_i32[p] = result;
    }
}

// classes 0 and 1: one pair per lane
template <int MAXLEN, int BLOCK, int MODE>
__global__ void __launch_bounds__(BLOCK) k_indel_pairs(PairSource src, const int32_t *__restrict__ pair_list, int64_t n, uint8_t *out_u8,
                                                       uint16_t *out_dist, int32_t *out_i32) {
    extern __shared__ __align__(16) unsigned char k2_raw[];
    typedef K2Smem<MAXLEN, BLOCK> Smem;
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
    int64_t tid_unused;
    load_pair(src, p, &a, &la, &b, &lb, &tid_unused);
    u64 *my_pm = sm.pm + threadIdx.x;
    const int total = la + lb;
    int d = 0;
    bool done = false;
    if (la <= MAXLEN && lb <= MAXLEN && (MODE == 1 || MAXLEN <= 64 || total <= 255)) {
        uint32_t *sa = sm.stage_a + threadIdx.x * Smem::STAGE_WORDS;
        uint32_t *sb = sm.stage_b + threadIdx.x * Smem::STAGE_WORDS;
        stage_string(a, la, sa);
        stage_string(b, lb, sb);
        bool ok;
        int lcs;
        if (MAXLEN <= 64) lcs = (la <= lb) ? lcs_one_block<BLOCK, MODE>(my_pm, sa, la, sb, lb, &ok) : lcs_one_block<BLOCK, MODE>(my_pm, sb, lb, sa, la, &ok);
        else lcs = (la <= lb) ? lcs_blocks<BLOCK, MODE>(my_pm, sa, la, sb, lb, &ok) : lcs_blocks<BLOCK, MODE>(my_pm, sb, lb, sa, la, &ok);
        d = total - 2 * lcs;
        done = ok;
        if (!ok) {   // a byte outside the table polluted the lane's column: clean it for the general path
            for (int c = 0; c < PM_CODES; ++c) my_pm[c * BLOCK] = 0;
        }
    }
    if (!done) {   // bytes outside the 40-symbol table (or an unsorted tiny batch): literal DP per lane
        if (MODE == 0) d = (lb <= 255) ? indel_u8_dp(a, la, b, lb) : indel_u8_dp(b, lb, a, la);
        else d = indel_true_dp(a, la, b, lb);
    }
    store_result<MODE>(total, d, p, out_u8, out_dist, out_i32);
}

// class 2 (MODE 0, la + lb > 255): the literal uint8 DP of feature_engineering.py:42-61, one pair per warp.
// Columns are processed in strips of 32 (lane = column); inside a strip the cells of an anti-diagonal are
// independent: at step s lane l owns row s - l + 1; `up` is its own previous cell, `left` / `diag` come from
// lane l - 1 (shuffle) or, for the strip's first column, from the previous strip's last column kept in
// shared memory.  Every cell is reduced mod 256 on store exactly like the uint8 matrix.
struct WrapSmem {
    uint8_t x[256];
    uint8_t y[256];
    uint8_t edge[2][260];
};

__global__ void __launch_bounds__(128) k_indel_wrap(PairSource src, const int32_t *__restrict__ pair_list, int64_t n, uint8_t *out_u8,
                                                    uint16_t *out_dist) {
    __shared__ WrapSmem smem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slot = (int64_t)blockIdx.x * 4 + warp;
    if (slot >= n) return;
    WrapSmem &sm = smem[warp];
    const int64_t p = pair_list ? (int64_t)pair_list[slot] : slot;
    const uint8_t *a, *b;
    int la, lb;
    int64_t tid_unused;
    load_pair(src, p, &a, &la, &b, &lb, &tid_unused);
    // x = rows (outer sweep), y = columns (strips); fewer strips with the shorter string on the columns
    const uint8_t *gx = (la >= lb) ? a : b, *gy = (la >= lb) ? b : a;
    const int lx = max(la, lb), ly = min(la, lb);
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
            } else if (i < 1) {
                prev_left = (lane == 0) ? left_col[0] : ((j - 1) & 0xff);
            }
        }
        __syncwarp();
        result = __shfl_sync(0xffffffffu, cur, last_lane);
        cur_edge ^= 1;
    }
    if (lane == 0) {
        out_u8[p] = (uint8_t)ratio_u8(la + lb, result);
        if (out_dist) out_dist[p] = (uint16_t)result;
    }
}

// ---------------------------------------------------------------------------------------------------
// K3: construct_features.  A warp owns 32 pairs:
//   phase 1 (lane = pair)   word counts, code check, lev_ratio(title, truth)
//   phase 2 (warp = pair)   for every truth word: the sliding windows over the space-less title are
//                           spread over the lanes (one shared match-mask table per warp); first
//                           strict maximum wins (feature_engineering.py:139-149); reconstruction
//   phase 3 (lane = pair)   lev_ratio(reconstructed title, truth), basic features
// ---------------------------------------------------------------------------------------------------
struct K3Smem {
    u64 pm_lane[PM_CODES * 32];
    u64 pm_word[PM_CODES];
    uint8_t a_ns[256];
    uint8_t b_cur[256];
    uint8_t recon[32][RECON_MAX];
    int recon_len[32];
};

__global__ void __launch_bounds__(K3_WARPS * 32) k_features(PairSource src, const uint32_t *__restrict__ counts, int counts_per_truth,
                                                            int space_code, uint32_t n_truth, int64_t n_pairs, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K3Smem &sm = reinterpret_cast<K3Smem *>(smem_raw)[warp];
    for (int i = lane; i < PM_CODES * 32; i += 32) sm.pm_lane[i] = 0;
    for (int i = lane; i < PM_CODES; i += 32) sm.pm_word[i] = 0;
    __syncwarp();

    const int64_t warp_id = (int64_t)blockIdx.x * K3_WARPS + warp;
    const int64_t p_base = warp_id * 32;
    if (p_base >= n_pairs) return;
    const int n_here = (int)min((int64_t)32, n_pairs - p_base);
    const float nan_f = __int_as_float(0x7fc00000);

    // ---- phase 1 ----
    const uint8_t *a = nullptr, *b = nullptr;
    int la = 0, lb = 0, words_a = 1, words_b = 1, lev = 0;
    int64_t truth_id = 0;
    bool ok = true;
    u64 *my_pm = sm.pm_lane + lane;
    if (lane < n_here) {
        load_pair(src, p_base + lane, &a, &la, &b, &lb, &truth_id);
        for (int i = 0; i < la; ++i) {
            const uint8_t c = a[i];
            words_a += (c == space_code);
            ok &= c < PM_CODES;
        }
        for (int i = 0; i < lb; ++i) {
            const uint8_t c = b[i];
            words_b += (c == space_code);
            ok &= c < PM_CODES;
        }
        lev = ratio_u8(la + lb, indel_distance_u8<32>(my_pm, a, la, b, lb, ok));
    }
    __syncwarp();

    // ---- phase 2 ----
    for (int pp = 0; pp < n_here; ++pp) {
        const uint8_t *pa = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, (u64)a, pp));
        const uint8_t *pb = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, (u64)b, pp));
        const int pla = __shfl_sync(0xffffffffu, la, pp), plb = __shfl_sync(0xffffffffu, lb, pp);
        const int p_words_b = __shfl_sync(0xffffffffu, words_b, pp);
        const bool p_ok = __shfl_sync(0xffffffffu, (int)ok, pp) != 0;
        const int64_t p_truth = __shfl_sync(0xffffffffu, (long long)truth_id, pp);
        // stage truth and the space-less title
        int n_ns = 0;
        for (int i0 = 0; i0 < max(pla, plb); i0 += 32) {
            const int i = i0 + lane;
            if (i < plb) sm.b_cur[i] = pb[i];
            uint8_t c = 0;
            bool keep = false;
            if (i < pla) {
                c = pa[i];
                keep = c != space_code;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (keep) sm.a_ns[n_ns + __popc(ballot & ((1u << lane) - 1))] = c;
            n_ns += __popc(ballot);
        }
        __syncwarp();

        float my_best = nan_f, my_wlen = nan_f, my_idf = nan_f;  // lane w < 15 owns truth word w
        int n_words = 0, last = 0, n_recon = 0;
        uint8_t *recon = sm.recon[pp];
        for (int base = 0; base <= plb && n_words < N_WORDS; base += 32) {
            const int pos_l = base + lane;
            const bool sep = pos_l <= plb && (pos_l == plb || sm.b_cur[pos_l] == space_code);
            unsigned sepmask = __ballot_sync(0xffffffffu, sep);
            while (sepmask != 0 && n_words < N_WORDS) {
                const int pos = base + __ffs(sepmask) - 1;
                sepmask &= sepmask - 1;
                const uint8_t *word = sm.b_cur + last;
                const int wl = pos - last;
                last = pos + 1;
                int best_ratio = 0, best_start = -1, best_len = 1;
                if (wl > 0 && n_ns > 0) {
                    const bool fast = p_ok && wl <= 64;
                    if (fast) {
                        for (int i = lane; i < wl; i += 32) atomicOr(&sm.pm_word[word[i]], 1ull << i);
                        __syncwarp();
                    }
                    for (int i0 = 0; i0 < n_ns; i0 += 32) {
                        const int i = i0 + lane;
                        int key = -1;
                        if (i < n_ns) {
                            const int pl = min(wl, n_ns - i);
                            const uint8_t *win = sm.a_ns + i;
                            int d;
                            if (fast) {
                                u64 v = ~0ull;
                                for (int j = 0; j < pl; ++j) {
                                    const u64 mm = sm.pm_word[win[j]];
                                    const u64 u = v & mm;
                                    v = (v + u) | (v & ~mm);
                                }
                                const u64 valid = (wl == 64) ? ~0ull : ((1ull << wl) - 1);
                                d = pl + wl - 2 * __popcll(~v & valid);
                            } else if (pl + wl <= 255 && p_ok) {
                                d = pl + wl - 2 * lcs_bitvector<32>(my_pm, win, pl, word, wl, 0);
                            } else {
                                d = indel_u8_dp(word, wl, win, pl);
                            }
                            const int r = ratio_u8(pl + wl, d);
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
                        for (int i = lane; i < wl; i += 32) sm.pm_word[word[i]] = 0;
                        __syncwarp();
                    }
                }
                if (lane == n_words) {
                    my_best = (float)best_ratio;
                    my_wlen = (float)wl;
                    const uint32_t cnt = counts[p_truth * counts_per_truth + n_words];
                    my_idf = (float)log(__ddiv_rn((double)n_truth, (double)cnt));
                }
                // reconstructed title: best window (or a single space) followed by a space (:154-155)
                if (best_start < 0) {
                    if (lane == 0) recon[n_recon] = (uint8_t)space_code;
                } else {
                    for (int i = lane; i < best_len; i += 32) recon[n_recon + i] = sm.a_ns[best_start + i];
                }
                n_recon += best_len;
                if (lane == 0) recon[n_recon] = (uint8_t)space_code;
                n_recon += 1;
                ++n_words;
            }
        }
        if (lane == 0) sm.recon_len[pp] = n_recon > 0 ? n_recon - 1 : 0;  // drop the trailing space (:161)
        // IDF ranks (:158): NaN for every slot unless all 15 word slots are filled
        float rank = nan_f;
        if (n_words == N_WORDS) {
            float mxv = (lane < N_WORDS) ? my_idf : -INFINITY;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mxv = fmaxf(mxv, __shfl_xor_sync(0xffffffffu, mxv, d));
            const float diff = __fsub_rn(mxv, my_idf);
            rank = (float)__dadd_rn(1.0, __ddiv_rn((double)diff, (double)p_words_b));
        }
        if (lane < N_WORDS) {
            float *o = out + (p_base + pp) * DS_N_FEATURES;
            o[6 + lane] = my_best;
            o[6 + N_WORDS + lane] = my_wlen;
            o[6 + 2 * N_WORDS + lane] = my_idf;
            o[6 + 3 * N_WORDS + lane] = rank;
        }
        __syncwarp();
    }

    // ---- phase 3 ----
    if (lane < n_here) {
        const int lr = sm.recon_len[lane];
        const uint8_t *recon = sm.recon[lane];
        bool ok_r = ok;  // recon only holds title characters and spaces
        ok_r &= space_code < PM_CODES;
        int d;
        if (lr + lb <= 255 && ok_r) {
            int lcs = (lr <= lb) ? lcs_bitvector<32>(my_pm, recon, lr, b, lb, 0) : lcs_bitvector<32>(my_pm, b, lb, recon, lr, 0);
            d = lr + lb - 2 * lcs;
        } else {
            d = indel_u8_dp(recon, lr, b, lb);
        }
        float *o = out + (p_base + lane) * DS_N_FEATURES;
        o[0] = (float)la;
        o[1] = (float)lb;
        o[2] = (float)words_a;
        o[3] = (float)words_b;
        o[4] = (float)lev;
        o[5] = (float)ratio_u8(lr + lb, d);
    }
}

template <int MAXLEN, int BLOCK, int MODE>
static int launch_indel_class(const PairSource &src, const int32_t *list, int64_t n, uint8_t *out_u8, uint16_t *out_dist,
                              int32_t *out_i32, cudaStream_t stream) {
    if (n <= 0) return DS_OK;
    const size_t smem = sizeof(K2Smem<MAXLEN, BLOCK>);
    static bool attr_done = false;
    if (!attr_done) {
        DS_CUDA((cudaFuncSetAttribute(k_indel_pairs<MAXLEN, BLOCK, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        attr_done = true;
    }
    k_indel_pairs<MAXLEN, BLOCK, MODE><<<(unsigned)ceil_div(n, BLOCK), BLOCK, smem, stream>>>(src, list, n, out_u8, out_dist, out_i32);
    DS_LAUNCHED("k_indel_pairs");
    return DS_OK;
}

template <int MODE>
static int launch_indel_mode(Workspace &ws, const PairSource &src, int64_t n, uint8_t *out_u8, uint16_t *out_dist, int32_t *out_i32) {
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
    DS_CHECK(ws.alloc(&d_counts, 3));
    DS_CUDA(cudaMemsetAsync(d_counts, 0, 3 * sizeof(int), stream));
    k_pair_keys<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(src, n, MODE, d_keys, d_ids, d_counts);
    DS_LAUNCHED("k_pair_keys");
    size_t temp_bytes = 0;
    DS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys, d_keys_sorted, d_ids, d_ids_sorted, (int)n, 0, 11, stream));
    unsigned char *d_temp = nullptr;
    DS_CHECK(ws.alloc(&d_temp, temp_bytes));
    DS_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys_sorted, d_ids, d_ids_sorted, (int)n, 0, 11, stream));
    g_kernel_launches.fetch_add(3);
    int h_counts[3] = {0, 0, 0};
    DS_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof(h_counts), cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    const int32_t *list = d_ids_sorted;
    DS_CHECK((launch_indel_class<64, K2_BLOCK, MODE>(src, list, h_counts[0], out_u8, out_dist, out_i32, stream)));
    DS_CHECK((launch_indel_class<255, 64, MODE>(src, list + h_counts[0], h_counts[1], out_u8, out_dist, out_i32, stream)));
    if (h_counts[2] > 0) {
        k_indel_wrap<<<(unsigned)ceil_div(h_counts[2], 4), 128, 0, stream>>>(src, list + h_counts[0] + h_counts[1], h_counts[2], out_u8,
                                                                             out_dist);
        DS_LAUNCHED("k_indel_wrap");
    }
    return DS_OK;
}

static int launch_indel(Workspace &ws, const PairSource &src, int64_t n, int mode, uint8_t *out_u8, uint16_t *out_dist, int32_t *out_i32) {
    return mode == 0 ? launch_indel_mode<0>(ws, src, n, out_u8, out_dist, out_i32)
                     : launch_indel_mode<1>(ws, src, n, out_u8, out_dist, out_i32);
}

static int launch_features(const PairSource &src, const uint32_t *counts, int counts_per_truth, uint8_t space_code, uint32_t n_truth,
                           int64_t n_pairs, float *out, cudaStream_t stream) {
    if (n_pairs <= 0) return DS_OK;
    const size_t smem = sizeof(K3Smem) * K3_WARPS;
    static bool attr_done = false;
    if (!attr_done) {
        DS_CUDA(cudaFuncSetAttribute(k_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int64_t warps = ceil_div(n_pairs, 32);
    k_features<<<(unsigned)ceil_div(warps, K3_WARPS), K3_WARPS * 32, smem, stream>>>(src, counts, counts_per_truth, space_code, n_truth,
                                                                                      n_pairs, out);
    DS_LAUNCHED("k_features");
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
static int stage_table(Workspace &ws, const uint8_t *bytes, const int64_t *offsets, int64_t n_titles, const uint8_t **d_bytes,
                       const int64_t **d_offsets) {
    DS_CHECK(ws.stage_in(d_offsets, offsets, (size_t)n_titles + 1));
    if (is_device_pointer(bytes)) {
        *d_bytes = bytes;
        return DS_OK;
    }
    if (is_device_pointer(offsets)) return fail(DS_ERR_BAD_ARG, "title bytes on the host need host offsets");
    return ws.stage_in(d_bytes, bytes, (size_t)offsets[n_titles]);
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
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src{};
    src.stride = stride;
    DS_CHECK(ws.stage_in(&src.a, a, (size_t)n * stride));
    DS_CHECK(ws.stage_in(&src.b, b, (size_t)n * stride));
    DS_CHECK(ws.stage_in(&src.la, la, (size_t)n));
    DS_CHECK(ws.stage_in(&src.lb, lb, (size_t)n));
    uint8_t *d_ratio = nullptr;
    uint16_t *d_dist = nullptr;
    DS_CHECK(ws.stage_out(&d_ratio, out_ratio, (size_t)n));
    DS_CHECK(ws.stage_out(&d_dist, out_dist, (size_t)n));
    DS_CHECK(launch_indel(ws, src, n, 0, d_ratio, d_dist, nullptr));
    return ws.finish_outputs();
}

static int pairs_source(Workspace &ws, const uint8_t *bytes_a, const int64_t *offsets_a, const uint8_t *bytes_b,
                        const int64_t *offsets_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n, int64_t n_titles_a,
                        int64_t n_titles_b, PairSource *src) {
    src->stride = 0;
    DS_CHECK(stage_table(ws, bytes_a, offsets_a, n_titles_a, &src->a, &src->off_a));
    DS_CHECK(stage_table(ws, bytes_b, offsets_b, n_titles_b, &src->b, &src->off_b));
    DS_CHECK(ws.stage_in(&src->idx_a, idx_a, (size_t)n));
    DS_CHECK(ws.stage_in(&src->idx_b, idx_b, (size_t)n));
    return DS_OK;
}


int ds_indel_ratio_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a, const uint8_t *bytes_b,
                           const int64_t *offsets_b, int64_t n_titles_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n,
                           uint8_t *out_ratio, uint16_t *out_dist, void *stream_) {
    if (n < 0) return fail(DS_ERR_BAD_ARG, "n < 0");
    if (n == 0) return DS_OK;
    if (!bytes_a || !offsets_a || !bytes_b || !offsets_b || !idx_a || !idx_b || !out_ratio) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src{};
    DS_CHECK(pairs_source(ws, bytes_a, offsets_a, bytes_b, offsets_b, idx_a, idx_b, n, n_titles_a, n_titles_b, &src));
    uint8_t *d_ratio = nullptr;
    uint16_t *d_dist = nullptr;
    DS_CHECK(ws.stage_out(&d_ratio, out_ratio, (size_t)n));
    DS_CHECK(ws.stage_out(&d_dist, out_dist, (size_t)n));
    DS_CHECK(launch_indel(ws, src, n, 0, d_ratio, d_dist, nullptr));
    return ws.finish_outputs();
}

int ds_levenshtein_ratio_pairs(const uint8_t *bytes_a, const int64_t *offsets_a, int64_t n_titles_a, const uint8_t *bytes_b,
                                 const int64_t *offsets_b, int64_t n_titles_b, const int32_t *idx_a, const int32_t *idx_b, int64_t n,
                                 int32_t *out, void *stream_) {
    if (n < 0) return fail(DS_ERR_BAD_ARG, "n < 0");
    if (n == 0) return DS_OK;
    if (!bytes_a || !offsets_a || !bytes_b || !offsets_b || !idx_a || !idx_b || !out) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src{};
    DS_CHECK(pairs_source(ws, bytes_a, offsets_a, bytes_b, offsets_b, idx_a, idx_b, n, n_titles_a, n_titles_b, &src));
    int32_t *d_out = nullptr;
    DS_CHECK(ws.stage_out(&d_out, out, (size_t)n));
    DS_CHECK(launch_indel(ws, src, n, 1, nullptr, nullptr, d_out));
    return ws.finish_outputs();
}

int ds_construct_features(const uint8_t *la, const uint8_t *lb, const uint8_t *a, const uint8_t *b, int64_t stride,
                          const uint32_t *counts, uint8_t space_code, uint32_t n_truth, int64_t n_pairs, float *out,
                          void *stream_) {
    if (n_pairs < 0 || stride <= 0) return fail(DS_ERR_BAD_ARG, "n_pairs < 0 or stride <= 0");
    if (n_pairs == 0) return DS_OK;
    if (!la || !lb || !a || !b || !counts || !out) return fail(DS_ERR_BAD_ARG, "NULL argument");
    DS_CHECK(require_device());
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src{};
    src.stride = stride;
    DS_CHECK(ws.stage_in(&src.a, a, (size_t)n_pairs * stride));
    DS_CHECK(ws.stage_in(&src.b, b, (size_t)n_pairs * stride));
    DS_CHECK(ws.stage_in(&src.la, la, (size_t)n_pairs));
    DS_CHECK(ws.stage_in(&src.lb, lb, (size_t)n_pairs));
    const uint32_t *d_counts = nullptr;
    DS_CHECK(ws.stage_in(&d_counts, counts, (size_t)n_pairs * DS_N_WORDS));
    float *d_out = nullptr;
    DS_CHECK(ws.stage_out(&d_out, out, (size_t)n_pairs * DS_N_FEATURES));
    DS_CHECK(launch_features(src, d_counts, DS_N_WORDS, space_code, n_truth, n_pairs, d_out, stream));
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
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    PairSource src{};
    DS_CHECK(pairs_source(ws, bytes_a, offsets_a, bytes_b, offsets_b, idx_a, idx_b, n_pairs, n_titles_a, n_titles_b, &src));
    const uint32_t *d_counts = nullptr;
    DS_CHECK(ws.stage_in(&d_counts, counts_b, (size_t)n_titles_b * DS_N_WORDS));
    float *d_out = nullptr;
    DS_CHECK(ws.stage_out(&d_out, out, (size_t)n_pairs * DS_N_FEATURES));
    DS_CHECK(launch_features(src, d_counts, DS_N_WORDS, space_code, n_truth, n_pairs, d_out, stream));
    return ws.finish_outputs();
}

}  // extern "C"
