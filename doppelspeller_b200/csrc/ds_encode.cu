// ds_encode.cu - index build, host-preprocessing half moved to the GPU:
// titles -> per-title trigram SETS -> canonical column ids -> document frequencies -> idf.
//
// Reference semantics (paths relative to /root/reference):
//   common.get_n_grams            doppelspeller/common.py:150-151   every 3-character window, as a set
//   common.get_n_grams_counter    doppelspeller/common.py:145-147   document frequency over the per-title sets
//   MatchMaker._idf_n_gram        doppelspeller/match_maker.py:135-142   idf = math.log(N / df)
//   MatchMaker._get_encoding_values  match_maker.py:149-153  query-only n-grams weigh max(idf)  (:95,:180-181)
// The reference numbers the columns by python set iteration order (:144-147), which depends on
// PYTHONHASHSEED; this encoder uses the canonical order instead (column id = rank of the trigram code
// c0 * 37^2 + c1 * 37 + c2 over the alphabet ' a..z0..9'), the order `encode.encode_canonical` defines on
// the host and every benchmark uses.  `MatchMaker.__init__` keeps the reference's own order for bit-parity
// with one particular reference process.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

#include "ds_common.cuh"

namespace ds {

constexpr int ENC_BASE = 37;
constexpr int ENC_CODES = ENC_BASE * ENC_BASE * ENC_BASE;   // 50,653 possible trigrams
constexpr int ENC_WARPS = 4;

__device__ __forceinline__ int enc_symbol(uint8_t c) {
    if (c == ' ') return 0;
    if (c >= 'a' && c <= 'z') return c - 'a' + 1;
    if (c >= '0' && c <= '9') return c - '0' + 27;
    return -1;
}

// One warp per title.  PASS 0: number of distinct trigrams.  PASS 1: the distinct trigram codes, ascending,
// at out[ptr[title] ..); presence flags; document frequencies (truth titles only).
template <int PASS>
__global__ void __launch_bounds__(ENC_WARPS * 32) k_trigrams(const uint8_t *__restrict__ bytes, const int64_t *__restrict__ offsets,
                                                            int64_t n_titles, int64_t *__restrict__ counts,
                                                            const int64_t *__restrict__ ptr, uint16_t *__restrict__ out,
                                                            int *__restrict__ present, int *__restrict__ df, int *__restrict__ bad) {
    __shared__ int s_code[ENC_WARPS][256];
    __shared__ bool s_first[ENC_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * ENC_WARPS + warp;
    if (t >= n_titles) return;
    int *code = s_code[warp];
    bool *is_first = s_first[warp];
    const int64_t o = offsets[t];
    const int len = (int)min((int64_t)DS_MAX_TITLE, offsets[t + 1] - o);
    const int g = max(len - 2, 0);
    for (int i = lane; i < g; i += 32) {
        const int a = enc_symbol(bytes[o + i]), b = enc_symbol(bytes[o + i + 1]), c = enc_symbol(bytes[o + i + 2]);
        if ((a | b | c) < 0) atomicExch(bad, 1);
        code[i] = (max(a, 0) * ENC_BASE + max(b, 0)) * ENC_BASE + max(c, 0);
    }
    __syncwarp();
    // set semantics (common.py:150-151): keep the first occurrence of every code
    int n_unique = 0;
    for (int i0 = 0; i0 < g; i0 += 32) {
        const int i = i0 + lane;
        bool first = false;
        if (i < g) {
            const int x = code[i];
            first = true;
            for (int j = 0; j < i; ++j) first &= code[j] != x;
            if (PASS == 1) is_first[i] = first;
        }
        n_unique += __popc(__ballot_sync(0xffffffffu, first));
    }
    if (PASS == 1) {
        __syncwarp();
        for (int i = lane; i < g; i += 32) {
            if (!is_first[i]) continue;
            const int x = code[i];
            int rank = 0;                                            // ascending position among the distinct codes
            for (int j = 0; j < g; ++j) rank += is_first[j] && code[j] < x;
            out[ptr[t] + rank] = (uint16_t)x;
            present[x] = 1;
            if (df != nullptr) atomicAdd(df + x, 1);
        }
    }
    if (PASS == 0 && lane == 0) counts[t] = n_unique;
}

__global__ void k_remap(uint16_t *__restrict__ cols, int64_t n, const int *__restrict__ new_id) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cols[i] = (uint16_t)new_id[cols[i]];
}

__global__ void k_vocab(const int *__restrict__ present, const int *__restrict__ new_id, const int *__restrict__ df,
                        int32_t *__restrict__ vocab_codes, int *__restrict__ df_by_col) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ENC_CODES && present[c]) {
        vocab_codes[new_id[c]] = c;
        df_by_col[new_id[c]] = df[c];
    }
}

static int exclusive_sum_i64(Workspace &ws, const int64_t *in, int64_t *out, int64_t n) {
    size_t temp_bytes = 0;
    DS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, in, out, (int)n, ws.stream()));
    unsigned char *temp = nullptr;
    DS_CHECK(ws.alloc(&temp, temp_bytes));
    DS_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, ws.stream()));
    g_kernel_launches.fetch_add(1);
    return DS_OK;
}

// trigram CSR of one title collection: row_ptr [n + 1], cols (trigram codes, ascending per title)
static int encode_side(Workspace &ws, const uint8_t *d_bytes, const int64_t *d_offsets, int64_t n, int64_t *d_row_ptr, uint16_t *d_cols,
                       int *d_present, int *d_df, int *d_bad, int64_t *nnz) {
    cudaStream_t stream = ws.stream();
    *nnz = 0;
    if (n == 0) {
        DS_CUDA(cudaMemsetAsync(d_row_ptr, 0, 8, stream));
        return DS_OK;
    }
    int64_t *d_counts = nullptr;
    DS_CHECK(ws.alloc(&d_counts, (size_t)n + 1));
    DS_CUDA(cudaMemsetAsync(d_counts + n, 0, 8, stream));
    const unsigned blocks = (unsigned)ceil_div(n, ENC_WARPS);
    k_trigrams<0><<<blocks, ENC_WARPS * 32, 0, stream>>>(d_bytes, d_offsets, n, d_counts, nullptr, nullptr, nullptr, nullptr, d_bad);
    DS_LAUNCHED("k_trigrams<0>");
    DS_CHECK(exclusive_sum_i64(ws, d_counts, d_row_ptr, n + 1));
    k_trigrams<1><<<blocks, ENC_WARPS * 32, 0, stream>>>(d_bytes, d_offsets, n, nullptr, d_row_ptr, d_cols, d_present, d_df, d_bad);
    DS_LAUNCHED("k_trigrams<1>");
    DS_CUDA(cudaMemcpyAsync(nnz, d_row_ptr + n, 8, cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    return DS_OK;
}


// ---------------------------------------------------------------------------------------------------
// f3: transform_title (common.py:20-47) for a batch of titles given as Unicode code points.
//   unicodedata.normalize('NFD') + encode('ascii', 'ignore'): every code point contributes the ASCII
//   characters of its canonical decomposition - at most one (table `ascii_of_cp`, built by the caller from
//   its Unicode database; code points beyond the table contribute nothing) - then lower(), '-' -> ' ',
//   keep [a-zA-Z0-9\s], collapse runs of ' ', strip(), remember the length, [:255].strip(), and left-pad
//   with '0' to N_GRAMS = 3 characters when the remembered length is below 3.
// One thread per title (a title is a few dozen characters); PASS 0 yields the lengths, PASS 1 the bytes.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_py_space(int c) {   // str.isspace() / regex \s over ASCII
    return c == 32 || (c >= 9 && c <= 13) || (c >= 28 && c <= 31);
}

template <int PASS>
__global__ void k_transform(const uint32_t *__restrict__ cps, const int64_t *__restrict__ offsets, int64_t n_titles,
                            const uint8_t *__restrict__ ascii_of_cp, int table_len, int32_t *__restrict__ raw_len,
                            int64_t *__restrict__ out_len, const int64_t *__restrict__ out_offsets, uint8_t *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_titles) return;
    int limit = 0, pad = 0;
    uint8_t *dst = nullptr;
    if (PASS == 1) {
        const int n = raw_len[t];
        pad = n < 3 ? 3 - n : 0;
        limit = (int)(out_offsets[t + 1] - out_offsets[t]) - pad;
        dst = out + out_offsets[t];
        for (int i = 0; i < pad; ++i) dst[i] = '0';
        dst += pad;
    }
    int kept = 0;          // characters after the collapse and the left strip
    int last_solid = 0;    // kept characters up to and including the last non-space one
    int last_solid_255 = 0;
    int prev = -1;
    for (int64_t i = offsets[t]; i < offsets[t + 1]; ++i) {
        const uint32_t cp = cps[i];
        int c = cp < (uint32_t)table_len ? ascii_of_cp[cp] : 0;
        if (c == 0) continue;
        if (c >= 'A' && c <= 'Z') c += 32;
        if (c == '-') c = ' ';
        const bool space = is_py_space(c);
        if (!(space || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9'))) continue;
        if (c == ' ' && prev == ' ') continue;   // ' +' -> ' ' (other white space is kept as it is)
        prev = c;
        if (kept == 0 && space) continue;        // left strip
        if (PASS == 1 && kept < limit) dst[kept] = (uint8_t)c;
        ++kept;
        if (!space) {
            last_solid = kept;
            if (kept <= 255) last_solid_255 = kept;
        }
    }
    if (PASS == 0) {
        const int n = last_solid;                                  // len(text) after strip()
        const int text = n <= 255 ? n : last_solid_255;            // text[:255].strip()
        raw_len[t] = n;
        out_len[t] = n < 3 ? 3 : text;
    }
}

}  // namespace ds

using namespace ds;

// ---------------------------------------------------------------------------------------------------
// f3: the pair kernels' per-title inputs from the transformed-title table, on the device
//   FeatureEngineering.encode_title            feature_engineering.py:298-307  (38-symbol codes)
//   common.get_words_counter                   common.py:140-142               (document frequency over per-title word SETS)
//   FeatureEngineering.get_truth_words_counts  feature_engineering.py:309-319  (df of each of a title's first 15 words)
// Words are the tokens of str.split(): runs of non-white-space bytes.  A word is identified by its 64-bit FNV-1a hash
// (1.7M words at 500k titles: collision probability ~1e-7); every (title, word) pair counts once, like the set does.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_space_byte(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

// PASS 0: number of words of every title.  PASS 1: hash of every word, whether it is the first occurrence inside its
// title, and the hashes of the title's first DS_N_WORDS words.
template <int PASS>
__global__ void k_words(const uint8_t *__restrict__ bytes, const int64_t *__restrict__ offsets, int64_t n_titles, int64_t *__restrict__ n_words,
                        const int64_t *__restrict__ word_ptr, unsigned long long *__restrict__ keys, uint32_t *__restrict__ first_in_title,
                        unsigned long long *__restrict__ head) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_titles) return;
    const uint8_t *p = bytes + offsets[t];
    const int len = (int)(offsets[t + 1] - offsets[t]);
    int64_t at = PASS == 1 ? word_ptr[t] : 0;
    const int64_t first = at;
    int count = 0;
    int i = 0;
    while (i < len) {
        while (i < len && is_space_byte(p[i])) ++i;
        if (i >= len) break;
        unsigned long long h = 14695981039346656037ull;
        while (i < len && !is_space_byte(p[i])) {
            h = (h ^ p[i]) * 1099511628211ull;
            ++i;
        }
        if (PASS == 1) {
            bool seen = false;
            for (int64_t j = first; j < at; ++j) seen |= keys[j] == h;
            keys[at] = h;
            first_in_title[at] = seen ? 0u : 1u;
            if (count < DS_N_WORDS) head[t * DS_N_WORDS + count] = h;
            ++at;
        }
        ++count;
    }
    if (PASS == 0) n_words[t] = count;
    else
        for (int c = count; c < DS_N_WORDS; ++c) head[t * DS_N_WORDS + c] = 0ull;
}

// counts[t][slot] = document frequency of the title's slot-th word (binary search among the distinct hashes), 0 past its words
__global__ void k_word_lookup(const unsigned long long *__restrict__ head, const int64_t *__restrict__ n_words, int64_t n_titles,
                              const unsigned long long *__restrict__ distinct, const uint32_t *__restrict__ df, const int *__restrict__ n_distinct,
                              uint32_t *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_titles * DS_N_WORDS) return;
    const int64_t t = i / DS_N_WORDS;
    const int slot = (int)(i % DS_N_WORDS);
    uint32_t value = 0;
    if (slot < n_words[t]) {
        const unsigned long long h = head[i];
        int lo = 0, hi = *n_distinct;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (distinct[mid] < h) lo = mid + 1;
            else hi = mid;
        }
        value = df[lo];
    }
    counts[i] = value;
}

__global__ void k_title_codes(const uint8_t *__restrict__ bytes, int64_t n, uint8_t *__restrict__ codes, int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t c = bytes[i];
    uint8_t code = 255;
    if (c >= 'a' && c <= 'z') code = (uint8_t)(c - 'a' + 2);
    else if (c >= '0' && c <= '9') code = (uint8_t)(c - '0' + 28);
    else if (c == ' ') code = 1;
    else if (c == '-') code = 0;
    if (code == 255) atomicOr(bad, 1);
    codes[i] = code;
}

extern "C" {

int32_t ds_encode_max_vocab(void) { return ENC_CODES; }

int ds_encode_trigrams(const uint8_t *truth_bytes, const int64_t *truth_offsets, int64_t n_truth, const uint8_t *query_bytes,
                       const int64_t *query_offsets, int64_t n_queries, int64_t *t_row_ptr, uint16_t *t_col_ids,
                       int64_t *q_row_ptr, uint16_t *q_col_ids, double *idf64_by_col, int32_t *vocab_codes, int32_t *out_n_vocab,
                       int64_t *out_truth_nnz, int64_t *out_query_nnz, int device, void *stream_) {
    if (n_truth < 1 || n_queries < 0) return fail(DS_ERR_BAD_ARG, "n_truth < 1 or n_queries < 0");
    if (!truth_bytes || !truth_offsets || !t_row_ptr || !t_col_ids || !q_row_ptr || !idf64_by_col || !vocab_codes || !out_n_vocab ||
        !out_truth_nnz || !out_query_nnz)
        return fail(DS_ERR_BAD_ARG, "NULL argument");
    if (n_queries > 0 && (!query_bytes || !query_offsets || !q_col_ids)) return fail(DS_ERR_BAD_ARG, "NULL query argument");
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(DS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    DeviceGuard guard(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);

    // byte totals (to size the staging of host inputs and the capacity of the outputs)
    auto last_offset = [&](const int64_t *offsets, int64_t n, int64_t *total) -> int {
        if (is_device_pointer(offsets)) {
            DS_CUDA(cudaMemcpyAsync(total, offsets + n, 8, cudaMemcpyDeviceToHost, stream));
            DS_CUDA(cudaStreamSynchronize(stream));
        } else {
            *total = offsets[n];
        }
        return DS_OK;
    };
    int64_t truth_total = 0, query_total = 0;
    DS_CHECK(last_offset(truth_offsets, n_truth, &truth_total));
    if (n_queries > 0) DS_CHECK(last_offset(query_offsets, n_queries, &query_total));

    const uint8_t *d_tb = nullptr, *d_qb = nullptr;
    const int64_t *d_to = nullptr, *d_qo = nullptr;
    DS_CHECK(ws.stage_in(&d_tb, truth_bytes, (size_t)std::max<int64_t>(1, truth_total)));
    DS_CHECK(ws.stage_in(&d_to, truth_offsets, (size_t)n_truth + 1));
    if (n_queries > 0) {
        DS_CHECK(ws.stage_in(&d_qb, query_bytes, (size_t)std::max<int64_t>(1, query_total)));
        DS_CHECK(ws.stage_in(&d_qo, query_offsets, (size_t)n_queries + 1));
    }
    // outputs: capacity of the column arrays = total title bytes (a title of L characters has <= L - 2 trigrams)
    int64_t *d_tp = nullptr, *d_qp = nullptr;
    uint16_t *d_tc = nullptr, *d_qc = nullptr;
    double *d_idf = nullptr;
    int32_t *d_vocab = nullptr;
    DS_CHECK(ws.stage_out(&d_tp, t_row_ptr, (size_t)n_truth + 1));
    DS_CHECK(ws.stage_out(&d_tc, t_col_ids, (size_t)std::max<int64_t>(1, truth_total)));
    DS_CHECK(ws.stage_out(&d_qp, q_row_ptr, (size_t)n_queries + 1));
    if (n_queries > 0) DS_CHECK(ws.stage_out(&d_qc, q_col_ids, (size_t)std::max<int64_t>(1, query_total)));
    DS_CHECK(ws.stage_out(&d_idf, idf64_by_col, (size_t)ENC_CODES));
    DS_CHECK(ws.stage_out(&d_vocab, vocab_codes, (size_t)ENC_CODES));

    int *d_present = nullptr, *d_df = nullptr, *d_bad = nullptr, *d_new_id = nullptr, *d_df_by_col = nullptr;
    DS_CHECK(ws.alloc(&d_present, ENC_CODES + 1));
    DS_CHECK(ws.alloc(&d_df, ENC_CODES));
    DS_CHECK(ws.alloc(&d_bad, 1));
    DS_CHECK(ws.alloc(&d_new_id, ENC_CODES + 1));
    DS_CHECK(ws.alloc(&d_df_by_col, ENC_CODES));
    DS_CUDA(cudaMemsetAsync(d_present, 0, (ENC_CODES + 1) * sizeof(int), stream));
    DS_CUDA(cudaMemsetAsync(d_df, 0, ENC_CODES * sizeof(int), stream));
    DS_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));

    int64_t truth_nnz = 0, query_nnz = 0;
    DS_CHECK(encode_side(ws, d_tb, d_to, n_truth, d_tp, d_tc, d_present, d_df, d_bad, &truth_nnz));
    DS_CHECK(encode_side(ws, d_qb, d_qo, n_queries, d_qp, d_qc, d_present, nullptr, d_bad, &query_nnz));

    // canonical column ids = ranks of the present trigram codes
    {
        size_t temp_bytes = 0;
        DS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_present, d_new_id, ENC_CODES + 1, stream));
        unsigned char *temp = nullptr;
        DS_CHECK(ws.alloc(&temp, temp_bytes));
        DS_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, d_present, d_new_id, ENC_CODES + 1, stream));
        g_kernel_launches.fetch_add(1);
    }
    int h_bad = 0, n_vocab = 0;
    DS_CUDA(cudaMemcpyAsync(&n_vocab, d_new_id + ENC_CODES, sizeof(int), cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    if (h_bad) return fail(DS_ERR_BAD_ARG, "a title holds a character outside ' a-z0-9' (titles must be transform_title output)");
    if (n_vocab > 65535) return fail(DS_ERR_UNSUPPORTED, "%d distinct trigrams exceed the 65535 u16 column ids", n_vocab);
    if (truth_nnz > 0) {
        k_remap<<<(unsigned)ceil_div(truth_nnz, 256), 256, 0, stream>>>(d_tc, truth_nnz, d_new_id);
        DS_LAUNCHED("k_remap");
    }
    if (query_nnz > 0) {
        k_remap<<<(unsigned)ceil_div(query_nnz, 256), 256, 0, stream>>>(d_qc, query_nnz, d_new_id);
        DS_LAUNCHED("k_remap");
    }
    k_vocab<<<(unsigned)ceil_div(ENC_CODES, 256), 256, 0, stream>>>(d_present, d_new_id, d_df, d_vocab, d_df_by_col);
    DS_LAUNCHED("k_vocab");

    // idf on the host with the C library's log, exactly what CPython's math.log(N / df) evaluates (match_maker.py:139)
    std::vector<int> h_df((size_t)std::max(1, n_vocab));
    DS_CUDA(cudaMemcpyAsync(h_df.data(), d_df_by_col, (size_t)n_vocab * sizeof(int), cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    std::vector<double> h_idf((size_t)std::max(1, n_vocab));
    double max_idf = 0.0;
    bool any = false;
    for (int c = 0; c < n_vocab; ++c) {
        if (h_df[(size_t)c] > 0) {
            h_idf[(size_t)c] = std::log((double)n_truth / (double)h_df[(size_t)c]);
            max_idf = any ? std::max(max_idf, h_idf[(size_t)c]) : h_idf[(size_t)c];
            any = true;
        }
    }
    for (int c = 0; c < n_vocab; ++c)
        if (h_df[(size_t)c] == 0) h_idf[(size_t)c] = max_idf;                    // query-only trigram (match_maker.py:151)
    DS_CUDA(cudaMemcpyAsync(d_idf, h_idf.data(), (size_t)n_vocab * sizeof(double), cudaMemcpyHostToDevice, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    *out_n_vocab = n_vocab;
    *out_truth_nnz = truth_nnz;
    *out_query_nnz = query_nnz;
    return ws.finish_outputs();
}

int ds_transform_titles(const uint32_t *codepoints, const int64_t *offsets, int64_t n_titles, const uint8_t *ascii_of_cp,
                        int32_t table_len, uint8_t *out_bytes, int64_t *out_offsets, int32_t *out_raw_len, int device,
                        void *stream_) {
    if (n_titles < 0 || (n_titles > 0 && offsets == nullptr)) return fail(DS_ERR_BAD_ARG, "bad n_titles / offsets");
    if (ascii_of_cp == nullptr || table_len < 128) return fail(DS_ERR_BAD_ARG, "ascii_of_cp must cover at least the ASCII range");
    if (out_bytes == nullptr || out_offsets == nullptr) return fail(DS_ERR_BAD_ARG, "out_bytes / out_offsets is NULL");
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(DS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    DeviceGuard guard(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    int64_t total = 0;
    if (n_titles > 0) {
        if (is_device_pointer(offsets)) {
            DS_CUDA(cudaMemcpyAsync(&total, offsets + n_titles, 8, cudaMemcpyDeviceToHost, stream));
            DS_CUDA(cudaStreamSynchronize(stream));
        } else {
            total = offsets[n_titles];
        }
    }
    if (total < 0) return fail(DS_ERR_BAD_ARG, "offsets must be non-decreasing from 0");
    const uint32_t *d_cps = nullptr;
    const int64_t *d_off = nullptr;
    const uint8_t *d_table = nullptr;
    DS_CHECK(ws.stage_in(&d_cps, codepoints, (size_t)std::max<int64_t>(1, total)));
    DS_CHECK(ws.stage_in(&d_off, offsets, (size_t)n_titles + 1));
    DS_CHECK(ws.stage_in(&d_table, ascii_of_cp, (size_t)table_len));
    uint8_t *d_out = nullptr;
    int64_t *d_out_off = nullptr, *d_len = nullptr;
    int32_t *d_raw = nullptr;
    DS_CHECK(ws.stage_out(&d_out, out_bytes, (size_t)(total + 3 * n_titles + 1)));   // a title never grows beyond max(len, 3)
    DS_CHECK(ws.stage_out(&d_out_off, out_offsets, (size_t)n_titles + 1));
    DS_CHECK(ws.stage_out(&d_raw, out_raw_len, (size_t)n_titles));
    if (d_raw == nullptr) DS_CHECK(ws.alloc(&d_raw, (size_t)std::max<int64_t>(1, n_titles)));
    DS_CHECK(ws.alloc(&d_len, (size_t)n_titles + 1));
    DS_CUDA(cudaMemsetAsync(d_len, 0, ((size_t)n_titles + 1) * 8, stream));
    if (n_titles > 0) {
        k_transform<0><<<(unsigned)ceil_div(n_titles, 128), 128, 0, stream>>>(d_cps, d_off, n_titles, d_table, table_len, d_raw, d_len,
                                                                            nullptr, nullptr);
        DS_LAUNCHED("k_transform");
    }
    size_t temp_bytes = 0;
    DS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_len, d_out_off, (int)(n_titles + 1), stream));
    unsigned char *temp = nullptr;
    DS_CHECK(ws.alloc(&temp, temp_bytes));
    DS_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, d_len, d_out_off, (int)(n_titles + 1), stream));
    g_kernel_launches.fetch_add(1);
    if (n_titles > 0) {
        k_transform<1><<<(unsigned)ceil_div(n_titles, 128), 128, 0, stream>>>(d_cps, d_off, n_titles, d_table, table_len, d_raw, nullptr,
                                                                            d_out_off, d_out);
        DS_LAUNCHED("k_transform");
    }
    return ws.finish_outputs();
}

int ds_title_features(const uint8_t *bytes, const int64_t *offsets, int64_t n_titles, uint8_t *out_codes, uint32_t *out_word_counts,
                      int device, void *stream_) {
    if (n_titles < 0) return fail(DS_ERR_BAD_ARG, "n_titles < 0");
    if (n_titles == 0) return DS_OK;
    if (!bytes || !offsets) return fail(DS_ERR_BAD_ARG, "NULL argument");
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(DS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    DeviceGuard guard(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    int64_t total = 0;
    if (is_device_pointer(offsets)) {
        DS_CUDA(cudaMemcpyAsync(&total, offsets + n_titles, 8, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
    } else {
        total = offsets[n_titles];
    }
    const uint8_t *d_bytes = nullptr;
    const int64_t *d_offsets = nullptr;
    DS_CHECK(ws.stage_in(&d_bytes, bytes, (size_t)std::max<int64_t>(1, total)));
    DS_CHECK(ws.stage_in(&d_offsets, offsets, (size_t)n_titles + 1));
    if (out_codes != nullptr && total > 0) {
        uint8_t *d_codes = nullptr;
        int *d_bad = nullptr;
        DS_CHECK(ws.stage_out(&d_codes, out_codes, (size_t)total));
        DS_CHECK(ws.alloc(&d_bad, 1));
        DS_CUDA(cudaMemsetAsync(d_bad, 0, 4, stream));
        k_title_codes<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(d_bytes, total, d_codes, d_bad);
        DS_LAUNCHED("k_title_codes");
        int h_bad = 0;
        DS_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
        if (h_bad) return fail(DS_ERR_BAD_ARG, "a title holds a character outside '- a-z0-9' (encode_title would fail on it)");
    }
    if (out_word_counts != nullptr) {
        int64_t *d_n_words = nullptr, *d_word_ptr = nullptr;
        DS_CHECK(ws.alloc(&d_n_words, (size_t)n_titles + 1));
        DS_CHECK(ws.alloc(&d_word_ptr, (size_t)n_titles + 1));
        DS_CUDA(cudaMemsetAsync(d_n_words, 0, ((size_t)n_titles + 1) * 8, stream));
        const unsigned blocks = (unsigned)ceil_div(n_titles, 128);
        k_words<0><<<blocks, 128, 0, stream>>>(d_bytes, d_offsets, n_titles, d_n_words, nullptr, nullptr, nullptr, nullptr);
        DS_LAUNCHED("k_words");
        DS_CHECK(exclusive_sum_i64(ws, d_n_words, d_word_ptr, n_titles + 1));
        int64_t n_words = 0;
        DS_CUDA(cudaMemcpyAsync(&n_words, d_word_ptr + n_titles, 8, cudaMemcpyDeviceToHost, stream));
        DS_CUDA(cudaStreamSynchronize(stream));
        if (n_words > INT32_MAX) return fail(DS_ERR_UNSUPPORTED, "more than 2^31-1 words");
        unsigned long long *d_keys = nullptr, *d_keys_sorted = nullptr, *d_head = nullptr, *d_distinct = nullptr;
        uint32_t *d_first = nullptr, *d_first_sorted = nullptr, *d_df = nullptr, *d_counts = nullptr;
        int *d_n_distinct = nullptr;
        const size_t words_alloc = (size_t)std::max<int64_t>(1, n_words);
        DS_CHECK(ws.alloc(&d_keys, words_alloc));
        DS_CHECK(ws.alloc(&d_keys_sorted, words_alloc));
        DS_CHECK(ws.alloc(&d_first, words_alloc));
        DS_CHECK(ws.alloc(&d_first_sorted, words_alloc));
        DS_CHECK(ws.alloc(&d_distinct, words_alloc + 1));
        DS_CHECK(ws.alloc(&d_df, words_alloc + 1));
        DS_CHECK(ws.alloc(&d_n_distinct, 1));
        DS_CHECK(ws.alloc(&d_head, (size_t)n_titles * DS_N_WORDS));
        DS_CHECK(ws.stage_out(&d_counts, out_word_counts, (size_t)n_titles * DS_N_WORDS));
        DS_CUDA(cudaMemsetAsync(d_n_distinct, 0, 4, stream));
        DS_CUDA(cudaMemsetAsync(d_df, 0, (words_alloc + 1) * 4, stream));
        k_words<1><<<blocks, 128, 0, stream>>>(d_bytes, d_offsets, n_titles, nullptr, d_word_ptr, d_keys, d_first, d_head);
        DS_LAUNCHED("k_words");
        if (n_words > 0) {
            size_t sort_bytes = 0, reduce_bytes = 0;
            DS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys_sorted, d_first, d_first_sorted, (int)n_words, 0, 64, stream));
            DS_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, reduce_bytes, d_keys_sorted, d_distinct, d_first_sorted, d_df, d_n_distinct,
                                                   ::cuda::std::plus<>(), (int)n_words, stream));
            unsigned char *d_temp = nullptr;
            DS_CHECK(ws.alloc(&d_temp, std::max(sort_bytes, reduce_bytes)));
            DS_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, sort_bytes, d_keys, d_keys_sorted, d_first, d_first_sorted, (int)n_words, 0, 64, stream));
            DS_CUDA(cub::DeviceReduce::ReduceByKey(d_temp, reduce_bytes, d_keys_sorted, d_distinct, d_first_sorted, d_df, d_n_distinct,
                                                   ::cuda::std::plus<>(), (int)n_words, stream));
            g_kernel_launches.fetch_add(2);
        }
        k_word_lookup<<<(unsigned)ceil_div(n_titles * DS_N_WORDS, 256), 256, 0, stream>>>(d_head, d_n_words, n_titles, d_distinct, d_df,
                                                                                          d_n_distinct, d_counts);
        DS_LAUNCHED("k_word_lookup");
    }
    return ws.finish_outputs();
}

}  // extern "C"
