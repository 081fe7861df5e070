// TEST INFRASTRUCTURE ONLY: sequential host versions of the CUB device-wide primitives the library calls, with CUB's
// two-phase calling convention (a null temporary buffer asks for its size).  The temporary buffer is written from
// end to end, so AddressSanitizer notices a caller that hands over less than it was asked for.
#pragma once
#include "../../cuda_emu.h"

namespace cub {

inline size_t emu_temp_bytes(size_t n, size_t per_item) { return 256 + n * per_item; }

struct DeviceScan {
    template <typename In, typename Out>
    static cudaError_t ExclusiveSum(void *temp, size_t &temp_bytes, In in, Out out, int n, cudaStream_t = nullptr) {
        if (temp == nullptr) {
            temp_bytes = emu_temp_bytes((size_t)(n > 0 ? n : 0), 4);
            return cudaSuccess;
        }
        memset(temp, 0xEE, temp_bytes);
        typedef typename std::remove_reference<decltype(out[0])>::type value_t;
        value_t running = value_t();
        for (int i = 0; i < n; ++i) {      // in == out is allowed
            const value_t x = (value_t)in[i];
            out[i] = running;
            running = (value_t)(running + x);
        }
        return cudaSuccess;
    }
};

struct DeviceRadixSort {
    // stable sort of (key, value) pairs by the key bits [begin_bit, end_bit)
    template <typename K, typename V>
    static cudaError_t SortPairs(void *temp, size_t &temp_bytes, const K *keys_in, K *keys_out, const V *values_in, V *values_out, int n,
                                 int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
        if (temp == nullptr) {
            temp_bytes = emu_temp_bytes((size_t)(n > 0 ? n : 0), sizeof(K) + sizeof(V));
            return cudaSuccess;
        }
        memset(temp, 0xEE, temp_bytes);
        if (begin_bit < 0 || end_bit > (int)sizeof(K) * 8 || begin_bit > end_bit) return cudaErrorInvalidValue;
        typedef typename std::make_unsigned<K>::type U;
        const int bits = end_bit - begin_bit;
        const U field = bits >= (int)sizeof(K) * 8 ? ~(U)0 : (U)((((U)1) << bits) - 1);
        std::vector<int> order((size_t)(n > 0 ? n : 0));
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return (((U)keys_in[a] >> begin_bit) & field) < (((U)keys_in[b] >> begin_bit) & field);
        });
        for (int i = 0; i < n; ++i) {
            keys_out[i] = keys_in[order[(size_t)i]];
            values_out[i] = values_in[order[(size_t)i]];
        }
        return cudaSuccess;
    }
};

struct DeviceReduce {
    template <typename KeyIn, typename KeyOut, typename ValIn, typename ValOut, typename Count, typename Op>
    static cudaError_t ReduceByKey(void *temp, size_t &temp_bytes, KeyIn keys_in, KeyOut unique_out, ValIn values_in, ValOut aggregates_out,
                                   Count num_runs_out, Op op, int n, cudaStream_t = nullptr) {
        if (temp == nullptr) {
            temp_bytes = emu_temp_bytes((size_t)(n > 0 ? n : 0), 8);
            return cudaSuccess;
        }
        memset(temp, 0xEE, temp_bytes);
        int runs = 0;
        for (int i = 0; i < n; ++i) {
            if (i > 0 && keys_in[i] == keys_in[i - 1]) {
                aggregates_out[runs - 1] = op(aggregates_out[runs - 1], values_in[i]);
            } else {
                unique_out[runs] = keys_in[i];
                aggregates_out[runs] = values_in[i];
                ++runs;
            }
        }
        *num_runs_out = runs;
        return cudaSuccess;
    }
};

}  // namespace cub
