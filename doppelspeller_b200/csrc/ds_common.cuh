// ds_common.cuh - shared plumbing of the C-ABI library: status/error handling, stream-ordered
// workspace, host/device pointer staging, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/doppelspeller_b200.h"

namespace ds {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_kernel_launches;

inline int fail(int status, const char *fmt, ...) {
    va_list args;
    va_start(args, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, args);
    va_end(args);
    return status;
}

#define DS_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            return ds::fail(DS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                            __FILE__, __LINE__);                                                   \
    } while (0)

#define DS_CHECK(expr)                      \
    do {                                    \
        int status__ = (expr);              \
        if (status__ != DS_OK) return status__; \
    } while (0)

#define DS_LAUNCHED(name)                                                                         \
    do {                                                                                          \
        ds::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                            \
        cudaError_t err__ = cudaGetLastError();                                                   \
        if (err__ != cudaSuccess)                                                                 \
            return ds::fail(DS_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(err__)); \
    } while (0)

// true when `p` can be dereferenced by a kernel running on the current device
inline bool is_device_pointer(const void *p) {
    if (p == nullptr) return false;
    cudaPointerAttributes attr;
    cudaError_t err = cudaPointerGetAttributes(&attr, p);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// device that owns a device pointer, -1 for host / unknown pointers
inline int pointer_device(const void *p) {
    if (p == nullptr) return -1;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? attr.device : -1;
}

// first device found among `pointers`: the device the call must run on (-1: all host buffers -> current device)
inline int owning_device(std::initializer_list<const void *> pointers) {
    for (const void *p : pointers) {
        const int d = pointer_device(p);
        if (d >= 0) return d;
    }
    return -1;
}

// Stream-ordered device allocation released in the destructor (cudaFreeAsync on the same stream).
class Workspace {
   public:
    explicit Workspace(cudaStream_t stream) : stream_(stream) { keep_pool_memory(); }
    ~Workspace() {
        for (void *p : blocks_) cudaFreeAsync(p, stream_);
    }
    Workspace(const Workspace &) = delete;
    Workspace &operator=(const Workspace &) = delete;

    template <typename T>
    int alloc(T **out, size_t count) {
        void *p = nullptr;
        // rounded up to 16 bytes: the string kernels fetch whole aligned 32-bit words around a string
        size_t bytes = (((count > 0 ? count : 1) * sizeof(T)) + 15) & ~(size_t)15;
        cudaError_t err = cudaMallocAsync(&p, bytes, stream_);
        if (err != cudaSuccess) {
            cudaGetLastError();
            return fail(err == cudaErrorMemoryAllocation ? DS_ERR_NO_MEMORY : DS_ERR_CUDA,
                        "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
        }
        blocks_.push_back(p);
        *out = static_cast<T *>(p);
        return DS_OK;
    }

    // Returns a device-usable pointer to `count` elements at `src`: `src` itself when it already is a
    // device pointer, otherwise a staged copy (H2D on the stream).
    template <typename T>
    int stage_in(const T **out, const T *src, size_t count) {
        if (src == nullptr || count == 0 || is_device_pointer(src)) {
            *out = src;
            return DS_OK;
        }
        T *dst = nullptr;
        DS_CHECK(alloc(&dst, count));
        DS_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, stream_));
        *out = dst;
        return DS_OK;
    }

    // Output staging: returns a device pointer to write to; `finish_outputs` copies back to host
    // destinations and reports whether a synchronisation is needed.
    template <typename T>
    int stage_out(T **out, T *dst, size_t count) {
        if (dst == nullptr || is_device_pointer(dst)) {
            *out = dst;
            return DS_OK;
        }
        T *tmp = nullptr;
        DS_CHECK(alloc(&tmp, count));
        pending_.push_back({dst, tmp, count * sizeof(T)});
        *out = tmp;
        return DS_OK;
    }

    int finish_outputs() {
        for (const Pending &p : pending_)
            DS_CUDA(cudaMemcpyAsync(p.host, p.device, p.bytes, cudaMemcpyDeviceToHost, stream_));
        if (!pending_.empty()) DS_CUDA(cudaStreamSynchronize(stream_));
        pending_.clear();
        return DS_OK;
    }

    cudaStream_t stream() const { return stream_; }

   private:
    // By default the stream-ordered pool hands freed memory back to the OS at every synchronisation, which
    // would make each call re-map its whole workspace; keep it cached in the pool instead.  The cached bytes are
    // bounded by DS_POOL_RELEASE_THRESHOLD_MB (environment, default: unbounded) and ds_trim() hands them back.
    static void keep_pool_memory() {
        static thread_local int done_for_device = -1;
        int device = 0;
        if (cudaGetDevice(&device) != cudaSuccess || device == done_for_device) return;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t threshold = UINT64_MAX;
            if (const char *mb = getenv("DS_POOL_RELEASE_THRESHOLD_MB")) threshold = (uint64_t)strtoull(mb, nullptr, 10) << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        done_for_device = device;
    }

    struct Pending {
        void *host;
        void *device;
        size_t bytes;
    };
    cudaStream_t stream_;
    std::vector<void *> blocks_;
    std::vector<Pending> pending_;
};

// Sets the device for the scope of a call and restores the previous one.
class DeviceGuard {
   public:
    explicit DeviceGuard(int device) {
        cudaGetDevice(&previous_);
        if (device >= 0 && device != previous_) {
            cudaSetDevice(device);
            changed_ = true;
        }
    }
    ~DeviceGuard() {
        if (changed_) cudaSetDevice(previous_);
    }

   private:
    int previous_ = 0;
    bool changed_ = false;
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Opt-in dynamic shared memory above 48 KB is a per-device attribute of a kernel: remembers, per kernel and
// per device, the largest size already granted.
inline int ensure_dynamic_smem(const void *kernel, size_t bytes) {
    static std::mutex lock;
    static std::map<std::pair<int, const void *>, size_t> granted;
    if (bytes <= 48 * 1024) return DS_OK;
    int device = 0;
    DS_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> guard(lock);
    size_t &have = granted[std::make_pair(device, kernel)];
    if (bytes > have) {
        DS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return DS_OK;
}

}  // namespace ds
