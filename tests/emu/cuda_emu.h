// cuda_emu.h - TEST INFRASTRUCTURE ONLY.  A host stand-in for the CUDA runtime and the device-side language
// subset the kernels of doppelspeller_b200/csrc use, so that the UNMODIFIED kernel sources can be compiled with
// g++ -fsanitize=address,undefined and executed thread by thread on the CPU (tests/emu/README.md).
//
// Why: compute-sanitizer is closed on the GPU pool this repository is measured on, so out-of-bounds shared /
// global memory accesses, misaligned vector loads, reads of freed workspaces, mismatched warp collectives and
// barrier deadlocks inside the kernels would otherwise only show up as wrong parity results.  Here every CUDA
// thread is a fiber with its own stack; __syncthreads / __syncwarp / shuffles / votes are rendezvous points between
// the fibers of a CTA; device allocations are plain malloc blocks (filled with garbage), so AddressSanitizer sees
// every kernel access with exact bounds.
//
// This is NOT a CPU fallback of the product: nothing under doppelspeller_b200/ refers to it, it is built only by
// tests/test_emulated_kernels.py into tests/emu/_build/ (git-ignored) and loaded only by that test's subprocess.
#pragma once

// every system header the translated sources may ask for comes first: the qualifier macros below would break
// __attribute__((__noinline__)) and friends inside them
#include <fenv.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <limits>
#include <map>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// CUDA puts the classification functions into the global namespace
using std::isfinite;
using std::isinf;
using std::isnan;

// ------------------------------------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static            /* one CTA runs at a time: a static is the CTA's shared variable */
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------------------------------------ vector types
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };

// ------------------------------------------------------------------------------------------------ runtime API
typedef int cudaError_t;
enum {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInvalidConfiguration = 9,
    cudaErrorInvalidDevice = 101,
    cudaErrorLaunchFailure = 719,
};
typedef struct ds_emu_stream *cudaStream_t;
typedef struct ds_emu_event *cudaEvent_t;
typedef struct ds_emu_pool *cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
    void *devicePointer;
    void *hostPointer;
};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };

cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaGetDevice(int *);
cudaError_t cudaSetDevice(int);
cudaError_t cudaMallocAsync(void **, size_t, cudaStream_t);
cudaError_t cudaFreeAsync(void *, cudaStream_t);
cudaError_t cudaFree(void *);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *, const void *);
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *, int);
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *);
cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t);
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
template <typename T>
inline cudaError_t cudaFuncSetAttribute(T *fn, cudaFuncAttribute attr, int value) {
    return cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), attr, value);
}

// ------------------------------------------------------------------------------------------------ execution engine
namespace ds_emu {

enum Op {
    OP_SYNCWARP, OP_BALLOT, OP_ALL, OP_ANY, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR,
    OP_RED_MAX_U, OP_RED_MAX_S, OP_RED_MIN_U, OP_RED_MIN_S, OP_RED_ADD, OP_RED_OR, OP_RED_AND, OP_RED_XOR,
};

struct ThreadCtx {
    uint3 thread_idx;
};
extern ThreadCtx *g_cur;          // the CUDA thread that is running
extern uint3 g_block_idx;
extern dim3 g_block_dim, g_grid_dim;
extern unsigned char *g_dyn_smem; // dynamic shared memory of the running CTA (exactly the launch's byte count)

// runs `body` once per CUDA thread of the grid (CTAs one after the other, the threads of a CTA interleaved at their
// synchronisation points).  `fn` identifies the kernel for the opt-in shared-memory check.
void launch(const void *fn, const char *name, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const std::function<void()> &body);
uint64_t warp_collective(int op, unsigned mask, uint64_t value, int arg, int width);
void cta_barrier();
inline unsigned char *dynamic_smem() { return g_dyn_smem; }

template <typename T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "collective payloads are at most 64 bits");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    return bits;
}
template <typename T>
inline T from_bits(uint64_t bits) {
    T v;
    memcpy(&v, &bits, sizeof(T));
    return v;
}

template <int MODE, typename F>
inline auto rounded(F f) -> decltype(f()) {
    const int old = fegetround();
    fesetround(MODE);
    auto r = f();
    fesetround(old);
    return r;
}
}  // namespace ds_emu

#define threadIdx (::ds_emu::g_cur->thread_idx)
#define blockIdx (::ds_emu::g_block_idx)
#define blockDim (::ds_emu::g_block_dim)
#define gridDim (::ds_emu::g_grid_dim)
static const int warpSize = 32;

// ------------------------------------------------------------------------------------------------ synchronisation, votes, shuffles
inline void __syncthreads() { ::ds_emu::cta_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { ::ds_emu::warp_collective(::ds_emu::OP_SYNCWARP, mask, 0, 0, 32); }
inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_BALLOT, mask, pred != 0, 0, 32); }
inline int __all_sync(unsigned mask, int pred) { return (int)::ds_emu::warp_collective(::ds_emu::OP_ALL, mask, pred != 0, 0, 32); }
inline int __any_sync(unsigned mask, int pred) { return (int)::ds_emu::warp_collective(::ds_emu::OP_ANY, mask, pred != 0, 0, 32); }
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return ::ds_emu::from_bits<T>(::ds_emu::warp_collective(::ds_emu::OP_SHFL, mask, ::ds_emu::to_bits(v), src, width));
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return ::ds_emu::from_bits<T>(::ds_emu::warp_collective(::ds_emu::OP_SHFL_UP, mask, ::ds_emu::to_bits(v), (int)delta, width));
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return ::ds_emu::from_bits<T>(::ds_emu::warp_collective(::ds_emu::OP_SHFL_DOWN, mask, ::ds_emu::to_bits(v), (int)delta, width));
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    return ::ds_emu::from_bits<T>(::ds_emu::warp_collective(::ds_emu::OP_SHFL_XOR, mask, ::ds_emu::to_bits(v), lane_mask, width));
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_RED_MAX_U, mask, v, 0, 32); }
inline int __reduce_max_sync(unsigned mask, int v) { return (int)::ds_emu::warp_collective(::ds_emu::OP_RED_MAX_S, mask, (uint64_t)(int64_t)v, 0, 32); }
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_RED_MIN_U, mask, v, 0, 32); }
inline int __reduce_min_sync(unsigned mask, int v) { return (int)::ds_emu::warp_collective(::ds_emu::OP_RED_MIN_S, mask, (uint64_t)(int64_t)v, 0, 32); }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_RED_ADD, mask, v, 0, 32); }
inline int __reduce_add_sync(unsigned mask, int v) { return (int)::ds_emu::warp_collective(::ds_emu::OP_RED_ADD, mask, (unsigned)v, 0, 32); }
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_RED_OR, mask, v, 0, 32); }
inline unsigned __reduce_and_sync(unsigned mask, unsigned v) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_RED_AND, mask, v, 0, 32); }
inline unsigned __reduce_xor_sync(unsigned mask, unsigned v) { return (unsigned)::ds_emu::warp_collective(::ds_emu::OP_RED_XOR, mask, v, 0, 32); }

// ------------------------------------------------------------------------------------------------ atomics (one fiber runs at a time)
// (relaxed __atomic builtins: the same code for one fiber at a time, and ThreadSanitizer knows the access is atomic)
template <typename T, typename U>
inline T atomicAdd(T *p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename U>
inline T atomicSub(T *p, U v) { return __atomic_fetch_sub(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename U>
inline T atomicOr(T *p, U v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename U>
inline T atomicAnd(T *p, U v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename U>
inline T atomicExch(T *p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename U, typename V>
inline T atomicCAS(T *p, U compare, V v) {
    T expected = (T)compare;
    __atomic_compare_exchange_n(p, &expected, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expected;
}
template <typename T, typename U>
inline T atomicMax(T *p, U v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while ((T)v > old && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <typename T, typename U>
inline T atomicMin(T *p, U v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while ((T)v < old && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ------------------------------------------------------------------------------------------------ loads, bit tricks
template <typename T>
inline T __ldg(const T *p) { return *p; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (shift & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (shift & 31)) >> 32); }
inline unsigned __vmaxu2(unsigned a, unsigned b) {
    return std::max(a & 0xffffu, b & 0xffffu) | (std::max(a >> 16, b >> 16) << 16);
}
inline unsigned __vminu2(unsigned a, unsigned b) {
    return std::min(a & 0xffffu, b & 0xffffu) | (std::min(a >> 16, b >> 16) << 16);
}
inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c) { return __vminu2(__vminu2(a, b), c); }

inline float __int_as_float(int x) { return ::ds_emu::from_bits<float>((uint32_t)x); }
inline float __uint_as_float(unsigned x) { return ::ds_emu::from_bits<float>(x); }
inline int __float_as_int(float x) { return (int)::ds_emu::to_bits(x); }
inline unsigned __float_as_uint(float x) { return (unsigned)::ds_emu::to_bits(x); }
inline double __longlong_as_double(long long x) { return ::ds_emu::from_bits<double>((uint64_t)x); }
inline long long __double_as_longlong(double x) { return (long long)::ds_emu::to_bits(x); }

// ------------------------------------------------------------------------------------------------ arithmetic with stated rounding
// (the sources are compiled with -ffp-contract=off -frounding-math: plain operators are IEEE round-to-nearest, never fused)
#define DS_EMU_BINARY(name, type, mode, expr)                                                   \
    inline type name(type a, type b) {                                                          \
        return ::ds_emu::rounded<mode>([&]() -> type { volatile type x = a, y = b; volatile type r = expr; return r; }); \
    }
DS_EMU_BINARY(__fadd_rn, float, FE_TONEAREST, x + y)
DS_EMU_BINARY(__fsub_rn, float, FE_TONEAREST, x - y)
DS_EMU_BINARY(__fmul_rn, float, FE_TONEAREST, x * y)
DS_EMU_BINARY(__fdiv_rn, float, FE_TONEAREST, x / y)
DS_EMU_BINARY(__fadd_ru, float, FE_UPWARD, x + y)
DS_EMU_BINARY(__fsub_ru, float, FE_UPWARD, x - y)
DS_EMU_BINARY(__fmul_ru, float, FE_UPWARD, x * y)
DS_EMU_BINARY(__fdiv_ru, float, FE_UPWARD, x / y)
DS_EMU_BINARY(__fadd_rd, float, FE_DOWNWARD, x + y)
DS_EMU_BINARY(__fsub_rd, float, FE_DOWNWARD, x - y)
DS_EMU_BINARY(__fmul_rd, float, FE_DOWNWARD, x * y)
DS_EMU_BINARY(__fdiv_rd, float, FE_DOWNWARD, x / y)
DS_EMU_BINARY(__dadd_rn, double, FE_TONEAREST, x + y)
DS_EMU_BINARY(__dsub_rn, double, FE_TONEAREST, x - y)
DS_EMU_BINARY(__dmul_rn, double, FE_TONEAREST, x * y)
DS_EMU_BINARY(__ddiv_rn, double, FE_TONEAREST, x / y)
#undef DS_EMU_BINARY
inline float __frcp_rn(float a) { return __fdiv_rn(1.0f, a); }
inline float __frcp_ru(float a) { return __fdiv_ru(1.0f, a); }
inline float __frcp_rd(float a) { return __fdiv_rd(1.0f, a); }
inline float __double2float_rn(double a) { return ::ds_emu::rounded<FE_TONEAREST>([&]() -> float { volatile double x = a; volatile float r = (float)x; return r; }); }
inline float __double2float_rd(double a) { return ::ds_emu::rounded<FE_DOWNWARD>([&]() -> float { volatile double x = a; volatile float r = (float)x; return r; }); }
inline float __double2float_ru(double a) { return ::ds_emu::rounded<FE_UPWARD>([&]() -> float { volatile double x = a; volatile float r = (float)x; return r; }); }
inline unsigned __float2uint_ru(float a) {   // saturating, NaN -> 0 (cvt.rpi.u32.f32)
    if (!(a > 0.0f)) return 0u;
    const float c = ceilf(a);
    return c >= 4294967296.0f ? 0xffffffffu : (unsigned)c;
}
inline unsigned __float2uint_rz(float a) {
    if (!(a > 0.0f)) return 0u;
    return a >= 4294967296.0f ? 0xffffffffu : (unsigned)a;
}
inline int __float2int_rz(float a) {
    if (a != a) return 0;
    if (a >= 2147483648.0f) return 2147483647;
    if (a <= -2147483648.0f) return (-2147483647 - 1);
    return (int)a;
}

// CUDA's overloaded min / max (integral promotions pick the int form for narrower types)
#define DS_EMU_MINMAX(type)                                          \
    inline type min(type a, type b) { return b < a ? b : a; }        \
    inline type max(type a, type b) { return a < b ? b : a; }
DS_EMU_MINMAX(int)
DS_EMU_MINMAX(unsigned)
DS_EMU_MINMAX(long)
DS_EMU_MINMAX(unsigned long)
DS_EMU_MINMAX(long long)
DS_EMU_MINMAX(unsigned long long)
#undef DS_EMU_MINMAX
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }

// ::cuda::std::plus<> (the reduction operator handed to cub::DeviceReduce::ReduceByKey)
namespace cuda {
namespace std {
template <typename T = void>
struct plus {
    template <typename A, typename B>
    auto operator()(const A &a, const B &b) const -> decltype(a + b) { return a + b; }
};
}  // namespace std
}  // namespace cuda
