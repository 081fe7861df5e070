// selftest.cu - TEST INFRASTRUCTURE ONLY: kernels with PLANTED faults, to prove that the host emulation + sanitizers
// (tests/emu/cuda_emu.h) report the classes of error they are relied on for.  Built and run by
// tests/test_emulated_kernels.py::test_emulator_reports_planted_faults; `selftest clean` must pass, every other case
// must be reported.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

__global__ void k_clean(const int *in, int n, int *out) {
    extern __shared__ __align__(16) unsigned char raw[];
    int *tile = reinterpret_cast<int *>(raw);
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31;
    const int i = blockIdx.x * blockDim.x + tid;
    int v = i < n ? in[i] : 0;
    tile[tid] = v;
    if (tid == 0) s_total = 0;
    __syncthreads();
    int sum = tile[tid ^ 1] + v;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const unsigned odd = __ballot_sync(0xffffffffu, v & 1);
    if (lane == 0) atomicAdd(&s_total, sum + __popc(odd));
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = s_total;
}

__global__ void k_dynamic_smem_overflow(int *out) {
    extern __shared__ __align__(16) unsigned char raw[];
    int *tile = reinterpret_cast<int *>(raw);
    tile[threadIdx.x + 1] = (int)threadIdx.x;     // the last thread writes one element past the launch's shared memory
    __syncthreads();
    out[threadIdx.x] = tile[threadIdx.x];
}

__global__ void k_static_smem_overflow(int *out, int shift) {
    __shared__ int s_tile[64];
    s_tile[threadIdx.x + shift] = (int)threadIdx.x;
    __syncthreads();
    out[threadIdx.x] = s_tile[threadIdx.x];
}

__global__ void k_global_overflow(const int *in, int n, int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i + 1];                  // in[n] does not exist
}

__global__ void k_mismatched_collectives(int *out) {
    const int lane = threadIdx.x & 31;
    int v;
    if (lane < 16) v = (int)__ballot_sync(0xffffffffu, lane & 1);
    else v = __shfl_sync(0xffffffffu, lane, 0);
    out[threadIdx.x] = v;
}

__global__ void k_barrier_deadlock(int *out) {
    if (threadIdx.x < 16) __syncthreads();          // half a warp waits for the CTA ...
    else __syncwarp(0xffffffffu);                   // ... the other half for the whole warp
    out[threadIdx.x] = 1;
}

__global__ void k_misaligned_vector_load(const unsigned char *bytes, uint4 *out) {
    out[threadIdx.x] = *reinterpret_cast<const uint4 *>(bytes + 4 + 16 * threadIdx.x);
}

// racecheck cases (ThreadSanitizer build): a neighbour's shared-memory value is read without the barrier that orders it
__global__ void k_race_missing_syncthreads(int *out) {
    __shared__ int s_tile[128];
    s_tile[threadIdx.x] = (int)threadIdx.x;
    out[threadIdx.x] = s_tile[(threadIdx.x + 32) & 127];      // another warp's element: needs __syncthreads()
}

__global__ void k_race_missing_syncwarp(int *out) {
    __shared__ int s_tile[32];
    s_tile[threadIdx.x] = (int)threadIdx.x;
    out[threadIdx.x] = s_tile[threadIdx.x ^ 1];               // the neighbouring lane's element: needs __syncwarp()
}

__global__ void k_ordered_by_syncwarp(int *out) {
    __shared__ int s_tile[32];
    s_tile[threadIdx.x] = (int)threadIdx.x;
    __syncwarp();
    out[threadIdx.x] = s_tile[threadIdx.x ^ 1];
}

__global__ void k_touch(const int *in, int *out) { out[threadIdx.x] = in[threadIdx.x]; }

#define CHECK(expr)                                                                         \
    do {                                                                                    \
        cudaError_t e = (expr);                                                             \
        if (e != cudaSuccess) {                                                             \
            fprintf(stderr, "selftest: %s -> %s\n", #expr, cudaGetErrorString(e));          \
            return 3;                                                                       \
        }                                                                                   \
    } while (0)

int main(int argc, char **argv) {
    const char *which = argc > 1 ? argv[1] : "clean";
    const int n = 1000, threads = 128, blocks = (n + threads - 1) / threads;
    int *d_in = nullptr, *d_out = nullptr;
    CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_in), n * sizeof(int), nullptr));
    CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_out), 1024 * sizeof(int), nullptr));
    int h_in[1000], h_out[1024];
    for (int i = 0; i < n; ++i) h_in[i] = i * 7 + 3;
    CHECK(cudaMemcpyAsync(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice, nullptr));
    if (!strcmp(which, "clean")) {
        k_clean<<<blocks, threads, threads * sizeof(int), nullptr>>>(d_in, n, d_out);
        CHECK(cudaGetLastError());
        CHECK(cudaMemcpyAsync(h_out, d_out, blocks * sizeof(int), cudaMemcpyDeviceToHost, nullptr));
        long long want = 0, got = 0;
        for (int b = 0; b < blocks; ++b) {
            got += h_out[b];
            for (int t = 0; t < threads; ++t) {
                const int i = b * threads + t, j = b * threads + (t ^ 1);
                const int v = i < n ? h_in[i] : 0, w = j < n ? h_in[j] : 0;
                want += v + w;                       // every lane's pair sum enters its warp's reduction once ...
                want += (v & 1) ? 1 : 0;             // ... and every odd value is counted once per warp
            }
        }
        // the reduction adds each warp's total once (lane 0), so the expected figure is sum(v + w) + #odd values
        if (got != want) {
            fprintf(stderr, "selftest: clean kernel computed %lld, expected %lld\n", got, want);
            return 4;
        }
    } else if (!strcmp(which, "dynamic-smem-overflow")) {
        k_dynamic_smem_overflow<<<1, threads, threads * sizeof(int), nullptr>>>(d_out);
    } else if (!strcmp(which, "static-smem-overflow")) {
        k_static_smem_overflow<<<1, 64, 0, nullptr>>>(d_out, 1);
    } else if (!strcmp(which, "global-overflow")) {
        k_global_overflow<<<blocks, threads, 0, nullptr>>>(d_in, n, d_out);
    } else if (!strcmp(which, "mismatched-collectives")) {
        k_mismatched_collectives<<<1, 32, 0, nullptr>>>(d_out);
    } else if (!strcmp(which, "barrier-deadlock")) {
        k_barrier_deadlock<<<1, 32, 0, nullptr>>>(d_out);
    } else if (!strcmp(which, "misaligned-vector-load")) {
        k_misaligned_vector_load<<<1, 8, 0, nullptr>>>(reinterpret_cast<const unsigned char *>(d_in), reinterpret_cast<uint4 *>(d_out));
    } else if (!strcmp(which, "race-missing-syncthreads")) {
        k_race_missing_syncthreads<<<1, 128, 0, nullptr>>>(d_out);
    } else if (!strcmp(which, "race-missing-syncwarp")) {
        k_race_missing_syncwarp<<<1, 32, 0, nullptr>>>(d_out);
    } else if (!strcmp(which, "ordered-by-syncwarp")) {
        k_ordered_by_syncwarp<<<2, 32, 0, nullptr>>>(d_out);
    } else if (!strcmp(which, "use-after-free")) {
        CHECK(cudaFreeAsync(d_in, nullptr));
        k_touch<<<1, 32, 0, nullptr>>>(d_in, d_out);
        d_in = nullptr;
    } else if (!strcmp(which, "smem-without-opt-in")) {
        k_clean<<<1, threads, 100 * 1024, nullptr>>>(d_in, n, d_out);    // > 48 KB needs cudaFuncSetAttribute first
        if (cudaGetLastError() == cudaSuccess) return 0;                   // (would be a miss)
        fprintf(stderr, "selftest: launch refused as expected\n");
        return 5;
    } else if (!strcmp(which, "empty-grid")) {
        k_touch<<<0, 32, 0, nullptr>>>(d_in, d_out);
        if (cudaGetLastError() == cudaSuccess) return 0;
        fprintf(stderr, "selftest: launch refused as expected\n");
        return 5;
    } else {
        fprintf(stderr, "selftest: unknown case %s\n", which);
        return 2;
    }
    if (d_in) CHECK(cudaFreeAsync(d_in, nullptr));
    CHECK(cudaFreeAsync(d_out, nullptr));
    printf("selftest %s: ok\n", which);
    return 0;
}
