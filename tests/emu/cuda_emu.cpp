// cuda_emu.cpp - TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the fiber engine that runs CUDA threads on the host, and
// the host stand-ins of the runtime calls the library makes.
#include "cuda_emu.h"

#include <sys/mman.h>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#include <sanitizer/common_interface_defs.h>
#define DS_EMU_ASAN 1
#else
#define DS_EMU_ASAN 0
#endif

// ThreadSanitizer build (build.py --tsan): the kernel sources are instrumented, this file is not.  Every CUDA thread is a
// TSan fiber (the 1,024 fibers are created once and reused CTA after CTA); fiber switches carry NO synchronisation, the
// happens-before edges are exactly CUDA's: launch -> every thread of a CTA -> end of the CTA (CTAs are ordered one after
// the other: the scope is that of compute-sanitizer's racecheck, hazards INSIDE a CTA), __syncthreads between the threads
// of the CTA, __syncwarp between the lanes it names.  Votes / shuffles / reductions order memory only when
// DS_EMU_COLLECTIVES_ORDER=1 (the programming guide promises it for __syncwarp alone).
#if defined(__SANITIZE_THREAD__) || defined(DS_EMU_TSAN)
#define DS_EMU_RACECHECK 1
extern "C" {
void *__tsan_get_current_fiber(void);
void *__tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void *fiber);
void __tsan_switch_to_fiber(void *fiber, unsigned flags);
void __tsan_acquire(void *addr);
void __tsan_release(void *addr);
}
#else
#define DS_EMU_RACECHECK 0
#endif

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace ds_emu {

ThreadCtx *g_cur = nullptr;
uint3 g_block_idx = {0, 0, 0};
dim3 g_block_dim, g_grid_dim;
unsigned char *g_dyn_smem = nullptr;

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;
constexpr int MAX_THREADS = 1024;
constexpr size_t SMEM_LIMIT = 227 * 1024;      // opt-in maximum of dynamic + static shared memory per CTA on sm_100
constexpr size_t SMEM_DEFAULT = 48 * 1024;     // without cudaFuncAttributeMaxDynamicSharedMemorySize

enum State { READY, WAIT_WARP, WAIT_CTA, DONE };

struct Fiber : ThreadCtx {
    void *sp = nullptr;
    char *stack = nullptr;
    int state = DONE;
    int lane = 0, warp = 0;
    // pending warp collective
    int op = 0, arg = 0, width = 32;
    unsigned mask = 0;
    uint64_t value = 0, result = 0;
    void *asan_fake_stack = nullptr;
    void *tsan_fiber = nullptr;
};

struct Engine {
    std::vector<Fiber> fibers;
    char *stacks = nullptr;
    int n_threads = 0, n_live = 0, n_at_barrier = 0;
    const std::function<void()> *body = nullptr;
    void *scheduler_sp = nullptr;
    void *scheduler_fake_stack = nullptr;
    const char *kernel = "";
    std::map<const void *, size_t> granted_smem;
    cudaError_t last_error = cudaSuccess;
    void *scheduler_tsan_fiber = nullptr;
    bool collectives_order_memory = false;
    // addresses ThreadSanitizer keeps vector clocks for (contents unused)
    char sync_cta_start = 0, sync_cta_done = 0, sync_barrier = 0;
    std::map<uint64_t, char> sync_warp;   // (warp, mask) -> clock of that group of lanes
    // statistics, printed when DS_EMU_STATS is set
    uint64_t launches = 0, ctas = 0, threads = 0, collectives = 0, switches = 0, reads_of_absent_lanes = 0;
};
Engine g;
std::recursive_mutex g_lock;

[[noreturn]] void die(const char *fmt, ...) {
    va_list args;
    va_start(args, fmt);
    fprintf(stderr, "ds_emu: FATAL (kernel %s, block (%u,%u,%u)): ", g.kernel, g_block_idx.x, g_block_idx.y, g_block_idx.z);
    vfprintf(stderr, fmt, args);
    fprintf(stderr, "\n");
    va_end(args);
    fflush(stderr);
    abort();
}

extern "C" void ds_emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl ds_emu_switch
    .type ds_emu_switch,@function
ds_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size ds_emu_switch,.-ds_emu_switch
)");

// stack of the host thread that runs the scheduler (learnt when a fiber is entered for the first time)
const void *g_main_stack_bottom = nullptr;
size_t g_main_stack_size = 0;

inline void to_scheduler(Fiber *self, bool dying) {
    ++g.switches;
#if DS_EMU_ASAN
    __sanitizer_start_switch_fiber(dying ? nullptr : &self->asan_fake_stack, g_main_stack_bottom, g_main_stack_size);
#endif
    (void)dying;
#if DS_EMU_RACECHECK
    __tsan_switch_to_fiber(g.scheduler_tsan_fiber, 1 /* no synchronisation */);
#endif
    ds_emu_switch(&self->sp, g.scheduler_sp);
#if DS_EMU_ASAN
    __sanitizer_finish_switch_fiber(self->asan_fake_stack, nullptr, nullptr);
#endif
}

void trampoline() {
#if DS_EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &g_main_stack_bottom, &g_main_stack_size);
#endif
    Fiber *self = static_cast<Fiber *>(g_cur);
#if DS_EMU_RACECHECK
    __tsan_acquire(&g.sync_cta_start);
#endif
    (*g.body)();
#if DS_EMU_RACECHECK
    __tsan_release(&g.sync_cta_done);
#endif
    self->state = DONE;
#if DS_EMU_ASAN
    __sanitizer_start_switch_fiber(nullptr, g_main_stack_bottom, g_main_stack_size);
#endif
    ++g.switches;
#if DS_EMU_RACECHECK
    __tsan_switch_to_fiber(g.scheduler_tsan_fiber, 1);
#endif
    ds_emu_switch(&self->sp, g.scheduler_sp);
    die("a finished CUDA thread was resumed");
}

void resume(Fiber *f) {
    g_cur = f;
#if DS_EMU_ASAN
    __sanitizer_start_switch_fiber(&g.scheduler_fake_stack, f->stack, STACK_BYTES);
#endif
#if DS_EMU_RACECHECK
    if (f->tsan_fiber == nullptr) f->tsan_fiber = __tsan_create_fiber(0);
    __tsan_switch_to_fiber(f->tsan_fiber, 1 /* no synchronisation */);
#endif
    ds_emu_switch(&g.scheduler_sp, f->sp);
#if DS_EMU_ASAN
    __sanitizer_finish_switch_fiber(g.scheduler_fake_stack, nullptr, nullptr);
#endif
    g_cur = nullptr;
}

void prepare(Fiber *f, int index, int linear) {
    if (g.stacks == nullptr) {
        g.stacks = static_cast<char *>(mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (g.stacks == MAP_FAILED) die("cannot map the fiber stacks");
    }
    f->stack = g.stacks + (size_t)index * STACK_BYTES;
#if DS_EMU_ASAN
    // the previous owner of this stack left through a switch, not through returns: its last frames are still poisoned
    __asan_unpoison_memory_region(f->stack + STACK_BYTES - 4096, 4096);
#endif
    uintptr_t top = (reinterpret_cast<uintptr_t>(f->stack) + STACK_BYTES) & ~(uintptr_t)15;
    void **sp = reinterpret_cast<void **>(top);
    *--sp = nullptr;                                   // return address of the trampoline (never used)
    *--sp = reinterpret_cast<void *>(&trampoline);     // `ret` of the first switch lands here
    for (int i = 0; i < 6; ++i) *--sp = nullptr;       // rbp rbx r12 r13 r14 r15
    f->sp = sp;
    f->state = READY;
    f->lane = linear & 31;
    f->warp = linear >> 5;
    f->asan_fake_stack = nullptr;
}

// ---------------------------------------------------------------------------------------------- warp collectives
inline bool lane_exists(int warp, int lane) { return warp * 32 + lane < g.n_threads; }

// All live lanes named by `mask` wait in the same collective -> compute every lane's result and release them.
bool try_complete(int warp, unsigned mask) {
    Fiber *lanes = &g.fibers[(size_t)warp * 32];
    int op = -1, first = -1;
    for (int l = 0; l < 32; ++l) {
        if (!(mask >> l & 1) || !lane_exists(warp, l) || lanes[l].state == DONE) continue;
        if (lanes[l].state != WAIT_WARP) return false;
        if (lanes[l].mask != mask) {
            // a lane named by this mask waits in a collective over ANOTHER set of lanes: on the GPU that is undefined
            die("warp %d: lane %d waits with mask %08x while lane(s) of mask %08x name it", warp, l, lanes[l].mask, mask);
        }
        if (op < 0) {
            op = lanes[l].op;
            first = l;
        } else if (lanes[l].op != op) {
            die("warp %d: lanes %d and %d meet in different collectives (%d vs %d) under mask %08x", warp, first, l, op, lanes[l].op, mask);
        }
    }
    if (op < 0) return true;
    ++g.collectives;
    auto present = [&](int l) { return l >= 0 && l < 32 && (mask >> l & 1) && lane_exists(warp, l) && lanes[l].state == WAIT_WARP; };
    uint64_t combined = 0;
    bool have = false;
    for (int l = 0; l < 32; ++l) {
        if (!present(l)) continue;
        const uint64_t v = lanes[l].value;
        if (!have) {
            combined = (op == OP_BALLOT) ? 0 : v;
            if (op == OP_BALLOT) combined = v ? (1ull << l) : 0;
            have = true;
            continue;
        }
        switch (op) {
            case OP_BALLOT: combined |= v ? (1ull << l) : 0; break;
            case OP_ALL: combined = combined && v; break;
            case OP_ANY: combined = combined || v; break;
            case OP_RED_MAX_U: combined = std::max((uint32_t)combined, (uint32_t)v); break;
            case OP_RED_MIN_U: combined = std::min((uint32_t)combined, (uint32_t)v); break;
            case OP_RED_MAX_S: combined = (uint64_t)std::max((int64_t)combined, (int64_t)v); break;
            case OP_RED_MIN_S: combined = (uint64_t)std::min((int64_t)combined, (int64_t)v); break;
            case OP_RED_ADD: combined = (uint32_t)(combined + v); break;
            case OP_RED_OR: combined |= v; break;
            case OP_RED_AND: combined &= v; break;
            case OP_RED_XOR: combined ^= v; break;
            default: break;
        }
    }
    for (int l = 0; l < 32; ++l) {
        if (!present(l)) continue;
        Fiber &f = lanes[l];
        const int w = f.width;
        const int seg = l & ~(w - 1);
        int src = l;
        switch (op) {
            case OP_SHFL: src = seg | (f.arg & (w - 1)); break;
            case OP_SHFL_UP: src = (l - f.arg >= seg) ? l - f.arg : l; break;
            case OP_SHFL_DOWN: src = (l + f.arg < seg + w) ? l + f.arg : l; break;
            case OP_SHFL_XOR: src = ((l ^ f.arg) < seg + w) ? (l ^ f.arg) : l; break;
            default: break;
        }
        if (op == OP_SHFL || op == OP_SHFL_UP || op == OP_SHFL_DOWN || op == OP_SHFL_XOR) {
            if (w <= 0 || w > 32 || (w & (w - 1))) die("shuffle width %d", w);
            if (!present(src)) {   // undefined value on the GPU
                ++g.reads_of_absent_lanes;
                src = l;
            }
            f.result = lanes[src].value;
        } else {
            f.result = combined;
        }
    }
    for (int l = 0; l < 32; ++l)
        if (present(l)) lanes[l].state = READY;
    return true;
}

void release_barrier_if_complete() {
    if (g.n_at_barrier > 0 && g.n_at_barrier == g.n_live) {
        for (int t = 0; t < g.n_threads; ++t)
            if (g.fibers[t].state == WAIT_CTA) g.fibers[t].state = READY;
        g.n_at_barrier = 0;
    }
}

void run_cta() {
#if DS_EMU_RACECHECK
    __tsan_release(&g.sync_cta_start);
#endif
    g.n_live = g.n_threads;
    g.n_at_barrier = 0;
    int pending = g.n_threads;
    while (pending > 0) {
        bool progress = false;
        for (int t = 0; t < g.n_threads; ++t) {
            Fiber *f = &g.fibers[t];
            while (f->state == READY) {
                progress = true;
                resume(f);
                if (f->state == DONE) {
                    --pending;
                    --g.n_live;
                    release_barrier_if_complete();
                    // lanes of this warp may have been waiting for it
                    Fiber *lanes = &g.fibers[(size_t)f->warp * 32];
                    for (int l = 0; l < 32; ++l)
                        if (lane_exists(f->warp, l) && lanes[l].state == WAIT_WARP && (lanes[l].mask >> f->lane & 1)) try_complete(f->warp, lanes[l].mask);
                }
            }
        }
        if (!progress) {
            fprintf(stderr, "ds_emu: deadlock in kernel %s, block (%u,%u,%u): %d threads left, %d at __syncthreads\n", g.kernel, g_block_idx.x,
                    g_block_idx.y, g_block_idx.z, pending, g.n_at_barrier);
            for (int t = 0; t < g.n_threads; ++t) {
                const Fiber &f = g.fibers[t];
                if (f.state == WAIT_WARP) fprintf(stderr, "  thread %d: warp collective %d, mask %08x\n", t, f.op, f.mask);
                if (f.state == WAIT_CTA) fprintf(stderr, "  thread %d: __syncthreads\n", t);
            }
            abort();
        }
    }
#if DS_EMU_RACECHECK
    __tsan_acquire(&g.sync_cta_done);
#endif
}

struct Allocation {
    size_t bytes;
};
std::map<uintptr_t, Allocation> g_device_memory;

const Allocation *find_allocation(const void *p, uintptr_t *base_out = nullptr) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_device_memory.upper_bound(a);
    if (it == g_device_memory.begin()) return nullptr;
    --it;
    if (a < it->first + std::max<size_t>(it->second.bytes, 1)) {
        if (base_out) *base_out = it->first;
        return &it->second;
    }
    return nullptr;
}

struct Stats {
    ~Stats() {
        const char *where = getenv("DS_EMU_STATS");   // "1": stderr; a path: appended there (one line per process)
        FILE *out = where == nullptr ? nullptr : (where[0] == '/' ? fopen(where, "a") : stderr);
        if (out != nullptr) {
            fprintf(out, "ds_emu: %llu launches, %llu CTAs, %llu threads, %llu warp collectives, %llu fiber switches, %llu shuffle reads of absent lanes, %zu live device blocks\n",
                    (unsigned long long)g.launches, (unsigned long long)g.ctas, (unsigned long long)g.threads, (unsigned long long)g.collectives,
                    (unsigned long long)g.switches, (unsigned long long)g.reads_of_absent_lanes, g_device_memory.size());
            if (out != stderr) fclose(out);
        }
    }
} g_stats;

}  // namespace

uint64_t warp_collective(int op, unsigned mask, uint64_t value, int arg, int width) {
    Fiber *self = static_cast<Fiber *>(g_cur);
    if (self == nullptr) die("warp collective outside a kernel");
    if (!(mask >> self->lane & 1)) die("lane %d calls a collective whose mask %08x does not name it", self->lane, mask);
    self->op = op;
    self->mask = mask;
    self->value = value;
    self->arg = arg;
    self->width = width;
    self->state = WAIT_WARP;
#if DS_EMU_RACECHECK
    char *clock = nullptr;
    if (op == OP_SYNCWARP || g.collectives_order_memory) {
        clock = &g.sync_warp[((uint64_t)self->warp << 32) | mask];
        __tsan_release(clock);
    }
#endif
    try_complete(self->warp, mask);
    if (self->state != READY) to_scheduler(self, false);
    if (self->state != READY) die("a waiting CUDA thread was resumed");
#if DS_EMU_RACECHECK
    if (clock != nullptr) __tsan_acquire(clock);
#endif
    return self->result;
}

void cta_barrier() {
    Fiber *self = static_cast<Fiber *>(g_cur);
    if (self == nullptr) die("__syncthreads outside a kernel");
    self->state = WAIT_CTA;
    ++g.n_at_barrier;
#if DS_EMU_RACECHECK
    __tsan_release(&g.sync_barrier);
#endif
    release_barrier_if_complete();
    if (self->state != READY) to_scheduler(self, false);
#if DS_EMU_RACECHECK
    __tsan_acquire(&g.sync_barrier);
#endif
}

void launch(const void *fn, const char *name, dim3 grid, dim3 block, size_t smem, cudaStream_t, const std::function<void()> &body) {
    std::lock_guard<std::recursive_mutex> guard(g_lock);
    if (g_cur != nullptr) die("kernel launch from inside a kernel");
    const uint64_t n_threads = (uint64_t)block.x * block.y * block.z;
    const uint64_t n_blocks = (uint64_t)grid.x * grid.y * grid.z;
    size_t allowed = SMEM_DEFAULT;
    auto it = g.granted_smem.find(fn);
    if (it != g.granted_smem.end()) allowed = std::max(allowed, it->second);
    if (n_blocks == 0 || n_threads == 0 || n_threads > MAX_THREADS || grid.y > 65535 || grid.z > 65535 || grid.x > 2147483647u || smem > allowed) {
        fprintf(stderr, "ds_emu: invalid launch configuration of %s: grid (%u,%u,%u) block (%u,%u,%u) dynamic smem %zu (allowed %zu)\n", name, grid.x,
                grid.y, grid.z, block.x, block.y, block.z, smem, allowed);
        g.last_error = cudaErrorInvalidConfiguration;
        return;
    }
    ++g.launches;
#if DS_EMU_RACECHECK
    g.scheduler_tsan_fiber = __tsan_get_current_fiber();
    g.collectives_order_memory = getenv("DS_EMU_COLLECTIVES_ORDER") != nullptr;
    g.sync_warp.clear();
#endif
    g.kernel = name;
    g.body = &body;
    g.n_threads = (int)n_threads;
    if (g.fibers.size() < (size_t)MAX_THREADS) g.fibers.resize(MAX_THREADS);
    g_block_dim = block;
    g_grid_dim = grid;
    unsigned char *dyn = static_cast<unsigned char *>(malloc(std::max<size_t>(smem, 1)));   // exact size: ASan flags any access past it
    g_dyn_smem = dyn;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_block_idx = {bx, by, bz};
                memset(dyn, 0xA5, smem);   // shared memory is not zeroed between CTAs
                int linear = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++linear) {
                            Fiber *f = &g.fibers[linear];
                            f->thread_idx = {tx, ty, tz};
                            prepare(f, linear, linear);
                        }
                ++g.ctas;
                g.threads += n_threads;
                run_cta();
            }
    g_dyn_smem = nullptr;
    free(dyn);
    g.body = nullptr;
    g.kernel = "";
}

}  // namespace ds_emu

// ------------------------------------------------------------------------------------------------ runtime API
using ds_emu::g;

cudaError_t cudaGetLastError() {
    cudaError_t e = g.last_error;
    g.last_error = cudaSuccess;
    return e;
}
const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument";
        case cudaErrorInvalidDevice: return "invalid device ordinal";
        default: return "emulated CUDA error";
    }
}
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }

cudaError_t cudaMallocAsync(void **out, size_t bytes, cudaStream_t) {
    std::lock_guard<std::recursive_mutex> guard(ds_emu::g_lock);
    void *p = nullptr;
    if (posix_memalign(&p, 256, std::max<size_t>(bytes, 1)) != 0) return cudaErrorMemoryAllocation;   // cudaMalloc alignment
    memset(p, 0xCB, bytes);                       // device memory is not zeroed either
    ds_emu::g_device_memory[reinterpret_cast<uintptr_t>(p)] = ds_emu::Allocation{bytes};
    *out = p;
    return cudaSuccess;
}
cudaError_t cudaFreeAsync(void *p, cudaStream_t) {
    if (p == nullptr) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> guard(ds_emu::g_lock);
    auto it = ds_emu::g_device_memory.find(reinterpret_cast<uintptr_t>(p));
    if (it == ds_emu::g_device_memory.end()) {
        fprintf(stderr, "ds_emu: cudaFree of %p, which is not a live device allocation\n", p);
        abort();
    }
    ds_emu::g_device_memory.erase(it);
    free(p);                                      // later kernel accesses are use-after-free reports
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { return cudaFreeAsync(p, nullptr); }

cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t) {
    std::lock_guard<std::recursive_mutex> guard(ds_emu::g_lock);
    if (bytes == 0) return cudaSuccess;
    const bool dst_dev = ds_emu::find_allocation(dst) != nullptr, src_dev = ds_emu::find_allocation(src) != nullptr;
    const bool ok = kind == cudaMemcpyDefault || (kind == cudaMemcpyHostToDevice && dst_dev && !src_dev) ||
                    (kind == cudaMemcpyDeviceToHost && !dst_dev && src_dev) || (kind == cudaMemcpyDeviceToDevice && dst_dev && src_dev) ||
                    (kind == cudaMemcpyHostToHost && !dst_dev && !src_dev);
    if (!ok) {
        fprintf(stderr, "ds_emu: cudaMemcpyAsync kind %d does not match its pointers (dst %s, src %s)\n", (int)kind, dst_dev ? "device" : "host",
                src_dev ? "device" : "host");
        g.last_error = cudaErrorInvalidValue;
        return cudaErrorInvalidValue;
    }
    memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t) {
    if (bytes > 0 && ds_emu::find_allocation(dst) == nullptr) {
        fprintf(stderr, "ds_emu: cudaMemsetAsync on a host pointer\n");
        return cudaErrorInvalidValue;
    }
    memset(dst, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *attr, const void *p) {
    std::lock_guard<std::recursive_mutex> guard(ds_emu::g_lock);
    const bool device = ds_emu::find_allocation(p) != nullptr;
    attr->type = device ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
    attr->device = device ? 0 : -1;
    attr->devicePointer = device ? const_cast<void *>(p) : nullptr;
    attr->hostPointer = device ? nullptr : const_cast<void *>(p);
    return cudaSuccess;
}
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *pool, int) { *pool = nullptr; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void *fn, cudaFuncAttribute attr, int value) {
    std::lock_guard<std::recursive_mutex> guard(ds_emu::g_lock);
    if (attr != cudaFuncAttributeMaxDynamicSharedMemorySize || value < 0 || (size_t)value > ds_emu::SMEM_LIMIT) return cudaErrorInvalidValue;
    g.granted_smem[fn] = (size_t)value;
    return cudaSuccess;
}

struct ds_emu_event {
    std::chrono::steady_clock::time_point at;
};
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new ds_emu_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->at = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->at - a->at).count();
    return cudaSuccess;
}

// device buffers for tests that hand the library DEVICE pointers (used in place instead of being staged)
// (sizes are rounded up to 16 bytes: the caller's allocator - cudaMalloc, torch - hands out 256-byte granules at least, and
// the string kernels rely on that when they fetch the aligned 32-bit word that holds a table's last byte; the library's OWN
// workspaces keep their exact size, Workspace::alloc rounds them itself)
extern "C" void *ds_emu_malloc(size_t bytes) {
    void *p = nullptr;
    return cudaMallocAsync(&p, (bytes + 15) & ~(size_t)15, nullptr) == cudaSuccess ? p : nullptr;
}
extern "C" void ds_emu_free(void *p) { cudaFreeAsync(p, nullptr); }

// tests read these: how much ran, and how many shuffles read a lane that was not part of the collective
extern "C" uint64_t ds_emu_reads_of_absent_lanes() { return g.reads_of_absent_lanes; }
extern "C" uint64_t ds_emu_kernel_threads() { return g.threads; }
extern "C" uint64_t ds_emu_live_device_blocks() { return ds_emu::g_device_memory.size(); }
