// Shared device/host helpers for the cnvkit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace cnvk {

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- error plumbing -------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define CNVK_CHECK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            cnvk::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                                 cudaGetErrorString(_e));                             \
            return (int)_e;                                                           \
        }                                                                             \
    } while (0)

extern std::atomic<unsigned long long> g_launch_count;  // kernels launched by this library (host-side tally)
#define CNVK_LAUNCH_CHECK()                 \
    do {                                    \
        cnvk::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
        CNVK_CHECK(cudaGetLastError());     \
    } while (0)

// ---- device scratch ---------------------------------------------------------
// Every entry point needs dozens of short-lived device buffers (a fix + CBS exome sample: ~300).  They come from a
// per-thread STACK over one slab: a Scratch scope remembers the stack top when it opens and restores it when it
// closes, so an allocation is a pointer bump.  This is safe for the same reason cudaFreeAsync / cudaMallocAsync are:
// all work of a scope is enqueued on ONE stream, so a later scope that reuses the bytes is ordered after it.  The slab
// is handed from stream to stream through an event recorded when the outermost scope closes; a scope opened on a
// different stream while another is open (a side stream) does not use the stack.  Requests that do not fit fall back
// to cudaMallocAsync, and the slab is regrown to the high-water mark when the next outermost scope opens.
// (Measured: with one cudaMallocAsync per buffer the per-sample host time grew with the state of the pool -- after a
// threaded cohort run the same call was 0.6-2.3 ms slower -- profiles/r02y_exome_leg_probe.log.)
struct ScratchArena {
    char* base = nullptr;
    size_t cap = 0, top = 0, vtop = 0, want = 0;
    int depth = 0;
    cudaStream_t stream = nullptr;      // stream of the open outermost scope (valid while depth > 0)
    cudaEvent_t ev = nullptr;           // recorded when the last outermost scope closed
    bool ev_pending = false;
    int device = -1;
    ScratchArena();
    void park();
    void adopt(int dev);
    ~ScratchArena();
};
// Slabs outlive their threads: a worker thread that exits parks its slab (with the event that orders its last use)
// for the next thread, so short-lived worker pools neither free device memory nor start without a slab.  Nothing is
// freed at process exit.
struct ScratchSlab {
    char* base;
    size_t cap;
    cudaEvent_t ev;
    bool ev_pending;
    int device;
};
void scratch_slab_park(const ScratchSlab& s);
bool scratch_slab_adopt(int device, ScratchSlab* s);
inline ScratchArena::ScratchArena() {}
inline void ScratchArena::park() {
    if (base || ev) scratch_slab_park(ScratchSlab{base, cap, ev, ev_pending, device});
    base = nullptr;
    ev = nullptr;
    cap = top = vtop = want = 0;
    ev_pending = false;
}
inline void ScratchArena::adopt(int dev) {
    ScratchSlab s;
    device = dev;
    if (scratch_slab_adopt(dev, &s)) { base = s.base; cap = s.cap; ev = s.ev; ev_pending = s.ev_pending; want = s.cap; }
}
inline ScratchArena::~ScratchArena() { park(); }
inline thread_local ScratchArena g_scratch_arena;

bool scratch_stack_enabled();
struct Scratch {
    cudaStream_t stream;
    void* ptrs[256];                    // fallback allocations of this scope
    int n;
    size_t mark, vmark;
    bool stacked;
    explicit Scratch(cudaStream_t s) : stream(s), n(0), mark(0), vmark(0), stacked(false) {
        ScratchArena& A = g_scratch_arena;
        if (!scratch_stack_enabled()) return;       // CNVK_SCRATCH_STACK=0: one cudaMallocAsync per buffer
        if (A.depth == 0) {
            int dev = -1;
            if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
            if (dev != A.device) {          // first use by this thread, or the thread moved to another GPU
                A.park();
                A.adopt(dev);
            }
            if (!A.ev && cudaEventCreateWithFlags(&A.ev, cudaEventDisableTiming) != cudaSuccess) { A.ev = nullptr; return; }
            if (A.ev_pending) {
                // whatever used the slab last has to finish before this stream touches it (no-op on the same stream)
                if (cudaStreamWaitEvent(s, A.ev, 0) != cudaSuccess) { cudaGetLastError(); return; }
                A.ev_pending = false;
            }
            if (A.want > A.cap) {
                if (A.base) cudaFreeAsync(A.base, s);
                A.base = nullptr;
                A.cap = 0;
                // generous: growing means new physical memory from the driver (~10 ms for a few hundred MB)
                const size_t ncap = std::max(A.want + A.want / 2, 2 * A.cap) + (1u << 20);
                void* p = nullptr;
                if (cudaMallocAsync(&p, ncap, s) == cudaSuccess) {
                    A.base = (char*)p;
                    A.cap = ncap;
                } else {
                    cudaGetLastError();
                }
            }
            A.top = A.vtop = 0;
            A.stream = s;
            stacked = true;
        } else {
            stacked = (s == A.stream);
        }
        if (stacked) {
            mark = A.top;
            vmark = A.vtop;
            A.depth++;
        }
    }
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], stream);
        if (!stacked) return;
        ScratchArena& A = g_scratch_arena;
        A.top = mark;
        A.vtop = vmark;
        if (--A.depth == 0) {
            if (A.ev && cudaEventRecord(A.ev, stream) == cudaSuccess) A.ev_pending = true;
            A.stream = nullptr;
        }
    }
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    template <typename T>
    cudaError_t alloc(T** out, size_t count) {
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        bytes = (bytes + 255) & ~(size_t)255;
        if (stacked) {
            ScratchArena& A = g_scratch_arena;
            A.vtop += bytes;
            if (A.vtop > A.want) A.want = A.vtop;
            if (A.top + bytes <= A.cap) {
                *out = (T*)(A.base + A.top);
                A.top += bytes;
                return cudaSuccess;
            }
        }
        void* p = nullptr;
        if (n >= 256) return cudaErrorMemoryAllocation;
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e != cudaSuccess) return e;
        ptrs[n++] = p;
        *out = (T*)p;
        return cudaSuccess;
    }
};

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers -------------------------------------------------------
// Monotone map double -> u64 such that unsigned order == numeric order,
// -0.0 and +0.0 collapse to one key and every NaN sorts last (numpy argsort
// semantics for kind="mergesort": NaN at the end, ties by original index).
__host__ __device__ __forceinline__ u64 f64_to_ordered(double x) {
    if (x != x) return ~0ull;
    if (x == 0.0) x = 0.0;  // -0.0 -> +0.0
#ifdef __CUDA_ARCH__
    u64 b = (u64)__double_as_longlong(x);
#else
    u64 b;
    memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ u32 warp_sum_u32(u32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace cnvk
