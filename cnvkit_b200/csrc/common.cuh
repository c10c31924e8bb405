// Shared device/host helpers for the cnvkit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

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

// RAII stream-ordered scratch allocation (cudaMallocAsync pool; cheap after warm-up).
struct Scratch {
    cudaStream_t stream;
    void* ptrs[64];
    int n;
    explicit Scratch(cudaStream_t s) : stream(s), n(0) {}
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], stream);
    }
    template <typename T>
    cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        if (n >= 64) return cudaErrorMemoryAllocation;
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
