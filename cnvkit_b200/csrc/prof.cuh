// Optional per-kernel-class device timing (CUDA events on the launching stream).  Off by default;
// bench.py turns it on to attribute the timed region to kernels for the roofline numbers.
#pragma once
#include "common.cuh"

namespace cnvk {

enum ProfClass {
    PC_SORT = 0,
    PC_WAVELET_BUILD,
    PC_WINDOW_QUERY,
    PC_CBS_PREP,
    PC_CBS_SCAN,
    PC_CBS_SEED,
    PC_CBS_BAND,
    PC_CBS_BOUND,
    PC_CBS_MEANS,
    PC_CBS_DECIDE,
    PC_CBS_PERM_HYBRID,
    PC_CBS_PERM_EXACT,
    PC_CBS_FLANK,
    PC_FIX,
    PC_HAAR,
    PC_BIWEIGHT,
    PC_OTHER,
    PC_COUNT
};

extern int g_prof_enabled;
extern int g_nvtx_enabled;      // CNVK_NVTX=1: an NVTX range per kernel class (for nsys / ncu --nvtx timelines)
void prof_begin(int cls, cudaStream_t stream, cudaEvent_t* start);
void prof_end(int cls, cudaStream_t stream, cudaEvent_t start);
void prof_nvtx_push(int cls);
void prof_nvtx_pop();

struct ProfScope {
    int cls;
    cudaStream_t stream;
    cudaEvent_t start;
    bool nvtx;
    ProfScope(int c, cudaStream_t s) : cls(c), stream(s), start(nullptr), nvtx(false) {
        if (g_nvtx_enabled) { prof_nvtx_push(cls); nvtx = true; }
        if (g_prof_enabled) prof_begin(cls, stream, &start);
    }
    ~ProfScope() {
        if (start) prof_end(cls, stream, start);
        if (nvtx) prof_nvtx_pop();
    }
};

}  // namespace cnvk
