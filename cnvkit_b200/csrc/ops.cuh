// Internal C++ entry points of the kernels (one per .cu file), called by capi.cu.
#pragma once
#include "prims.cuh"

namespace cnvk {

// rolling.cu
int rolling_window_stat(const double* d_x, long long n, long long wing, int mode, double q, int min_periods,
                        double* d_out, cudaStream_t stream);

}  // namespace cnvk
