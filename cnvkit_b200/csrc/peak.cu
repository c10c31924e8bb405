// FP64 pipe peak microbenchmark (dependent-chain-free DFMA stream), used by bench.py as the
// roofline denominator for the CBS arc scan -- MEASURED_PEAKS.json has no FP64 figure.
#include "ops.cuh"

namespace cnvk {

__global__ void __launch_bounds__(256) fp64_fma_kernel(double* __restrict__ out, int iters, double b, double c) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int k = 0; k < iters; ++k) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int fp64_peak(int iters, int reps, double* tflops, cudaStream_t stream) {
    int dev = 0, sms = 148;
    CNVK_CHECK(cudaGetDevice(&dev));
    CNVK_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 8;
    Scratch scratch(stream);
    double* out;
    CNVK_CHECK(scratch.alloc(&out, (size_t)grid * 256));
    cudaEvent_t a, b;
    CNVK_CHECK(cudaEventCreate(&a));
    CNVK_CHECK(cudaEventCreate(&b));
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        CNVK_CHECK(cudaEventRecord(a, stream));
        fp64_fma_kernel<<<grid, 256, 0, stream>>>(out, iters, 0.999999, 1e-6);
        CNVK_LAUNCH_CHECK();
        CNVK_CHECK(cudaEventRecord(b, stream));
        CNVK_CHECK(cudaEventSynchronize(b));
        float ms = 0.f;
        CNVK_CHECK(cudaEventElapsedTime(&ms, a, b));
        const double tf = 2.0 * 8.0 * (double)iters * (double)grid * 256.0 / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return 0;
}

}  // namespace cnvk
