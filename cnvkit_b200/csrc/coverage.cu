// Coverage finalisation -- the numeric tail of cnvlib/coverage.py interval_coverages_pileup (:637-643) and
// interval_coverages_bedgraph (:403-409): per bin, depth = basecount / (end - start) for bins of positive width
// (zero-width or reversed user bins keep depth 0), log2 = log2(depth) where depth > 0, else NULL_LOG2_COVERAGE.
// One coalesced pass: 24 B read + 16 B written per bin (HBM-bound).
#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

__global__ void __launch_bounds__(256)
coverage_finalize_kernel(const double* __restrict__ basecount, const long long* __restrict__ start,
                         const long long* __restrict__ end, long long n, double null_log2, double* __restrict__ depth,
                         double* __restrict__ log2v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long span = end[i] - start[i];
    double d = 0.0;
    if (span > 0) d = __ddiv_rn(basecount[i], (double)span);
    depth[i] = d;
    log2v[i] = (d > 0.0) ? log2(d) : null_log2;      // NaN depth (missing basecount) compares false, as in pandas
}

int coverage_finalize(const double* d_basecount, const long long* d_start, const long long* d_end, long long n,
                      double null_log2, double* d_depth, double* d_log2, cudaStream_t stream) {
    if (n <= 0) return 0;
    ProfScope ps(PC_FIX, stream);
    coverage_finalize_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_basecount, d_start, d_end, n, null_log2, d_depth, d_log2);
    CNVK_LAUNCH_CHECK();
    return 0;
}

}  // namespace cnvk
