// Internal C++ entry points of the kernels (one per .cu file), called by capi.cu.
#pragma once
#include "prims.cuh"

namespace cnvk {

// rolling.cu
int rolling_window_stat(const double* d_x, long long n, long long wing, int mode, double q, int min_periods,
                        double* d_out, cudaStream_t stream);

// peak.cu
int fp64_peak(int iters, int reps, double* tflops, cudaStream_t stream);

// cbs.cu
struct CbsParams {
    double alpha;
    int nperm, min_width, kmax, nmin;
    double eta;
    int hybrid;  // p.method == "hybrid" (DNAcopy default)
    int prune;   // block-bound pruning of the arc scan (results identical either way)
    u64 seed;
};
struct CbsSeg {
    int arm;
    long long lo, hi;
    double mean;
};
struct CbsStats {
    long long nodes, levels, obs_subtiles_total, obs_subtiles_scanned, obs_arcs, perm_arcs, perm_nodes, perms_run,
        flank_perm_tests;
};
}  // namespace cnvk
#include <vector>
namespace cnvk {
int cbs_segment(const double* d_x, const double* d_w, const long long* h_off, int n_arms, const CbsParams& P,
                std::vector<CbsSeg>& segs, CbsStats* stats, cudaStream_t stream);
void cbs_getbdry(double eta, int nperm, int max_ones, std::vector<int>& out);

}  // namespace cnvk
