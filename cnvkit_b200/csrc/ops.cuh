// Internal C++ entry points of the kernels (one per .cu file), called by capi.cu.
#pragma once
#include <vector>

#include "prims.cuh"

namespace cnvk {

// rolling.cu
int rolling_window_stat(const double* d_x, long long n, long long wing, int mode, double q, int min_periods,
                        double* d_out, cudaStream_t stream);

// select.cu
int segmented_median(const double* d_x, const unsigned char* d_use, const long long* d_seg_off, int nseg,
                     double* d_med, int* d_cnt, cudaStream_t stream);
int median_of_medians(const double* d_med, const int* d_cnt, int nseg, double* d_out, double sign,
                      cudaStream_t stream);

// biweight.cu
int biweight_of_sorted(const double* d_sorted, const u32* d_nvalid, int n, const double* d_initial, double* d_out,
                       cudaStream_t stream);
int biweight_columns(const double* d_mat, int rows, long long ncols, int mode, double* d_loc, double* d_var,
                     cudaStream_t stream);

// fix.cu
int fix_mask(const double* d_sample_log2, const double* d_ref_log2, const double* d_ref_spread,
             const double* d_ref_depth, const double* d_ref_gc, long long n, unsigned char* d_keep,
             cudaStream_t stream);
int center_all(double* d_log2, const double* d_depth, const unsigned char* d_sel, long long n,
               const long long* d_chrom_off, int nchrom, int skip_low, double* d_shift_out, cudaStream_t stream);
int center_by_window(double* d_log2, const double* d_key, const u32* d_perm, long long n, long long wing,
                     cudaStream_t stream);
int trait_order(const double* d_key, const u32* d_perm, long long n, u32* d_src_of, cudaStream_t stream);
int center_by_order(double* d_log2, const u32* d_src_of, long long n, long long wing, cudaStream_t stream);
int fix_group_ordered(double* d_log2, const double* d_depth, const unsigned char* d_auto_sel, long long n,
                      const long long* d_chrom_off, int nchrom, const u32* const* d_orders, int n_orders, int skip_low,
                      int do_center, long long wing, int* h_skipped, cudaStream_t stream);
int fix_group_ordered_enqueue(double* d_log2, const double* d_depth, const unsigned char* d_auto_sel, long long n,
                              const long long* d_chrom_off, int nchrom, const u32* const* d_orders, int n_orders,
                              int skip_low, int do_center, long long wing, unsigned long long* d_cnt,
                              cudaStream_t stream);
int edge_bias(const long long* d_start, const long long* d_end, const int* d_chrom_id, long long n, long long margin,
              double* d_out, cudaStream_t stream);
int fix_group(double* d_log2, const double* d_depth, const unsigned char* d_auto_sel, long long n,
              const long long* d_chrom_off, int nchrom, const int* d_chrom_id, const long long* d_start,
              const long long* d_end, const double* d_gc, const double* d_rmask, const double* d_edge_in,
              int do_edge, int skip_low, int do_center, long long wing, const u32* d_perm, long long margin,
              const char* order, const double* d_flat, const unsigned char* d_xy, int is_xx, int* h_skipped,
              cudaStream_t stream);
int fix_finish(double* d_log2, const double* d_depth, const double* d_ref_log2, const double* d_ref_spread,
               const long long* d_start, const long long* d_end, const int* d_chrom_id,
               const unsigned char* d_is_anti, const unsigned char* d_auto_sel, long long n,
               const long long* d_chrom_off, int nchrom, double* d_weight, double* d_info, cudaStream_t stream);

// segutil.cu
int outlier_mask(const double* d_x, const long long* h_off, int n_arms, double q, double factor,
                 unsigned char* d_mask, cudaStream_t stream);
int cbs_input_filter_enqueue(const double* d_log2, const double* d_weight, const unsigned char* d_outlier,
                             const long long* d_arm_id, const long long* d_arm_off, int n_arms, long long n,
                             double* d_x_out, double* d_w_out, long long* d_src_out, long long* d_units_out,
                             unsigned int* d_flags_out, cudaStream_t stream);
int outlier_tables(const long long* h_off, int n_arms, std::vector<int>* ex_arm, std::vector<int>* tile_off);
int outlier_mask_enqueue(const double* d_x, long long N, const long long* d_off, const int* d_tile_off, const int* d_ex,
                         int n_ex, int tiles, double q, double factor, unsigned char* d_mask, cudaStream_t stream);
int round_sig(const double* d_x, long long n, int digits, double* d_out, unsigned long long* h_unhandled,
              cudaStream_t stream);
int segment_bin_sums(const double* d_depth, const double* d_weight, const long long* d_lo, const long long* d_hi,
                     const long long* d_bin_end, const long long* d_seg_start, int nseg, double* d_seg_weight,
                     double* d_seg_depth, cudaStream_t stream);

// panel.cu (cnvk_panel is declared in include/cnvkit_b200.h)
}  // namespace cnvk
struct cnvk_panel;
namespace cnvk {
int panel_prepare(const cnvk_panel* P, int ns, const double* const* in, double* const* out, int want_cbs,
                  double skip_outliers, double* const* cbs_out, long long* const* src_out, long long* h_units,
                  unsigned int* h_status, cudaStream_t stream);
int panel_transfer(const cnvk_panel* P, int ns, int nseg, const long long* h_seg, const long long* const* src_rows,
                   const double* const* depth_cols, const double* const* weight_cols, long long* h_bounds,
                   double* h_sums, cudaStream_t stream);

// bootstrap.cu
int bootstrap_ci(const double* d_values, const double* d_weights, const long long* h_lo, const long long* h_hi,
                 int n_segs, double alpha, int bootstraps, int smooth_upto, const u64 pcg[4], u64 noise_seed,
                 double* d_ci_lo, double* d_ci_hi, cudaStream_t stream);

// fasta.cu
int fasta_gc_lo(const unsigned char* d_bytes, long long nbytes, const long long* d_lo, const long long* d_hi,
                long long n, double* d_gc, double* d_rm, cudaStream_t stream);

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
        flank_perm_tests, prep_bins;
};
}  // namespace cnvk
#include <vector>
namespace cnvk {
struct CbsDiag;   // cbs.cu: per-node diagnostics of one level (cnvk_cbs_node_diag)
int cbs_segment(const double* d_x, const double* d_w, const long long* h_off, int n_arms, const CbsParams& P,
                std::vector<CbsSeg>& segs, CbsStats* stats, cudaStream_t stream, CbsDiag* diag = nullptr);
int cbs_node_diag(const double* d_x, const double* d_w, const long long* h_off, int n_nodes, const CbsParams& P,
                  double* out6, int* iout5, cudaStream_t stream);
int cbs_input_filter(const double* d_log2, const double* d_weight, const unsigned char* d_outlier, const long long* d_arm_id,
                     const long long* h_arm_off, int n_arms, long long n, double* d_x_out, double* d_w_out,
                     long long* d_src_out, long long* h_units, unsigned int* h_flags, cudaStream_t stream);
// coverage.cu
int coverage_finalize(const double* d_basecount, const long long* d_start, const long long* d_end, long long n,
                      double null_log2, double* d_depth, double* d_log2, cudaStream_t stream);
// smooth_cna.cu
int smooth_cna(const double* d_x, const long long* h_off, int n_arms, int k, double outlier_scale, double smooth_scale,
               double trim, double* d_out, double* d_sd_out, cudaStream_t stream);
// haar.cu (segments reuse CbsSeg)
int haar_segment(const double* d_x, const double* d_w, const long long* h_off, int n_arms, double q,
                 std::vector<CbsSeg>& segs, cudaStream_t stream, const double* h_sigma_override = nullptr);
int haar_sigma(const double* d_x, const long long* h_off, int n_arms, double* h_sigma, cudaStream_t stream);
int haar_segment_means(const double* d_x, const double* d_w, const long long* h_lo, const long long* h_hi, int nseg,
                       double* h_mean, cudaStream_t stream);
void cbs_getbdry(double eta, int nperm, int max_ones, std::vector<int>& out);

}  // namespace cnvk
