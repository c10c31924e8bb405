/* cnvkit_b200 -- C ABI of the B200-native CNVkit hot path (libcnvkit_b200.so).
 *
 * The reference (etal/cnvkit) is pure Python and has no FFI layer; the seam it offers is a set
 * of Python callables.  Every entry point below replaces the numeric body of one of those
 * callables (file:line cited per function, relative to the reference tree) and is what a
 * ctypes/cffi stub inside the reference would bind -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / pandas types.
 *   - `*_dev` functions take DEVICE pointers and a cudaStream_t passed as void* (0 = legacy
 *     default stream); work is enqueued on that stream and the call returns without
 *     synchronising unless documented.  Functions without the suffix take HOST pointers, do
 *     their own H2D/D2H copies and return after the result is in host memory.
 *   - float columns are float64, coordinate/count columns int64, row order is the caller's.
 *   - return value 0 = success; >0 = cudaError_t; <0 = argument/shape error.  The message of
 *     the last failure on this thread is available from cnvk_last_error().
 *   - there is NO CPU fallback: every function needs a CUDA device and fails otherwise.
 */
#ifndef CNVKIT_B200_H
#define CNVKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ----------------------------------------------------------------------- */
const char* cnvk_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int cnvk_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench "gpu_launches"). */
unsigned long long cnvk_launch_count(void);
/* Device properties the host layer wants without importing torch: SM count, free/total bytes. */
int cnvk_device_info(int device, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes);

/* Measured FP64 FMA throughput of the device (TFLOP/s, best of `reps` runs of a register-only DFMA
 * stream): the roofline denominator bench.py uses for the CBS arc scan. */
int cnvk_fp64_peak(int iters, int reps, double* tflops);

/* Per-kernel-class device timing (CUDA events on the launching stream), off by default.
 * cnvk_prof_get synchronises on the recorded events and returns accumulated milliseconds and the
 * number of timed scopes per class; class names come from cnvk_prof_name(0..cnvk_prof_count()-1). */
void cnvk_prof_enable(int on);
/* NVTX range per kernel class around the launches (also switched on by the environment variable CNVK_NVTX=1), for
 * `ncu --nvtx` / timeline tools. */
void cnvk_nvtx_enable(int on);
void cnvk_prof_reset(void);
int cnvk_prof_count(void);
const char* cnvk_prof_name(int cls);
int cnvk_prof_get(double* ms, unsigned long long* scopes, int n);

/* ---- smoothing --------------------------------------------------------------------- */
/* cnvlib/smoothing.py:77-87 rolling_median  (pandas rolling(2*wing+1, 1, center=True).median()
 * over the mirror padding of smoothing.py:72-74).  `wing` is the integer half-window the host
 * derives with smoothing._width2wing (smoothing.py:53-69); requires n >= 2, 1 <= wing <= n-1.
 * NaN inputs are skipped as pandas does (even valid counts average the two middle values;
 * an all-NaN window yields NaN). */
int cnvk_rolling_median_dev(const double* d_x, int64_t n, int64_t wing, double* d_out, void* stream);
int cnvk_rolling_median(const double* x, int64_t n, int64_t wing, double* out);

/* cnvlib/smoothing.py:135-141 rolling_quantile (pandas rolling(2*wing+1, 2, center=True)
 * .quantile(q), linear interpolation) over the same mirror padding. */
int cnvk_rolling_quantile_dev(const double* d_x, int64_t n, int64_t wing, double q, double* d_out,
                              void* stream);
int cnvk_rolling_quantile(const double* x, int64_t n, int64_t wing, double q, double* out);

/* ---- sorting ------------------------------------------------------------------------ */
/* np.argsort(x, kind="mergesort") as used by cnvlib/fix.py:414 -- stable, NaN last. */
int cnvk_argsort_f64_dev(const double* d_x, int64_t n, uint32_t* d_order, void* stream);

/* ---- fix: bin-level normalisation (device-pointer primitives) ---------------------------- */
/* cnvlib/fix.py:295-321 mask_bad_bins OR'ed with the sample-NaN test of fix.py:244; keep[i]=1 for
 * bins that survive.  d_sample_log2, d_ref_depth, d_ref_gc may be NULL (column absent). */
int cnvk_fix_mask_dev(const double* d_sample_log2, const double* d_ref_log2, const double* d_ref_spread,
                      const double* d_ref_depth, const double* d_ref_gc, int64_t n, uint8_t* d_keep,
                      void* stream);

/* cnvlib/cnary.py:253-313 CopyNumArray.center_all(estimator=median, by_chrom=True, skip_low):
 * log2 += -(median over chromosomes of the per-chromosome median of the selected bins).
 * d_sel (nullable) is the host-derived autosome(+PAR) row mask, d_chrom_offsets[nchrom+1] the
 * device CSR of chromosome runs; with skip_low the drop_low_coverage test (cnary.py:315-335) is
 * applied on the fly.  *d_shift_out (nullable, device) receives the applied shift. */
int cnvk_center_all_dev(double* d_log2, const double* d_depth, const uint8_t* d_sel, int64_t n,
                        const int64_t* d_chrom_offsets, int32_t nchrom, int32_t skip_low,
                        double* d_shift_out, void* stream);

/* cnvlib/fix.py:373-427 center_by_window with bias_smoother="median": d_perm is
 * np.random.default_rng(0xA5EED).permutation(n) (fix.py:404-405, a pure function of n, generated
 * on the host), wing = smoothing._width2wing(fraction, n).  In place on d_log2, genomic order
 * is preserved (the reference re-sorts at fix.py:426). */
int cnvk_center_by_window_dev(double* d_log2, const double* d_key, const uint32_t* d_perm, int64_t n,
                              int64_t wing, void* stream);
int cnvk_center_by_window(double* log2, const double* key, const uint32_t* perm, int64_t n, int64_t wing);

/* cnvlib/fix.py:430-512 get_edge_bias(cnarr, margin): gains - losses per bin; d_chrom_id marks
 * chromosome runs so neighbours never cross a chromosome boundary. */
int cnvk_edge_bias_dev(const int64_t* d_start, const int64_t* d_end, const int32_t* d_chrom_id, int64_t n,
                       int64_t margin, double* d_out, void* stream);

/* One bin set (targets or antitargets) of cnvlib/fix.py:249-291 load_adjust_coverages after row
 * filtering -- and of cnvlib/reference.py:641-666 bias_correct_logr: [center_all(skip_low)] ->
 * "mostly empty" guard (fix.py:252-258; *skipped = 1 and nothing else is done) -> centring by
 * window on the traits named in `order` ('g' = d_gc, 'e' = edge bias, 'r' = d_rmask; fix uses
 * "ger", reference "gre"); NULL trait columns / do_edge = 0 are skipped.  d_edge_key (nullable)
 * supplies a precomputed edge-bias column (reference.py:563), otherwise it is computed from
 * start/end/chrom_id with `margin` (params.INSERT_SIZE).  Synchronises the stream once. */
int cnvk_fix_group_dev(double* d_log2, const double* d_depth, const uint8_t* d_auto_sel, int64_t n,
                       const int64_t* d_chrom_offsets, int32_t nchrom, const int32_t* d_chrom_id,
                       const int64_t* d_start, const int64_t* d_end, const double* d_gc, const double* d_rmask,
                       const double* d_edge_key, int32_t do_edge, int32_t skip_low, int32_t do_center,
                       int64_t wing, const uint32_t* d_perm, int64_t margin, const char* order,
                       int32_t* skipped, void* stream);

/* Cohort form of center_by_window: the order in which fix.py:404-414 visits the bins (fixed shuffle,
 * then stable argsort by the trait) depends on the bin set only.  cnvk_trait_order_dev computes it once
 * (d_order[t] = row visited t-th); cnvk_center_by_order_dev applies fix.py:417-426 along it (rolling
 * median of log2, subtracted in place); cnvk_fix_group_ordered_dev is cnvk_fix_group_dev with up to three
 * precomputed orders (NULL = skip), applied in the order given. */
int cnvk_trait_order_dev(const double* d_key, const uint32_t* d_perm, int64_t n, uint32_t* d_order, void* stream);
int cnvk_center_by_order_dev(double* d_log2, const uint32_t* d_order, int64_t n, int64_t wing, void* stream);
int cnvk_fix_group_ordered_dev(double* d_log2, const double* d_depth, const uint8_t* d_auto_sel, int64_t n,
                               const int64_t* d_chrom_offsets, int32_t nchrom, const uint32_t* d_order0,
                               const uint32_t* d_order1, const uint32_t* d_order2, int32_t skip_low, int32_t do_center,
                               int64_t wing, int32_t* skipped, void* stream);

/* cnvlib/reference.py:625-667 bias_correct_logr for one normal sample: center_all(skip_low) ->
 * shift_sex_chroms (reference.py:670-714: log2 += flat_logr; XX sample: chrY bins = -1, XY
 * sample: chrX/chrY bins += 1; d_sex_chrom bit0 = chrX row, bit1 = chrY row) -> "mostly empty"
 * guard -> centring by window (fraction 0.1 => wing from the host) on GC, RepeatMasker, then the
 * precomputed edge-bias column (reference.py:563); NULL columns are skipped. */
int cnvk_bias_correct_logr_dev(double* d_log2, const double* d_depth, const uint8_t* d_auto_sel, int64_t n,
                               const int64_t* d_chrom_offsets, int32_t nchrom, const double* d_gc,
                               const double* d_rmask, const double* d_edge_key, int32_t skip_low, int64_t wing,
                               const uint32_t* d_perm, const double* d_flat_logr, const uint8_t* d_sex_chrom,
                               int32_t is_xx, int32_t* skipped, void* stream);

/* cnvlib/fix.py:212-214: log2 -= ref_log2; weight = apply_weights(...) (fix.py:515-617, epsilon
 * 1e-4); center_all(skip_low=True).  d_is_anti = gene in ANTITARGET_ALIASES (host string test).
 * d_info (nullable, device, 8 doubles) receives [tgt_loc, tgt_sd, anti_loc, anti_sd,
 * anti_sum_sqrt_size, n_anti, n_anti_ok, 0] for the log messages of fix.py:571-595. */
int cnvk_fix_finish_dev(double* d_log2, const double* d_depth, const double* d_ref_log2,
                        const double* d_ref_spread, const int64_t* d_start, const int64_t* d_end,
                        const int32_t* d_chrom_id, const uint8_t* d_is_anti, const uint8_t* d_auto_sel,
                        int64_t n, const int64_t* d_chrom_offsets, int32_t nchrom, double* d_weight,
                        double* d_info, void* stream);

/* ---- robust statistics ------------------------------------------------------------------- */
/* Exact per-segment medians (pandas Series.median: NaN skipped, even counts averaged) of the
 * rows with d_use != 0 (nullable = all rows); d_cnt receives the contributing counts. */
int cnvk_segmented_median_dev(const double* d_x, const uint8_t* d_use, const int64_t* d_seg_offsets,
                              int32_t nseg, double* d_med, int32_t* d_cnt, void* stream);

/* cnvlib/descriptives.py:96-142 biweight_location and :188-222 biweight_midvariance of one array
 * (host buffers).  out[0] = location (or *initial if given), out[1] = midvariance about out[0]. */
int cnvk_biweight(const double* x, int64_t n, const double* initial, double* out);

/* cnvlib/reference.py:717-741 summarize_info building block: per column of the row-major
 * [rows, ncols] matrix (rows = samples): mode&1 -> biweight location, mode&2 -> midvariance about
 * that location. */
int cnvk_biweight_columns_dev(const double* d_mat, int32_t rows, int64_t ncols, int32_t mode, double* d_loc,
                              double* d_var, void* stream);
int cnvk_biweight_columns(const double* mat, int32_t rows, int64_t ncols, int32_t mode, double* loc,
                          double* var);

/* ---- segmentation pre-filters / post-processing -------------------------------------------- */
/* cnvlib/segmentation/__init__.py:372-398 drop_outliers(cnarr, 50, factor) via
 * smoothing.rolling_outlier_quantile(x, 50, q, factor) (smoothing.py:439-463) for every arm of a
 * batch: d_mask[i] = 1 for outlier bins; arms of <= 50 bins are never flagged.  arm_offsets is a
 * HOST array.  Synchronises the stream. */
int cnvk_outlier_mask_dev(const double* d_log2, const int64_t* arm_offsets, int32_t n_arms, double q,
                          double factor, uint8_t* d_mask, void* stream);

/* float("%.{digits}g" % x) element-wise (exact decimal round-half-even): the quantisation every
 * float column undergoes on its way to DNAcopy (segmentation/__init__.py:299-301, float_format
 * "%.6g").  *unhandled counts elements outside the exactly-handled range (|x| < 1e-16 or > 1e21),
 * which are copied through for the host to finish.  Synchronises when unhandled != NULL. */
int cnvk_round_sig_dev(const double* d_x, int64_t n, int32_t digits, double* d_out, uint64_t* unhandled,
                       void* stream);

/* Numeric part of transfer_fields (segmentation/__init__.py:489-501): per segment over its bin
 * rows [d_lo, d_hi): seg_weight = nansum(weight) (weight NULL = 1 per bin), seg_depth = weighted
 * mean of depth over non-NaN weights, 0 when seg_weight == 0. */
int cnvk_segment_bin_sums_dev(const double* d_depth, const double* d_weight, const int64_t* d_lo,
                              const int64_t* d_hi, int32_t nseg, double* d_seg_weight, double* d_seg_depth,
                              void* stream);
/* The same for tables whose bins nest (`end` not ascending within a chromosome): skgenome/intersect.py
 * _irange_nested, mode "outer" -- of the rows [d_lo, d_hi) only those with d_bin_end[row] > d_seg_start[seg]
 * belong to the segment. */
int cnvk_segment_bin_sums_masked_dev(const double* d_depth, const double* d_weight, const int64_t* d_lo,
                                     const int64_t* d_hi, const int64_t* d_bin_end, const int64_t* d_seg_start,
                                     int32_t nseg, double* d_seg_weight, double* d_seg_depth, void* stream);

/* ---- circular binary segmentation ------------------------------------------------------ */
/* Replaces the Rscript + DNAcopy::segment(cna, weights, alpha) subprocess of
 * cnvlib/segmentation/__init__.py:292-322 (R script cnvlib/segmentation/cbs.py:1-78).  Field
 * defaults are DNAcopy's (cbs.py:60 passes only weights and alpha): nperm 10000, min_width 2,
 * kmax 25, nmin 200, eta 0.05, hybrid 1; seed is cbs.py:49's 0xA5EED. */
typedef struct cnvk_cbs_params {
    double alpha;
    int32_t nperm, min_width, kmax, nmin;
    double eta;
    int32_t hybrid; /* p.method == "hybrid" */
    int32_t prune;  /* bit 0: skip arc tiles whose upper bound cannot reach the current best;
                     * bit 1: always sum the tail-probability series (no closed-form bracket).
                     * Results are identical for every value; 1 is the fast setting. */
    uint64_t seed;
} cnvk_cbs_params;

typedef struct cnvk_cbs_stats {
    int64_t nodes, levels, obs_subtiles_total, obs_subtiles_scanned, obs_arcs, perm_arcs, perm_nodes,
        perms_run, flank_perm_tests;
    int64_t prep_bins; /* bins centred + prefix-summed, summed over recursion levels (ABI >= 2) */
} cnvk_cbs_stats;

/* One call segments every arm of a batch: bins of arm a are [arm_offsets[a], arm_offsets[a+1])
 * of the (already filtered, %.6g-quantised) log2/weight columns; weights must be > 0
 * (cbs.py:22-35).  Segments come back in arm order as half-open bin ranges [seg_lo, seg_hi)
 * with seg_mean = weighted.mean(log2, weight) (cbs.py:70-71).  arm_offsets and all outputs are
 * HOST arrays (the table is tiny); *n_segs receives the count, which must be <= capacity
 * (capacity = number of bins is always enough). */
int cnvk_cbs_segment_dev(const double* d_log2, const double* d_weight, const int64_t* arm_offsets,
                         int32_t n_arms, const cnvk_cbs_params* params, int64_t capacity,
                         int32_t* seg_arm, int64_t* seg_lo, int64_t* seg_hi, double* seg_mean,
                         int64_t* n_segs, cnvk_cbs_stats* stats, void* stream);
int cnvk_cbs_segment(const double* log2, const double* weight, const int64_t* arm_offsets,
                     int32_t n_arms, const cnvk_cbs_params* params, int64_t capacity, int32_t* seg_arm,
                     int64_t* seg_lo, int64_t* seg_hi, double* seg_mean, int64_t* n_segs,
                     cnvk_cbs_stats* stats);
/* Everything between a .cnr table and DNAcopy's input, fused for device-resident columns: the per-arm bin filters
 * of cnvlib/segmentation/__init__.py:239-282 (outlier mask from cnvk_outlier_mask_dev or NULL, weight == 0), the
 * "%.6g" TSV quantisation (:299-301) and the R script's row filters (cbs.py:22-35: per arm, no non-NA weight ->
 * weight 1.0 for every bin, else NA / <= 0 weights dropped; NA log2 dropped), survivors compacted in order.
 * d_arm_id[n] = arm of every bin, arm_offsets (HOST) the unfiltered arm offsets.  Outputs: d_log2_out / d_weight_out /
 * d_src_rows (capacity n), unit_offsets[n_arms + 1] (HOST) = arm offsets of the survivors, *flags (HOST): bit 0 a
 * log2 is not finite (caller must drop those bins first), bit 1 a value outside the device quantiser's range. */
int cnvk_cbs_input_filter_dev(const double* d_log2, const double* d_weight, const uint8_t* d_outlier,
                              const int64_t* d_arm_id, const int64_t* arm_offsets, int32_t n_arms, int64_t n,
                              double* d_log2_out, double* d_weight_out, int64_t* d_src_rows, int64_t* unit_offsets,
                              uint32_t* flags, void* stream);

/* ---- cohort fast path: a batch of samples on one shared bin set -------------------------------------------- */
/* cnvlib/batch.py:524-566 (batch_run_sample) runs do_fix + do_segmentation for every sample against the SAME targets,
 * antitargets and reference.  A cnvk_panel holds what depends on the bins only; all d_* members are device pointers
 * owned by the caller, arm_off is a HOST array.  One bin set (targets or antitargets) after the reference masks of
 * load_adjust_coverages (fix.py:226-262): */
typedef struct cnvk_panel_group {
    int64_t n0;                 /* rows of the sample's coverage table */
    int64_t n;                  /* rows kept by the reference masks */
    const int64_t* d_keep_idx;  /* [n] coverage-table row of every kept row */
    const uint8_t* d_auto;      /* [n] autosome selector of center_all */
    const int64_t* d_chrom_off; /* [nchrom + 1] chromosome offsets of the kept rows */
    int32_t nchrom;
    int32_t skip_low;           /* center_all(skip_low=...) of this group */
    const uint32_t* d_order[3]; /* visiting orders of center_by_window for GC, edge, RepeatMasker (cnvk_trait_order_dev;
                                   NULL = correction not applied) */
    int64_t wing;               /* rolling-median half window of the corrections */
} cnvk_panel_group;

typedef struct cnvk_panel {
    cnvk_panel_group grp[2];    /* targets, antitargets */
    int64_t n;                  /* rows of the .cnr = grp[0].n + grp[1].n */
    const int64_t* d_order;     /* [n] row of cat(targets, antitargets) found at every .cnr row (genomic order) */
    const double* d_ref_log2;   /* [n] reference columns in .cnr row order */
    const double* d_ref_spread;
    const int64_t* d_start;     /* [n] */
    const int64_t* d_end;
    const int32_t* d_chrom_id;
    const uint8_t* d_is_anti;
    const uint8_t* d_auto;
    const int64_t* d_chrom_off; /* [nchrom + 1] */
    int32_t nchrom;
    int32_t n_arms;
    const int64_t* d_arm_id;    /* [n] arm of every .cnr row (GenomicArray.by_arm) */
    const int64_t* arm_off;     /* HOST [n_arms + 1] */
} cnvk_panel;

/* status bits of cnvk_panel_prepare_dev: a flagged sample's outputs are to be discarded and the sample redone through
 * the step-by-step entry points (which implement the reference's handling of these cases) */
#define CNVK_PANEL_NAN_INPUT 1u       /* a NaN log2 in a coverage table: the kept set changes (fix.py:244) */
#define CNVK_PANEL_LOW_COVERAGE 2u    /* "most bins have no or very low coverage": corrections skipped (fix.py:275) */
#define CNVK_PANEL_NONFINITE 4u       /* a non-finite ratio reached the CBS input filter */
#define CNVK_PANEL_QUANTISER_RANGE 8u /* a value outside the device %.6g quantiser's range */

/* sizeof(cnvk_panel) as this library was compiled: a binding checks its own struct definition against it */
int64_t cnvk_panel_sizeof(void);

/* do_fix (fix.py:22-85) for n_samples samples, optionally followed by the CBS input preparation of
 * segmentation/__init__.py:239-301 + cbs.py:22-35 (cnvk_outlier_mask_dev with factor skip_outliers when > 0, then
 * cnvk_cbs_input_filter_dev), with ONE synchronisation for the whole batch.
 * d_in[4 s + {0,1,2,3}]  : target log2, target depth, antitarget log2, antitarget depth of sample s (device, grp.n0 rows)
 * d_out[3 s + {0,1,2}]   : .cnr log2, depth, weight of sample s (device, capacity n)
 * d_cbs[2 s + {0,1}]     : quantised survivors log2 / weight (device, capacity n); d_src[s]: their .cnr rows
 * unit_offsets (HOST)    : [n_samples][n_arms + 1] arm offsets of the survivors;  status (HOST): [n_samples] bits above
 * d_cbs / d_src / unit_offsets may be NULL when want_cbs == 0. */
int cnvk_panel_prepare_dev(const cnvk_panel* panel, int32_t n_samples, const double* const* d_in, double* const* d_out,
                           int32_t want_cbs, double skip_outliers, double* const* d_cbs, int64_t* const* d_src,
                           int64_t* unit_offsets, uint32_t* status, void* stream);

/* Numeric part of transfer_fields (segmentation/__init__.py:401-506) for all segments of a batch.  seg (HOST) is
 * [4][n_seg]: sample, arm, first and one-past-last survivor index (the CBS output, relative to the sample's survivors),
 * sorted by (sample, arm, position).  Per segment: source rows of its first / last bin, start and end after the per-arm
 * stretching (:431-437), the .cnr row range it spans (intersect.py:323-346, mode "outer"; bins must not nest) and
 * weight = nansum(weight), depth = weighted mean depth over that range.
 * bounds (HOST) [6][n_seg]: r_lo, r_hi, seg_start, seg_end, bin_lo, bin_hi;  sums (HOST) [2][n_seg]: weight, depth. */
int cnvk_panel_transfer_dev(const cnvk_panel* panel, int32_t n_samples, int32_t n_seg, const int64_t* seg,
                            const int64_t* const* d_src, const double* const* d_depth, const double* const* d_weight,
                            int64_t* bounds, double* sums, void* stream);

/* ---- coverage finalisation (the step before the path) ------------------------------ */
/* cnvlib/coverage.py:637-643 (interval_coverages_pileup) and :403-409 (interval_coverages_bedgraph): depth =
 * basecount / (end - start) for bins of positive width, else 0; log2 = log2(depth) where depth > 0, else null_log2
 * (params.NULL_LOG2_COVERAGE = -20).  basecount is float64 (bedcov counts are far below 2^53). */
int cnvk_coverage_finalize_dev(const double* d_basecount, const int64_t* d_start, const int64_t* d_end, int64_t n,
                               double null_log2, double* d_depth, double* d_log2, void* stream);
int cnvk_coverage_finalize(const double* basecount, const int64_t* start, const int64_t* end, int64_t n,
                           double null_log2, double* depth, double* log2);

/* DNAcopy::smooth.CNA -- the `--smooth-cbs` option: cnvlib/segmentation/cbs.py:51-58 runs `cna = smooth.CNA(cna)`
 * (DNAcopy defaults: smooth_region 10, outlier_sd_scale 4, smooth_sd_scale 2, trim 0.025) on the arm's quantised
 * log2 before segment(); segment means are still taken from the unsmoothed values (cbs.py:64-73).  Per arm:
 * trimmed.SD from the smallest round((1 - 2 trim)(n - 1)) absolute first differences, then every bin that lies more
 * than outlier_sd_scale trimmed.SD above (below) all other bins within smooth_region bins is moved to the window
 * median +- smooth_sd_scale trimmed.SD.  arm_offsets is a HOST array; d_sd / sd (one trimmed.SD per arm) may be NULL. */
int cnvk_smooth_cna_dev(const double* d_log2, const int64_t* arm_offsets, int32_t n_arms, int32_t smooth_region,
                        double outlier_sd_scale, double smooth_sd_scale, double trim, double* d_out, double* d_sd,
                        void* stream);
int cnvk_smooth_cna(const double* log2, const int64_t* arm_offsets, int32_t n_arms, int32_t smooth_region,
                    double outlier_sd_scale, double smooth_sd_scale, double trim, double* out, double* sd);
/* Diagnostics of the CBS decision chain (DNAcopy wfindcpt; call site
 * cnvlib/segmentation/cbs.py:49-60): every interval [node_offsets[k], node_offsets[k+1]) is tested ONCE as a
 * node, nothing is decided early -- all nperm permutations run (no sequential stopping) and both ternary-split
 * flank tests run their permutations.  Host buffers.  Per node k:
 *   out6[6k..]  = sqrt(T^2) observed, tailp (hybrid nodes, else -1), getmncwt delta (hybrid, else -1),
 *                 0.99999 T^2 (what permutation statistics are compared with), flank p-values 1 and 2 (-1 when
 *                 the maximal arc touches an end of the node)
 *   iout5[5k..] = arc i, arc j (prefix points relative to the node), permutations whose statistic reached the
 *                 observed one, 1 = hybrid (n > nmin) / 0 = exact, 1 = constant data (no test).
 * This is what tests/test_cbs_stats_gpu.py compares with true-shuffle Monte-Carlo estimates (north_star:
 * "permutation p-values agree within Monte Carlo standard error"). */
int cnvk_cbs_node_diag(const double* log2, const double* weight, const int64_t* node_offsets,
                       int32_t n_nodes, const cnvk_cbs_params* params, double* out6, int32_t* iout5);
/* DNAcopy getbdry(eta, nperm, max.ones): sequential stopping boundary,
 * out[max_ones*(max_ones+1)/2]. Host-only. */
int cnvk_cbs_getbdry(double eta, int32_t nperm, int32_t max_ones, int32_t* out);

/* ---- reference build: GC / RepeatMasker content from the genome ------------------------------ */
/* cnvlib/reference.py:859-880 get_fasta_stats / calculate_gc_lo: for byte range r of a FASTA file
 * image (the bin's sequence under the .fai line geometry; line breaks and non-ACGT letters do not
 * count), gc[r] = (#G + #C) / #ACGT and rmask[r] = #lowercase ACGT / #ACGT in either case, 0 when
 * the range holds no ACGT.  Ranges are clamped to [0, nbytes). */
int cnvk_fasta_gc_lo_dev(const uint8_t* d_bytes, int64_t nbytes, const int64_t* d_range_lo,
                         const int64_t* d_range_hi, int64_t n_ranges, double* d_gc, double* d_rmask,
                         void* stream);
int cnvk_fasta_gc_lo(const uint8_t* bytes, int64_t nbytes, const int64_t* range_lo, const int64_t* range_hi,
                     int64_t n_ranges, double* gc, double* rmask);

/* ---- segment metrics ------------------------------------------------------------------------ */
/* cnvlib/segmetrics.py:176-245 confidence_interval_bootstrap (the 'ci' interval statistic of
 * do_segmetrics, :24-121, which `cnvkit.py batch` runs after segmentation, batch.py:559-566) for every
 * segment of a table.  Segment s resamples the bins [seg_lo[s], seg_hi[s]) of values / weights with
 * numpy's default_rng(0xA5EED).integers(0, k, (bootstraps, k)) -- the PCG64 stream and Lemire's
 * bounded-integer rejection are reproduced draw for draw -- and takes percentiles of the weighted
 * means, each summed in np.average's pairwise order.  bootstraps is raised to ceil(2/alpha) like the
 * reference does.  smooth_upto: -1 never smooth (BCa correction, :292-351), otherwise smooth
 * segments of <= smooth_upto bins (INT32_MAX = always): N(0, k^-1/4 sqrt(1-w)) noise on every
 * resampled value (:248-289; the reference's noise generator is unseeded, ours is counter-based
 * on noise_seed).  pcg_state / pcg_inc = {high, low} words of the generator's initial 128-bit state
 * and increment; all zero selects numpy's PCG64(0xA5EED).  k == 0 -> NaN, k == 1 -> the bin's value.
 * seg_lo / seg_hi are HOST arrays. */
typedef struct cnvk_boot_params {
    double alpha;
    int32_t bootstraps;
    int32_t smooth_upto;
    uint64_t pcg_state[2];
    uint64_t pcg_inc[2];
    uint64_t noise_seed;
} cnvk_boot_params;
int cnvk_bootstrap_ci_dev(const double* d_values, const double* d_weights, const int64_t* seg_lo,
                          const int64_t* seg_hi, int32_t n_segs, const cnvk_boot_params* params,
                          double* d_ci_lo, double* d_ci_hi, void* stream);
int cnvk_bootstrap_ci(const double* values, const double* weights, int64_t n, const int64_t* seg_lo,
                      const int64_t* seg_hi, int32_t n_segs, const cnvk_boot_params* params, double* ci_lo,
                      double* ci_hi);

/* ---- HaarSeg ------------------------------------------------------------------------------ */
/* cnvlib/segmentation/haar.py:257-371 haarSeg(signal, breaksFdrQ, weights) for every arm of a batch
 * (haarStartLevel 1, haarEndLevel 5; the arms are those of segment_haar's by_arm(), haar.py:80-82).
 * d_weight may be NULL (no weight column).  Same batching and output convention as
 * cnvk_cbs_segment; seg_mean is SegmentByPeaks' (weighted) mean (haar.py:405-451). */
int cnvk_haar_segment_dev(const double* d_log2, const double* d_weight, const int64_t* arm_offsets,
                          int32_t n_arms, double fdr_q, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo,
                          int64_t* seg_hi, double* seg_mean, int64_t* n_segs, void* stream);
int cnvk_haar_segment(const double* log2, const double* weight, const int64_t* arm_offsets, int32_t n_arms,
                      double fdr_q, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo, int64_t* seg_hi,
                      double* seg_mean, int64_t* n_segs);
/* The depth + BAF breakpoint union of cnvlib/segmentation/haar.py:188-225 (one_chrom_baf) needs three more pieces:
 * haarSeg with a caller-supplied noise estimate per arm (sigma_override of haar.py:334-339; NaN = use the signal's
 * own estimate), that estimate for the compacted observed BAF values (haar.py:154-170 _observed_baf_sigma:
 * 1.4826 * median |HaarConv(values, None, 1)|, floored at 1e-10; arms of fewer than 2 values get the floor), and
 * SegmentByPeaks' (weighted) means for breakpoints merged on the host (haar.py:405-451).  Host buffers. */
int cnvk_haar_segment_sigma(const double* signal, const double* weight, const int64_t* arm_offsets, int32_t n_arms,
                            double fdr_q, const double* sigma_override, int64_t capacity, int32_t* seg_arm,
                            int64_t* seg_lo, int64_t* seg_hi, double* seg_mean, int64_t* n_segs);
int cnvk_haar_sigma(const double* signal, const int64_t* arm_offsets, int32_t n_arms, double* sigma);
int cnvk_haar_segment_means(const double* signal, const double* weight, int64_t n, const int64_t* seg_lo,
                            const int64_t* seg_hi, int32_t n_segs, double* seg_mean);

/* ---- table I/O (host only) ------------------------------------------------------------------ */
/* Reader for the tab-separated .cnn/.cnr/.cns tables: skgenome/tabio/tab.py:18-33 read_tab
 * (pandas.read_csv(sep="\t", dtype={"chromosome": str})).  cnvk_tsv_open indexes the file (header +
 * rows, blank lines skipped, minimal quoting); columns are then converted on demand.  Missing values
 * follow pandas' default na_values; cnvk_tsv_col_kind is pandas' dtype inference (1 int64, 2 float64,
 * 0 str).  The caller applies read_tab's "drop rows without log2" rule on the columns. */
typedef struct cnvk_tsv cnvk_tsv;
int cnvk_tsv_open(const char* path, cnvk_tsv** out);
void cnvk_tsv_close(cnvk_tsv* h);
int cnvk_tsv_shape(const cnvk_tsv* h, int64_t* nrows, int32_t* ncols);
const char* cnvk_tsv_colname(const cnvk_tsv* h, int32_t col);
int cnvk_tsv_col_kind(const cnvk_tsv* h, int32_t col);
int cnvk_tsv_read_f64(const cnvk_tsv* h, int32_t col, double* out);
int cnvk_tsv_read_i64(const cnvk_tsv* h, int32_t col, int64_t* out);
/* String column as bytes + offsets[nrows+1]; buf == NULL only reports *nbytes. */
int cnvk_tsv_read_str(const cnvk_tsv* h, int32_t col, int64_t* offsets, char* buf, int64_t cap, int64_t* nbytes);
/* FNV-1a of the raw field bytes of the given columns: equality test of bin sets without strings. */
uint64_t cnvk_tsv_hash_cols(const cnvk_tsv* h, const int32_t* cols, int32_t n);
/* String column dictionary-coded in order of first appearance: codes[nrows] and the distinct values
 * (offsets[n_unique + 1] into buf).  Call with codes == NULL for the sizes first. */
int cnvk_tsv_read_str_dict(const cnvk_tsv* h, int32_t col, int32_t* codes, int64_t* offsets, char* buf,
                           int64_t cap, int64_t* n_unique, int64_t* nbytes);

/* Writer: skgenome/tabio/__init__.py:175-195 write (DataFrame.to_csv(sep="\t", index=False,
 * float_format="%.6g")).  kinds[c]: 0 str (ptrs[c] = int64 offsets[nrows+1], ptrs2[c] = bytes),
 * 1 int64, 2 float64 (NaN -> empty field), 3 dictionary-coded str (ptrs[c] = int32 codes[nrows]
 * into the distinct values of ptrs2[c] = cnvk_tsv_dict*, -1 = missing -> empty field).  Rows are
 * formatted by several host threads and written in order. */
typedef struct cnvk_tsv_dict {
    const int64_t* offsets; /* n + 1 offsets into blob */
    const char* blob;       /* the distinct values, concatenated (UTF-8) */
    int64_t n;
} cnvk_tsv_dict;
int cnvk_tsv_write(const char* path, int32_t ncols, const char* const* names, const int32_t* kinds,
                   const void* const* ptrs, const void* const* ptrs2, int64_t nrows);

/* skgenome/combiners.py:58-94 join_strings per bin range -- the string half of transfer_fields
 * (cnvlib/segmentation/__init__.py:467-501), host only: for every row range [lo[s], hi[s]) of a dictionary-coded
 * gene column the distinct codes in order of first appearance; codes < 0 (missing) and codes with keep[code] == 0
 * (ignored placeholder / antitarget names) are skipped.  mask_end / mask_start (both NULL or both given) restrict a
 * range to rows with mask_end[row] > mask_start[s] (nested bins).  CSR output out_off[nseg + 1], out_codes[capacity]. */
int cnvk_ordered_unique_ranges(const int64_t* codes, const int64_t* lo, const int64_t* hi, int32_t nseg, int64_t ncodes,
                               const uint8_t* keep, const int64_t* mask_end, const int64_t* mask_start, int64_t* out_off,
                               int64_t* out_codes, int64_t capacity);
/* The same pass with the names joined natively ("a,b,c" per range, "-" when nothing remains): name_blob /
 * name_off[ncodes + 1] carry one kept name per code (empty = no name).  For gene columns in which every code holds at
 * most one name and no name is shared between codes. */
int cnvk_join_names_ranges(const int64_t* codes, const int64_t* lo, const int64_t* hi, int32_t nseg, int64_t ncodes,
                           const char* name_blob, const int64_t* name_off, const int64_t* mask_end,
                           const int64_t* mask_start, char* out_blob, int64_t* out_off, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* CNVKIT_B200_H */
