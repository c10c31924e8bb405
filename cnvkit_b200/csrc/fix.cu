// Bin-level normalisation kernels behind cnvlib.fix.do_fix / reference.bias_correct_logr.
//
//   fix_mask            fix.py:295-321 mask_bad_bins + the sample-NaN test of fix.py:244
//   center_all          cnary.py:253-313 (+ drop_low_coverage cnary.py:315-335)
//   center_by_window    fix.py:373-427  (shuffle -> stable argsort by trait -> rolling median ->
//                        subtract -> back to genomic order), one call, everything on device
//   edge_bias           fix.py:430-512  get_edge_bias / edge_losses / edge_gains
//   fix_finish          fix.py:212-214  ratio = log2 - ref_log2, apply_weights (fix.py:515-617),
//                        final center_all(skip_low=True)
// Elementwise kernels are one coalesced pass each (HBM-bound); all order statistics are exact.
#include "ops.cuh"
#include "prof.cuh"
#include <algorithm>

namespace cnvk {

typedef unsigned char u8;

static const double MIN_REF_COVERAGE = -5.0;   // cnvlib/params.py
static const double MAX_REF_SPREAD = 1.0;
static const double NULL_LOG2_COVERAGE = -20.0;
static const double GC_MIN_FRACTION = 0.3, GC_MAX_FRACTION = 0.7;

// ------------------------------------------------------------------ mask_bad_bins
__global__ void fix_mask_kernel(const double* __restrict__ sample_log2, const double* __restrict__ ref_log2,
                                const double* __restrict__ ref_spread, const double* __restrict__ ref_depth,
                                const double* __restrict__ ref_gc, long long n, u8* __restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double l = ref_log2[i], s = ref_spread[i];
    bool bad = (l < MIN_REF_COVERAGE) || (l > -MIN_REF_COVERAGE) || (s > MAX_REF_SPREAD) || (l != l) || (s != s);
    if (ref_depth) { const double d = ref_depth[i]; bad = bad || (d == 0.0) || (d != d); }
    if (ref_gc) { const double g = ref_gc[i]; bad = bad || (g > GC_MAX_FRACTION) || (g < GC_MIN_FRACTION); }
    if (sample_log2) { const double x = sample_log2[i]; bad = bad || (x != x); }
    keep[i] = bad ? 0 : 1;
}

int fix_mask(const double* d_sample_log2, const double* d_ref_log2, const double* d_ref_spread,
             const double* d_ref_depth, const double* d_ref_gc, long long n, u8* d_keep, cudaStream_t stream) {
    if (n <= 0) return 0;
    ProfScope ps(PC_FIX, stream);
    fix_mask_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_sample_log2, d_ref_log2, d_ref_spread, d_ref_depth, d_ref_gc, n, d_keep);
    CNVK_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ center_all
__global__ void use_mask_kernel(const double* __restrict__ log2, const double* __restrict__ depth,
                                const u8* __restrict__ sel, long long n, int skip_low, u8* __restrict__ use) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool u = sel ? (sel[i] != 0) : true;
    if (skip_low) {
        const double x = log2[i];
        bool drop = (x < NULL_LOG2_COVERAGE - MIN_REF_COVERAGE) || (x != x);
        if (depth) { const double d = depth[i]; drop = drop || (d == 0.0) || (d != d); }
        u = u && !drop;
    }
    use[i] = u ? 1 : 0;
}

__global__ void add_scalar_kernel(double* __restrict__ x, long long n, const double* __restrict__ shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __dadd_rn(x[i], *shift);
}

int center_all(double* d_log2, const double* d_depth, const u8* d_sel, long long n, const long long* d_chrom_off,
               int nchrom, int skip_low, double* d_shift_out, cudaStream_t stream) {
    if (n <= 0 || nchrom <= 0) {
        if (d_shift_out) CNVK_CHECK(cudaMemsetAsync(d_shift_out, 0, sizeof(double), stream));
        return 0;
    }
    Scratch scratch(stream);
    u8* use;
    double *med, *shift;
    int* cnt;
    CNVK_CHECK(scratch.alloc(&use, (size_t)n));
    CNVK_CHECK(scratch.alloc(&med, (size_t)nchrom));
    CNVK_CHECK(scratch.alloc(&cnt, (size_t)nchrom));
    CNVK_CHECK(scratch.alloc(&shift, 1));
    {
        ProfScope ps(PC_FIX, stream);
        use_mask_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, d_depth, d_sel, n, skip_low, use);
        CNVK_LAUNCH_CHECK();
    }
    int rc = segmented_median(d_log2, use, d_chrom_off, nchrom, med, cnt, stream);
    if (rc) return rc;
    rc = median_of_medians(med, cnt, nchrom, shift, -1.0, stream);
    if (rc) return rc;
    {
        ProfScope ps(PC_FIX, stream);
        add_scalar_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, n, shift);
        CNVK_LAUNCH_CHECK();
    }
    if (d_shift_out) CNVK_CHECK(cudaMemcpyAsync(d_shift_out, shift, sizeof(double), cudaMemcpyDeviceToDevice, stream));
    return 0;
}

// ------------------------------------------------------------------ center_by_window
__global__ void gather_f64_kernel(const double* __restrict__ src, const u32* __restrict__ idx, long long n,
                                  double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = src[idx[t]];
}

// src_of[t] = perm[order[t]];  xs[t] = log2[src_of[t]]
__global__ void compose_gather_kernel(const u32* __restrict__ perm, const u32* __restrict__ order,
                                      const double* __restrict__ log2, long long n, u32* __restrict__ src_of,
                                      double* __restrict__ xs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const u32 s = perm[order[t]];
        src_of[t] = s;
        xs[t] = log2[s];
    }
}

__global__ void scatter_sub_kernel(double* __restrict__ log2, const u32* __restrict__ src_of,
                                   const double* __restrict__ xs, const double* __restrict__ bias, long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) log2[src_of[t]] = __dsub_rn(xs[t], bias[t]);
}

__global__ void compose_kernel(const u32* __restrict__ perm, const u32* __restrict__ order, long long n,
                               u32* __restrict__ src_of) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) src_of[t] = perm[order[t]];
}

// The order in which center_by_window visits the bins -- shuffle by the fixed permutation, then stable
// argsort by the trait (fix.py:404-414) -- depends on the bin set only, not on the sample: a cohort
// computes it once per (trait, bin set) and reuses it for every sample.
int trait_order(const double* d_key, const u32* d_perm, long long n, u32* d_src_of, cudaStream_t stream) {
    if (n <= 0) return 0;
    Scratch scratch(stream);
    double* kperm;
    u32* order;
    CNVK_CHECK(scratch.alloc(&kperm, (size_t)n));
    CNVK_CHECK(scratch.alloc(&order, (size_t)n));
    {
        ProfScope ps(PC_FIX, stream);
        gather_f64_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_key, d_perm, n, kperm);
        CNVK_LAUNCH_CHECK();
    }
    int rc = argsort_f64(kperm, n, order, stream);
    if (rc) return rc;
    ProfScope ps(PC_FIX, stream);
    compose_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_perm, order, n, d_src_of);
    CNVK_LAUNCH_CHECK();
    return 0;
}

// center_by_window given the visiting order: rolling median of log2 along it, subtracted in place
int center_by_order(double* d_log2, const u32* d_src_of, long long n, long long wing, cudaStream_t stream) {
    if (n <= 0) return 0;
    Scratch scratch(stream);
    double *xs, *bias;
    CNVK_CHECK(scratch.alloc(&xs, (size_t)n));
    CNVK_CHECK(scratch.alloc(&bias, (size_t)n));
    {
        ProfScope ps(PC_FIX, stream);
        gather_f64_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, d_src_of, n, xs);
        CNVK_LAUNCH_CHECK();
    }
    if (n < 2) {
        CNVK_CHECK(cudaMemcpyAsync(bias, xs, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    } else {
        int rc = rolling_window_stat(xs, n, wing, 0, 0.5, 1, bias, stream);
        if (rc) return rc;
    }
    ProfScope ps(PC_FIX, stream);
    scatter_sub_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, d_src_of, xs, bias, n);
    CNVK_LAUNCH_CHECK();
    return 0;
}

int center_by_window(double* d_log2, const double* d_key, const u32* d_perm, long long n, long long wing,
                     cudaStream_t stream) {
    if (n <= 0) return 0;
    Scratch scratch(stream);
    double *kperm, *xs, *bias;
    u32 *order, *src_of;
    CNVK_CHECK(scratch.alloc(&kperm, (size_t)n));
    CNVK_CHECK(scratch.alloc(&xs, (size_t)n));
    CNVK_CHECK(scratch.alloc(&bias, (size_t)n));
    CNVK_CHECK(scratch.alloc(&order, (size_t)n));
    CNVK_CHECK(scratch.alloc(&src_of, (size_t)n));
    {
        ProfScope ps(PC_FIX, stream);
        gather_f64_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_key, d_perm, n, kperm);
        CNVK_LAUNCH_CHECK();
    }
    int rc = argsort_f64(kperm, n, order, stream);
    if (rc) return rc;
    {
        ProfScope ps(PC_FIX, stream);
        compose_gather_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_perm, order, d_log2, n, src_of, xs);
        CNVK_LAUNCH_CHECK();
    }
    if (n < 2) {
        // smoothing.rolling_median returns the input unchanged for len < 2 (smoothing.py:79-82)
        CNVK_CHECK(cudaMemcpyAsync(bias, xs, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    } else {
        rc = rolling_window_stat(xs, n, wing, 0, 0.5, 1, bias, stream);
        if (rc) return rc;
    }
    {
        ProfScope ps(PC_FIX, stream);
        scatter_sub_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, src_of, xs, bias, n);
        CNVK_LAUNCH_CHECK();
    }
    return 0;
}

// ------------------------------------------------------------------ edge bias
// per bin, with its left/right neighbours on the same chromosome (chrom_id marks the runs)
__global__ void edge_bias_kernel(const long long* __restrict__ start, const long long* __restrict__ end,
                                 const int* __restrict__ chrom_id, long long n, long long margin,
                                 double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = start[i], e = end[i], t = e - s;
    const long long I = margin;
    // edge_losses: i/(2t) - (i-t)^2/(2it) if t < i
    double loss = (double)I / (double)(2 * t);
    if (t < I) loss = loss - (double)((I - t) * (I - t)) / (double)(2 * I * t);
    double gains = 0.0;
    // left neighbour (gap between its end and our start)
    if (i > 0 && chrom_id[i - 1] == chrom_id[i]) {
        long long g = s - end[i - 1];
        if (g < I) {
            if (g < 0) g = 0;
            double gn = (double)((I - g) * (I - g)) / (double)(4 * I * t);
            if (t + g < I) gn = gn - (double)((I - t - g) * (I - t - g)) / (double)(4 * I * t);
            gains = gains + gn;
        }
    }
    if (i + 1 < n && chrom_id[i + 1] == chrom_id[i]) {
        long long g = start[i + 1] - e;
        if (g < I) {
            if (g < 0) g = 0;
            double gn = (double)((I - g) * (I - g)) / (double)(4 * I * t);
            if (t + g < I) gn = gn - (double)((I - t - g) * (I - t - g)) / (double)(4 * I * t);
            gains = gains + gn;
        }
    }
    out[i] = gains - loss;
}

int edge_bias(const long long* d_start, const long long* d_end, const int* d_chrom_id, long long n, long long margin,
              double* d_out, cudaStream_t stream) {
    if (n <= 0) return 0;
    ProfScope ps(PC_FIX, stream);
    edge_bias_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_start, d_end, d_chrom_id, n, margin, d_out);
    CNVK_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ one bin set of load_adjust_coverages
__global__ void __launch_bounds__(256)
count_gt_kernel(const double* __restrict__ x, long long n, double thr, unsigned long long* __restrict__ out) {
    u32 c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        c += (x[i] > thr) ? 1u : 0u;
    c = warp_sum_u32(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// fix.py:249-291 (and reference.py:641-666 with shifts applied by the caller in between):
// center_all, the "mostly empty" guard, then GC / edge / RepeatMasker centring by window in the
// order given by `order` (fix: gc,edge,rmask = "ger"; reference: gc,rmask,edge = "gre").
// reference.py:670-714 shift_sex_chroms: log2 += flat; XX sample: chrY = -1; XY sample: X,Y += 1
__global__ void shift_sex_kernel(double* __restrict__ log2, const double* __restrict__ flat,
                                 const u8* __restrict__ xy, long long n, int is_xx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = __dadd_rn(log2[i], flat[i]);
    const u8 s = xy[i];   // bit0: chrX (PAR-aware), bit1: chrY
    if (is_xx) { if (s & 2) v = -1.0; }
    else if (s & 3) v = __dadd_rn(v, 1.0);
    log2[i] = v;
}

// the same with the visiting orders precomputed (trait_order): one rolling median per trait, no sort
int fix_group_ordered(double* d_log2, const double* d_depth, const u8* d_auto_sel, long long n,
                      const long long* d_chrom_off, int nchrom, const u32* const* d_orders, int n_orders, int skip_low,
                      int do_center, long long wing, int* h_skipped, cudaStream_t stream) {
    if (h_skipped) *h_skipped = 0;
    if (n <= 0) return 0;
    int rc;
    if (do_center) {
        rc = center_all(d_log2, d_depth, d_auto_sel, n, d_chrom_off, nchrom, skip_low, nullptr, stream);
        if (rc) return rc;
    }
    Scratch scratch(stream);
    unsigned long long* d_cnt;
    CNVK_CHECK(scratch.alloc(&d_cnt, 1));
    CNVK_CHECK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), stream));
    {
        ProfScope ps(PC_FIX, stream);
        count_gt_kernel<<<std::min(div_up(n, 256), 1184), 256, 0, stream>>>(d_log2, n, NULL_LOG2_COVERAGE - MIN_REF_COVERAGE, d_cnt);
        CNVK_LAUNCH_CHECK();
    }
    unsigned long long h_cnt = 0;
    CNVK_CHECK(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    if ((long long)h_cnt <= n / 2) {
        if (h_skipped) *h_skipped = 1;   // "most bins have no or very low coverage"
        return 0;
    }
    for (int k = 0; k < n_orders; ++k) {
        if (!d_orders[k]) continue;
        rc = center_by_order(d_log2, d_orders[k], n, wing, stream);
        if (rc) return rc;
    }
    return 0;
}

// The same without the host round trip of the "mostly empty" guard: the count of bins above the coverage floor goes to
// *d_cnt (device) and the corrections run regardless; the caller reads the count later and, for the pathological sample
// whose count is <= n / 2, discards the result (cnvk_panel_prepare_dev).
int fix_group_ordered_enqueue(double* d_log2, const double* d_depth, const u8* d_auto_sel, long long n,
                              const long long* d_chrom_off, int nchrom, const u32* const* d_orders, int n_orders,
                              int skip_low, int do_center, long long wing, unsigned long long* d_cnt,
                              cudaStream_t stream) {
    if (n <= 0) return 0;
    int rc;
    if (do_center) {
        rc = center_all(d_log2, d_depth, d_auto_sel, n, d_chrom_off, nchrom, skip_low, nullptr, stream);
        if (rc) return rc;
    }
    {
        ProfScope ps(PC_FIX, stream);
        count_gt_kernel<<<std::min(div_up(n, 256), 1184), 256, 0, stream>>>(d_log2, n, NULL_LOG2_COVERAGE - MIN_REF_COVERAGE, d_cnt);
        CNVK_LAUNCH_CHECK();
    }
    for (int k = 0; k < n_orders; ++k) {
        if (!d_orders[k]) continue;
        rc = center_by_order(d_log2, d_orders[k], n, wing, stream);
        if (rc) return rc;
    }
    return 0;
}

int fix_group(double* d_log2, const double* d_depth, const u8* d_auto_sel, long long n,
              const long long* d_chrom_off, int nchrom, const int* d_chrom_id, const long long* d_start,
              const long long* d_end, const double* d_gc, const double* d_rmask, const double* d_edge_in,
              int do_edge, int skip_low, int do_center, long long wing, const u32* d_perm, long long margin,
              const char* order, const double* d_flat, const u8* d_xy, int is_xx, int* h_skipped,
              cudaStream_t stream) {
    if (h_skipped) *h_skipped = 0;
    if (n <= 0) return 0;
    int rc;
    if (do_center) {
        rc = center_all(d_log2, d_depth, d_auto_sel, n, d_chrom_off, nchrom, skip_low, nullptr, stream);
        if (rc) return rc;
    }
    if (d_flat && d_xy) {
        ProfScope ps(PC_FIX, stream);
        shift_sex_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, d_flat, d_xy, n, is_xx);
        CNVK_LAUNCH_CHECK();
    }
    Scratch scratch(stream);
    unsigned long long* d_cnt;
    CNVK_CHECK(scratch.alloc(&d_cnt, 1));
    CNVK_CHECK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), stream));
    {
        ProfScope ps(PC_FIX, stream);
        count_gt_kernel<<<std::min(div_up(n, 256), 1184), 256, 0, stream>>>(d_log2, n, NULL_LOG2_COVERAGE - MIN_REF_COVERAGE, d_cnt);
        CNVK_LAUNCH_CHECK();
    }
    unsigned long long h_cnt = 0;
    CNVK_CHECK(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    if ((long long)h_cnt <= n / 2) {
        if (h_skipped) *h_skipped = 1;   // "most bins have no or very low coverage"
        return 0;
    }
    for (const char* p = order; *p; ++p) {
        if (*p == 'g' && d_gc) {
            rc = center_by_window(d_log2, d_gc, d_perm, n, wing, stream);
            if (rc) return rc;
        } else if (*p == 'r' && d_rmask) {
            rc = center_by_window(d_log2, d_rmask, d_perm, n, wing, stream);
            if (rc) return rc;
        } else if (*p == 'e' && do_edge) {
            const double* key = d_edge_in;
            double* edge = nullptr;
            if (!key) {
                CNVK_CHECK(scratch.alloc(&edge, (size_t)n));
                rc = edge_bias(d_start, d_end, d_chrom_id, n, margin, edge, stream);
                if (rc) return rc;
                key = edge;
            }
            rc = center_by_window(d_log2, key, d_perm, n, wing, stream);
            if (rc) return rc;
        }
    }
    return 0;
}

// ------------------------------------------------------------------ ratio + weights + final centring
__global__ void sub_kernel(double* __restrict__ x, const double* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __dsub_rn(x[i], y[i]);
}

// group mask: which==0 -> targets (!is_anti), 1 -> antitargets
__global__ void group_mask_kernel(const u8* __restrict__ is_anti, int which, long long n, u8* __restrict__ sel) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sel[i] = ((is_anti[i] != 0) == (which != 0)) ? 1 : 0;
}

// resid[i] = log2[i] - median(chrom of i) for used bins, NaN otherwise (NaN sorts last)
__global__ void residual_kernel(const double* __restrict__ log2, const u8* __restrict__ use,
                                const int* __restrict__ chrom_id, const double* __restrict__ med, long long n,
                                double* __restrict__ resid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    resid[i] = use[i] ? __dsub_rn(log2[i], med[chrom_id[i]]) : __longlong_as_double(0x7ff8000000000000ll);
}

// sum of sqrt(end-start) and count over a group; also counts used (non-low) bins.  Two stages with a
// fixed accumulation order (per-CTA partials, then one warp adds them left to right): run-to-run
// deterministic, unlike floating-point atomics, and not limited to one SM's bandwidth.
static const int GS_CTAS = 128;

__global__ void __launch_bounds__(256)
group_size_part_kernel(const long long* __restrict__ start, const long long* __restrict__ end,
                       const u8* __restrict__ sel, const u8* __restrict__ use, long long n, double* __restrict__ part) {
    __shared__ double s_s[8], s_c[8], s_u[8];
    double s = 0.0, c = 0.0, u = 0.0;
    // contiguous slice per CTA, strided inside it: the order of every partial sum is fixed
    const long long per = (n + GS_CTAS - 1) / GS_CTAS;
    const long long i0 = (long long)blockIdx.x * per, i1 = min(n, i0 + per);
    for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
        if (sel[i]) {
            s += sqrt((double)(end[i] - start[i]));
            c += 1.0;
            if (use[i]) u += 1.0;
        }
    }
    s = warp_sum(s); c = warp_sum(c); u = warp_sum(u);
    if ((threadIdx.x & 31) == 0) { s_s[threadIdx.x >> 5] = s; s_c[threadIdx.x >> 5] = c; s_u[threadIdx.x >> 5] = u; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { s += s_s[k]; c += s_c[k]; u += s_u[k]; }
        part[3 * blockIdx.x + 0] = s; part[3 * blockIdx.x + 1] = c; part[3 * blockIdx.x + 2] = u;
    }
}

__global__ void group_size_final_kernel(const double* __restrict__ part, double* __restrict__ acc) {
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int k = 0; k < GS_CTAS; ++k) v += part[3 * k + threadIdx.x];
        acc[threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256)
ref_flags_kernel(const double* __restrict__ spread, const double* __restrict__ ref_log2, long long n, double eps,
                 int* __restrict__ flags /*[2]*/) {
    int a = 0, b = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (spread[i] > eps) a = 1;
        // np.mod(x, 1): result has the sign of the divisor -> in [0, 1)
        const double r = ref_log2[i];
        double m = fmod(r, 1.0);
        if (m < 0.0) m += 1.0;
        if (fabs(m) > eps) b = 1;
    }
    if (a) flags[0] = 1;
    if (b) flags[1] = 1;
}

__global__ void weights_kernel(const long long* __restrict__ start, const long long* __restrict__ end,
                               const u8* __restrict__ is_anti, const double* __restrict__ spread,
                               const double* __restrict__ bw /*[2][2]: loc, midvar per group*/,
                               const double* __restrict__ acc /*[2][3]*/, const int* __restrict__ flags, long long n,
                               double eps, double* __restrict__ weight) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = is_anti[i] ? 1 : 0;
    const double sd = bw[g * 2 + 1];
    const double var = sd * sd;
    const double mean_sz = acc[g * 3 + 0] / acc[g * 3 + 1];
    const double bin_sz = sqrt((double)(end[i] - start[i]));
    const double simple = 1.0 - var / (bin_sz / mean_sz);
    double wt = simple;
    if (flags[0] && flags[1]) {
        const double fancy = 1.0 - spread[i] * spread[i];
        const double xx = 0.9;
        wt = xx * fancy + (1 - xx) * simple;
    }
    if (wt != wt) wt = eps;
    weight[i] = fmin(fmax(wt, eps), 1.0);
}

int fix_finish(double* d_log2, const double* d_depth, const double* d_ref_log2, const double* d_ref_spread,
               const long long* d_start, const long long* d_end, const int* d_chrom_id, const u8* d_is_anti,
               const u8* d_auto_sel, long long n, const long long* d_chrom_off, int nchrom, double* d_weight,
               double* d_info /*[8] optional: tgt_var, anti_var, n_anti, n_anti_ok*/, cudaStream_t stream) {
    if (n <= 0) return 0;
    const double eps = 1e-4;
    Scratch scratch(stream);
    u8 *sel, *use;
    double *med, *resid, *bw, *acc;
    int *cnt, *flags;
    CNVK_CHECK(scratch.alloc(&sel, (size_t)n));
    CNVK_CHECK(scratch.alloc(&use, (size_t)n));
    CNVK_CHECK(scratch.alloc(&med, (size_t)nchrom));
    CNVK_CHECK(scratch.alloc(&cnt, (size_t)nchrom));
    CNVK_CHECK(scratch.alloc(&resid, (size_t)n));
    CNVK_CHECK(scratch.alloc(&bw, 4));
    CNVK_CHECK(scratch.alloc(&acc, 6));
    double* gs_part;
    CNVK_CHECK(scratch.alloc(&gs_part, (size_t)3 * GS_CTAS));
    CNVK_CHECK(scratch.alloc(&flags, 2));
    CNVK_CHECK(cudaMemsetAsync(acc, 0, 6 * sizeof(double), stream));
    CNVK_CHECK(cudaMemsetAsync(flags, 0, 2 * sizeof(int), stream));
    const int g256 = div_up(n, 256);
    {
        ProfScope ps(PC_FIX, stream);
        sub_kernel<<<g256, 256, 0, stream>>>(d_log2, d_ref_log2, n);   // fix.py:212
        CNVK_LAUNCH_CHECK();
        ref_flags_kernel<<<std::min(g256, 1184), 256, 0, stream>>>(d_ref_spread, d_ref_log2, n, eps, flags);
        CNVK_LAUNCH_CHECK();
    }
    // scratch for the sort of residuals
    u64 *k0, *k1;
    u32 *v0, *v1, *nvalid;
    double* sorted;
    CNVK_CHECK(scratch.alloc(&k0, (size_t)n));
    CNVK_CHECK(scratch.alloc(&k1, (size_t)n));
    CNVK_CHECK(scratch.alloc(&v0, (size_t)n));
    CNVK_CHECK(scratch.alloc(&v1, (size_t)n));
    CNVK_CHECK(scratch.alloc(&nvalid, 1));
    CNVK_CHECK(scratch.alloc(&sorted, (size_t)n));
    for (int g = 0; g < 2; ++g) {
        {
            ProfScope ps(PC_FIX, stream);
            group_mask_kernel<<<g256, 256, 0, stream>>>(d_is_anti, g, n, sel);
            CNVK_LAUNCH_CHECK();
            use_mask_kernel<<<g256, 256, 0, stream>>>(d_log2, d_depth, sel, n, 1, use);   // drop_low_coverage
            CNVK_LAUNCH_CHECK();
            group_size_part_kernel<<<GS_CTAS, 256, 0, stream>>>(d_start, d_end, sel, use, n, gs_part);
            CNVK_LAUNCH_CHECK();
            group_size_final_kernel<<<1, 32, 0, stream>>>(gs_part, acc + 3 * g);
            CNVK_LAUNCH_CHECK();
        }
        int rc = segmented_median(d_log2, use, d_chrom_off, nchrom, med, cnt, stream);
        if (rc) return rc;
        {
            ProfScope ps(PC_FIX, stream);
            residual_kernel<<<g256, 256, 0, stream>>>(d_log2, use, d_chrom_id, med, n, resid);
            CNVK_LAUNCH_CHECK();
        }
        rc = sort_values_f64(resid, n, k0, k1, v0, v1, sorted, nvalid, stream);
        if (rc) return rc;
        rc = biweight_of_sorted(sorted, nvalid, 0, nullptr, bw + 2 * g, stream);
        if (rc) return rc;
    }
    {
        ProfScope ps(PC_FIX, stream);
        weights_kernel<<<g256, 256, 0, stream>>>(d_start, d_end, d_is_anti, d_ref_spread, bw, acc, flags, n, eps, d_weight);
        CNVK_LAUNCH_CHECK();
    }
    if (d_info) {
        CNVK_CHECK(cudaMemcpyAsync(d_info, bw, 4 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        CNVK_CHECK(cudaMemcpyAsync(d_info + 4, acc + 3, 3 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    }
    // fix.py:214 cnarr.center_all(skip_low=True)
    return center_all(d_log2, d_depth, d_auto_sel, n, d_chrom_off, nchrom, 1, nullptr, stream);
}

}  // namespace cnvk
