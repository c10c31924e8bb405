// Cohort fast path: fix + CBS input preparation, and the numeric part of transfer_fields, for a BATCH of samples that
// share one bin set (cnvlib/batch.py:524-566 runs do_fix + do_segmentation per sample on the same targets /
// antitargets / reference).  Everything that depends on the bins only lives in a cnvk_panel descriptor (device
// pointers owned by the caller); a sample is four coverage columns.
//
// The per-sample host round trips of the step-by-step entry points (NaN test, the "mostly empty" guard of
// load_adjust_coverages, the survivor counts of the CBS input filter) are folded into ONE synchronisation per batch:
// the guards are evaluated speculatively on the device and reported as status bits, and a flagged sample is simply
// redone by the caller through the general entry points.
//
//   cnvk_panel_prepare_dev : fix.py:22-85 do_fix for the panel's bin set (load_adjust_coverages for both groups, row
//                            order of the .cnr, apply_weights, center_all) + segmentation/__init__.py:239-301 (outlier
//                            mask, zero-weight drop, %.6g) + cbs.py:22-35 row filters
//   cnvk_panel_transfer_dev: segmentation/__init__.py:401-506 transfer_fields per arm -- arm-end stretching, the bin
//                            range of every segment (intersect.py:323-346 _irange_simple, mode "outer") and the
//                            weight / depth sums, for all segments of all samples of the batch in two launches
#include "ops.cuh"
#include "prof.cuh"

#include "../../include/cnvkit_b200.h"

namespace cnvk {

__global__ void nan_any_kernel(const double* __restrict__ x, long long n, unsigned long long* __restrict__ flag) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        bad |= (v != v);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1ull);
}

__global__ void gather_rows2_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                    const long long* __restrict__ idx, long long n, double* __restrict__ oa,
                                    double* __restrict__ ob) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const long long s = idx[t];
        oa[t] = a[s];
        ob[t] = b[s];
    }
}

// .cnr row t takes row order[t] of cat(targets, antitargets)
__global__ void merge_order_kernel(const double* __restrict__ tl, const double* __restrict__ td,
                                   const double* __restrict__ al, const double* __restrict__ ad, long long nt,
                                   const long long* __restrict__ order, long long n, double* __restrict__ ol,
                                   double* __restrict__ od) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const long long s = order[t];
        if (s < nt) { ol[t] = tl[s]; od[t] = td[s]; }
        else { ol[t] = al[s - nt]; od[t] = ad[s - nt]; }
    }
}

int panel_prepare(const cnvk_panel* P, int ns, const double* const* in, double* const* out, int want_cbs,
                  double skip_outliers, double* const* cbs_out, long long* const* src_out, long long* h_units,
                  unsigned int* h_status, cudaStream_t stream) {
    if (ns <= 0) return 0;
    const long long n = P->n;
    const int n_arms = P->n_arms;
    Scratch scratch(stream);
    // per sample: [0] NaN seen, [1] / [2] bins above the coverage floor per group; then the filter flags
    unsigned long long* d_stat;
    unsigned int* d_fflags;
    long long *d_units, *d_arm_off;
    CNVK_CHECK(scratch.alloc(&d_stat, (size_t)ns * 3));
    CNVK_CHECK(scratch.alloc(&d_fflags, (size_t)ns));
    CNVK_CHECK(scratch.alloc(&d_units, (size_t)ns * (n_arms + 1)));
    CNVK_CHECK(scratch.alloc(&d_arm_off, (size_t)n_arms + 1));
    CNVK_CHECK(cudaMemsetAsync(d_stat, 0, (size_t)ns * 3 * sizeof(unsigned long long), stream));
    CNVK_CHECK(cudaMemsetAsync(d_fflags, 0, (size_t)ns * sizeof(unsigned int), stream));
    CNVK_CHECK(cudaMemsetAsync(d_units, 0, (size_t)ns * (n_arms + 1) * sizeof(long long), stream));
    CNVK_CHECK(cudaMemcpyAsync(d_arm_off, P->arm_off, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    std::vector<long long> h_arm_off(P->arm_off, P->arm_off + n_arms + 1);
    const long long nt = P->grp[0].n, na = P->grp[1].n;
    double* g[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    for (int k = 0; k < 2; ++k) {
        const long long m = P->grp[k].n;
        CNVK_CHECK(scratch.alloc(&g[k][0], (size_t)std::max<long long>(m, 1)));
        CNVK_CHECK(scratch.alloc(&g[k][1], (size_t)std::max<long long>(m, 1)));
    }
    unsigned char* d_mask = nullptr;
    if (want_cbs && skip_outliers > 0.0) CNVK_CHECK(scratch.alloc(&d_mask, (size_t)std::max<long long>(n, 1)));
    for (int s = 0; s < ns; ++s) {
        for (int k = 0; k < 2; ++k) {
            const cnvk_panel_group& G = P->grp[k];
            if (G.n0 <= 0 || G.n <= 0) continue;
            const double* in_l = in[4 * s + 2 * k];
            const double* in_d = in[4 * s + 2 * k + 1];
            {
                ProfScope ps(PC_FIX, stream);
                nan_any_kernel<<<(int)std::min<long long>(div_up(G.n0, 256), 592), 256, 0, stream>>>(in_l, G.n0, d_stat + 3 * s);
                CNVK_LAUNCH_CHECK();
                gather_rows2_kernel<<<div_up(G.n, 256), 256, 0, stream>>>(in_l, in_d, (const long long*)G.d_keep_idx, G.n, g[k][0], g[k][1]);
                CNVK_LAUNCH_CHECK();
            }
            const u32* orders[3] = {G.d_order[0], G.d_order[1], G.d_order[2]};
            int rc = fix_group_ordered_enqueue(g[k][0], g[k][1], G.d_auto, G.n, (const long long*)G.d_chrom_off, G.nchrom,
                                               orders, 3, G.skip_low, 1, G.wing, d_stat + 3 * s + 1 + k, stream);
            if (rc) return rc;
        }
        double *o_l = out[3 * s], *o_d = out[3 * s + 1], *o_w = out[3 * s + 2];
        if (n > 0) {
            {
                ProfScope ps(PC_FIX, stream);
                merge_order_kernel<<<div_up(n, 256), 256, 0, stream>>>(g[0][0], g[0][1], g[1][0], g[1][1], nt,
                                                                       (const long long*)P->d_order, n, o_l, o_d);
                CNVK_LAUNCH_CHECK();
            }
            int rc = fix_finish(o_l, o_d, P->d_ref_log2, P->d_ref_spread, (const long long*)P->d_start,
                                (const long long*)P->d_end, P->d_chrom_id, P->d_is_anti, P->d_auto, n,
                                (const long long*)P->d_chrom_off, P->nchrom, o_w, nullptr, stream);
            if (rc) return rc;
            if (want_cbs) {
                if (d_mask) {
                    rc = outlier_mask(o_l, h_arm_off.data(), n_arms, 0.95, skip_outliers, d_mask, stream);
                    if (rc) return rc;
                }
                rc = cbs_input_filter_enqueue(o_l, o_w, d_mask, (const long long*)P->d_arm_id, d_arm_off, n_arms, n,
                                              cbs_out[2 * s], cbs_out[2 * s + 1], src_out[s],
                                              d_units + (size_t)s * (n_arms + 1), d_fflags + s, stream);
                if (rc) return rc;
            }
        }
    }
    std::vector<unsigned long long> h_stat((size_t)ns * 3);
    std::vector<unsigned int> h_ff((size_t)ns);
    CNVK_CHECK(cudaMemcpyAsync(h_stat.data(), d_stat, h_stat.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaMemcpyAsync(h_ff.data(), d_fflags, h_ff.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    if (want_cbs)
        CNVK_CHECK(cudaMemcpyAsync(h_units, d_units, (size_t)ns * (n_arms + 1) * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    (void)na;
    for (int s = 0; s < ns; ++s) {
        unsigned int st = 0;
        if (h_stat[3 * s]) st |= CNVK_PANEL_NAN_INPUT;
        for (int k = 0; k < 2; ++k) {
            const cnvk_panel_group& G = P->grp[k];
            if (G.n0 > 0 && G.n > 0 && (long long)h_stat[3 * s + 1 + k] <= G.n / 2) st |= CNVK_PANEL_LOW_COVERAGE;
        }
        if (h_ff[s] & 1u) st |= CNVK_PANEL_NONFINITE;
        if (h_ff[s] & 2u) st |= CNVK_PANEL_QUANTISER_RANGE;
        h_status[s] = st;
    }
    return 0;
}

// ------------------------------------------------------------------ transfer_fields, numeric part
// one thread per segment (segments sorted by sample, arm, position)
__global__ void panel_seg_bounds_kernel(const long long* __restrict__ seg_sample, const long long* __restrict__ seg_arm,
                                        const long long* __restrict__ s_lo, const long long* __restrict__ s_hi, int nseg,
                                        const long long* const* __restrict__ src_rows,
                                        const long long* __restrict__ start, const long long* __restrict__ end,
                                        const long long* __restrict__ arm_off,
                                        long long* __restrict__ out /* [6][nseg]: r_lo, r_hi, seg_start, seg_end, bin_lo, bin_hi */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    const long long smp = seg_sample[i], arm = seg_arm[i];
    const long long* src = src_rows[smp];
    const long long r_lo = src[s_lo[i]], r_hi = src[s_hi[i] - 1];
    const bool first = (i == 0) || seg_sample[i - 1] != smp || seg_arm[i - 1] != arm;
    const bool last = (i == nseg - 1) || seg_sample[i + 1] != smp || seg_arm[i + 1] != arm;
    const long long a_lo = arm_off[arm], a_hi = arm_off[arm + 1];
    const long long seg_start = first ? start[a_lo] : start[r_lo];
    const long long seg_end = last ? end[a_hi - 1] : end[r_hi];
    // bin_hi: first row of the arm whose start >= seg_end (searchsorted left); bin_lo: first row whose end > seg_start
    long long lo = a_lo, hi = a_hi;
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (start[m] < seg_end) lo = m + 1; else hi = m; }
    long long bin_hi = lo;
    lo = a_lo; hi = a_hi;
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (end[m] <= seg_start) lo = m + 1; else hi = m; }
    const long long bin_lo = lo;
    if (bin_hi < bin_lo) bin_hi = bin_lo;
    out[i] = r_lo;
    out[(size_t)nseg + i] = r_hi;
    out[(size_t)2 * nseg + i] = seg_start;
    out[(size_t)3 * nseg + i] = seg_end;
    out[(size_t)4 * nseg + i] = bin_lo;
    out[(size_t)5 * nseg + i] = bin_hi;
}

// one CTA per segment; the columns of the segment's own sample
__global__ void __launch_bounds__(256)
panel_seg_sums_kernel(const long long* __restrict__ seg_sample, const long long* __restrict__ bounds, int nseg,
                      const double* const* __restrict__ depth_cols, const double* const* __restrict__ weight_cols,
                      double* __restrict__ out /* [2][nseg]: weight, depth */) {
    __shared__ double s_w[8], s_dw[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    const long long smp = seg_sample[s];
    const double* __restrict__ depth = depth_cols[smp];
    const double* __restrict__ weight = weight_cols[smp];
    const long long k0 = bounds[(size_t)4 * nseg + s], k1 = bounds[(size_t)5 * nseg + s];
    double sw = 0.0, sdw = 0.0;
    for (long long kb = k0 + threadIdx.x; kb < k1; kb += 4 * 256) {
        double w[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long k = kb + u * 256;
            const bool ok = k < k1;
            w[u] = ok ? weight[k] : __longlong_as_double(0x7ff8000000000000ll);
            d[u] = ok ? depth[k] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (w[u] == w[u]) { sw += w[u]; sdw += d[u] * w[u]; }
    }
    sw = warp_sum(sw);
    sdw = warp_sum(sdw);
    if (lane == 0) { s_w[warp] = sw; s_dw[warp] = sdw; }
    __syncthreads();
    if (threadIdx.x == 0) {
        sw = 0.0; sdw = 0.0;
        for (int k = 0; k < 8; ++k) { sw += s_w[k]; sdw += s_dw[k]; }
        out[s] = sw;
        out[(size_t)nseg + s] = (sw > 0.0) ? sdw / sw : 0.0;
    }
}

int panel_transfer(const cnvk_panel* P, int ns, int nseg, const long long* h_seg /* [4][nseg] sample, arm, lo, hi */,
                   const long long* const* src_rows, const double* const* depth_cols, const double* const* weight_cols,
                   long long* h_bounds /* [6][nseg] */, double* h_sums /* [2][nseg] */, cudaStream_t stream) {
    if (nseg <= 0) return 0;
    ProfScope ps(PC_OTHER, stream);
    Scratch scratch(stream);
    long long *d_seg, *d_bounds, *d_arm_off;
    const void** d_ptrs;
    double* d_sums;
    CNVK_CHECK(scratch.alloc(&d_seg, (size_t)4 * nseg));
    CNVK_CHECK(scratch.alloc(&d_bounds, (size_t)6 * nseg));
    CNVK_CHECK(scratch.alloc(&d_sums, (size_t)2 * nseg));
    CNVK_CHECK(scratch.alloc(&d_arm_off, (size_t)P->n_arms + 1));
    CNVK_CHECK(scratch.alloc(&d_ptrs, (size_t)3 * ns));
    std::vector<const void*> h_ptrs((size_t)3 * ns);
    for (int s = 0; s < ns; ++s) {
        h_ptrs[s] = src_rows[s];
        h_ptrs[(size_t)ns + s] = depth_cols[s];
        h_ptrs[(size_t)2 * ns + s] = weight_cols[s];
    }
    CNVK_CHECK(cudaMemcpyAsync(d_seg, h_seg, (size_t)4 * nseg * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemcpyAsync(d_arm_off, P->arm_off, ((size_t)P->n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemcpyAsync(d_ptrs, h_ptrs.data(), h_ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice, stream));
    panel_seg_bounds_kernel<<<div_up(nseg, 128), 128, 0, stream>>>(
        d_seg, d_seg + nseg, d_seg + (size_t)2 * nseg, d_seg + (size_t)3 * nseg, nseg, (const long long* const*)d_ptrs,
        (const long long*)P->d_start, (const long long*)P->d_end, d_arm_off, d_bounds);
    CNVK_LAUNCH_CHECK();
    panel_seg_sums_kernel<<<nseg, 256, 0, stream>>>(d_seg, d_bounds, nseg, (const double* const*)(d_ptrs + ns),
                                                    (const double* const*)(d_ptrs + (size_t)2 * ns), d_sums);
    CNVK_LAUNCH_CHECK();
    CNVK_CHECK(cudaMemcpyAsync(h_bounds, d_bounds, (size_t)6 * nseg * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaMemcpyAsync(h_sums, d_sums, (size_t)2 * nseg * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

}  // namespace cnvk
