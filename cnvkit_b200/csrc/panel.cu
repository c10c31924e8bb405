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
#include <vector>

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

// ---- one sample: everything from the four coverage columns to the CBS inputs, enqueued on `stream` without a
// synchronisation and without touching host memory (the caller synchronises once per batch)
struct PanelTables {            // device copies of the panel's small host tables
    long long* d_arm_off = nullptr;
    int *d_tile_off = nullptr, *d_ex = nullptr;
    int n_ex = 0, tiles = 0;
};
struct SampleIO {
    const double* in[4];        // target log2 / depth, antitarget log2 / depth (coverage-table rows)
    double *o_l, *o_d, *o_w;    // .cnr columns
    double *xq, *wq;            // CBS inputs (survivors)
    long long* src;             // .cnr rows of the survivors
    unsigned long long* stat;   // [3]: NaN seen, bins above the coverage floor per group
    unsigned int* fflags;       // [1]
    long long* units;           // [n_arms + 1]
    double* g[2][2];            // per-group work columns (grp.n each)
    unsigned char* mask;        // outlier mask (n), nullptr = no outlier test
};

static int prepare_one(const cnvk_panel* P, const PanelTables& T, const SampleIO& io, int want_cbs, double skip_outliers,
                       cudaStream_t stream) {
    const long long n = P->n;
    const int n_arms = P->n_arms;
    CNVK_CHECK(cudaMemsetAsync(io.stat, 0, 3 * sizeof(unsigned long long), stream));
    CNVK_CHECK(cudaMemsetAsync(io.fflags, 0, sizeof(unsigned int), stream));
    if (want_cbs) CNVK_CHECK(cudaMemsetAsync(io.units, 0, ((size_t)n_arms + 1) * sizeof(long long), stream));
    for (int k = 0; k < 2; ++k) {
        const cnvk_panel_group& G = P->grp[k];
        if (G.n0 <= 0 || G.n <= 0) continue;
        const double* in_l = io.in[2 * k];
        const double* in_d = io.in[2 * k + 1];
        {
            ProfScope ps(PC_FIX, stream);
            nan_any_kernel<<<(int)std::min<long long>(div_up(G.n0, 256), 592), 256, 0, stream>>>(in_l, G.n0, io.stat);
            CNVK_LAUNCH_CHECK();
            gather_rows2_kernel<<<div_up(G.n, 256), 256, 0, stream>>>(in_l, in_d, (const long long*)G.d_keep_idx, G.n, io.g[k][0], io.g[k][1]);
            CNVK_LAUNCH_CHECK();
        }
        const u32* orders[3] = {G.d_order[0], G.d_order[1], G.d_order[2]};
        int rc = fix_group_ordered_enqueue(io.g[k][0], io.g[k][1], G.d_auto, G.n, (const long long*)G.d_chrom_off, G.nchrom,
                                           orders, 3, G.skip_low, 1, G.wing, io.stat + 1 + k, stream);
        if (rc) return rc;
    }
    if (n <= 0) return 0;
    {
        ProfScope ps(PC_FIX, stream);
        merge_order_kernel<<<div_up(n, 256), 256, 0, stream>>>(io.g[0][0], io.g[0][1], io.g[1][0], io.g[1][1], P->grp[0].n,
                                                               (const long long*)P->d_order, n, io.o_l, io.o_d);
        CNVK_LAUNCH_CHECK();
    }
    int rc = fix_finish(io.o_l, io.o_d, P->d_ref_log2, P->d_ref_spread, (const long long*)P->d_start,
                        (const long long*)P->d_end, P->d_chrom_id, P->d_is_anti, P->d_auto, n,
                        (const long long*)P->d_chrom_off, P->nchrom, io.o_w, nullptr, stream);
    if (rc) return rc;
    if (!want_cbs) return 0;
    if (io.mask) {
        rc = outlier_mask_enqueue(io.o_l, n, T.d_arm_off, T.d_tile_off, T.d_ex, T.n_ex, T.tiles, 0.95, skip_outliers,
                                  io.mask, stream);
        if (rc) return rc;
    }
    return cbs_input_filter_enqueue(io.o_l, io.o_w, io.mask, (const long long*)P->d_arm_id, T.d_arm_off, n_arms, n, io.xq,
                                    io.wq, io.src, io.units, io.fflags, stream);
}

// ---- per-thread plan: device copies of the panel's small host tables and the work columns of one sample, kept while
// the same panel is served.  (Replaying the chain of a sample as a captured CUDA graph was tried -- one launch per
// sample instead of ~110 -- and did not move the cohort throughput (811 vs 796 samples/s with 4 host threads,
// profiles/r02aj_panel_graph.log): the cohort is not bound by the launch rate.  Removed again.)
struct PanelPlan {
    cnvk_panel key;
    bool have_key = false;
    int want_cbs = -1;
    double skip_outliers = -1.0;
    std::vector<long long> arm_off;
    PanelTables T;
    SampleIO io;
    void* bufs[16];
    int nbufs = 0;
    void release() {
        for (int i = 0; i < nbufs; ++i) cudaFree(bufs[i]);
        nbufs = 0;
        have_key = false;
    }
    ~PanelPlan() {}             // thread exit: buffers are left to the process (CUDA may already be shutting down)
    template <typename T_>
    cudaError_t buf(T_** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T_));
        if (e != cudaSuccess) return e;
        bufs[nbufs++] = p;
        *out = (T_*)p;
        return cudaSuccess;
    }
};
static thread_local PanelPlan g_plan;

static int plan_build(PanelPlan& L, const cnvk_panel* P, int want_cbs, double skip_outliers) {
    L.release();
    L.key = *P;
    L.arm_off.assign(P->arm_off, P->arm_off + P->n_arms + 1);
    L.have_key = true;
    L.want_cbs = want_cbs;
    L.skip_outliers = skip_outliers;
    const long long n = std::max<long long>(P->n, 1);
    const int n_arms = P->n_arms;
    std::vector<int> ex_arm, tile_off;
    L.T.tiles = outlier_tables(L.arm_off.data(), n_arms, &ex_arm, &tile_off);
    L.T.n_ex = (int)ex_arm.size();
    CNVK_CHECK(L.buf(&L.T.d_arm_off, (size_t)n_arms + 1));
    CNVK_CHECK(L.buf(&L.T.d_tile_off, tile_off.size()));
    CNVK_CHECK(L.buf(&L.T.d_ex, ex_arm.size() + 1));
    CNVK_CHECK(cudaMemcpy(L.T.d_arm_off, L.arm_off.data(), ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    CNVK_CHECK(cudaMemcpy(L.T.d_tile_off, tile_off.data(), tile_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (L.T.n_ex) CNVK_CHECK(cudaMemcpy(L.T.d_ex, ex_arm.data(), ex_arm.size() * sizeof(int), cudaMemcpyHostToDevice));
    SampleIO& io = L.io;
    for (int k = 0; k < 2; ++k) {
        CNVK_CHECK(L.buf(&io.g[k][0], (size_t)std::max<long long>(P->grp[k].n, 1)));
        CNVK_CHECK(L.buf(&io.g[k][1], (size_t)std::max<long long>(P->grp[k].n, 1)));
    }
    io.mask = nullptr;
    if (want_cbs && skip_outliers > 0.0) CNVK_CHECK(L.buf(&io.mask, (size_t)n));
    return 0;
}

int panel_prepare(const cnvk_panel* P, int ns, const double* const* in, double* const* out, int want_cbs,
                  double skip_outliers, double* const* cbs_out, long long* const* src_out, long long* h_units,
                  unsigned int* h_status, cudaStream_t stream) {
    if (ns <= 0) return 0;
    const long long n = P->n;
    const int n_arms = P->n_arms;
    PanelPlan& L = g_plan;
    const bool same = L.have_key && memcmp(&L.key, P, sizeof(cnvk_panel)) == 0 && L.want_cbs == want_cbs &&
                      L.skip_outliers == skip_outliers && (int)L.arm_off.size() == n_arms + 1 &&
                      memcmp(L.arm_off.data(), P->arm_off, ((size_t)n_arms + 1) * sizeof(long long)) == 0;
    if (!same) {
        int rc = plan_build(L, P, want_cbs, skip_outliers);
        if (rc) { L.release(); return rc; }
    }
    Scratch scratch(stream);
    // per sample: [0] NaN seen, [1] / [2] bins above the coverage floor per group; then the filter flags
    unsigned long long* d_stat;
    unsigned int* d_fflags;
    long long* d_units;
    CNVK_CHECK(scratch.alloc(&d_stat, (size_t)ns * 3));
    CNVK_CHECK(scratch.alloc(&d_fflags, (size_t)ns));
    CNVK_CHECK(scratch.alloc(&d_units, (size_t)ns * (n_arms + 1)));
    for (int s = 0; s < ns; ++s) {
        SampleIO io = L.io;
        for (int j = 0; j < 4; ++j) io.in[j] = in[4 * s + j];
        io.o_l = out[3 * s]; io.o_d = out[3 * s + 1]; io.o_w = out[3 * s + 2];
        if (want_cbs) { io.xq = cbs_out[2 * s]; io.wq = cbs_out[2 * s + 1]; io.src = src_out[s]; }
        io.stat = d_stat + 3 * s;
        io.fflags = d_fflags + s;
        io.units = d_units + (size_t)s * (n_arms + 1);
        int rc = prepare_one(P, L.T, io, want_cbs, skip_outliers, stream);
        if (rc) return rc;
    }
    std::vector<unsigned long long> h_stat((size_t)ns * 3);
    std::vector<unsigned int> h_ff((size_t)ns);
    CNVK_CHECK(cudaMemcpyAsync(h_stat.data(), d_stat, h_stat.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaMemcpyAsync(h_ff.data(), d_fflags, h_ff.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    if (want_cbs)
        CNVK_CHECK(cudaMemcpyAsync(h_units, d_units, (size_t)ns * (n_arms + 1) * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
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
