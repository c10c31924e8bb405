// Tukey biweight location / midvariance (cnvlib/descriptives.py:96-142, 188-222).
//
// Both estimators need median(|a - c|) for a sequence of centres c.  With the values sorted once,
// the m elements nearest to c form a contiguous window found by binary search, so every MAD is
// O(log n) and exact (it is the rounded |a_i - c| of an input element, as numpy computes it);
// the weighted sums are plain parallel reductions (tolerance 1e-12).
//
//  * biweight_sorted_kernel : one cluster of 8 CTAs, one long array (apply_weights: residuals of ~10^5 bins)
//  * biweight_columns_kernel: one warp per bin across S samples (reference.summarize_info,
//    cnvlib/reference.py:717-741), columns staged through shared memory and sorted per warp.
#include <cooperative_groups.h>

#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

#define NANQ __longlong_as_double(0x7ff8000000000000ll)

// distance of the m-th nearest element of sorted a[0,n) to c  (1 <= m <= n)
template <typename Acc>
__device__ __forceinline__ double mth_nearest(const Acc& a, int n, double c, int m) {
    int lo = 0, hi = n - m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c - a(mid) > a(mid + m) - c) lo = mid + 1; else hi = mid;
    }
    return fmax(fabs(a(lo) - c), fabs(a(lo + m - 1) - c));
}

template <typename Acc>
__device__ __forceinline__ double median_abs_dev(const Acc& a, int n, double c) {
    if (n & 1) return mth_nearest(a, n, c, n / 2 + 1);
    return (mth_nearest(a, n, c, n / 2) + mth_nearest(a, n, c, n / 2 + 1)) / 2.0;
}

template <typename Acc>
__device__ __forceinline__ double median_sorted(const Acc& a, int n) {
    return (n & 1) ? a(n / 2) : (a(n / 2 - 1) + a(n / 2)) / 2.0;
}

struct GlobalAcc {
    const double* p;
    __device__ __forceinline__ double operator()(int i) const { return p[i]; }
};

// ------------------------------------------------------------------ one thread-block cluster, one long array
// apply_weights' residuals: ~10^5 values, up to six passes (five location iterations + the midvariance), each a
// MAD (a search in the sorted array) followed by weighted sums over all values.  One CTA spent 160 us on it (130
// serial loads per thread and pass, 17-step binary searches of dependent loads by one thread); here a cluster of
// BW_CLUSTER CTAs shares each pass -- partial sums meet through distributed shared memory in rank order, so the
// result does not depend on scheduling -- and the MAD search is 32-ary (one probe per lane, 4 steps).
static const int BW_CLUSTER = 8;
static const int BW_THREADS = 512;

// warp-cooperative mth_nearest: first lo in [0, n-m] with !(c - a[lo] > a[lo+m] - c); all lanes get the distance
__device__ __forceinline__ double mth_nearest_warp(const double* __restrict__ a, int n, double c, int m) {
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = n - m;
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int probe = lo + (lane + 1) * step - 1;
        bool p = false;
        if (probe < hi) p = (c - __ldg(a + probe)) > (__ldg(a + probe + m) - c);
        const int t = __popc(__ballot_sync(0xffffffffu, p));     // the predicate is monotone: a prefix of the lanes
        const int nhi = min(lo + (t + 1) * step - 1, hi);
        lo += t * step;
        hi = nhi;
    }
    return fmax(fabs(__ldg(a + lo) - c), fabs(__ldg(a + lo + m - 1) - c));
}

__device__ __forceinline__ double median_abs_dev_warp(const double* __restrict__ a, int n, double c) {
    if (n & 1) return mth_nearest_warp(a, n, c, n / 2 + 1);
    return (mth_nearest_warp(a, n, c, n / 2) + mth_nearest_warp(a, n, c, n / 2 + 1)) / 2.0;
}

// sums of NV per-thread values over the whole cluster, identical in every thread of every CTA
template <int NV>
__device__ __forceinline__ void cluster_sum(double (&v)[NV], double* s_warp, double* s_cta) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double t = warp_sum(v[k]);
        if (lane == 0) s_warp[warp * NV + k] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int w = 0; w < BW_THREADS / 32; ++w) t += s_warp[w * NV + threadIdx.x];
        s_cta[threadIdx.x] = t;
    }
    cluster.sync();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (unsigned r = 0; r < cluster.num_blocks(); ++r) t += cluster.map_shared_rank(s_cta, r)[k];
        v[k] = t;
    }
    cluster.sync();                                              // s_cta may be rewritten by the next pass
}

// out[0] = biweight_location(a) (or the given initial), out[1] = biweight_midvariance(a, initial=out[0])
__global__ void __launch_bounds__(BW_THREADS)
biweight_sorted_kernel(const double* __restrict__ a, const u32* __restrict__ nvalid_p, int n_given,
                       const double* __restrict__ initial_p, double* __restrict__ out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_warp[(BW_THREADS / 32) * 4];
    __shared__ double s_cta[4];
    __shared__ double s_b;
    const int n = nvalid_p ? (int)*nvalid_p : n_given;
    const bool writer = cluster.block_rank() == 0 && threadIdx.x == 0;
    if (n == 0) { if (writer) { out[0] = NANQ; out[1] = NANQ; } return; }
    if (n == 1) { if (writer) { out[0] = a[0]; out[1] = 0.0; } return; }
    const int tid = (int)cluster.block_rank() * BW_THREADS + threadIdx.x;
    const int nthr = (int)cluster.num_blocks() * BW_THREADS;
    GlobalAcc A{a};
    double loc;
    if (initial_p) {
        loc = *initial_p;
    } else {
        double initial = median_sorted(A, n);
        double result = initial;
        for (int it = 0; it < 5; ++it) {
            if (threadIdx.x < 32) {                              // every CTA finds the same MAD on its own
                const double m_ = median_abs_dev_warp(a, n, initial);
                if (threadIdx.x == 0) s_b = m_;
            }
            __syncthreads();
            const double mad = s_b;
            double scale = 0.0;
            if (mad == 0.0) {
                result = initial;
            } else {
                scale = 6.0 * mad;
                double acc[2] = {0.0, 0.0};
                for (int i = tid; i < n; i += nthr) {
                    const double d = a[i] - initial;
                    const double u = d / scale;
                    const double t = 1.0 - u * u;
                    const double w = t * t;
                    if (w < 1.0) { acc[0] += w; acc[1] += d * w; }
                }
                cluster_sum<2>(acc, s_warp, s_cta);
                result = (acc[0] == 0.0) ? initial : initial + acc[1] / acc[0];
            }
            __syncthreads();
            if (fabs(result - initial) <= 1e-3 * scale) break;   // uniform: every thread holds the same numbers
            initial = result;
        }
        loc = result;
    }
    // midvariance about loc
    if (threadIdx.x < 32) {
        const double m_ = median_abs_dev_warp(a, n, loc);
        if (threadIdx.x == 0) s_b = m_;
    }
    __syncthreads();
    const double mad = s_b;
    const double denom = fmax(9.0 * mad, 1e-3);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};                        // sum w, count, numerator, denominator
    for (int i = tid; i < n; i += nthr) {
        const double d = a[i] - loc;
        const double w = d / denom;
        if (fabs(w) < 1.0) {
            acc[0] += w;
            acc[1] += 1.0;
            const double w2 = w * w;
            const double om = 1.0 - w2;
            const double om2 = om * om;
            acc[2] += (d * d) * (om2 * om2);
            acc[3] += om * (1.0 - 5.0 * w2);
        }
    }
    cluster_sum<4>(acc, s_warp, s_cta);
    if (writer) {
        out[0] = loc;
        out[1] = (acc[0] == 0.0) ? mad * 1.4826 : sqrt((acc[1] * acc[2]) / (acc[3] * acc[3]));
    }
}

int biweight_of_sorted(const double* d_sorted, const u32* d_nvalid, int n, const double* d_initial, double* d_out,
                       cudaStream_t stream) {
    ProfScope ps(PC_BIWEIGHT, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(BW_CLUSTER);
    cfg.blockDim = dim3(BW_THREADS);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = BW_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CNVK_CHECK(cudaLaunchKernelEx(&cfg, biweight_sorted_kernel, d_sorted, d_nvalid, n, d_initial, d_out));
    CNVK_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ one warp per column
struct SmemAcc {
    const double* p;
    __device__ __forceinline__ double operator()(int i) const { return p[i]; }
};

// sort m (power of two, <= 1024) doubles in shared memory with one warp (bitonic)
__device__ void warp_bitonic_sort(double* v, int m) {
    const int lane = threadIdx.x & 31;
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < m; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double x = v[i], y = v[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { v[i] = y; v[ixj] = x; }
                }
            }
            __syncwarp();
        }
    }
}

__device__ double warp_biweight_location(const double* v, int n, double* scale_out) {
    SmemAcc A{v};
    const int lane = threadIdx.x & 31;
    double initial = median_sorted(A, n);
    double result = initial, scale = 0.0;
    for (int it = 0; it < 5; ++it) {
        const double mad = median_abs_dev(A, n, initial);  // every lane computes the same value
        scale = 0.0;
        if (mad == 0.0) {
            result = initial;
        } else {
            scale = 6.0 * mad;
            double ws = 0.0, dws = 0.0;
            for (int i = lane; i < n; i += 32) {
                const double d = v[i] - initial;
                const double u = d / scale;
                const double t = 1.0 - u * u;
                const double w = t * t;
                if (w < 1.0) { ws += w; dws += d * w; }
            }
            ws = warp_sum(ws);
            dws = warp_sum(dws);
            result = (ws == 0.0) ? initial : initial + dws / ws;
        }
        if (fabs(result - initial) <= 1e-3 * scale) break;
        initial = result;
    }
    *scale_out = scale;
    return result;
}

__device__ double warp_biweight_midvariance(const double* v, int n, double loc) {
    SmemAcc A{v};
    const int lane = threadIdx.x & 31;
    const double mad = median_abs_dev(A, n, loc);
    const double denom = fmax(9.0 * mad, 1e-3);
    double sw = 0.0, cnt = 0.0, num = 0.0, den = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double d = v[i] - loc;
        const double w = d / denom;
        if (fabs(w) < 1.0) {
            sw += w;
            cnt += 1.0;
            const double w2 = w * w, om = 1.0 - w2, om2 = om * om;
            num += (d * d) * (om2 * om2);
            den += om * (1.0 - 5.0 * w2);
        }
    }
    sw = warp_sum(sw); cnt = warp_sum(cnt); num = warp_sum(num); den = warp_sum(den);
    return (sw == 0.0) ? mad * 1.4826 : sqrt((cnt * num) / (den * den));
}

// mat is [rows, ncols] row-major (rows = samples).  One warp per column; 4 warps per CTA.
// mode bit0: write location, bit1: write midvariance (about the location just computed)
__global__ void __launch_bounds__(128)
biweight_columns_kernel(const double* __restrict__ mat, int rows, long long ncols, int padded, int mode,
                        double* __restrict__ loc_out, double* __restrict__ var_out) {
    extern __shared__ double s_buf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long col0 = (long long)blockIdx.x * 4;
    // coalesced staging: consecutive threads read consecutive columns of a row
    for (int idx = threadIdx.x; idx < rows * 4; idx += 128) {
        const int r = idx >> 2, c = idx & 3;
        const long long col = col0 + c;
        s_buf[c * padded + r] = (col < ncols) ? mat[(size_t)r * ncols + col] : NANQ;
    }
    for (int idx = threadIdx.x; idx < (padded - rows) * 4; idx += 128) {
        const int r = rows + (idx >> 2), c = idx & 3;
        s_buf[c * padded + r] = NANQ;
    }
    __syncthreads();
    const long long col = col0 + warp;
    if (col >= ncols) return;
    double* v = s_buf + warp * padded;
    // NaN -> +inf-like key so they sort last: replace by a sentinel and count
    int nan_cnt = 0;
    for (int i = lane; i < padded; i += 32) {
        if (v[i] != v[i]) { v[i] = INFINITY; if (i < rows) ++nan_cnt; }
    }
    nan_cnt = (int)warp_sum_u32((u32)nan_cnt);
    __syncwarp();
    warp_bitonic_sort(v, padded);
    const int n = rows - nan_cnt;
    double loc, var;
    if (n == 0) { loc = NANQ; var = NANQ; }
    else if (n == 1) { loc = v[0]; var = 0.0; }
    else {
        double scale;
        loc = warp_biweight_location(v, n, &scale);
        var = (mode & 2) ? warp_biweight_midvariance(v, n, loc) : 0.0;
    }
    if (lane == 0) {
        if (mode & 1) loc_out[col] = loc;
        if (mode & 2) var_out[col] = var;
    }
}

int biweight_columns(const double* d_mat, int rows, long long ncols, int mode, double* d_loc, double* d_var,
                     cudaStream_t stream) {
    if (ncols <= 0) return 0;
    if (rows < 1 || rows > 4096) { set_last_error("biweight_columns: rows=%d unsupported (1..4096)", rows); return -2; }
    int padded = 1;
    while (padded < rows) padded <<= 1;
    const size_t smem = (size_t)padded * 4 * sizeof(double);
    ProfScope ps(PC_BIWEIGHT, stream);
    if (smem > 48 * 1024)
        CNVK_CHECK(cudaFuncSetAttribute(biweight_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    biweight_columns_kernel<<<div_up(ncols, 4), 128, smem, stream>>>(d_mat, rows, ncols, padded, mode, d_loc, d_var);
    CNVK_LAUNCH_CHECK();
    return 0;
}

}  // namespace cnvk
