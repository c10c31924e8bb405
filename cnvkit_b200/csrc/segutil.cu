// Per-arm pre-filters and post-processing of cnvlib.segmentation.do_segmentation:
//
//   outlier_mask     segmentation/__init__.py:372-398 drop_outliers -> smoothing.rolling_outlier_quantile
//                    (smoothing.py:439-463): trend = savgol(x, 50) (seven passes of scipy's 7-tap
//                    order-3 Savitzky-Golay FIR over the 25-mirror-padded arm, smoothing.py:236-331),
//                    q = rolling 0.95-quantile (window 51) of |x - trend|, outlier = |x-trend| > q*factor.
//   round_sig        the "%.6g" quantisation applied to every float column of the TSV handed to
//                    DNAcopy (segmentation/__init__.py:299-301), exact round-half-even in decimal.
//   segment_bin_sums transfer_fields' numeric part (segmentation/__init__.py:489-501): per segment
//                    nansum of bin weights and the weighted mean of bin depths.
#include "ops.cuh"
#include "prims.cuh"
#include "prof.cuh"
#include <vector>

namespace cnvk {

typedef unsigned char u8;

static const int OL_WING = 25;      // _width2wing(50, x) for len(x) > 50
static const int OL_PASSES = 7;     // total_width 51 // window_width 7
static const int OL_TILE = 512;
static const int OL_TOP = 8;        // upper-tail quantiles needing at most this many of the largest values take the register path
static const int OL_HALO = 3 * OL_PASSES;   // 21 points each side feed the 7 passes

// scipy.signal.savgol_coeffs(7, 3) as produced by this image's scipy/LAPACK, in the order
// scipy.ndimage.correlate1d's symmetric branch consumes them (centre, then pairs at +-3, +-2, +-1).
__device__ __constant__ double SG_C0 = 0x1.555555555555ap-2;
__device__ __constant__ double SG_C3 = -0x1.861861861861bp-4;
__device__ __constant__ double SG_C2 = 0x1.2492492492495p-3;
__device__ __constant__ double SG_C1 = 0x1.2492492492496p-2;

__device__ __forceinline__ int mirror_pad(int t, int n, int wing) {   // padded index -> arm-local index
    if (t < wing) return wing - 1 - t;
    if (t < wing + n) return t - wing;
    return n - 1 - (t - wing - n);
}

// dists[i] = |x[i] - savgol_trend[i]| for arms longer than 50 bins
__global__ void __launch_bounds__(256)
savgol_dists_kernel(const double* __restrict__ x, const long long* __restrict__ off, const int* __restrict__ tile_off,
                    const int* __restrict__ ex_arm, int n_ex, double* __restrict__ dists) {
    __shared__ double s_a[OL_TILE + 2 * OL_HALO], s_b[OL_TILE + 2 * OL_HALO];
    int a = 0, b = n_ex;
    while (b - a > 1) { const int m = (a + b) >> 1; if (tile_off[m] <= (int)blockIdx.x) a = m; else b = m; }
    const long long lo = off[ex_arm[a]];
    const int n = (int)(off[ex_arm[a] + 1] - lo);
    const int i0 = ((int)blockIdx.x - tile_off[a]) * OL_TILE;
    const int cnt = min(OL_TILE, n - i0);
    const int L = n + 2 * OL_WING;
    // padded positions [p0, p0 + cnt + 2*HALO), p0 = i0 + WING - HALO
    const int p0 = i0 + OL_WING - OL_HALO;
    const int span = cnt + 2 * OL_HALO;
    for (int t = threadIdx.x; t < span; t += 256) {
        const int p = p0 + t;
        s_a[t] = (p >= 0 && p < L) ? x[lo + mirror_pad(p, n, OL_WING)] : 0.0;
    }
    __syncthreads();
    double* in = s_a;
    double* out = s_b;
    for (int pass = 0; pass < OL_PASSES; ++pass) {
        const int margin = 3 * (pass + 1);
        for (int t = threadIdx.x; t < span; t += 256) {
            double v = in[t];
            if (t >= margin && t < span - margin) {
                double acc = __dmul_rn(in[t], SG_C0);
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[t - 3], in[t + 3]), SG_C3));
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[t - 2], in[t + 2]), SG_C2));
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(in[t - 1], in[t + 1]), SG_C1));
                v = acc;
            }
            out[t] = v;
        }
        __syncthreads();
        double* tmp = in; in = out; out = tmp;
    }
    for (int t = threadIdx.x; t < cnt; t += 256) {
        const int i = i0 + t;
        dists[lo + i] = fabs(__dsub_rn(x[lo + i], in[t + OL_HALO]));
    }
}

// the M-th largest (*vlow) and the (M-1)-th largest (*vhigh) of v[0, W)
template <int M>
__device__ __forceinline__ void top_m(const double* v_, int W, double* vlow, double* vhigh) {
    double top[M];
#pragma unroll
    for (int r = 0; r < M; ++r) top[r] = -INFINITY;
    for (int j = 0; j < W; ++j) {
        double v = v_[j];
        if (v > top[M - 1]) {
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const double hi_ = fmax(top[r], v), lo_ = fmin(top[r], v);
                top[r] = hi_;
                v = lo_;
            }
        }
    }
    *vlow = top[M - 1];
    *vhigh = (M >= 2) ? top[M >= 2 ? M - 2 : 0] : 0.0;
}

// mask[i] = dists[i] > quantile_q(window 2*wing+1 of mirror-padded dists) * factor
__global__ void __launch_bounds__(256)
quantile_mask_kernel(const double* __restrict__ dists, const long long* __restrict__ off,
                     const int* __restrict__ tile_off, const int* __restrict__ ex_arm, int n_ex, double q,
                     double factor, u8* __restrict__ mask) {
    __shared__ double s_v[OL_TILE + 2 * OL_WING];
    int a = 0, b = n_ex;
    while (b - a > 1) { const int m = (a + b) >> 1; if (tile_off[m] <= (int)blockIdx.x) a = m; else b = m; }
    const long long lo = off[ex_arm[a]];
    const int n = (int)(off[ex_arm[a] + 1] - lo);
    const int i0 = ((int)blockIdx.x - tile_off[a]) * OL_TILE;
    const int cnt = min(OL_TILE, n - i0);
    const int W = 2 * OL_WING + 1;
    for (int t = threadIdx.x; t < cnt + 2 * OL_WING; t += 256)
        s_v[t] = dists[lo + mirror_pad(i0 + t, n, OL_WING)];
    __syncthreads();
    const double idxf = q * (double)(W - 1);
    const int idx = (int)idxf;
    const double frac = idxf - (double)idx;
    const int m = W - idx;                    // the order statistic `idx` is the m-th largest of the window
    for (int t = threadIdx.x; t < cnt; t += 256) {
        double vlow = 0.0, vhigh = 0.0;
        if (m <= OL_TOP) {
            // upper-tail quantile (0.95 of 51 values: the 4th and 3rd largest): keep the m largest in registers,
            // W * m compare-exchanges instead of W * W rank comparisons
            switch (m) {
                case 1: top_m<1>(s_v + t, W, &vlow, &vhigh); break;
                case 2: top_m<2>(s_v + t, W, &vlow, &vhigh); break;
                case 3: top_m<3>(s_v + t, W, &vlow, &vhigh); break;
                case 4: top_m<4>(s_v + t, W, &vlow, &vhigh); break;
                case 5: top_m<5>(s_v + t, W, &vlow, &vhigh); break;
                case 6: top_m<6>(s_v + t, W, &vlow, &vhigh); break;
                case 7: top_m<7>(s_v + t, W, &vlow, &vhigh); break;
                default: top_m<8>(s_v + t, W, &vlow, &vhigh); break;
            }
        } else {
            for (int j = 0; j < W; ++j) {
                const double v = s_v[t + j];
                int rank = 0;
                for (int k = 0; k < W; ++k) {
                    const double u = s_v[t + k];
                    rank += (u < v || (u == v && k < j)) ? 1 : 0;
                }
                if (rank == idx) vlow = v;
                if (rank == idx + 1) vhigh = v;
            }
        }
        const double quant = (frac == 0.0) ? vlow : __dadd_rn(vlow, __dmul_rn(__dsub_rn(vhigh, vlow), frac));
        mask[lo + i0 + t] = (dists[lo + i0 + t] > __dmul_rn(quant, factor)) ? 1 : 0;
    }
}

// tables of the arms that are examined (more than 50 bins, smoothing.py:458-459): tile -> arm lookup
int outlier_tables(const long long* h_off, int n_arms, std::vector<int>* ex_arm, std::vector<int>* tile_off) {
    int tiles = 0;
    ex_arm->clear();
    tile_off->clear();
    for (int a = 0; a < n_arms; ++a) {
        const long long n = h_off[a + 1] - h_off[a];
        if (n > 50) {
            ex_arm->push_back(a);
            tile_off->push_back(tiles);
            tiles += div_up(n, OL_TILE);
        }
    }
    tile_off->push_back(tiles);
    return tiles;
}

// device part with the tables already resident (no host memory involved: can be captured into a CUDA graph)
int outlier_mask_enqueue(const double* d_x, long long N, const long long* d_off, const int* d_tile_off, const int* d_ex,
                         int n_ex, int tiles, double q, double factor, u8* d_mask, cudaStream_t stream) {
    if (N <= 0) return 0;
    ProfScope ps(PC_OTHER, stream);
    CNVK_CHECK(cudaMemsetAsync(d_mask, 0, (size_t)N, stream));
    if (n_ex == 0) return 0;
    Scratch scratch(stream);
    double* d_dists;
    CNVK_CHECK(scratch.alloc(&d_dists, (size_t)N));
    savgol_dists_kernel<<<tiles, 256, 0, stream>>>(d_x, d_off, d_tile_off, d_ex, n_ex, d_dists);
    CNVK_LAUNCH_CHECK();
    quantile_mask_kernel<<<tiles, 256, 0, stream>>>(d_dists, d_off, d_tile_off, d_ex, n_ex, q, factor, d_mask);
    CNVK_LAUNCH_CHECK();
    return 0;
}

int outlier_mask(const double* d_x, const long long* h_off, int n_arms, double q, double factor, u8* d_mask,
                 cudaStream_t stream) {
    if (n_arms <= 0) return 0;
    const long long N = h_off[n_arms];
    if (N <= 0) return 0;
    std::vector<int> ex_arm, tile_off;
    const int tiles = outlier_tables(h_off, n_arms, &ex_arm, &tile_off);
    const int n_ex = (int)ex_arm.size();
    Scratch scratch(stream);
    long long* d_off;
    int *d_tile_off, *d_ex;
    CNVK_CHECK(scratch.alloc(&d_off, (size_t)n_arms + 1));
    CNVK_CHECK(scratch.alloc(&d_tile_off, tile_off.size()));
    CNVK_CHECK(scratch.alloc(&d_ex, ex_arm.size() + 1));
    // an asynchronous copy FROM pageable host memory returns once the source has been staged, so the host vectors
    // may go out of scope without a synchronisation
    CNVK_CHECK(cudaMemcpyAsync(d_off, h_off, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemcpyAsync(d_tile_off, tile_off.data(), tile_off.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (n_ex) CNVK_CHECK(cudaMemcpyAsync(d_ex, ex_arm.data(), ex_arm.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    return outlier_mask_enqueue(d_x, N, d_off, d_tile_off, d_ex, n_ex, tiles, q, factor, d_mask, stream);
}

// ------------------------------------------------------------------ "%.{digits}g" quantisation
__device__ __forceinline__ double pow10_exact(int k) {   // 10^k, exact for 0 <= k <= 22
    const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                          1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
}

// nearest integer to the exact value hi+lo (|lo| <= ulp(hi)/2), ties to even
__device__ __forceinline__ double round_exact_sum(double hi, double lo) {
    double f = floor(hi);
    const double d = (hi - f) - 0.5;          // exact
    if (d > 0.0 || (d == 0.0 && lo > 0.0)) return f + 1.0;
    if (d < 0.0 || (d == 0.0 && lo < 0.0)) return f;
    // exact tie
    return (fmod(f, 2.0) == 0.0) ? f : f + 1.0;
}

// float("%.{digits}g" % v): returns false when the value is outside the handled range (|v| < 1e-16 or > 1e21)
__device__ __forceinline__ bool round_sig_value(double v, int digits, double* out) {
    if (v == 0.0 || !(fabs(v) < INFINITY)) { *out = v; return true; }   // 0, inf, NaN print as themselves
    const double a = fabs(v);
    int e = (int)floor(log10(a));
    double m = 0.0;
    bool ok = false;
    for (int attempt = 0; attempt < 3 && !ok; ++attempt) {
        const int k = digits - 1 - e;          // scale by 10^k
        double hi, lo;
        if (k >= 0 && k <= 22) {
            const double s = pow10_exact(k);
            hi = a * s;
            lo = fma(a, s, -hi);
        } else if (k < 0 && k >= -22) {
            const double s = pow10_exact(-k);
            hi = a / s;
            lo = -fma(hi, s, -a) / s;           // residual of the division (sign/magnitude only matter at ties)
        } else {
            break;
        }
        m = round_exact_sum(hi, lo);
        const double top = pow10_exact(digits), bot = pow10_exact(digits - 1);
        if (m >= top) { if (m == top) { m = bot; ++e; ok = true; } else { ++e; } }
        else if (m < bot) { --e; }
        else ok = true;
    }
    if (!ok) { *out = v; return false; }
    // value of the decimal m * 10^(e-digits+1), correctly rounded (Clinger's fast path)
    const int p = e - digits + 1;
    double r;
    if (p >= 0 && p <= 22) r = m * pow10_exact(p);
    else if (p < 0 && p >= -22) r = m / pow10_exact(-p);
    else { *out = v; return false; }
    *out = copysign(r, v);
    return true;
}

__global__ void round_sig_kernel(const double* __restrict__ x, long long n, int digits, double* __restrict__ out,
                                 unsigned long long* __restrict__ unhandled) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    if (!round_sig_value(x[i], digits, &r)) atomicAdd(unhandled, 1ull);
    out[i] = r;
}

int round_sig(const double* d_x, long long n, int digits, double* d_out, unsigned long long* h_unhandled,
              cudaStream_t stream) {
    if (h_unhandled) *h_unhandled = 0;
    if (n <= 0) return 0;
    if (digits < 1 || digits > 15) { set_last_error("round_sig: digits=%d unsupported", digits); return -2; }
    ProfScope ps(PC_OTHER, stream);
    Scratch scratch(stream);
    unsigned long long* d_cnt;
    CNVK_CHECK(scratch.alloc(&d_cnt, 1));
    CNVK_CHECK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), stream));
    round_sig_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_x, n, digits, d_out, d_cnt);
    CNVK_LAUNCH_CHECK();
    if (h_unhandled) {
        CNVK_CHECK(cudaMemcpyAsync(h_unhandled, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CNVK_CHECK(cudaStreamSynchronize(stream));
    }
    return 0;
}

// ------------------------------------------------------------------ CBS input filter
// Everything between a .cnr table and DNAcopy's input, fused (cnvlib/segmentation/__init__.py:239-282 per-arm
// filters + :299-301 "%.6g" TSV + the R row filters cbs.py:22-35):
//   keep a bin unless it is an outlier (mask from cnvk_outlier_mask_dev) or its weight is 0 (NaN weights survive
//   this gate, as in the reference); quantise log2 and weight to 6 significant digits; per ARM, if no bin has a
//   non-NA weight every weight becomes 1.0, else bins with NA or <= 0 weight are dropped; bins with NA log2 are
//   dropped; the survivors are compacted in order.  Outputs: quantised columns, the source row of every survivor,
//   the survivors' arm offsets (n_arms + 1) and a flag word (bit 0: a log2 is not finite -- the caller takes the
//   general path; bit 1: a value outside the range of the device quantiser).
__global__ void __launch_bounds__(256)
cbs_filter_mark_kernel(const double* __restrict__ log2v, const double* __restrict__ weight, const unsigned char* __restrict__ outlier,
                       const long long* __restrict__ arm_id, long long n, double* __restrict__ xq, double* __restrict__ wq,
                       unsigned char* __restrict__ keep1, int* __restrict__ arm_has_w, unsigned int* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = log2v[i], w = weight ? weight[i] : 1.0;
    unsigned f = 0;
    if (!(fabs(x) < INFINITY)) f |= 1u;
    const bool k1 = !(outlier && outlier[i]) && !(w == 0.0);
    double rx, rw;
    if (!round_sig_value(x, 6, &rx)) f |= 2u;
    if (!round_sig_value(w, 6, &rw)) f |= 2u;
    xq[i] = rx;
    wq[i] = rw;
    keep1[i] = k1 ? 1 : 0;
    if (k1 && rw == rw) arm_has_w[arm_id[i]] = 1;          // benign race: every writer stores 1
    if (f) atomicOr(flags, f);
}

__global__ void __launch_bounds__(256)
cbs_filter_ok_kernel(const double* __restrict__ xq, const double* __restrict__ wq, const unsigned char* __restrict__ keep1,
                     const long long* __restrict__ arm_id, const int* __restrict__ arm_has_w, long long n, int have_weight,
                     u32* __restrict__ ok) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { ok[n] = 0; return; }
    bool o = keep1[i] != 0;
    if (o && have_weight && arm_has_w[arm_id[i]]) o = (wq[i] == wq[i]) && wq[i] > 0.0;   // cbs.py:22-24
    if (o) o = xq[i] == xq[i];                                                          // cbs.py:30-33
    ok[i] = o ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
cbs_filter_scatter_kernel(const double* __restrict__ xq, const double* __restrict__ wq, const u32* __restrict__ ok,
                          const u32* __restrict__ pos, const long long* __restrict__ arm_id, const int* __restrict__ arm_has_w,
                          long long n, int have_weight, double* __restrict__ x_out, double* __restrict__ w_out,
                          long long* __restrict__ src_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !ok[i]) return;
    const u32 p = pos[i];
    x_out[p] = xq[i];
    w_out[p] = (have_weight && arm_has_w[arm_id[i]]) ? wq[i] : 1.0;                       // cbs.py:25-28
    src_out[p] = i;
}

__global__ void cbs_filter_units_kernel(const u32* __restrict__ pos, const long long* __restrict__ arm_off, int n_arms,
                                        long long* __restrict__ units) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a <= n_arms) units[a] = (long long)pos[arm_off[a]];     // pos has n + 1 entries: pos[n] = number of survivors
}

// device part: survivors, source rows, *d_units_out[n_arms + 1] and *d_flags_out stay on the device (no synchronisation)
int cbs_input_filter_enqueue(const double* d_log2, const double* d_weight, const unsigned char* d_outlier,
                             const long long* d_arm_id, const long long* d_arm_off, int n_arms, long long n,
                             double* d_x_out, double* d_w_out, long long* d_src_out, long long* d_units_out,
                             unsigned int* d_flags_out, cudaStream_t stream) {
    if (n >= (1ll << 31)) { set_last_error("cbs_input_filter: too many bins"); return -2; }
    ProfScope ps(PC_OTHER, stream);
    Scratch scratch(stream);
    double *xq, *wq;
    unsigned char* keep1;
    int* arm_has_w;
    u32 *ok, *pos;
    CNVK_CHECK(scratch.alloc(&xq, (size_t)n));
    CNVK_CHECK(scratch.alloc(&wq, (size_t)n));
    CNVK_CHECK(scratch.alloc(&keep1, (size_t)n));
    CNVK_CHECK(scratch.alloc(&arm_has_w, (size_t)n_arms + 2));
    CNVK_CHECK(scratch.alloc(&ok, (size_t)n + 1));
    CNVK_CHECK(scratch.alloc(&pos, (size_t)n + 1));
    CNVK_CHECK(cudaMemsetAsync(arm_has_w, 0, ((size_t)n_arms + 2) * sizeof(int), stream));
    CNVK_CHECK(cudaMemsetAsync(d_flags_out, 0, sizeof(unsigned int), stream));
    const int have_weight = d_weight ? 1 : 0;
    cbs_filter_mark_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_log2, d_weight, d_outlier, d_arm_id, n, xq, wq, keep1, arm_has_w, d_flags_out);
    CNVK_LAUNCH_CHECK();
    cbs_filter_ok_kernel<<<div_up(n + 1, 256), 256, 0, stream>>>(xq, wq, keep1, d_arm_id, arm_has_w, n, have_weight, ok);
    CNVK_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(ok, pos, n + 1, nullptr, stream);
    if (rc) return rc;
    cbs_filter_scatter_kernel<<<div_up(n, 256), 256, 0, stream>>>(xq, wq, ok, pos, d_arm_id, arm_has_w, n, have_weight, d_x_out, d_w_out, d_src_out);
    CNVK_LAUNCH_CHECK();
    cbs_filter_units_kernel<<<div_up(n_arms + 1, 128), 128, 0, stream>>>(pos, d_arm_off, n_arms, d_units_out);
    CNVK_LAUNCH_CHECK();
    return 0;
}

int cbs_input_filter(const double* d_log2, const double* d_weight, const unsigned char* d_outlier, const long long* d_arm_id,
                     const long long* h_arm_off, int n_arms, long long n, double* d_x_out, double* d_w_out,
                     long long* d_src_out, long long* h_units, unsigned int* h_flags, cudaStream_t stream) {
    *h_flags = 0;
    for (int a = 0; a <= n_arms; ++a) h_units[a] = 0;
    if (n <= 0 || n_arms <= 0) return 0;
    Scratch scratch(stream);
    unsigned int* flags;
    long long *d_off, *d_units;
    CNVK_CHECK(scratch.alloc(&flags, 1));
    CNVK_CHECK(scratch.alloc(&d_off, (size_t)n_arms + 1));
    CNVK_CHECK(scratch.alloc(&d_units, (size_t)n_arms + 1));
    CNVK_CHECK(cudaMemcpyAsync(d_off, h_arm_off, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    int rc = cbs_input_filter_enqueue(d_log2, d_weight, d_outlier, d_arm_id, d_off, n_arms, n, d_x_out, d_w_out, d_src_out,
                                      d_units, flags, stream);
    if (rc) return rc;
    CNVK_CHECK(cudaMemcpyAsync(h_units, d_units, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaMemcpyAsync(h_flags, flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

// ------------------------------------------------------------------ transfer_fields sums
// one CTA per segment over its bin range [lo, hi): weight = nansum(w); depth = sum(d*w)/sum(w)
// over non-NaN weights, or 0 when the segment carries no weight.  (A segment is up to a whole chromosome arm: 256
// threads with four independent loads in flight each, instead of one warp walking 10^4 bins.)
__global__ void __launch_bounds__(256)
segment_bin_sums_kernel(const double* __restrict__ depth, const double* __restrict__ weight,
                        const long long* __restrict__ lo, const long long* __restrict__ hi,
                        const long long* __restrict__ bin_end, const long long* __restrict__ seg_start, int nseg,
                        double* __restrict__ seg_weight, double* __restrict__ seg_depth) {
    __shared__ double s_w[8], s_dw[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    double sw = 0.0, sdw = 0.0;
    // nested bins (intersect.py _irange_nested, mode "outer"): rows [lo, hi) whose end lies beyond the
    // segment start; without the two optional columns every row of the range counts (_irange_simple)
    const long long s0 = bin_end ? seg_start[s] : 0;
    const long long k0 = lo[s], k1 = hi[s];
    for (long long kb = k0 + threadIdx.x; kb < k1; kb += 4 * 256) {
        double w[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long k = kb + u * 256;
            const bool ok = k < k1 && !(bin_end && !(bin_end[k] > s0));
            w[u] = ok ? (weight ? weight[k] : 1.0) : __longlong_as_double(0x7ff8000000000000ll);
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
        seg_weight[s] = sw;
        seg_depth[s] = (sw > 0.0) ? sdw / sw : 0.0;
    }
}

int segment_bin_sums(const double* d_depth, const double* d_weight, const long long* d_lo, const long long* d_hi,
                     const long long* d_bin_end, const long long* d_seg_start, int nseg, double* d_seg_weight,
                     double* d_seg_depth, cudaStream_t stream) {
    if (nseg <= 0) return 0;
    ProfScope ps(PC_OTHER, stream);
    segment_bin_sums_kernel<<<nseg, 256, 0, stream>>>(d_depth, d_weight, d_lo, d_hi, d_bin_end, d_seg_start, nseg, d_seg_weight, d_seg_depth);
    CNVK_LAUNCH_CHECK();
    return 0;
}

}  // namespace cnvk
