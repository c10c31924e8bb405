// HaarSeg on B200 -- restates cnvlib/segmentation/haar.py:257-371 (haarSeg) and its kernels
// HaarConv (:460-558), FindLocalPeaks (:561-612), FDRThres (:374-402), UnifyLevels (:615-663),
// SegmentByPeaks (:405-451), batched over every arm of every sample in one call.
//
// Bit-exactness (north star: Haar breakpoints bit-exact) pins the arithmetic order:
//   * the unweighted convolution is np.cumsum of (x[hi]+x[lo])-2x[k-1] -- a strictly sequential
//     FP64 add chain -- divided by sqrt(2h); the weighted one keeps four sequential running sums.
//     One CTA per (arm, level): all threads stage a chunk of increments in shared memory, one
//     elected thread runs the chain(s), all threads finish (divide/scale/store).  Parallelism
//     comes from arms x 6 chains (x samples), never from inside a chain.
//   * the initial window sums of the weighted convolution replay numpy's pairwise summation.
//   * peaks are found per element (the reference's plateau state machine is stateless once
//     restated), compacted in order by a device-wide scan; the FDR threshold needs the peaks of
//     each (arm, level) ranked by |value|: one global radix sort + a stable pass by segment id.
//   * level unification is a per-arm sorted-list merge with the reference's distance rule.
#include <vector>

#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

static const int HAAR_LEVELS = 5;
static const int HC_CHUNK = 1024;
typedef unsigned char u8;

__device__ __forceinline__ int mirror_hi(int i, int n) { return i >= n ? n - 1 - (i - n) : i; }
__device__ __forceinline__ int mirror_lo(int i) { return i < 0 ? -i - 1 : i; }

// numpy's float64 add.reduce over a contiguous array of m <= 128 elements (pairwise_sum base case)
__device__ double np_sum_small(const double* a, int m) {
    if (m < 8) {
        double r = 0.0;
        for (int i = 0; i < m; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < m - (m % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < m; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// grid (n_arms, 6): chain 0 = unweighted h=1 (noise estimate), chains 1..5 = levels 1..5
__global__ void __launch_bounds__(128)
haar_conv_kernel(const double* __restrict__ x, const double* __restrict__ w, const long long* __restrict__ off,
                 long long N, double* __restrict__ conv) {
    __shared__ double s_a[HC_CHUNK], s_b[HC_CHUNK], s_c[HC_CHUNK], s_d[HC_CHUNK];
    __shared__ double s_init[2];
    const int chain = blockIdx.y, arm = blockIdx.x;
    const long long lo = off[arm];
    const int n = (int)(off[arm + 1] - lo);
    if (n <= 0) return;
    const double* xs = x + lo;
    double* out = conv + (size_t)chain * N + lo;
    const int h = chain == 0 ? 1 : (1 << chain);
    const int tid = threadIdx.x;
    if (h > n) {  // haar.py:479-488: scale too coarse for this arm -> zeros
        for (int k = tid; k < n; k += 128) out[k] = 0.0;
        return;
    }
    const bool weighted = (chain > 0) && (w != nullptr);
    if (!weighted) {
        const double norm = sqrt(2.0 * (double)h);
        double run = 0.0;
        if (tid == 0) out[0] = 0.0;
        for (int c0 = 1; c0 < n; c0 += HC_CHUNK) {
            const int m = min(HC_CHUNK, n - c0);
            for (int t = tid; t < m; t += 128) {
                const int k = c0 + t;
                const int hi = mirror_hi(k + h - 1, n), lw = mirror_lo(k - h - 1);
                s_a[t] = __dsub_rn(__dadd_rn(xs[hi], xs[lw]), __dmul_rn(2.0, xs[k - 1]));
            }
            __syncthreads();
            if (tid == 0) {
#pragma unroll 8
                for (int t = 0; t < m; ++t) {
                    run = __dadd_rn(run, s_a[t]);
                    s_a[t] = run;
                }
            }
            __syncthreads();
            for (int t = tid; t < m; t += 128) out[c0 + t] = __ddiv_rn(s_a[t], norm);
            __syncthreads();
        }
        return;
    }
    const double* ws = w + lo;
    if (tid == 0) {
        double tw[32], tp[32];
        for (int i = 0; i < h; ++i) { tw[i] = ws[i]; tp[i] = __dmul_rn(ws[i], xs[i]); }
        s_init[0] = np_sum_small(tw, h);
        s_init[1] = np_sum_small(tp, h);
    }
    __syncthreads();
    double highW = s_init[0], highNN = s_init[1];
    double lowW = highW, lowNN = -highNN;
    const double norm = sqrt((double)h / 2.0);
    if (tid == 0) out[0] = 0.0;
    for (int c0 = 1; c0 < n; c0 += HC_CHUNK) {
        const int m = min(HC_CHUNK, n - c0);
        for (int t = tid; t < m; t += 128) {
            const int k = c0 + t;
            const int hi = mirror_hi(k + h - 1, n), lw = mirror_lo(k - h - 1);
            const double mid = __dmul_rn(xs[k - 1], ws[k - 1]);
            s_a[t] = __dsub_rn(__dmul_rn(xs[lw], ws[lw]), mid);   // lowNonNormed increment
            s_b[t] = __dsub_rn(__dmul_rn(xs[hi], ws[hi]), mid);   // highNonNormed increment
            s_c[t] = __dsub_rn(ws[k - 1], ws[lw]);                // lowWeightSum increment
            s_d[t] = __dsub_rn(ws[hi], ws[k - 1]);                // highWeightSum increment
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll 4
            for (int t = 0; t < m; ++t) {
                lowNN = __dadd_rn(lowNN, s_a[t]);
                highNN = __dadd_rn(highNN, s_b[t]);
                lowW = __dadd_rn(lowW, s_c[t]);
                highW = __dadd_rn(highW, s_d[t]);
                s_a[t] = lowNN; s_b[t] = highNN; s_c[t] = lowW; s_d[t] = highW;
            }
        }
        __syncthreads();
        for (int t = tid; t < m; t += 128)
            out[c0 + t] = __dmul_rn(norm, __dadd_rn(__ddiv_rn(s_a[t], s_c[t]), __ddiv_rn(s_b[t], s_d[t])));
        __syncthreads();
    }
}

__global__ void abs_kernel(const double* __restrict__ in, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fabs(in[i]);
}

__global__ void sigma_kernel(const double* __restrict__ med, int n_arms, double* __restrict__ sigma) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_arms) {
        const double m = med[a];
        double s = (m == m) ? __dmul_rn(m, 1.4826) : 0.0;   // empty -> 0.0 (haar.py:310-311)
        sigma[a] = fmax(s, 1e-10);                           // SIGMA_FLOOR
    }
}

// FindLocalPeaks restated per element; flags[(level-1)*N + g]
__global__ void __launch_bounds__(256)
haar_peak_flags_kernel(const double* __restrict__ conv, const long long* __restrict__ off, const int* __restrict__ arm_of,
                       long long N, u32* __restrict__ flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)HAAR_LEVELS * N) return;
    const int lev = (int)(idx / N);
    const long long g = idx - (long long)lev * N;
    const int arm = arm_of[g];
    const long long lo = off[arm];
    const int n = (int)(off[arm + 1] - lo);
    const int k = (int)(g - lo);
    u32 f = 0;
    if (k >= 1 && k <= n - 2) {
        const double* s = conv + (size_t)(lev + 1) * N + lo;
        const double c = s[k], p = s[k - 1], q = s[k + 1];
        if (c > 0.0) {
            if (c > p && c > q) f = 1;
            else if (c > p && c == q) {
                int m = k + 1;
                while (m <= n - 2 && s[m + 1] == c) ++m;      // walk the plateau
                if (m <= n - 2 && c > s[m + 1]) f = 1;         // ends in a fall inside the scanned range
            }
        } else if (c < 0.0) {
            if (c < p && c < q) f = 1;
            else if (c < p && c == q) {
                int m = k + 1;
                while (m <= n - 2 && s[m + 1] == c) ++m;
                if (m <= n - 2 && c < s[m + 1]) f = 1;
            }
        }
    }
    flags[idx] = f;
}

// compact peaks in (level, position) order; key = ordered |conv| for the ranking sort
__global__ void __launch_bounds__(256)
haar_peak_compact_kernel(const double* __restrict__ conv, const u32* __restrict__ flags, const u32* __restrict__ scan,
                         const int* __restrict__ arm_of, long long N, int n_arms, u32* __restrict__ pk_g,
                         u32* __restrict__ pk_seg, u64* __restrict__ pk_key, u32* __restrict__ pk_val) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)HAAR_LEVELS * N || !flags[idx]) return;
    const int lev = (int)(idx / N);
    const long long g = idx - (long long)lev * N;
    const u32 pos = scan[idx];
    const double v = fabs(conv[(size_t)(lev + 1) * N + g]);
    pk_g[pos] = (u32)g;
    pk_seg[pos] = (u32)(lev * n_arms + arm_of[g]);
    pk_key[pos] = f64_to_ordered(v);
    pk_val[pos] = pos;
}

__global__ void gather_seg_keys_kernel(const u32* __restrict__ order, const u32* __restrict__ pk_seg, long long n,
                                       u64* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (u64)pk_seg[order[i]];
}

// scipy.stats.norm.cdf -> ndtr (cephes): 0.5+0.5*erf(x/sqrt2) near 0, else via erfc
__device__ __forceinline__ double ndtr_dev(double a) {
    const double xx = a * 0.70710678118654752440;
    const double z = fabs(xx);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(xx);
    double yy = 0.5 * erfc(z);
    if (xx > 0) yy = 1.0 - yy;
    return yy;
}

// per sorted peak (ascending |value| inside each (level, arm) segment): does it pass the FDR line?
__global__ void __launch_bounds__(256)
haar_fdr_kernel(const u32* __restrict__ order, const u32* __restrict__ pk_seg, const u64* __restrict__ pk_key,
                const u32* __restrict__ scan, const long long* __restrict__ off, long long N, int n_arms,
                const double* __restrict__ sigma, double q, long long npk, u32* __restrict__ jmin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npk) return;
    const u32 src = order[i];
    const u32 seg = pk_seg[src];
    const int lev = (int)(seg / (u32)n_arms), arm = (int)(seg % (u32)n_arms);
    const u32 b = scan[(size_t)lev * N + off[arm]], e = scan[(size_t)lev * N + off[arm + 1]];
    const u32 M = e - b;
    if (M < 2) return;
    const u32 j = (u32)i - b;                     // ascending position inside the segment
    const u64 key = pk_key[src];
    const double v = __longlong_as_double((long long)(key & 0x7fffffffffffffffull));  // |value| >= 0
    const double p = 2.0 * (1.0 - ndtr_dev(v / sigma[arm]));
    const double mq = ((double)(M - j) / (double)M) * q;
    if (p <= mq) atomicMin(&jmin[seg], j);
}

// T per (level, arm) and the "passes the threshold" flag per peak (in list order)
__global__ void __launch_bounds__(256)
haar_keep_kernel(const u32* __restrict__ order, const u32* __restrict__ pk_seg, const u64* __restrict__ pk_key,
                 const u32* __restrict__ scan, const long long* __restrict__ off, long long N, int n_arms,
                 const u32* __restrict__ jmin, long long npk, u8* __restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npk) return;
    const u32 seg = pk_seg[i];
    const int lev = (int)(seg / (u32)n_arms), arm = (int)(seg % (u32)n_arms);
    const u32 b = scan[(size_t)lev * N + off[arm]], e = scan[(size_t)lev * N + off[arm + 1]];
    const u32 M = e - b;
    bool k;
    if (M < 2) {
        k = true;                                  // FDRThres returns 0 -> everything passes
    } else {
        const u32 j = jmin[seg];
        if (j == 0xffffffffu) k = false;           // T = nextafter(max) admits nothing
        else k = pk_key[i] >= pk_key[order[b + j]];
    }
    keep[i] = k ? 1 : 0;
}

// UnifyLevels over levels 1..5 for one arm (one CTA); lists are arm-local bin indices
__device__ __forceinline__ int lower_bound_i(const int* a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

__global__ void __launch_bounds__(256)
haar_unify_kernel(const long long* __restrict__ off, const u32* __restrict__ scan, const u32* __restrict__ pk_g,
                  const u8* __restrict__ keep, long long N, int* __restrict__ bufA, int* __restrict__ bufB,
                  int* __restrict__ bufT, int* __restrict__ bp_out, int* __restrict__ bp_count) {
    __shared__ int s_warp[8];
    __shared__ int s_total;
    const int arm = blockIdx.x;
    const long long lo = off[arm], hi = off[arm + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* cur = bufA + lo;
    int* nxt = bufB + lo;
    int* add = bufT + lo;
    int nb = 0;
    for (int lev = 0; lev < HAAR_LEVELS; ++lev) {
        const u32 b = scan[(size_t)lev * N + lo], e = scan[(size_t)lev * N + hi];
        const int window = 1 << lev;               // 2^(level-1), level = lev+1
        int running = 0;
        for (u32 c = b; c < e; c += 256) {
            const u32 t = c + tid;
            bool k = (t < e) && keep[t];
            int posl = 0;
            if (k) {
                posl = (int)((long long)pk_g[t] - lo);
                if (nb > 0) {
                    const int ip = lower_bound_i(cur, nb, posl);
                    const bool left = ip > 0 && (posl - cur[ip - 1]) <= window;
                    const bool right = ip < nb && (cur[ip] - posl) <= window;
                    if (left || right) k = false;
                }
            }
            const u32 bal = __ballot_sync(0xffffffffu, k);
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            int base = running;
            for (int ww = 0; ww < warp; ++ww) base += s_warp[ww];
            if (k) add[base + __popc(bal & ((1u << lane) - 1u))] = posl;
            if (tid == 0) { int tot = 0; for (int ww = 0; ww < 8; ++ww) tot += s_warp[ww]; s_total = tot; }
            __syncthreads();
            running += s_total;
            __syncthreads();
        }
        const int na = running;
        if (na > 0) {
            // merge cur[0,nb) and add[0,na) (disjoint, both ascending) into nxt
            for (int i = tid; i < nb; i += 256) nxt[i + lower_bound_i(add, na, cur[i])] = cur[i];
            for (int i = tid; i < na; i += 256) nxt[i + lower_bound_i(cur, nb, add[i])] = add[i];
            __syncthreads();
            int* tmp = cur; cur = nxt; nxt = tmp;
            nb += na;
        }
        __syncthreads();
    }
    for (int i = tid; i < nb; i += 256) bp_out[lo + i] = cur[i];
    if (tid == 0) bp_count[arm] = nb;
}

// SegmentByPeaks means; one warp per segment
__global__ void __launch_bounds__(128)
haar_seg_mean_kernel(const double* __restrict__ x, const double* __restrict__ w, const long long* __restrict__ lo,
                     const long long* __restrict__ hi, int nseg, double* __restrict__ mean) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 4 + warp;
    if (s >= nseg) return;
    double sw = 0.0, swx = 0.0, sx = 0.0;
    for (long long k = lo[s] + lane; k < hi[s]; k += 32) {
        const double xv = x[k];
        sx += xv;
        if (w) {
            const double wv = w[k];
            if (wv == wv) { sw += wv; swx += xv * wv; }
        }
    }
    sw = warp_sum(sw); swx = warp_sum(swx); sx = warp_sum(sx);
    if (lane == 0) mean[s] = (w && sw > 0.0) ? swx / sw : sx / (double)(hi[s] - lo[s]);
}

__global__ void arm_of_kernel(const long long* __restrict__ off, int n_arms, int* __restrict__ arm_of) {
    const int a = blockIdx.x;
    for (long long g = off[a] + threadIdx.x; g < off[a + 1]; g += blockDim.x) arm_of[g] = a;
}

__global__ void fill_u32_kernel(u32* p, long long n, u32 v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void sigma_override_kernel(const double* __restrict__ ov, int n_arms, double* __restrict__ sigma) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_arms && ov[a] == ov[a]) sigma[a] = fmax(ov[a], 1e-10);      // haar.py:334-339 sigma_override, floored
}

// peakSigmaEst of haarSeg (haar.py:307-333) for every arm: 1.4826 * median |HaarConv(x, None, 1)|, floored
int haar_sigma(const double* d_x, const long long* h_off, int n_arms, double* h_sigma, cudaStream_t stream) {
    if (n_arms <= 0) return 0;
    const long long N = h_off[n_arms];
    for (int a = 0; a < n_arms; ++a) h_sigma[a] = 1e-10;
    if (N <= 0) return 0;
    ProfScope pscope(PC_HAAR, stream);
    Scratch scratch(stream);
    long long* d_off;
    double *conv, *absv, *med, *sigma;
    int* cnt;
    CNVK_CHECK(scratch.alloc(&d_off, (size_t)n_arms + 1));
    CNVK_CHECK(scratch.alloc(&conv, (size_t)N));
    CNVK_CHECK(scratch.alloc(&absv, (size_t)N));
    CNVK_CHECK(scratch.alloc(&med, (size_t)n_arms));
    CNVK_CHECK(scratch.alloc(&sigma, (size_t)n_arms));
    CNVK_CHECK(scratch.alloc(&cnt, (size_t)n_arms));
    CNVK_CHECK(cudaMemcpyAsync(d_off, h_off, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    haar_conv_kernel<<<dim3(n_arms, 1), 128, 0, stream>>>(d_x, nullptr, d_off, N, conv);     // chain 0 only
    CNVK_LAUNCH_CHECK();
    abs_kernel<<<div_up(N, 256), 256, 0, stream>>>(conv, N, absv);
    CNVK_LAUNCH_CHECK();
    int rc = segmented_median(absv, nullptr, d_off, n_arms, med, cnt, stream);
    if (rc) return rc;
    sigma_kernel<<<div_up(n_arms, 128), 128, 0, stream>>>(med, n_arms, sigma);
    CNVK_LAUNCH_CHECK();
    std::vector<double> tmp((size_t)n_arms);
    CNVK_CHECK(cudaMemcpyAsync(tmp.data(), sigma, (size_t)n_arms * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    for (int a = 0; a < n_arms; ++a)
        if (h_off[a + 1] - h_off[a] >= 2) h_sigma[a] = tmp[(size_t)a];       // fewer than 2 values: the floor (haar.py:167-168)
    return 0;
}

// SegmentByPeaks means (haar.py:405-451) for given segments
int haar_segment_means(const double* d_x, const double* d_w, const long long* h_lo, const long long* h_hi, int nseg,
                       double* h_mean, cudaStream_t stream) {
    if (nseg <= 0) return 0;
    ProfScope pscope(PC_HAAR, stream);
    Scratch scratch(stream);
    long long *d_lo, *d_hi;
    double* d_mean;
    CNVK_CHECK(scratch.alloc(&d_lo, (size_t)nseg));
    CNVK_CHECK(scratch.alloc(&d_hi, (size_t)nseg));
    CNVK_CHECK(scratch.alloc(&d_mean, (size_t)nseg));
    CNVK_CHECK(cudaMemcpyAsync(d_lo, h_lo, (size_t)nseg * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemcpyAsync(d_hi, h_hi, (size_t)nseg * sizeof(long long), cudaMemcpyHostToDevice, stream));
    haar_seg_mean_kernel<<<div_up(nseg, 4), 128, 0, stream>>>(d_x, d_w, d_lo, d_hi, nseg, d_mean);
    CNVK_LAUNCH_CHECK();
    CNVK_CHECK(cudaMemcpyAsync(h_mean, d_mean, (size_t)nseg * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

int haar_segment(const double* d_x, const double* d_w, const long long* h_off, int n_arms, double q,
                 std::vector<CbsSeg>& segs, cudaStream_t stream, const double* h_sigma_override) {
    segs.clear();
    if (n_arms <= 0) return 0;
    const long long N = h_off[n_arms];
    if (N <= 0) return 0;
    if ((long long)HAAR_LEVELS * N >= (1ll << 31)) { set_last_error("haar_segment: too many bins (%lld)", N); return -2; }
    if ((long long)HAAR_LEVELS * n_arms >= (1ll << 24)) { set_last_error("haar_segment: too many arms"); return -2; }
    ProfScope pscope(PC_HAAR, stream);
    Scratch scratch(stream);
    long long* d_off;
    int* arm_of;
    double *conv, *absv, *med, *sigma;
    int* cnt;
    CNVK_CHECK(scratch.alloc(&d_off, (size_t)n_arms + 1));
    CNVK_CHECK(scratch.alloc(&arm_of, (size_t)N));
    CNVK_CHECK(scratch.alloc(&conv, (size_t)6 * N));
    CNVK_CHECK(scratch.alloc(&absv, (size_t)N));
    CNVK_CHECK(scratch.alloc(&med, (size_t)n_arms));
    CNVK_CHECK(scratch.alloc(&sigma, (size_t)n_arms));
    CNVK_CHECK(scratch.alloc(&cnt, (size_t)n_arms));
    CNVK_CHECK(cudaMemcpyAsync(d_off, h_off, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    arm_of_kernel<<<n_arms, 256, 0, stream>>>(d_off, n_arms, arm_of);
    CNVK_LAUNCH_CHECK();
    haar_conv_kernel<<<dim3(n_arms, 6), 128, 0, stream>>>(d_x, d_w, d_off, N, conv);
    CNVK_LAUNCH_CHECK();
    abs_kernel<<<div_up(N, 256), 256, 0, stream>>>(conv, N, absv);
    CNVK_LAUNCH_CHECK();
    int rc = segmented_median(absv, nullptr, d_off, n_arms, med, cnt, stream);
    if (rc) return rc;
    sigma_kernel<<<div_up(n_arms, 128), 128, 0, stream>>>(med, n_arms, sigma);
    CNVK_LAUNCH_CHECK();
    if (h_sigma_override) {
        double* d_ov;
        CNVK_CHECK(scratch.alloc(&d_ov, (size_t)n_arms));
        CNVK_CHECK(cudaMemcpyAsync(d_ov, h_sigma_override, (size_t)n_arms * sizeof(double), cudaMemcpyHostToDevice, stream));
        sigma_override_kernel<<<div_up(n_arms, 128), 128, 0, stream>>>(d_ov, n_arms, sigma);
        CNVK_LAUNCH_CHECK();
    }
    // peaks
    const long long LN = (long long)HAAR_LEVELS * N;
    u32 *flags, *scan, *total;
    CNVK_CHECK(scratch.alloc(&flags, (size_t)LN + 1));
    CNVK_CHECK(scratch.alloc(&scan, (size_t)LN + 1));
    CNVK_CHECK(scratch.alloc(&total, 1));
    haar_peak_flags_kernel<<<div_up(LN, 256), 256, 0, stream>>>(conv, d_off, arm_of, N, flags);
    CNVK_LAUNCH_CHECK();
    CNVK_CHECK(cudaMemsetAsync(flags + LN, 0, sizeof(u32), stream));
    rc = exclusive_scan_u32(flags, scan, LN + 1, total, stream);  // scan[LN] = number of peaks
    if (rc) return rc;
    u32 npk32 = 0;
    CNVK_CHECK(cudaMemcpyAsync(&npk32, scan + LN, sizeof(u32), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    const long long npk = npk32;
    const size_t pk_cap = (size_t)std::max<long long>(npk, 1);
    u32 *pk_g, *pk_seg, *v0, *v1, *jmin;
    u64 *k0, *k1;
    u8* keep;
    CNVK_CHECK(scratch.alloc(&pk_g, pk_cap));
    CNVK_CHECK(scratch.alloc(&pk_seg, pk_cap));
    CNVK_CHECK(scratch.alloc(&v0, pk_cap));
    CNVK_CHECK(scratch.alloc(&v1, pk_cap));
    CNVK_CHECK(scratch.alloc(&k0, pk_cap));
    CNVK_CHECK(scratch.alloc(&k1, pk_cap));
    CNVK_CHECK(scratch.alloc(&keep, pk_cap));
    CNVK_CHECK(scratch.alloc(&jmin, (size_t)HAAR_LEVELS * n_arms));
    u64* pk_key;  // ordered |value| in list order (kept besides the sort's ping-pong buffers)
    CNVK_CHECK(scratch.alloc(&pk_key, pk_cap));
    if (npk > 0) {
        haar_peak_compact_kernel<<<div_up(LN, 256), 256, 0, stream>>>(conv, flags, scan, arm_of, N, n_arms, pk_g, pk_seg, k0, v0);
        CNVK_LAUNCH_CHECK();
        CNVK_CHECK(cudaMemcpyAsync(pk_key, k0, (size_t)npk * sizeof(u64), cudaMemcpyDeviceToDevice, stream));
        u64* ok;
        u32* ov;
        rc = radix_sort_pairs(k0, k1, v0, v1, npk, 0, 64, &ok, &ov, stream);   // by |value|
        if (rc) return rc;
        u64* kk = (ok == k0) ? k1 : k0;
        u32* vv = (ov == v0) ? v1 : v0;
        gather_seg_keys_kernel<<<div_up(npk, 256), 256, 0, stream>>>(ov, pk_seg, npk, ok);
        CNVK_LAUNCH_CHECK();
        int seg_bits = 8;
        while ((1ll << seg_bits) < (long long)HAAR_LEVELS * n_arms) seg_bits += 8;
        u64* ok2;
        u32* order;
        rc = radix_sort_pairs(ok, kk, ov, vv, npk, 0, seg_bits, &ok2, &order, stream);  // stable, by (level, arm)
        if (rc) return rc;
        fill_u32_kernel<<<div_up((long long)HAAR_LEVELS * n_arms, 256), 256, 0, stream>>>(jmin, (long long)HAAR_LEVELS * n_arms, 0xffffffffu);
        CNVK_LAUNCH_CHECK();
        haar_fdr_kernel<<<div_up(npk, 256), 256, 0, stream>>>(order, pk_seg, pk_key, scan, d_off, N, n_arms, sigma, q, npk, jmin);
        CNVK_LAUNCH_CHECK();
        haar_keep_kernel<<<div_up(npk, 256), 256, 0, stream>>>(order, pk_seg, pk_key, scan, d_off, N, n_arms, jmin, npk, keep);
        CNVK_LAUNCH_CHECK();
    }
    // unify levels per arm
    int *bufA, *bufB, *bufT, *bp, *bp_count;
    CNVK_CHECK(scratch.alloc(&bufA, (size_t)N));
    CNVK_CHECK(scratch.alloc(&bufB, (size_t)N));
    CNVK_CHECK(scratch.alloc(&bufT, (size_t)N));
    CNVK_CHECK(scratch.alloc(&bp, (size_t)N));
    CNVK_CHECK(scratch.alloc(&bp_count, (size_t)n_arms));
    haar_unify_kernel<<<n_arms, 256, 0, stream>>>(d_off, scan, pk_g, keep, N, bufA, bufB, bufT, bp, bp_count);
    CNVK_LAUNCH_CHECK();
    std::vector<int> h_cnt(n_arms), h_bp((size_t)N);
    CNVK_CHECK(cudaMemcpyAsync(h_cnt.data(), bp_count, (size_t)n_arms * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaMemcpyAsync(h_bp.data(), bp, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    std::vector<long long> slo, shi;
    for (int a = 0; a < n_arms; ++a) {
        const long long lo = h_off[a], hi = h_off[a + 1];
        if (hi <= lo) continue;
        long long prev = lo;
        for (int k = 0; k < h_cnt[a]; ++k) {
            const long long b = lo + h_bp[(size_t)lo + k];
            CbsSeg s; s.arm = a; s.lo = prev; s.hi = b; s.mean = 0.0;
            segs.push_back(s); slo.push_back(prev); shi.push_back(b);
            prev = b;
        }
        CbsSeg s; s.arm = a; s.lo = prev; s.hi = hi; s.mean = 0.0;
        segs.push_back(s); slo.push_back(prev); shi.push_back(hi);
    }
    const int ns = (int)segs.size();
    if (ns > 0) {
        long long *d_lo, *d_hi;
        double* d_mean;
        CNVK_CHECK(scratch.alloc(&d_lo, (size_t)ns));
        CNVK_CHECK(scratch.alloc(&d_hi, (size_t)ns));
        CNVK_CHECK(scratch.alloc(&d_mean, (size_t)ns));
        CNVK_CHECK(cudaMemcpyAsync(d_lo, slo.data(), (size_t)ns * sizeof(long long), cudaMemcpyHostToDevice, stream));
        CNVK_CHECK(cudaMemcpyAsync(d_hi, shi.data(), (size_t)ns * sizeof(long long), cudaMemcpyHostToDevice, stream));
        haar_seg_mean_kernel<<<div_up(ns, 4), 128, 0, stream>>>(d_x, d_w, d_lo, d_hi, ns, d_mean);
        CNVK_LAUNCH_CHECK();
        std::vector<double> means(ns);
        CNVK_CHECK(cudaMemcpyAsync(means.data(), d_mean, (size_t)ns * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CNVK_CHECK(cudaStreamSynchronize(stream));
        for (int s = 0; s < ns; ++s) segs[s].mean = means[s];
    }
    return 0;
}

}  // namespace cnvk
