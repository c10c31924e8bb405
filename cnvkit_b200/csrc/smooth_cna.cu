// DNAcopy::smooth.CNA on the device -- the `--smooth-cbs` option of cnvlib/segmentation/cbs.py:51-58
// (`cna = smooth.CNA(cna)` before `segment()`, DNAcopy defaults smooth.region = 10, outlier.SD.scale = 4,
// smooth.SD.scale = 2, trim = 0.025).  Like the rest of the CBS path this restates a dependency whose source is
// not in /root/reference (PARITY UNPINNED; oracle twin: cbso_smooth_cna in oracle/cbs_oracle.c):
//   trimmed.SD = sqrt( inflfact(trim) * sum( sort(|diff(x)|)[1:n.keep]^2 ) / (2 n.keep) ),  n.keep = round((1-2 trim)(n-1))
//   smoothLR   : a bin more than 4 trimmed.SD above (below) EVERY other bin within 10 bins on either side (window
//                clipped to the arm) is moved to median(window incl. itself) +- 2 trimmed.SD.
// Kernels: (1) one CTA per arm finds the n.keep-th smallest |diff| exactly by an MSB-first radix select over the
// bit patterns (8 histogram passes over the arm, L2-resident) and sums the squares below it as a double-double
// (R's sum() is a long-double accumulation: the correctly rounded sum, independent of order);
// (2) one thread per bin applies smoothLR.
#include <cmath>
#include <vector>

#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

static const int SM_THREADS = 1024;

struct DD2 {
    double hi, lo;
};
__device__ __forceinline__ void dd2_add(DD2& a, double p) {          // a += p  (Knuth two-sum)
    const double s = __dadd_rn(a.hi, p);
    const double bb = __dsub_rn(s, a.hi);
    const double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(p, bb));
    a.hi = s;
    a.lo = __dadd_rn(a.lo, e);
}
__device__ __forceinline__ void dd2_merge(DD2& a, double bhi, double blo) {
    dd2_add(a, bhi);
    a.lo = __dadd_rn(a.lo, blo);
}

__device__ __forceinline__ unsigned long long absdiff_bits(const double* __restrict__ x, long long i) {
    return (unsigned long long)__double_as_longlong(fabs(__dsub_rn(x[i + 1], x[i])));
}

__global__ void __launch_bounds__(SM_THREADS)
trimmed_sd_kernel(const double* __restrict__ x, const long long* __restrict__ off, int n_arms, double trim, double infl,
                  double* __restrict__ sd_out) {
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ long long s_rank;
    __shared__ double s_hi[32], s_lo[32];
    __shared__ long long s_cnt[32];
    const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long lo = off[a], n = off[a + 1] - lo;
    if (n < 2) { if (tid == 0) sd_out[a] = 0.0; return; }
    const long long m = n - 1;                                            // number of differences
    long long nkeep = (long long)rint((1.0 - 2.0 * trim) * (double)m);    // R's round(): half to even
    nkeep = max(1ll, min(nkeep, m));
    const double* xa = x + lo;
    if (tid == 0) { s_prefix = 0ull; s_rank = nkeep; }
    __syncthreads();
    for (int pass = 7; pass >= 0; --pass) {
        for (int b = tid; b < 256; b += SM_THREADS) s_hist[b] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const int sh = 8 * pass;
        for (long long i = tid; i < m; i += SM_THREADS) {
            const unsigned long long bits = absdiff_bits(xa, i);
            if (pass == 7 || (bits >> (sh + 8)) == prefix) atomicAdd(&s_hist[(unsigned)((bits >> sh) & 255ull)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            long long rank = s_rank, cum = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (cum + (long long)s_hist[b] >= rank) break;
                cum += s_hist[b];
            }
            s_rank = rank - cum;
            s_prefix = (prefix << 8) | (unsigned long long)b;
        }
        __syncthreads();
    }
    const unsigned long long vk_bits = s_prefix;                          // bit pattern of the n.keep-th smallest |diff|
    DD2 acc = {0.0, 0.0};
    long long less = 0;
    for (long long i = tid; i < m; i += SM_THREADS) {
        const unsigned long long bits = absdiff_bits(xa, i);
        if (bits < vk_bits) {
            const double d = __longlong_as_double((long long)bits);
            dd2_add(acc, __dmul_rn(d, d));
            ++less;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double h = __shfl_xor_sync(0xffffffffu, acc.hi, o), l = __shfl_xor_sync(0xffffffffu, acc.lo, o);
        dd2_merge(acc, h, l);
        less += __shfl_xor_sync(0xffffffffu, less, o);
    }
    if (lane == 0) { s_hi[warp] = acc.hi; s_lo[warp] = acc.lo; s_cnt[warp] = less; }
    __syncthreads();
    if (tid == 0) {
        DD2 t = {s_hi[0], s_lo[0]};
        long long c = s_cnt[0];
        for (int k = 1; k < SM_THREADS / 32; ++k) { dd2_merge(t, s_hi[k], s_lo[k]); c += s_cnt[k]; }
        // the remaining (n.keep - c) kept values all equal v_k: add c_k * v_k^2 exactly (two-product)
        const double vk = __longlong_as_double((long long)vk_bits), p = __dmul_rn(vk, vk);
        const double mult = (double)(nkeep - c);
        const double ph = __dmul_rn(mult, p), pl = __fma_rn(mult, p, -ph);
        dd2_merge(t, ph, pl);
        const double total = __dadd_rn(t.hi, t.lo);
        sd_out[a] = sqrt(__dmul_rn(infl, __ddiv_rn(total, __dmul_rn(2.0, (double)nkeep))));
    }
}

static const int SM_KMAX = 31;

__global__ void __launch_bounds__(256)
smooth_lr_kernel(const double* __restrict__ x, const long long* __restrict__ off, int n_arms, const double* __restrict__ sd,
                 int k, double outlier_scale, double smooth_scale, double* __restrict__ out) {
    const int a = blockIdx.y;
    const long long lo = off[a], n = off[a + 1] - lo;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* xa = x + lo;
    const double xi = xa[i];
    double res = xi;
    if (n >= 2) {
        const double oSD = outlier_scale * sd[a], sSD = smooth_scale * sd[a];
        const long long jlo = max(0ll, i - k), jhi = min(n - 1, i + (long long)k);
        double mx = -INFINITY, mn = INFINITY;
        for (long long j = jlo; j <= jhi; ++j) {
            if (j == i) continue;
            const double v = xa[j];
            mx = fmax(mx, v);
            mn = fmin(mn, v);
        }
        const double above = xi - mx, below = mn - xi;
        if (above > oSD || below > oSD) {
            double nb[2 * SM_KMAX + 1];
            int m = 0;
            for (long long j = jlo; j <= jhi; ++j) {                      // insertion sort of the window
                const double v = xa[j];
                int p = m++;
                while (p > 0 && nb[p - 1] > v) { nb[p] = nb[p - 1]; --p; }
                nb[p] = v;
            }
            const double med = (m & 1) ? nb[m / 2] : (nb[m / 2 - 1] + nb[m / 2]) / 2.0;
            res = above > oSD ? med + sSD : med - sSD;
        }
    }
    out[lo + i] = res;
}

// Wichura's AS241 (PPND16), the algorithm behind R's qnorm
static double qnorm_as241(double p) {
    const double q = p - 0.5;
    double r, val;
    if (std::fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                       + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                     + 133.14166789178437745) * r + 3.387132872796366608)
              / (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r
                     + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
                   + 42.313330701600911252) * r + 1.0);
        return val;
    }
    r = q < 0 ? p : 1.0 - p;
    r = std::sqrt(-std::log(r));
    if (r <= 5.0) {
        r -= 1.6;
        val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r
                   + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
                 + 4.6303378461565452959) * r + 1.42343711074968357734)
              / (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r
                     + 0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r
                   + 2.05319162663775882187) * r + 1.0);
    } else {
        r -= 5.0;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r
                   + 0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r
                 + 5.4637849111641143699) * r + 6.6579046435011037772)
              / (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                     + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r
                   + 0.59983224954205144545) * r + 1.0);
    }
    return q < 0.0 ? -val : val;
}

// DNAcopy inflfact(trim): 1 / variance of the standard normal truncated at +-qnorm(1 - trim) (midpoint rule, 10 000 cells)
double smooth_inflfact(double trim) {
    const double a = qnorm_as241(1.0 - trim);
    const int m = 10001;
    double s = 0.0;
    for (int i = 0; i < m - 1; ++i) {
        const double x0 = -a + (2.0 * a) * (double)i / (double)(m - 1), x1 = -a + (2.0 * a) * (double)(i + 1) / (double)(m - 1);
        const double xm = (x0 + x1) / 2.0;
        s += xm * xm * (std::exp(-0.5 * xm * xm) * 0.3989422804014327) / (1.0 - 2.0 * trim);
    }
    return 1.0 / (s * (2.0 * a / 10000.0));
}

int smooth_cna(const double* d_x, const long long* h_off, int n_arms, int k, double outlier_scale, double smooth_scale,
               double trim, double* d_out, double* d_sd_out, cudaStream_t stream) {
    if (n_arms <= 0) return 0;
    if (k < 1 || k > SM_KMAX) { set_last_error("smooth_cna: smooth.region must be in [1, %d]", SM_KMAX); return -2; }
    if (!(trim >= 0.0 && trim < 0.5)) { set_last_error("smooth_cna: trim must be in [0, 0.5)"); return -2; }
    Scratch scratch(stream);
    long long* d_off;
    double* d_sd = d_sd_out;
    CNVK_CHECK(scratch.alloc(&d_off, (size_t)n_arms + 1));
    if (!d_sd) CNVK_CHECK(scratch.alloc(&d_sd, (size_t)n_arms));
    CNVK_CHECK(cudaMemcpyAsync(d_off, h_off, ((size_t)n_arms + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    long long maxn = 0;
    for (int a = 0; a < n_arms; ++a) maxn = std::max(maxn, h_off[a + 1] - h_off[a]);
    if (maxn == 0) return 0;
    ProfScope ps(PC_OTHER, stream);
    trimmed_sd_kernel<<<n_arms, SM_THREADS, 0, stream>>>(d_x, d_off, n_arms, trim, smooth_inflfact(trim), d_sd);
    CNVK_LAUNCH_CHECK();
    for (int a0 = 0; a0 < n_arms; a0 += 32768) {
        const int na = std::min(32768, n_arms - a0);
        smooth_lr_kernel<<<dim3(div_up(maxn, 256), na), 256, 0, stream>>>(d_x, d_off + a0, na, d_sd + a0, k, outlier_scale,
                                                                          smooth_scale, d_out);
        CNVK_LAUNCH_CHECK();
    }
    // d_off was copied from pageable host memory: make sure the copy is done before h_off can change
    CNVK_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

}  // namespace cnvk
