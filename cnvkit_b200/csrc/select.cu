// Segmented exact medians by MSB-first radix select (one CTA per segment, no sort, no compaction).
//
// Used for CopyNumArray.center_all (cnvlib/cnary.py:253-313: per-chromosome medians of the
// autosomes, then the median of those) and CopyNumArray.residuals() (cnary.py:688-692: log2 minus
// the chromosome median).  pandas' Series.median skips NaN and averages the two middle values of
// an even count -- reproduced exactly: the select returns input *elements*, the average is one
// (a+b)/2.
#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

__device__ __forceinline__ double ordered_to_f64(u64 k) {
    const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) ordered key among used elements of [lo,hi); all threads get the result.
__device__ u64 cta_radix_select(const double* __restrict__ x, const unsigned char* __restrict__ use,
                                long long lo, long long hi, long long k, u32* s_hist, u64* s_state) {
    u64 prefix = 0, pmask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int d = threadIdx.x; d < 256; d += blockDim.x) s_hist[d] = 0;
        __syncthreads();
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            if (use && !use[i]) continue;
            const u64 key = f64_to_ordered(x[i]);
            if (key == ~0ull) continue;  // NaN
            if ((key & pmask) == prefix) {
                // warp-aggregated: in the leading passes every key of a segment shares the digit, and
                // 1024 atomics on one shared-memory word serialise
                const u32 d = (u32)(key >> shift) & 255u;
                const u32 active = __activemask();
                const u32 peers = __match_any_sync(active, d);
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[d], (u32)__popc(peers));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // bucket holding the k-th element: lane l owns buckets 8l..8l+7
            const int lane = threadIdx.x;
            u32 c[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = s_hist[8 * lane + q]; tot += c[q]; }
            u32 inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t_ = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t_;
            }
            const u32 owners = __ballot_sync(0xffffffffu, (long long)inc > k);
            const int owner = __ffs(owners) - 1;
            if (lane == owner) {
                long long kk = k - (long long)(inc - tot);
                int q = 0;
                for (; q < 7; ++q) {
                    if (kk < (long long)c[q]) break;
                    kk -= c[q];
                }
                s_state[0] = (u64)(8 * lane + q);
                s_state[1] = (u64)kk;
            }
        }
        __syncthreads();
        prefix |= s_state[0] << shift;
        pmask |= 0xffull << shift;
        k = (long long)s_state[1];
        __syncthreads();
    }
    return prefix;
}

// Register-resident variant for segments of at most SEL_K * 1024 elements (a chromosome of an exome panel, an arm of
// a Haar level): every thread keeps its keys in registers, so the eight digit passes touch no memory but the 256-entry
// shared histogram, and the loads of the single read pass are all in flight at once.  (The memory-walking select above
// re-reads the segment once per digit with one dependent load per iteration: 60-90 us for a 13 k-bin chromosome
// against ~10 us here.)
static const int SEL_K = 16;

__device__ u64 cta_radix_select_regs(const u64 (&key)[SEL_K], u32 alive, long long k, u32* s_hist, u64* s_state) {
    u64 prefix = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int d = threadIdx.x; d < 256; d += blockDim.x) s_hist[d] = 0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < SEL_K; ++q) {
            if (alive & (1u << q)) {
                const u32 d = (u32)(key[q] >> shift) & 255u;
                const u32 peers = __match_any_sync(__activemask(), d);
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[d], (u32)__popc(peers));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            u32 c[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = s_hist[8 * lane + q]; tot += c[q]; }
            u32 inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t_ = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t_;
            }
            const u32 owners = __ballot_sync(0xffffffffu, (long long)inc > k);
            const int owner = __ffs(owners) - 1;
            if (lane == owner) {
                long long kk = k - (long long)(inc - tot);
                int q = 0;
                for (; q < 7; ++q) {
                    if (kk < (long long)c[q]) break;
                    kk -= c[q];
                }
                s_state[0] = (u64)(8 * lane + q);
                s_state[1] = (u64)kk;
            }
        }
        __syncthreads();
        const u32 dsel = (u32)s_state[0];
        prefix |= (u64)dsel << shift;
        k = (long long)s_state[1];
#pragma unroll
        for (int q = 0; q < SEL_K; ++q)
            if ((alive & (1u << q)) && ((u32)(key[q] >> shift) & 255u) != dsel) alive &= ~(1u << q);
        __syncthreads();
    }
    return prefix;
}

__global__ void __launch_bounds__(1024)
segmented_median_kernel(const double* __restrict__ x, const unsigned char* __restrict__ use,
                        const long long* __restrict__ seg_off, double* __restrict__ med, int* __restrict__ cnt_out) {
    __shared__ u32 s_hist[256];
    __shared__ u64 s_state[2];
    __shared__ long long s_cnt;
    __shared__ u64 s_min_gt;
    __shared__ u32 s_cle;
    const long long lo = seg_off[blockIdx.x], hi = seg_off[blockIdx.x + 1];
    if (threadIdx.x == 0) { s_cnt = 0; s_min_gt = ~0ull; s_cle = 0; }
    __syncthreads();
    const bool in_regs = hi - lo <= (long long)SEL_K * 1024;
    u64 key[SEL_K];
    u32 alive = 0;
    long long c = 0;
    if (in_regs) {
#pragma unroll
        for (int q = 0; q < SEL_K; ++q) {
            const long long i = lo + (long long)q * 1024 + threadIdx.x;
            const bool ok = i < hi && (!use || use[i]);
            key[q] = ok ? f64_to_ordered(x[i]) : ~0ull;          // NaN maps to ~0 as well
            if (key[q] != ~0ull) alive |= 1u << q;
        }
        c = __popc(alive);
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            if (use && !use[i]) continue;
            const double v = x[i];
            if (v == v) ++c;
        }
    }
    c = (long long)warp_sum_u32((u32)c);                         // a thread counts at most 2^31 / 1024 elements
    if ((threadIdx.x & 31) == 0 && c) atomicAdd((unsigned long long*)&s_cnt, (unsigned long long)c);
    __syncthreads();
    const long long cnt = s_cnt;
    if (cnt == 0) {
        if (threadIdx.x == 0) { med[blockIdx.x] = __longlong_as_double(0x7ff8000000000000ll); cnt_out[blockIdx.x] = 0; }
        return;
    }
    const long long k1 = (cnt - 1) / 2;  // lower middle
    const u64 key1 = in_regs ? cta_radix_select_regs(key, alive, k1, s_hist, s_state)
                             : cta_radix_select(x, use, lo, hi, k1, s_hist, s_state);
    double result = ordered_to_f64(key1);
    if ((cnt & 1) == 0) {
        // upper middle: key1 again if enough duplicates, else the smallest key above it
        u32 cle = 0;
        u64 mgt = ~0ull;
        if (in_regs) {
#pragma unroll
            for (int q = 0; q < SEL_K; ++q) {
                if (key[q] == ~0ull) continue;
                if (key[q] <= key1) ++cle; else if (key[q] < mgt) mgt = key[q];
            }
        } else {
            for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                if (use && !use[i]) continue;
                const u64 kk = f64_to_ordered(x[i]);
                if (kk == ~0ull) continue;
                if (kk <= key1) ++cle; else if (kk < mgt) mgt = kk;
            }
        }
        if (cle) atomicAdd(&s_cle, cle);
        if (mgt != ~0ull) atomicMin((unsigned long long*)&s_min_gt, (unsigned long long)mgt);
        __syncthreads();
        const u64 key2 = ((long long)s_cle >= k1 + 2) ? key1 : s_min_gt;
        result = (ordered_to_f64(key1) + ordered_to_f64(key2)) / 2.0;
    }
    if (threadIdx.x == 0) { med[blockIdx.x] = result; cnt_out[blockIdx.x] = (int)min(cnt, (long long)0x7fffffff); }
}

int segmented_median(const double* d_x, const unsigned char* d_use, const long long* d_seg_off, int nseg,
                     double* d_med, int* d_cnt, cudaStream_t stream) {
    if (nseg <= 0) return 0;
    ProfScope ps(PC_FIX, stream);
    segmented_median_kernel<<<nseg, 1024, 0, stream>>>(d_x, d_use, d_seg_off, d_med, d_cnt);
    CNVK_LAUNCH_CHECK();
    return 0;
}

// median of the medians of non-empty segments (pandas: estimator(pd.Series(values))); NaN if none.
// single warp, rank counting (segment counts are tens to a few thousands).
__global__ void median_of_medians_kernel(const double* __restrict__ med, const int* __restrict__ cnt, int nseg,
                                         double* __restrict__ out, double sign) {
    const int lane = threadIdx.x;
    int m = 0;
    for (int s = lane; s < nseg; s += 32) m += (cnt[s] > 0 && med[s] == med[s]) ? 1 : 0;
    m = (int)warp_sum_u32((u32)m);
    if (m == 0) {
        if (lane == 0) *out = 0.0;
        return;
    }
    const int k1 = (m - 1) / 2, k2 = m / 2;
    double v1 = 0.0, v2 = 0.0;
    bool f1 = false, f2 = false;
    for (int s = lane; s < nseg; s += 32) {
        if (!(cnt[s] > 0 && med[s] == med[s])) continue;
        const double v = med[s];
        int less = 0, eq_before = 0;
        for (int t = 0; t < nseg; ++t) {
            if (!(cnt[t] > 0 && med[t] == med[t])) continue;
            if (med[t] < v) ++less;
            else if (med[t] == v && t < s) ++eq_before;
        }
        const int r = less + eq_before;
        if (r == k1) { v1 = v; f1 = true; }
        if (r == k2) { v2 = v; f2 = true; }
    }
    // exactly one lane holds each
    for (int o = 16; o > 0; o >>= 1) {
        const double o1 = __shfl_xor_sync(0xffffffffu, v1, o);
        const bool of1 = __shfl_xor_sync(0xffffffffu, (int)f1, o);
        const double o2 = __shfl_xor_sync(0xffffffffu, v2, o);
        const bool of2 = __shfl_xor_sync(0xffffffffu, (int)f2, o);
        if (!f1 && of1) { v1 = o1; f1 = true; }
        if (!f2 && of2) { v2 = o2; f2 = true; }
    }
    if (lane == 0) *out = sign * ((k1 == k2) ? v1 : (v1 + v2) / 2.0);
}

int median_of_medians(const double* d_med, const int* d_cnt, int nseg, double* d_out, double sign,
                      cudaStream_t stream) {
    median_of_medians_kernel<<<1, 32, 0, stream>>>(d_med, d_cnt, nseg, d_out, sign);
    CNVK_LAUNCH_CHECK();
    return 0;
}

}  // namespace cnvk
