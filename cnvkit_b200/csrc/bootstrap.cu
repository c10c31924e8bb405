// Bootstrap confidence interval of the weighted segment mean, every segment of a table at once:
// cnvlib/segmetrics.py:176-245 confidence_interval_bootstrap (+ _smooth_samples_by_weight :248-289,
// _bca_correct_alpha :292-351), the 'ci' interval statistic of do_segmetrics (:24-121) that
// `cnvkit.py batch` runs right after segmentation (batch.py:559-566).
//
// The reference draws, per segment of k bins, rng = default_rng(0xA5EED); rng.integers(0, k, (B, k)):
// numpy's PCG64 stream read as 32-bit halves (low half first) through Lemire's bounded-integer
// rejection (a draw d is kept iff low32(d * k) >= 2^32 mod k).  PCG64 is an LCG, so any position of the
// raw stream is reachable by a jump; rejections only shift positions.  Pass 1 counts the kept draws per
// chunk of the raw stream, pass 2 (one warp per bootstrap replicate) finds its first raw position from
// the chunk prefix and walks on, so the resampled indices are numpy's, draw for draw.  Each replicate's
// weighted mean is accumulated in numpy's pairwise order (8 strided accumulators per block of <= 128
// terms, blocks combined along the n/2-rounded-to-8 recursion), which makes the bootstrap distribution
// -- and with it the count of replicates below the observed mean that BCa starts from -- bit-equal to
// np.average's.  The smoothing noise of the reference comes from an unseeded generator (not
// reproducible by construction); here it is a counter-based normal keyed by (segment, element).
#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

typedef unsigned __int128 u128;

static const int BOOT_CHUNK = 4096;      // raw 32-bit draws per counted chunk
static const int BOOT_LEAF = 128;        // numpy PW_BLOCKSIZE
static const int BOOT_MAX_B = 4096;      // replicates per segment the CI kernel sorts in shared memory
static const int BOOT_WARPS = 4;

struct BootSeg {
    long long lo;         // first bin (row of values / weights)
    long long chunk_off;  // first chunk of this segment in the count / prefix arrays
    int k;                // bins
    int nchunks;
    int smooth;           // smoothed bootstrap for this segment
    int out;              // row of the caller's segment table
};

__host__ __device__ __forceinline__ u128 mk128(u64 hi, u64 lo) { return ((u128)hi << 64) | (u128)lo; }

// PCG_DEFAULT_MULTIPLIER_128
__host__ __device__ __forceinline__ u128 pcg_mult() { return mk128(0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull); }

// coefficients of `delta` LCG steps: s -> am * s + ap
__host__ __device__ inline void pcg_jump(u128 inc, u64 delta, u128& am, u128& ap) {
    u128 acc_m = 1, acc_p = 0, cur_m = pcg_mult(), cur_p = inc;
    while (delta) {
        if (delta & 1) {
            acc_m = acc_m * cur_m;
            acc_p = acc_p * cur_m + cur_p;
        }
        cur_p = (cur_m + 1) * cur_p;
        cur_m = cur_m * cur_m;
        delta >>= 1;
    }
    am = acc_m;
    ap = acc_p;
}

// XSL-RR output of a 128-bit state
__device__ __forceinline__ u64 pcg_out(u128 s) {
    const u64 hi = (u64)(s >> 64), lo = (u64)s;
    const u64 x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

// One lane's view of the raw stream: positions base + lane, base + lane + 32, ...
struct RawLane {
    u128 st;      // state whose output holds this lane's current position
    u128 a16, c16;
    int half;     // 0: low 32 bits, 1: high 32 bits
    __device__ void seek(u128 s0, u128 inc, u64 raw_pos) {
        u128 am, ap;
        pcg_jump(inc, (raw_pos >> 1) + 1, am, ap);   // the first output comes after one step
        st = am * s0 + ap;
        half = (int)(raw_pos & 1);
    }
    __device__ __forceinline__ u32 next() {
        const u64 o = pcg_out(st);
        st = a16 * st + c16;
        return half ? (u32)(o >> 32) : (u32)o;
    }
};

// ---- pass 1: kept draws per chunk of the raw stream -------------------------------------------
__global__ void __launch_bounds__(32 * BOOT_WARPS)
boot_count_kernel(const BootSeg* __restrict__ segs, int nact, long long total_chunks, u64 s_hi, u64 s_lo,
                  u64 i_hi, u64 i_lo, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * BOOT_WARPS + (threadIdx.x >> 5);
    if (c >= total_chunks) return;
    int a = 0, hi_ = nact;
    while (hi_ - a > 1) {
        const int m = (a + hi_) >> 1;
        if (segs[m].chunk_off <= c) a = m; else hi_ = m;
    }
    const u32 k = (u32)segs[a].k;
    const u32 thr = (u32)((0x100000000ull - (u64)k) % (u64)k);   // == (UINT32_MAX - (k - 1)) % k
    if (thr == 0) {
        if (lane == 0) counts[c] = BOOT_CHUNK;
        return;
    }
    const u128 s0 = mk128(s_hi, s_lo), inc = mk128(i_hi, i_lo);
    RawLane rl;
    pcg_jump(inc, 16, rl.a16, rl.c16);
    rl.seek(s0, inc, (u64)(c - segs[a].chunk_off) * BOOT_CHUNK + lane);
    int cnt = 0;
    for (int it = 0; it < BOOT_CHUNK / 32; ++it) {
        const u64 m = (u64)rl.next() * (u64)k;
        cnt += ((u32)m >= thr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) counts[c] = cnt;
}

// exclusive prefix of the chunk counts, one warp per segment
__global__ void __launch_bounds__(32 * BOOT_WARPS)
boot_prefix_kernel(const BootSeg* __restrict__ segs, int nact, const int* __restrict__ counts,
                   long long* __restrict__ pre, long long need_per_k, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * BOOT_WARPS + (threadIdx.x >> 5);
    if (a >= nact) return;
    const long long co = segs[a].chunk_off;
    const int nc = segs[a].nchunks;
    long long run = 0;
    for (int b = 0; b < nc; b += 32) {
        const int i = b + lane;
        const long long v = (i < nc) ? counts[co + i] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < nc) pre[co + i] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0 && run < need_per_k * (long long)segs[a].k) atomicExch(err, 1);
}

// ---- pass 2: one warp per (segment, replicate) ------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

// standard normal number `c` of the stream `key` (Box-Muller on two hashed uniforms)
__device__ __forceinline__ double boot_normal(u64 key, u64 c) {
    const u64 h1 = mix64(key + (c + 1) * 0x9E3779B97F4A7C15ull);
    const u64 h2 = mix64(h1 ^ 0xD1B54A32D192ED03ull);
    const double u1 = ((double)(h1 >> 11) + 0.5) * 0x1.0p-53;
    const double u2 = (double)(h2 >> 11) * 0x1.0p-53;
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// numpy pairwise_sum of one block of n <= 128 staged terms: lanes 0-7 sum pa, lanes 8-15 sum pb.
// Returns the sum of pa on lane 0 and of pb on lane 8.
__device__ __forceinline__ double leaf_sum(const double* pa, const double* pb, int n, int lane) {
    const double* p = (lane & 8) ? pb : pa;
    const int j = lane & 7;
    double res = 0.0;
    if (n < 8) {
        if (j == 0)
            for (int i = 0; i < n; ++i) res = __dadd_rn(res, p[i]);
        return res;
    }
    const int nfull = n - (n & 7);
    double r = p[j];
    for (int i = 8; i < nfull; i += 8) r = __dadd_rn(r, p[i + j]);
    r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1));   // even j: r[j] + r[j+1]
    r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 2));   // j % 4 == 0: (..) + (..)
    r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 4));   // j == 0
    if (j == 0)
        for (int i = nfull; i < n; ++i) r = __dadd_rn(r, p[i]);
    return r;
}

__global__ void __launch_bounds__(32 * BOOT_WARPS)
boot_replicate_kernel(const double* __restrict__ values, const double* __restrict__ weights,
                      const BootSeg* __restrict__ segs, int nact, int B, const int* __restrict__ counts,
                      const long long* __restrict__ pre, u64 s_hi, u64 s_lo, u64 i_hi, u64 i_lo, u64 noise_seed,
                      double* __restrict__ boot, double* __restrict__ orig_svw, double* __restrict__ orig_sw) {
    __shared__ double s_pa[BOOT_WARPS][BOOT_LEAF], s_pb[BOOT_WARPS][BOOT_LEAF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long task = (long long)blockIdx.x * BOOT_WARPS + warp;
    if (task >= (long long)nact * (B + 1)) return;
    const int a = (int)(task / (B + 1)), r = (int)(task % (B + 1));
    const BootSeg sg = segs[a];
    const u32 k = (u32)sg.k;
    const double* val = values + sg.lo;
    const double* wt = weights + sg.lo;
    double* pa = s_pa[warp];
    double* pb = s_pb[warp];
    const bool identity = (r == B);
    const bool smooth = sg.smooth != 0 && !identity;
    const u32 thr = (u32)((0x100000000ull - (u64)k) % (u64)k);
    const u32 lt = (1u << lane) - 1u;
    const double bw = smooth ? pow((double)k, -0.25) : 0.0;
    const u64 nkey = mix64(noise_seed ^ mix64((u64)sg.out + 0x632BE59BD9B4E019ull));

    RawLane rl;
    u32 avail = 0;     // lanes of the current 32-draw step that are kept and not yet consumed
    u64 m = 0;         // this lane's draw * k
    if (!identity) {
        const u128 s0 = mk128(s_hi, s_lo), inc = mk128(i_hi, i_lo);
        pcg_jump(inc, 16, rl.a16, rl.c16);
        const long long o0 = (long long)r * (long long)k;    // first output index of this replicate
        long long lo_c = sg.chunk_off, hi_c = sg.chunk_off + sg.nchunks;
        while (hi_c - lo_c > 1) {
            const long long mid = (lo_c + hi_c) >> 1;
            if (pre[mid] <= o0) lo_c = mid; else hi_c = mid;
        }
        long long skip = o0 - pre[lo_c];
        u64 raw = (u64)(lo_c - sg.chunk_off) * BOOT_CHUNK;
        if (counts[lo_c] == BOOT_CHUNK) {   // nothing rejected in this chunk: the position is known
            raw += (u64)skip;
            skip = 0;
        }
        rl.seek(s0, inc, raw + lane);
        while (skip > 0) {                   // discard kept draws up to the replicate's first one
            if (!avail) {
                m = (u64)rl.next() * (u64)k;
                avail = __ballot_sync(0xffffffffu, (u32)m >= thr);
            }
            const int cnt = __popc(avail);
            const int take = (int)min((long long)cnt, skip);
            const bool mine = (avail >> lane) & 1u;
            const bool use = mine && __popc(avail & lt) < take;
            avail = __ballot_sync(0xffffffffu, mine && !use);
            skip -= take;
        }
    }

    // numpy pairwise recursion over the k terms, leaves filled in order
    int st_n[24];
    unsigned char st_stage[24];
    double st_a[24];
    int d = 0;
    st_n[0] = (int)k;
    st_stage[0] = 0;
    double ret = 0.0;            // lane 0: sum of v*w; lane 8: sum of w
    long long elem = 0;          // terms produced so far
    while (d >= 0) {
        const int n = st_n[d];
        if (st_stage[d] == 0) {
            if (n <= BOOT_LEAF) {
                if (identity) {
                    for (int e = lane; e < n; e += 32) {
                        const double v = val[elem + e], w = wt[elem + e];
                        pa[e] = __dmul_rn(v, w);
                        pb[e] = w;
                    }
                } else {
                    int got = 0;
                    while (got < n) {
                        if (!avail) {
                            m = (u64)rl.next() * (u64)k;
                            avail = __ballot_sync(0xffffffffu, (u32)m >= thr);
                        }
                        const int cnt = __popc(avail);
                        const int take = min(cnt, n - got);
                        const bool mine = (avail >> lane) & 1u;
                        const int rank = __popc(avail & lt);
                        const bool use = mine && rank < take;
                        if (use) {
                            const u32 idx = (u32)(m >> 32);
                            double v = val[idx];
                            const double w = wt[idx];
                            if (smooth) {
                                const u64 c = (u64)r * (u64)k + (u64)(elem + got + rank);
                                v = __dadd_rn(v, __dmul_rn(__dmul_rn(bw, sqrt(__dsub_rn(1.0, w))), boot_normal(nkey, c)));
                            }
                            pa[got + rank] = __dmul_rn(v, w);
                            pb[got + rank] = w;
                        }
                        avail = __ballot_sync(0xffffffffu, mine && !use);
                        got += take;
                    }
                }
                __syncwarp();
                ret = leaf_sum(pa, pb, n, lane);
                __syncwarp();
                elem += n;
                --d;
            } else {
                int n2 = n / 2;
                n2 -= n2 % 8;
                st_stage[d] = 1;
                st_n[d + 1] = n2;
                st_stage[d + 1] = 0;
                ++d;
            }
        } else if (st_stage[d] == 1) {
            st_a[d] = ret;
            int n2 = n / 2;
            n2 -= n2 % 8;
            st_stage[d] = 2;
            st_n[d + 1] = n - n2;
            st_stage[d + 1] = 0;
            ++d;
        } else {
            ret = __dadd_rn(st_a[d], ret);
            --d;
        }
    }
    const double svw = __shfl_sync(0xffffffffu, ret, 0), sw = __shfl_sync(0xffffffffu, ret, 8);
    if (lane == 0) {
        if (identity) {
            orig_svw[a] = svw;
            orig_sw[a] = sw;
        } else {
            boot[(long long)a * B + r] = __ddiv_rn(svw, sw);
        }
    }
}

// ---- pass 3: percentiles (with BCa where the segment was not smoothed), one CTA per segment ----
__device__ __forceinline__ double block_sum128(double v, double* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    return (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
}

// numpy.percentile(method="linear") of the sorted sample, q given as a fraction
__device__ double np_percentile_sorted(const double* s, int B, double alpha_frac) {
    const double q = (100.0 * alpha_frac) / 100.0;                       // list(100 * alphas), then / 100
    const double vi = ((double)B * q + (1.0 + q * -1.0)) - 1.0;           // n*q + (alpha + q*(1-alpha-beta)) - 1
    if (!(vi < (double)(B - 1))) return s[B - 1];
    if (vi < 0.0) return s[0];
    const double fl = floor(vi);
    const int i0 = (int)fl;
    const double t = vi - fl;
    const double x = s[i0], y = s[i0 + 1];
    const double diff = y - x;
    return (t >= 0.5) ? y - diff * (1.0 - t) : x + diff * t;
}

__global__ void __launch_bounds__(128)
boot_ci_kernel(const double* __restrict__ values, const double* __restrict__ weights,
               const long long* __restrict__ seg_lo, const long long* __restrict__ seg_hi,
               const int* __restrict__ act_of, const BootSeg* __restrict__ segs, int B, double alpha,
               const double* __restrict__ boot, const double* __restrict__ orig_svw,
               const double* __restrict__ orig_sw, double* __restrict__ ci_lo, double* __restrict__ ci_hi) {
    __shared__ double s_v[BOOT_MAX_B];
    __shared__ double s_red[4];
    __shared__ int s_cnt[4];
    const int s = blockIdx.x, tid = threadIdx.x;
    const long long lo = seg_lo[s], k = seg_hi[s] - seg_lo[s];
    if (k <= 0) {
        if (tid == 0) { ci_lo[s] = __longlong_as_double(0x7ff8000000000000ll); ci_hi[s] = ci_lo[s]; }
        return;
    }
    if (k == 1) {
        if (tid == 0) { ci_lo[s] = values[lo]; ci_hi[s] = values[lo]; }
        return;
    }
    const int a = act_of[s];
    int P2 = 1;
    while (P2 < B) P2 <<= 1;
    for (int i = tid; i < P2; i += 128) s_v[i] = (i < B) ? boot[(long long)a * B + i] : INFINITY;
    __syncthreads();
    for (int kk = 2; kk <= P2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += 128) {
                const int l = i ^ j;
                if (l > i) {
                    const double x = s_v[i], y = s_v[l];
                    const bool up = (i & kk) == 0;
                    if ((x > y) == up) { s_v[i] = y; s_v[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    double a0 = alpha / 2.0, a1 = 1.0 - alpha / 2.0;
    if (!segs[a].smooth) {
        // _bca_correct_alpha (segmetrics.py:292-351); any failed check keeps the plain alphas
        const double svw = orig_svw[a], sw = orig_sw[a];
        const double orig_mean = svw / sw;
        int below = 0;
        for (int i = tid; i < B; i += 128) below += (s_v[i] < orig_mean);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if ((tid & 31) == 0) s_cnt[tid >> 5] = below;
        __syncthreads();
        below = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
        const double prop = (double)below / (double)B;
        // jackknife: u_i = weighted mean without bin i
        double part = 0.0;
        for (long long i = tid; i < k; i += 128) {
            const double v = values[lo + i], w = weights[lo + i];
            part += (svw - v * w) / (sw - w);
        }
        const double umean = block_sum128(part, s_red) / (double)k;
        double p2 = 0.0, p3 = 0.0;
        for (long long i = tid; i < k; i += 128) {
            const double v = values[lo + i], w = weights[lo + i];
            const double uu = umean - (svw - v * w) / (sw - w);
            p2 += uu * uu;
            p3 += uu * uu * uu;
        }
        const double uu_var = block_sum128(p2, s_red);
        const double uu_cub = block_sum128(p3, s_red);
        if (prop != 0.0 && prop != 1.0 && !(uu_var < 1e-10)) {
            const double z0 = normcdfinv(prop);
            const double za0 = normcdfinv(a0), za1 = normcdfinv(a1);
            const double acc = uu_cub / (6.0 * pow(uu_var, 1.5));
            const double den0 = 1.0 - acc * (z0 + za0), den1 = 1.0 - acc * (z0 + za1);
            if (den0 > 0.0 && den1 > 0.0) {
                const double n0 = normcdf(z0 + (z0 + za0) / den0), n1 = normcdf(z0 + (z0 + za1) / den1);
                if (n0 > 0.0 && n0 < 1.0 && n1 > 0.0 && n1 < 1.0) { a0 = n0; a1 = n1; }
            }
        }
    }
    if (tid == 0) {
        ci_lo[s] = np_percentile_sorted(s_v, B, a0);
        ci_hi[s] = np_percentile_sorted(s_v, B, a1);
    }
}

int bootstrap_ci(const double* d_values, const double* d_weights, const long long* h_lo, const long long* h_hi,
                 int n_segs, double alpha, int bootstraps, int smooth_upto, const u64 pcg[4], u64 noise_seed,
                 double* d_ci_lo, double* d_ci_hi, cudaStream_t stream) {
    if (!(alpha > 0.0 && alpha < 1.0)) {
        set_last_error("alpha must be between 0 and 1; got %g", alpha);
        return -2;
    }
    int B = bootstraps;
    if ((double)B <= 2.0 / alpha) B = (int)ceil(2.0 / alpha);     // segmetrics.py:206-214
    if (B > BOOT_MAX_B) {
        set_last_error("bootstrap_ci: %d bootstraps exceed the supported %d", B, BOOT_MAX_B);
        return -2;
    }
    if (n_segs <= 0) return 0;
    ProfScope prof(PC_OTHER, stream);
    std::vector<BootSeg> act;
    std::vector<int> act_of((size_t)n_segs, -1);
    long long total_chunks = 0;
    for (int s = 0; s < n_segs; ++s) {
        const long long k = h_hi[s] - h_lo[s];
        if (k < 2) continue;
        if (k > 0x7fffffffll / (B + 1)) {
            set_last_error("bootstrap_ci: segment %d has %lld bins, too many for %d bootstraps", s, k, B);
            return -2;
        }
        BootSeg g;
        g.lo = h_lo[s];
        g.k = (int)k;
        g.smooth = (smooth_upto < 0) ? 0 : (k <= (long long)smooth_upto);
        g.out = s;
        const long long N = (long long)B * k;
        const unsigned long long thr = (0x100000000ull - (unsigned long long)k) % (unsigned long long)k;
        const long long expect = (long long)((double)N * (double)thr / 4294967296.0) + 1;
        g.nchunks = (int)((N + 8 * expect + 256 + BOOT_CHUNK - 1) / BOOT_CHUNK);
        g.chunk_off = total_chunks;
        total_chunks += g.nchunks;
        act_of[(size_t)s] = (int)act.size();
        act.push_back(g);
    }
    const int nact = (int)act.size();
    Scratch scratch(stream);
    long long *d_lo, *d_hi, *d_pre;
    int *d_act_of, *d_counts, *d_err;
    BootSeg* d_segs;
    double *d_boot, *d_svw, *d_sw;
    CNVK_CHECK(scratch.alloc(&d_lo, (size_t)n_segs));
    CNVK_CHECK(scratch.alloc(&d_hi, (size_t)n_segs));
    CNVK_CHECK(scratch.alloc(&d_act_of, (size_t)n_segs));
    CNVK_CHECK(scratch.alloc(&d_segs, (size_t)nact + 1));
    CNVK_CHECK(scratch.alloc(&d_counts, (size_t)total_chunks + 1));
    CNVK_CHECK(scratch.alloc(&d_pre, (size_t)total_chunks + 1));
    CNVK_CHECK(scratch.alloc(&d_err, 1));
    CNVK_CHECK(scratch.alloc(&d_boot, (size_t)nact * (size_t)B + 1));
    CNVK_CHECK(scratch.alloc(&d_svw, (size_t)nact + 1));
    CNVK_CHECK(scratch.alloc(&d_sw, (size_t)nact + 1));
    CNVK_CHECK(cudaMemcpyAsync(d_lo, h_lo, (size_t)n_segs * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemcpyAsync(d_hi, h_hi, (size_t)n_segs * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemcpyAsync(d_act_of, act_of.data(), (size_t)n_segs * sizeof(int), cudaMemcpyHostToDevice, stream));
    CNVK_CHECK(cudaMemsetAsync(d_err, 0, sizeof(int), stream));
    if (nact) {
        CNVK_CHECK(cudaMemcpyAsync(d_segs, act.data(), (size_t)nact * sizeof(BootSeg), cudaMemcpyHostToDevice, stream));
        boot_count_kernel<<<div_up(total_chunks, BOOT_WARPS), 32 * BOOT_WARPS, 0, stream>>>(
            d_segs, nact, total_chunks, pcg[0], pcg[1], pcg[2], pcg[3], d_counts);
        CNVK_LAUNCH_CHECK();
        boot_prefix_kernel<<<div_up(nact, BOOT_WARPS), 32 * BOOT_WARPS, 0, stream>>>(d_segs, nact, d_counts, d_pre, B, d_err);
        CNVK_LAUNCH_CHECK();
        const long long tasks = (long long)nact * (B + 1);
        boot_replicate_kernel<<<div_up(tasks, BOOT_WARPS), 32 * BOOT_WARPS, 0, stream>>>(
            d_values, d_weights, d_segs, nact, B, d_counts, d_pre, pcg[0], pcg[1], pcg[2], pcg[3], noise_seed, d_boot,
            d_svw, d_sw);
        CNVK_LAUNCH_CHECK();
    }
    boot_ci_kernel<<<n_segs, 128, 0, stream>>>(d_values, d_weights, d_lo, d_hi, d_act_of, d_segs, B, alpha, d_boot, d_svw,
                                               d_sw, d_ci_lo, d_ci_hi);
    CNVK_LAUNCH_CHECK();
    int h_err = 0;
    CNVK_CHECK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CNVK_CHECK(cudaStreamSynchronize(stream));
    if (h_err) {
        set_last_error("bootstrap_ci: the counted part of the random stream was too short");
        return -5;
    }
    return 0;
}

}  // namespace cnvk
