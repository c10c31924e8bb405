// Mirror-padded centred rolling median / quantile, exact order statistics.
//
// Replaces cnvlib/smoothing.py:77-87 (rolling_median) and :135-141 (rolling_quantile), i.e.
// pandas' Series.rolling(2*wing+1, min_periods, center=True).median()/.quantile(q) over the
// `_pad_array` mirror padding (smoothing.py:72-74).
//
// Design (window-size independent, O(n log n), all HBM/L2 traffic coalesced or L2-resident):
//   1. rank transform: stable radix argsort of the n values -> every element gets a distinct
//      rank in [0,n); NaN keys sort last, so ranks >= n_valid are the NaNs pandas skips.
//   2. the padded rank sequence (length M = n + 2*wing, mirrored indices) is indexed by a
//      wavelet matrix: per bit level one bitvector + per-word cumulative popcount directory.
//   3. one thread per output walks the ceil(log2 n) levels (two directory reads per level) to
//      find the k-th smallest rank inside its window, then maps rank -> value through the
//      sorted value table.  Selecting an *element* keeps the result bit-exact; even valid
//      counts (only possible when NaNs are present) average the two middle elements as
//      pandas' roll_median_c does.
#include "prims.cuh"
#include "prof.cuh"

#include <cooperative_groups.h>

#include <algorithm>

namespace cnvk {

static const int WM_BLOCK = 1024;  // elements per build block (32 words)

__global__ void count_valid_kernel(const u64* __restrict__ sorted_keys, long long n, u32* __restrict__ nvalid) {
    // sorted ascending with NaN == ~0 last: n_valid = first index holding ~0
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const bool me = sorted_keys[i] == ~0ull;
        const bool prev = (i > 0) && (sorted_keys[i - 1] == ~0ull);
        if (me && !prev) *nvalid = (u32)i;
    }
}

__global__ void rank_tables_kernel(const double* __restrict__ x, const u32* __restrict__ order, long long n,
                                   double* __restrict__ sorted_vals, u32* __restrict__ rank) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) {
        const u32 src = order[p];
        sorted_vals[p] = x[src];
        rank[src] = (u32)p;
    }
}

// padded position -> source index, per smoothing._pad_array
__device__ __forceinline__ long long mirror_index(long long t, long long n, long long wing) {
    if (t < wing) return wing - 1 - t;
    if (t < wing + n) return t - wing;
    return n - 1 - (t - wing - n);
}

__global__ void padded_ranks_kernel(const u32* __restrict__ rank, long long n, long long wing,
                                    u32* __restrict__ seq, long long M) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < M) seq[t] = rank[mirror_index(t, n, wing)];
}

__global__ void __launch_bounds__(WM_BLOCK)
wm_count_kernel(const u32* __restrict__ seq, long long M, int bit, u32* __restrict__ block_ones) {
    __shared__ u32 s_cnt[32];
    const long long t = (long long)blockIdx.x * WM_BLOCK + threadIdx.x;
    const bool b = (t < M) && ((seq[t] >> bit) & 1u);
    const u32 word = __ballot_sync(0xffffffffu, b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = __popc(word);
    __syncthreads();
    if (warp == 0) {
        u32 c = warp_sum_u32(s_cnt[lane]);
        if (lane == 0) block_ones[blockIdx.x] = c;
    }
}

__global__ void __launch_bounds__(WM_BLOCK)
wm_scatter_kernel(const u32* __restrict__ seq, u32* __restrict__ seq_out, long long M, int bit,
                  const u32* __restrict__ block_prefix, const u32* __restrict__ total_ones,
                  uint2* __restrict__ dir, u32* __restrict__ zeros_out, long long nwords) {
    __shared__ u32 s_cnt[32];
    const long long t = (long long)blockIdx.x * WM_BLOCK + threadIdx.x;
    const bool in = t < M;
    const u32 v = in ? seq[t] : 0u;
    const bool b = in && ((v >> bit) & 1u);
    const u32 word = __ballot_sync(0xffffffffu, b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = __popc(word);
    __syncthreads();
    if (warp == 0) {
        u32 c = s_cnt[lane];
        u32 inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        s_cnt[lane] = inc - c;
    }
    __syncthreads();
    const u32 ones_total = *total_ones;
    const u32 Z = (u32)M - ones_total;
    const u32 cum = block_prefix[blockIdx.x] + s_cnt[warp];  // ones before this word
    const long long w = t >> 5;
    if (lane == 0 && w < nwords) dir[w] = make_uint2(word, cum);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        dir[nwords] = make_uint2(0u, ones_total);
        *zeros_out = Z;
    }
    if (in) {
        const u32 ones_before = cum + __popc(word & ((1u << lane) - 1u));
        const u32 pos = b ? (Z + ones_before) : ((u32)t - ones_before);
        seq_out[pos] = v;
    }
}

// All levels of the wavelet matrix in ONE cooperative launch (the exome-sized sequences used to pay
// ~6 launches per level for a few microseconds of work each).  Per level: every CTA counts the ones of
// its blocks, grid.sync, every CTA sums the counts of the blocks before its own (and the total) straight
// from the table and scatters, grid.sync.  Same bitvectors, directories and order as the per-level path.
static const int WM_COOP_MAX_BLOCKS = 2048;

__global__ void __launch_bounds__(WM_BLOCK)
wm_build_coop_kernel(u32* seq_a, u32* seq_b, long long M, int levels, u32* __restrict__ block_ones,
                     uint2* __restrict__ dir, u32* __restrict__ zeros_out, long long stride, long long nwords,
                     int nblocks) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 s_cnt[32];
    __shared__ u32 s_a[32], s_b[32];
    __shared__ u32 s_pre[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32* sin = seq_a;
    u32* sout = seq_b;
    for (int lev = 0; lev < levels; ++lev) {
        const int bit = levels - 1 - lev;
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const long long t = (long long)blk * WM_BLOCK + threadIdx.x;
            const bool b = (t < M) && ((sin[t] >> bit) & 1u);
            const u32 word = __ballot_sync(0xffffffffu, b);
            if (lane == 0) s_cnt[warp] = __popc(word);
            __syncthreads();
            if (warp == 0) {
                const u32 c = warp_sum_u32(s_cnt[lane]);
                if (lane == 0) block_ones[blk] = c;
            }
            __syncthreads();
        }
        grid.sync();
        uint2* ldir = dir + (size_t)lev * stride;
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            // ones in the blocks before this one, and in all blocks
            u32 before = 0, total = 0;
            for (int q = threadIdx.x; q < nblocks; q += WM_BLOCK) {
                const u32 c = block_ones[q];
                total += c;
                if (q < blk) before += c;
            }
            before = warp_sum_u32(before);
            total = warp_sum_u32(total);
            if (lane == 0) { s_a[warp] = before; s_b[warp] = total; }
            __syncthreads();
            if (warp == 0) {
                const u32 x = warp_sum_u32(s_a[lane]), y = warp_sum_u32(s_b[lane]);
                if (lane == 0) { s_pre[0] = x; s_pre[1] = y; }
            }
            __syncthreads();
            const u32 block_prefix = s_pre[0], ones_total = s_pre[1];
            const long long t = (long long)blk * WM_BLOCK + threadIdx.x;
            const bool in = t < M;
            const u32 v = in ? sin[t] : 0u;
            const bool b = in && ((v >> bit) & 1u);
            const u32 word = __ballot_sync(0xffffffffu, b);
            if (lane == 0) s_cnt[warp] = __popc(word);
            __syncthreads();
            if (warp == 0) {
                const u32 c = s_cnt[lane];
                u32 inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                s_cnt[lane] = inc - c;
            }
            __syncthreads();
            const u32 Z = (u32)M - ones_total;
            const u32 cum = block_prefix + s_cnt[warp];  // ones before this word
            const long long w = t >> 5;
            if (lane == 0 && w < nwords) ldir[w] = make_uint2(word, cum);
            if (blk == 0 && threadIdx.x == 0) {
                ldir[nwords] = make_uint2(0u, ones_total);
                zeros_out[lev] = Z;
            }
            if (in) {
                const u32 ones_before = cum + __popc(word & ((1u << lane) - 1u));
                const u32 pos = b ? (Z + ones_before) : ((u32)t - ones_before);
                sout[pos] = v;
            }
            __syncthreads();
        }
        grid.sync();
        u32* tmp = sin; sin = sout; sout = tmp;
    }
}

static int wm_coop_grid_limit() {
    static thread_local int cached = -1;
    if (cached < 0) {
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wm_build_coop_kernel, WM_BLOCK, 0);
        cached = coop ? sms * per_sm : 0;
    }
    return cached;
}

struct WmView {
    const uint2* dir;    // [levels][stride]
    const u32* zeros;    // [levels]
    long long stride;
    int levels;
};

__device__ __forceinline__ u32 wm_rank1(const uint2* __restrict__ d, u32 p) {
    const uint2 e = d[p >> 5];
    return e.y + __popc(e.x & ((1u << (p & 31u)) - 1u));
}

// k-th smallest (0-based) value in seq[l, r)
__device__ __forceinline__ u32 wm_kth(const WmView& wm, u32 l, u32 r, u32 k) {
    u32 val = 0;
    for (int lev = 0; lev < wm.levels; ++lev) {
        const uint2* d = wm.dir + (size_t)lev * wm.stride;
        const u32 ol = wm_rank1(d, l), orr = wm_rank1(d, r);
        const u32 zl = l - ol, zr = r - orr;
        const u32 c0 = zr - zl;
        val <<= 1;
        if (k < c0) {
            l = zl;
            r = zr;
        } else {
            k -= c0;
            const u32 Z = wm.zeros[lev];
            l = Z + ol;
            r = Z + orr;
            val |= 1u;
        }
    }
    return val;
}

// number of values < v in seq[l, r)
__device__ __forceinline__ u32 wm_count_lt(const WmView& wm, u32 l, u32 r, u32 v) {
    if (wm.levels < 32 && (v >> wm.levels)) return r - l;
    u32 cnt = 0;
    for (int lev = 0; lev < wm.levels; ++lev) {
        const uint2* d = wm.dir + (size_t)lev * wm.stride;
        const u32 ol = wm_rank1(d, l), orr = wm_rank1(d, r);
        const u32 zl = l - ol, zr = r - orr;
        if ((v >> (wm.levels - 1 - lev)) & 1u) {
            cnt += zr - zl;
            const u32 Z = wm.zeros[lev];
            l = Z + ol;
            r = Z + orr;
        } else {
            l = zl;
            r = zr;
        }
    }
    return cnt;
}

// mode 0: median (pandas roll_median_c), mode 1: linear-interpolated quantile (roll_quantile)
__global__ void __launch_bounds__(256)
wm_window_query_kernel(WmView wm, const double* __restrict__ sorted_vals, const u32* __restrict__ nvalid_p,
                       long long n, long long wing, int mode, double q, int min_periods,
                       double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 l = (u32)i, r = (u32)(i + 2 * wing + 1);
    const u32 nvalid = *nvalid_p;
    u32 c = r - l;
    if (nvalid < (u32)n) c = wm_count_lt(wm, l, r, nvalid);
    double res;
    if (c == 0 || (int)c < min_periods) {
        res = __longlong_as_double(0x7ff8000000000000ll);
    } else if (mode == 0) {
        const u32 mid = c >> 1;
        const double a = sorted_vals[wm_kth(wm, l, r, mid)];
        if (c & 1u) {
            res = a;
        } else {
            const double b = sorted_vals[wm_kth(wm, l, r, mid - 1)];
            res = (a + b) / 2.0;
        }
    } else {
        const double idxf = q * (double)(c - 1);
        const u32 idx = (u32)idxf;
        const double vlow = sorted_vals[wm_kth(wm, l, r, idx)];
        if (idxf == (double)idx) {
            res = vlow;
        } else {
            const double vhigh = sorted_vals[wm_kth(wm, l, r, idx + 1)];
            res = __dadd_rn(vlow, __dmul_rn(__dsub_rn(vhigh, vlow), __dsub_rn(idxf, (double)idx)));
        }
    }
    out[i] = res;
}

// ------------------------------------------------------------------------------------------------
// Shared-memory sliding window for moderate windows (the fix path: 2*wing+1 <= 6145).
//
// One CTA owns T = L - 2*wing consecutive outputs and loads the L padded inputs they need.  The tile
// is rank-transformed in shared memory (bitonic sort of (ordered key, local index)), then every warp
// slides its own window over a contiguous chunk of outputs: the window is a BITSET over tile ranks
// (one bit per element, NaNs never set), entering / leaving elements flip one bit each, and the
// k-th smallest member is a popcount select -- per-lane word counts kept incrementally, a warp scan
// over 32 counts, then __fns inside one word.  Selecting an ELEMENT keeps the result bit-exact; an
// even number of valid values (only with NaNs) averages the two middle elements, as pandas does.
// One launch, no global scratch: ~20 us for a 150 k-bin column, against ~350 us for the wavelet path.
static const int SLIDE_MAX_OUT = 1024;             // outputs per CTA: short per-warp chains, every SM busy

template <int L>
struct SlideCfg {
    static const int THREADS = L / 8;              // 512 (L = 4096) or 1024 (L = 8192)
    static const int WARPS = THREADS / 32;
    static const int WORDS = L / 32;               // bitset words per warp
    static const int WPL = WORDS / 32;             // words per lane
    static const size_t SMEM = (size_t)L * 8       // keys (aliased by the per-warp bitsets after the sort)
                               + (size_t)L * 8     // values
                               + (size_t)L * 2     // sorted position -> local index
                               + (size_t)L * 2;    // local index -> rank
};

// ---- in-tile bitonic sort of (ordered key, local index), 8 consecutive elements per thread in registers.
// Partner distance j < 8: inside the thread; 8 <= j < 256: another lane of the warp (shuffles);
// j >= 256: another warp (one round trip through shared memory).  Only 10 of the 78 stages of a 4096-slot
// tile (15 of 91 for 8192) touch shared memory and a barrier.
__device__ __forceinline__ void sort_cx(u64& k, unsigned short& ix, u64 pk, unsigned short pix, bool keep_min) {
    const bool mine_greater = k > pk || (k == pk && ix > pix);
    if (mine_greater == keep_min) { k = pk; ix = pix; }
}

template <int J>
__device__ __forceinline__ void sort_in_thread(u64 (&key)[8], unsigned short (&idx)[8], bool asc_hi, int kk) {
    // kk < 8: the direction depends on the register slot; otherwise it is the thread's (asc_hi)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int p = r ^ J;
        if (p > r) {
            const bool asc = kk < 8 ? ((r & kk) == 0) : asc_hi;
            const bool gt = key[r] > key[p] || (key[r] == key[p] && idx[r] > idx[p]);
            if (gt == asc) {
                const u64 tk = key[r]; key[r] = key[p]; key[p] = tk;
                const unsigned short ti = idx[r]; idx[r] = idx[p]; idx[p] = ti;
            }
        }
    }
}

template <int L>
__device__ __forceinline__ void tile_bitonic_sort(u64 (&key)[8], unsigned short (&idx)[8], u64* s_key,
                                                  unsigned short* s_idx) {
    const int t = threadIdx.x;
    for (int k = 2; k <= L; k <<= 1) {
        const bool asc_hi = ((8 * t) & k) == 0;        // direction of this thread's elements when k >= 8
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 256) {
#pragma unroll
                for (int r = 0; r < 8; ++r) { s_key[8 * t + r] = key[r]; s_idx[8 * t + r] = idx[r]; }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = 8 * t + r, p = i ^ j;
                    sort_cx(key[r], idx[r], s_key[p], s_idx[p], (i < p) == asc_hi);
                }
                __syncthreads();
            } else if (j >= 8) {
                const int lm = j >> 3;
                const bool keep_min = ((t & lm) == 0) == asc_hi;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const u64 pk = __shfl_xor_sync(0xffffffffu, key[r], lm);
                    const unsigned short pix = (unsigned short)__shfl_xor_sync(0xffffffffu, (unsigned)idx[r], lm);
                    sort_cx(key[r], idx[r], pk, pix, keep_min);
                }
            } else if (j == 4) {
                sort_in_thread<4>(key, idx, asc_hi, k);
            } else if (j == 2) {
                sort_in_thread<2>(key, idx, asc_hi, k);
            } else {
                sort_in_thread<1>(key, idx, asc_hi, k);
            }
        }
    }
}

template <int L>
__global__ void __launch_bounds__(SlideCfg<L>::THREADS)
rolling_slide_kernel(const double* __restrict__ x, long long n, long long wing, int mode, double q, int min_periods,
                     double* __restrict__ out) {
    typedef SlideCfg<L> C;
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_key = reinterpret_cast<u64*>(s_raw);
    double* s_val = reinterpret_cast<double*>(s_raw + (size_t)L * 8);
    unsigned short* s_idx = reinterpret_cast<unsigned short*>(s_raw + (size_t)L * 16);
    unsigned short* s_rank = s_idx + L;
    __shared__ int s_nvalid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = min(L - (int)(2 * wing), SLIDE_MAX_OUT);     // outputs per CTA
    const long long o0 = (long long)blockIdx.x * T;            // first output = first padded index
    const long long M = n + 2 * wing;
    const int nload = (int)min((long long)(T + 2 * wing), M - o0);   // padded elements o0 .. o0+nload-1
    const int nout = (int)min((long long)T, n - o0);
    // ---- load: thread t owns slots 8t..8t+7; slots past the end of the sequence sort after everything
    u64 key[8];
    unsigned short idx[8];
    int myvalid = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int k = 8 * tid + r;
        double v = 0.0;
        u64 kk = ~0ull;
        if (k < nload) {
            v = x[mirror_index(o0 + k, n, wing)];
            kk = f64_to_ordered(v);
        }
        s_val[k] = v;
        key[r] = kk;
        idx[r] = (unsigned short)k;
        myvalid += kk != ~0ull;
    }
    if (tid == 0) s_nvalid = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) myvalid += __shfl_xor_sync(0xffffffffu, myvalid, o);
    if (lane == 0 && myvalid) atomicAdd(&s_nvalid, myvalid);   // NaN keys (and empty slots) sort last
    tile_bitonic_sort<L>(key, idx, s_key, s_idx);
    __syncthreads();
    // ---- sorted position -> local index, local index -> rank
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        s_idx[8 * tid + r] = idx[r];
        s_rank[idx[r]] = (unsigned short)(8 * tid + r);
    }
    __syncthreads();
    const int nvalid = s_nvalid;
    __syncthreads();                                           // keys are dead: their memory becomes bitsets
    u32* bits = reinterpret_cast<u32*>(s_raw) + (size_t)warp * C::WORDS;
    for (int wd = lane; wd < C::WORDS; wd += 32) bits[wd] = 0u;
    __syncwarp();
    // ---- this warp's chunk of outputs
    const int per = (nout + C::WARPS - 1) / C::WARPS;
    const int c0 = warp * per, c1 = min(nout, c0 + per);
    if (c0 >= c1) return;
    const int win = (int)(2 * wing + 1);
    // initial window [c0, c0 + win): lanes set bits (several lanes may hit one word -> atomicOr)
    for (int e = c0 + lane; e < c0 + win; e += 32) {
        const int r = s_rank[e];
        if (r < nvalid) atomicOr(&bits[r >> 5], 1u << (r & 31));
    }
    __syncwarp();
    int mycnt = 0;                                             // set bits in this lane's words
#pragma unroll
    for (int t = 0; t < C::WPL; ++t) mycnt += __popc(bits[lane * C::WPL + t]);
    int cnt = mycnt;                                           // valid elements in the window
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    // rank of the k-th smallest member (0-based) of the bitset, warp-cooperative (used once per warp)
    auto select_rank = [&](int k) -> int {
        int inc = mycnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t_ = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t_;
        }
        const unsigned owners = __ballot_sync(0xffffffffu, inc > k);
        const int owner = __ffs(owners) - 1;
        int r = 0;
        if (lane == owner) {
            int rem = k - (inc - mycnt);
            for (int t = 0; t < C::WPL; ++t) {
                const u32 wv = bits[lane * C::WPL + t];
                const int pc = __popc(wv);
                if (rem < pc) {
                    int pos = 0;                               // (rem+1)-th set bit by popcount bisection
#pragma unroll
                    for (int half = 16; half > 0; half >>= 1) {
                        const int cl = __popc(wv & (((1u << half) - 1u) << pos));
                        if (rem >= cl) { rem -= cl; pos += half; }
                    }
                    r = ((lane * C::WPL + t) << 5) + pos;
                    break;
                }
                rem -= pc;
            }
        }
        return __shfl_sync(0xffffffffu, r, owner);
    };
    // which order statistic a window of c valid values needs: the median's upper middle, or the
    // quantile's lower neighbour
    auto kof = [&](int c) -> int { return mode == 0 ? (c >> 1) : (int)(q * (double)(c - 1)); };
    int m = cnt > 0 ? select_rank(kof(cnt)) : L;               // rank of the k-th member; `below` = k
    int below = cnt > 0 ? kof(cnt) : 0;
    if (lane != 0) return;
    // ---- lane 0 slides the window: each step flips at most two bits and moves m by at most two members
    auto next_ge = [&](int start) -> int {                     // smallest member >= start, L if none
        if (start >= L) return L;
        int wd = start >> 5;
        u32 v = bits[wd] & (~0u << (start & 31));
        while (!v) {
            if (++wd >= C::WORDS) return L;
            v = bits[wd];
        }
        return (wd << 5) + __ffs(v) - 1;
    };
    auto prev_le = [&](int e) -> int {                         // largest member <= e, -1 if none
        if (e < 0) return -1;
        int wd = e >> 5;
        u32 v = bits[wd] & (0xffffffffu >> (31 - (e & 31)));
        while (!v) {
            if (--wd < 0) return -1;
            v = bits[wd];
        }
        return (wd << 5) + 31 - __clz(v);
    };
    for (int o = c0; o < c1; ++o) {
        double res;
        if (cnt == 0 || cnt < min_periods) {
            res = __longlong_as_double(0x7ff8000000000000ll);
        } else if (mode == 0) {
            const double a = s_val[s_idx[m]];
            if (cnt & 1) res = a;
            else res = (a + s_val[s_idx[prev_le(m - 1)]]) / 2.0;
        } else {
            const double idxf = q * (double)(cnt - 1);
            const int idx = (int)idxf;
            const double vlow = s_val[s_idx[m]];
            if (idxf == (double)idx) res = vlow;
            else {
                const double vhigh = s_val[s_idx[next_ge(m + 1)]];
                res = __dadd_rn(vlow, __dmul_rn(__dsub_rn(vhigh, vlow), __dsub_rn(idxf, (double)idx)));
            }
        }
        out[o0 + o] = res;
        if (o + 1 >= c1) break;
        // element o leaves, element o + win enters
        const int rl = s_rank[o], re = s_rank[o + win];
        if (rl < nvalid) {
            bits[rl >> 5] &= ~(1u << (rl & 31));
            --cnt;
            if (rl < m) --below;
            else if (rl == m) m = next_ge(m + 1);              // `below` still counts the members under m
        }
        if (re < nvalid) {
            bits[re >> 5] |= 1u << (re & 31);
            ++cnt;
            if (re < m) ++below;
        }
        if (cnt > 0) {
            const int k = kof(cnt);
            if (m >= L) { m = prev_le(L - 1); below = cnt - 1; }   // ran off the top: restart from the maximum
            while (below < k) { m = next_ge(m + 1); ++below; }
            while (below > k) { m = prev_le(m - 1); --below; }
        } else {
            m = L;
            below = 0;
        }
    }
}

template <int L>
static int launch_slide(const double* d_x, long long n, long long wing, int mode, double q, int min_periods,
                        double* d_out, cudaStream_t stream) {
    typedef SlideCfg<L> C;
    static thread_local bool configured = false;
    if (!configured) {
        CNVK_CHECK(cudaFuncSetAttribute(rolling_slide_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)C::SMEM));
        configured = true;
    }
    const int T = std::min(L - (int)(2 * wing), SLIDE_MAX_OUT);
    ProfScope qscope(PC_WINDOW_QUERY, stream);
    rolling_slide_kernel<L><<<div_up(n, T), C::THREADS, C::SMEM, stream>>>(d_x, n, wing, mode, q, min_periods, d_out);
    CNVK_LAUNCH_CHECK();
    return 0;
}

int rolling_window_stat(const double* d_x, long long n, long long wing, int mode, double q, int min_periods,
                        double* d_out, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (n < 2 || wing < 1 || wing > n - 1) {
        set_last_error("rolling_window_stat: need n >= 2 and 1 <= wing <= n-1 (n=%lld wing=%lld)", n, wing);
        return -2;
    }
    const long long M = n + 2 * wing;
    if (M >= (1ll << 31)) {
        set_last_error("rolling_window_stat: padded length %lld too large", M);
        return -2;
    }
    // moderate windows: shared-memory sliding window, one launch
    if (2 * wing + 1 <= 2048) return launch_slide<4096>(d_x, n, wing, mode, q, min_periods, d_out, stream);
    if (2 * wing + 1 <= 6145) return launch_slide<8192>(d_x, n, wing, mode, q, min_periods, d_out, stream);
    Scratch scratch(stream);
    // 1. rank transform
    u64 *k0, *k1;
    u32 *v0, *v1;
    CNVK_CHECK(scratch.alloc(&k0, (size_t)n));
    CNVK_CHECK(scratch.alloc(&k1, (size_t)n));
    CNVK_CHECK(scratch.alloc(&v0, (size_t)n));
    CNVK_CHECK(scratch.alloc(&v1, (size_t)n));
    int rc = make_sort_keys(d_x, n, k0, v0, stream);
    if (rc) return rc;
    u64* sk;
    u32* order;
    rc = radix_sort_pairs(k0, k1, v0, v1, n, 0, 64, &sk, &order, stream);
    if (rc) return rc;
    u32* nvalid;
    CNVK_CHECK(scratch.alloc(&nvalid, 1));
    const u32 n32 = (u32)n;
    CNVK_CHECK(cudaMemcpyAsync(nvalid, &n32, sizeof(u32), cudaMemcpyHostToDevice, stream));
    count_valid_kernel<<<div_up(n, 256), 256, 0, stream>>>(sk, n, nvalid);
    CNVK_LAUNCH_CHECK();
    double* sorted_vals;
    u32* rank;
    CNVK_CHECK(scratch.alloc(&sorted_vals, (size_t)n));
    CNVK_CHECK(scratch.alloc(&rank, (size_t)n));
    rank_tables_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_x, order, n, sorted_vals, rank);
    CNVK_LAUNCH_CHECK();
    // 2. wavelet matrix over the padded rank sequence
    int levels = 1;
    while ((1ll << levels) < n) ++levels;
    const long long nwords = (M + 31) / 32;
    const long long stride = nwords + 1;
    const int nblocks = div_up(M, WM_BLOCK);
    u32 *seq0, *seq1, *block_ones, *total_ones, *zeros;
    uint2* dir;
    CNVK_CHECK(scratch.alloc(&seq0, (size_t)M));
    CNVK_CHECK(scratch.alloc(&seq1, (size_t)M));
    CNVK_CHECK(scratch.alloc(&block_ones, (size_t)nblocks));
    CNVK_CHECK(scratch.alloc(&total_ones, 1));
    CNVK_CHECK(scratch.alloc(&zeros, (size_t)levels));
    CNVK_CHECK(scratch.alloc(&dir, (size_t)levels * stride));
    padded_ranks_kernel<<<div_up(M, 256), 256, 0, stream>>>(rank, n, wing, seq0, M);
    CNVK_LAUNCH_CHECK();
    u32* sin = seq0;
    u32* sout = seq1;
    ProfScope* wscope = new ProfScope(PC_WAVELET_BUILD, stream);
    const int coop_limit = nblocks <= WM_COOP_MAX_BLOCKS ? wm_coop_grid_limit() : 0;
    if (coop_limit > 0) {
        int grid = std::min(nblocks, coop_limit);
        long long M_ = M, stride_ = stride, nwords_ = nwords;
        int levels_ = levels, nblocks_ = nblocks;
        void* args[] = {&sin, &sout, &M_, &levels_, &block_ones, &dir, &zeros, &stride_, &nwords_, &nblocks_};
        CNVK_CHECK(cudaLaunchCooperativeKernel((const void*)wm_build_coop_kernel, dim3(grid), dim3(WM_BLOCK), args, 0,
                                               stream));
        CNVK_LAUNCH_CHECK();
    }
    for (int lev = 0; lev < levels && coop_limit <= 0; ++lev) {
        const int bit = levels - 1 - lev;
        wm_count_kernel<<<nblocks, WM_BLOCK, 0, stream>>>(sin, M, bit, block_ones);
        CNVK_LAUNCH_CHECK();
        rc = exclusive_scan_u32(block_ones, block_ones, nblocks, total_ones, stream);
        if (rc) return rc;
        wm_scatter_kernel<<<nblocks, WM_BLOCK, 0, stream>>>(sin, sout, M, bit, block_ones, total_ones,
                                                             dir + (size_t)lev * stride, zeros + lev, nwords);
        CNVK_LAUNCH_CHECK();
        u32* t = sin; sin = sout; sout = t;
    }
    delete wscope;
    // 3. queries
    ProfScope qscope(PC_WINDOW_QUERY, stream);
    WmView wm;
    wm.dir = dir;
    wm.zeros = zeros;
    wm.stride = stride;
    wm.levels = levels;
    wm_window_query_kernel<<<div_up(n, 256), 256, 0, stream>>>(wm, sorted_vals, nvalid, n, wing, mode, q,
                                                               min_periods, d_out);
    CNVK_LAUNCH_CHECK();
    return 0;
}

}  // namespace cnvk
