// Stable LSD radix sort (8-bit digits) and u32 exclusive scan for sm_100a.
//
// Sort layout per pass (n keys, tiles of 256 threads x 16 keys):
//   digit_hist     : per-tile digit histogram -> counts[digit * ntiles + tile]
//   exclusive scan : over the digit-major counter array = global base of every (digit, tile)
//   scatter        : warp-striped tile walk; __match_any_sync gives the in-warp stable rank,
//                    per-warp digit counters + a cross-warp scan give the in-tile stable rank.
// HBM traffic per pass: keys+vals read twice (hist + scatter) and written once.
#include <cooperative_groups.h>

#include <algorithm>

#include "prims.cuh"
#include "prof.cuh"

namespace cnvk {

static const int SORT_THREADS = 256;
static const int SORT_ITEMS = 16;
static const int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096
static const int SORT_WARPS = SORT_THREADS / 32;

// ---------------------------------------------------------------------------
// exclusive scan
// ---------------------------------------------------------------------------
static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 16;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* smem_warp, u32* block_total) {
    // inclusive warp scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = (lane < (int)(blockDim.x >> 5)) ? smem_warp[lane] : 0u;
        u32 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (int)(blockDim.x >> 5)) smem_warp[lane] = winc - w;  // exclusive warp base
        if (lane == 31) *block_total = winc;
    }
    __syncthreads();
    return smem_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const u32* __restrict__ in, u32* __restrict__ out, u32* __restrict__ tile_sums, long long n) {
    __shared__ u32 s_warp[32];
    __shared__ u32 s_total;
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u32 sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        long long i = base + k;
        v[k] = (i < n) ? in[i] : 0u;
        sum += v[k];
    }
    u32 ex = block_exclusive_scan(sum, s_warp, &s_total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        long long i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(u32* __restrict__ out, const u32* __restrict__ tile_prefix, long long n) {
    const u32 add = tile_prefix[blockIdx.x];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS) {
        long long i = base + k;
        if (i < n) out[i] += add;
    }
}

__global__ void scan_total_kernel(const u32* prefix_last, const u32* in_last, u32* total) {
    *total = *prefix_last + *in_last;
}

int exclusive_scan_u32(const u32* d_in, u32* d_out, long long n, u32* d_total, cudaStream_t stream) {
    if (n <= 0) {
        if (d_total) CNVK_CHECK(cudaMemsetAsync(d_total, 0, sizeof(u32), stream));
        return 0;
    }
    Scratch scratch(stream);
    const int ntiles = div_up(n, SCAN_TILE);
    u32* tile_sums = nullptr;
    CNVK_CHECK(scratch.alloc(&tile_sums, (size_t)ntiles + 1));
    // the input's last element is needed for the total when scanning in place
    u32* last_in = nullptr;
    if (d_total) {
        CNVK_CHECK(scratch.alloc(&last_in, 1));
        CNVK_CHECK(cudaMemcpyAsync(last_in, d_in + (n - 1), sizeof(u32), cudaMemcpyDeviceToDevice, stream));
    }
    scan_tile_kernel<<<ntiles, SCAN_THREADS, 0, stream>>>(d_in, d_out, tile_sums, n);
    CNVK_LAUNCH_CHECK();
    if (ntiles > 1) {
        int rc = exclusive_scan_u32(tile_sums, tile_sums, ntiles, nullptr, stream);
        if (rc) return rc;
        scan_add_kernel<<<ntiles, SCAN_THREADS, 0, stream>>>(d_out, tile_sums, n);
        CNVK_LAUNCH_CHECK();
    }
    if (d_total) {
        scan_total_kernel<<<1, 1, 0, stream>>>(d_out + (n - 1), last_in, d_total);
        CNVK_LAUNCH_CHECK();
    }
    return 0;
}

// ---------------------------------------------------------------------------
// radix sort
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
digit_hist_kernel(const u64* __restrict__ keys, long long n, int shift, u32* __restrict__ counts, int ntiles) {
    __shared__ u32 s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int k = threadIdx.x; k < SORT_TILE; k += SORT_THREADS) {
        long long i = base + k;
        if (i < n) atomicAdd(&s_hist[(u32)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(SORT_THREADS)
digit_scatter_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals, u64* __restrict__ keys_out,
                     u32* __restrict__ vals_out, long long n, int shift, const u32* __restrict__ bases,
                     int ntiles) {
    __shared__ u32 s_warp_hist[SORT_WARPS][256];
    __shared__ u32 s_base[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = lane; d < 256; d += 32) s_warp_hist[warp][d] = 0;
    s_base[threadIdx.x] = bases[(size_t)threadIdx.x * ntiles + blockIdx.x];
    __syncwarp();

    const long long wbase = (long long)blockIdx.x * SORT_TILE + (long long)warp * (SORT_ITEMS * 32);
    u64 k[SORT_ITEMS];
    u32 local[SORT_ITEMS];
    const u32 lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        long long i = wbase + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? keys[i] : 0ull;
        const u32 d = ok ? ((u32)(k[r] >> shift) & 255u) : 0xffffffffu;
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (ok && lane == leader) {
            old = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        local[r] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    // cross-warp exclusive prefix per digit (thread d owns digit d)
    {
        const int d = threadIdx.x;
        u32 run = s_base[d];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            u32 c = s_warp_hist[w][d];
            s_warp_hist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        long long i = wbase + r * 32 + lane;
        if (i < n) {
            const u32 d = (u32)(k[r] >> shift) & 255u;
            const u32 pos = s_warp_hist[warp][d] + local[r];
            keys_out[pos] = k[r];
            vals_out[pos] = vals[i];
        }
    }
}

// Whole sort in ONE cooperative launch for arrays of up to SORT_COOP_MAX_TILES tiles (the sizes of the
// exome path: 50 k - 1 M keys, where eight passes x five launches were pure launch latency).  Per pass:
// every CTA histograms its tiles, grid.sync, every CTA derives the global base of its (digit, tile)
// pairs straight from the counter table (digit totals + the counts of the tiles before it -- no
// separate scan kernel) and scatters, grid.sync.  Same stable ranking as the multi-launch path.
static const int SORT_COOP_MAX_TILES = 256;

__global__ void __launch_bounds__(SORT_THREADS)
radix_sort_coop_kernel(u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, long long n, int begin_bit, int end_bit,
                       u32* __restrict__ counts, int ntiles) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 s_warp_hist[SORT_WARPS][256];
    __shared__ u32 s_base[256];
    __shared__ u32 s_scan[32];
    __shared__ u32 s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    u64* kin = keys_a;
    u64* kout = keys_b;
    u32* vin = vals_a;
    u32* vout = vals_b;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        // ---- histogram of every tile this CTA owns
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            s_base[threadIdx.x] = 0;
            __syncthreads();
            const long long base = (long long)tile * SORT_TILE;
#pragma unroll 4
            for (int k = threadIdx.x; k < SORT_TILE; k += SORT_THREADS) {
                const long long i = base + k;
                if (i < n) atomicAdd(&s_base[(u32)(kin[i] >> shift) & 255u], 1u);
            }
            __syncthreads();
            counts[(size_t)threadIdx.x * ntiles + tile] = s_base[threadIdx.x];
            __syncthreads();
        }
        grid.sync();
        // ---- scatter
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int d = lane; d < 256; d += 32) s_warp_hist[warp][d] = 0;
            {
                // thread d: total of digit d over all tiles and the part that lies in earlier tiles
                const u32* row = counts + (size_t)threadIdx.x * ntiles;
                u32 before = 0, total = 0;
                for (int t = 0; t < ntiles; ++t) {
                    const u32 c = row[t];
                    if (t < tile) before += c;
                    total += c;
                }
                const u32 ex = block_exclusive_scan(total, s_scan, &s_total);   // digits before d
                s_base[threadIdx.x] = ex + before;
            }
            __syncthreads();
            const long long wbase = (long long)tile * SORT_TILE + (long long)warp * (SORT_ITEMS * 32);
            u64 k[SORT_ITEMS];
            u32 local[SORT_ITEMS];
#pragma unroll
            for (int r = 0; r < SORT_ITEMS; ++r) {
                const long long i = wbase + r * 32 + lane;
                const bool ok = i < n;
                k[r] = ok ? kin[i] : 0ull;
                const u32 d = ok ? ((u32)(k[r] >> shift) & 255u) : 0xffffffffu;
                const u32 peers = __match_any_sync(0xffffffffu, d);
                const int leader = __ffs(peers) - 1;
                u32 old = 0;
                if (ok && lane == leader) {
                    old = s_warp_hist[warp][d];
                    s_warp_hist[warp][d] = old + __popc(peers);
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                local[r] = old + __popc(peers & lt_mask);
                __syncwarp();
            }
            __syncthreads();
            {
                const int d = threadIdx.x;
                u32 run = s_base[d];
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) {
                    const u32 c = s_warp_hist[w][d];
                    s_warp_hist[w][d] = run;
                    run += c;
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < SORT_ITEMS; ++r) {
                const long long i = wbase + r * 32 + lane;
                if (i < n) {
                    const u32 d = (u32)(k[r] >> shift) & 255u;
                    const u32 pos = s_warp_hist[warp][d] + local[r];
                    kout[pos] = k[r];
                    vout[pos] = vin[i];
                }
            }
            __syncthreads();
        }
        grid.sync();
        u64* tk = kin; kin = kout; kout = tk;
        u32* tv = vin; vin = vout; vout = tv;
    }
}

static int coop_grid_limit() {
    static thread_local int cached = -1;
    if (cached < 0) {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radix_sort_coop_kernel, SORT_THREADS, 0);
        cached = coop ? sms * per_sm : 0;
    }
    return cached;
}

int radix_sort_pairs(u64* keys, u64* keys_alt, u32* vals, u32* vals_alt, long long n, int begin_bit,
                     int end_bit, u64** out_keys, u32** out_vals, cudaStream_t stream) {
    *out_keys = keys;
    *out_vals = vals;
    if (n <= 1) return 0;
    if (n >= (1ll << 32)) {
        set_last_error("radix_sort_pairs: n=%lld exceeds 2^32", n);
        return -1;
    }
    ProfScope pscope(PC_SORT, stream);
    Scratch scratch(stream);
    const int ntiles = div_up(n, SORT_TILE);
    u32* counts = nullptr;
    CNVK_CHECK(scratch.alloc(&counts, (size_t)256 * ntiles));
    u64* kin = keys;
    u64* kout = keys_alt;
    u32* vin = vals;
    u32* vout = vals_alt;
    const int coop_limit = ntiles <= SORT_COOP_MAX_TILES ? coop_grid_limit() : 0;
    if (coop_limit > 0) {
        int grid = std::min(ntiles, coop_limit);
        long long n_ = n;
        int bb = begin_bit, eb = end_bit, nt = ntiles;
        void* args[] = {&kin, &kout, &vin, &vout, &n_, &bb, &eb, &counts, &nt};
        CNVK_CHECK(cudaLaunchCooperativeKernel((const void*)radix_sort_coop_kernel, dim3(grid), dim3(SORT_THREADS), args,
                                               0, stream));
        CNVK_LAUNCH_CHECK();
        const int passes = (end_bit - begin_bit + 7) / 8;
        if (passes & 1) { *out_keys = kout; *out_vals = vout; } else { *out_keys = kin; *out_vals = vin; }
        return 0;
    }
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        digit_hist_kernel<<<ntiles, SORT_THREADS, 0, stream>>>(kin, n, shift, counts, ntiles);
        CNVK_LAUNCH_CHECK();
        int rc = exclusive_scan_u32(counts, counts, (long long)256 * ntiles, nullptr, stream);
        if (rc) return rc;
        digit_scatter_kernel<<<ntiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, counts, ntiles);
        CNVK_LAUNCH_CHECK();
        u64* tk = kin; kin = kout; kout = tk;
        u32* tv = vin; vin = vout; vout = tv;
    }
    *out_keys = kin;
    *out_vals = vin;
    return 0;
}

__global__ void make_sort_keys_kernel(const double* __restrict__ x, long long n, u64* __restrict__ keys,
                                      u32* __restrict__ idx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = f64_to_ordered(x[i]);
        idx[i] = (u32)i;
    }
}

int make_sort_keys(const double* d_x, long long n, u64* d_keys, u32* d_idx, cudaStream_t stream) {
    if (n <= 0) return 0;
    make_sort_keys_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_x, n, d_keys, d_idx);
    CNVK_LAUNCH_CHECK();
    return 0;
}

__global__ void keys_to_values_kernel(const u64* __restrict__ keys, long long n, double* __restrict__ out,
                                      u32* __restrict__ nvalid) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 k = keys[i];
    const bool isnan_ = (k == ~0ull);
    const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    out[i] = isnan_ ? __longlong_as_double(0x7ff8000000000000ll) : __longlong_as_double((long long)b);
    const bool prev_nan = (i > 0) && (keys[i - 1] == ~0ull);
    if (isnan_ && !prev_nan) *nvalid = (u32)i;
}

__global__ void set_u32_kernel(u32* p, u32 v) { *p = v; }

// ascending values with NaNs last; *d_nvalid = number of non-NaN values
int sort_values_f64(const double* d_x, long long n, u64* k0, u64* k1, u32* v0, u32* v1, double* d_sorted,
                    u32* d_nvalid, cudaStream_t stream) {
    set_u32_kernel<<<1, 1, 0, stream>>>(d_nvalid, (u32)(n > 0 ? n : 0));
    CNVK_LAUNCH_CHECK();
    if (n <= 0) return 0;
    int rc = make_sort_keys(d_x, n, k0, v0, stream);
    if (rc) return rc;
    u64* ok;
    u32* ov;
    rc = radix_sort_pairs(k0, k1, v0, v1, n, 0, 64, &ok, &ov, stream);
    if (rc) return rc;
    keys_to_values_kernel<<<div_up(n, 256), 256, 0, stream>>>(ok, n, d_sorted, d_nvalid);
    CNVK_LAUNCH_CHECK();
    return 0;
}

int argsort_f64(const double* d_x, long long n, u32* d_order, cudaStream_t stream) {
    if (n <= 0) return 0;
    Scratch scratch(stream);
    u64 *k0, *k1;
    u32 *v0, *v1;
    CNVK_CHECK(scratch.alloc(&k0, (size_t)n));
    CNVK_CHECK(scratch.alloc(&k1, (size_t)n));
    CNVK_CHECK(scratch.alloc(&v0, (size_t)n));
    CNVK_CHECK(scratch.alloc(&v1, (size_t)n));
    make_sort_keys_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_x, n, k0, v0);
    CNVK_LAUNCH_CHECK();
    u64* ok;
    u32* ov;
    int rc = radix_sort_pairs(k0, k1, v0, v1, n, 0, 64, &ok, &ov, stream);
    if (rc) return rc;
    CNVK_CHECK(cudaMemcpyAsync(d_order, ov, (size_t)n * sizeof(u32), cudaMemcpyDeviceToDevice, stream));
    return 0;
}

}  // namespace cnvk
