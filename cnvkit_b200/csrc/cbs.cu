// Weighted circular binary segmentation (CBS) on B200 -- replaces the Rscript/DNAcopy
// subprocess of cnvlib/segmentation/__init__.py:292-322 (R script cnvlib/segmentation/cbs.py:1-78,
// i.e. DNAcopy::segment(CNA(log2,...), weights=w, alpha=threshold) with DNAcopy defaults).
//
// Structure (level-synchronous over ALL arms of ALL samples in the batch):
//   prep   : one CTA (1024 threads) per open interval ("node"): sum w, sum x*w, sum w*xc^2 as
//            double-double reductions (R-level sum() calls), weighted centring, and the prefix sums
//            S_i = sum w xc, cw_i = cumsum(w)/sqrt(sum w) as a blocked FP64 scan with one fixed
//            association (thread-sequential x8, Kogge-Stone per warp, warp totals and tile carry
//            left to right) that the CPU oracle replays bit for bit; per-256-point block min/max
//            (+ arg) of S for pruning, getmncwt's delta.
//   seed   : every pair of 256-point blocks proposes the arcs between its extreme prefix points.
//   band   : arcs whose end points lie in the same or adjacent 256-point blocks, pruned on 32x32
//            sub-blocks out of shared memory (they cannot be pruned at 256-point granularity).
//   bound  : upper bound (block min/max of S, extreme c) of every remaining 256x256 sub-tile
//            against the node's running best; survivors go to a device-wide tile queue -- the
//            GPU form of DNAcopy's block-pruned wtmaxo.
//   scan   : persistent CTAs pop 1024x256 tiles; every thread keeps 4 arc starts in registers and
//            streams 256 arc ends from shared memory; BSS_ij = (S_j-S_i)^2 / (c (W-c)),
//            c = cw_j-cw_i, is compared division-free against the running best (exact IEEE
//            division only on the rare update).  Warp-shuffle argmax, then one lock-protected
//            update of the node's (q,i,j), ties -> smallest (i,j).
//   decide : one warp per node: T^2, the 0.1 / 7.0 shortcuts, Siegmund tail approximation
//            (hybrid, n > nmin), and the number of exceedances tolerated.
//   perms  : reference distribution for undecided nodes, evaluated in SM-resident tiles with
//            counter-based permutations (cbs_rng.cuh): hybrid = circular arcs of width <= kmax
//            (n > nmin), each arc dropped after add + multiply + compare against a per-node, per-width
//            limit (the weights are not permuted, so the smallest denominator is a constant of the
//            node); exact = all arcs (n <= nmin); DNAcopy's sequential stopping boundary is replayed
//            on the exceedance flags chunk by chunk.
//   flank  : ternary-split confirmation (wtpermp) for arcs interior to the node.
// Host code only moves a few bytes of per-node decisions per level and builds the next frontier.
#include <algorithm>
#include <cmath>
#include <vector>

#include "cbs_rng.cuh"
#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

static const int CBS_BLK = 256;           // points per bound block / j-block
static const int CBS_IW = 4;              // arc starts per thread
static const int CBS_IBLK = CBS_BLK * CBS_IW;
static const size_t PREP_SMEM = (size_t)2 * (2048 + 256) * sizeof(double);   // two padded 2048-term tiles
static const int HYB_TP = 2048;           // arc starts per CTA in the hybrid permutation kernel
static const int HYB_KMAX = 64;           // max supported kmax
static const int EXACT_NMAX = 256;        // max supported nmin

enum { ST_FINAL = 0, ST_SPLIT = 1, ST_PERM_HYBRID = 2, ST_PERM_EXACT = 3 };

struct Node {
    int lo, hi, base, flat;
    double avg, tss, rsw, total;
    double best_q;
    int best_i, best_j;
    int lock, state, nrejc, blk_off;
    double ostat, ostat1;
    double delta;  // getmncwt: lightest circular arc of kmax+1 probes / total weight (n > nmin only)
    double tailp;  // DNAcopy tailp of the node when the series was summed, else -1 (diagnostics)
};

struct CbsDiag {
    std::vector<Node> nodes;       // after the decision kernels of the (single) level
    std::vector<int> nrej;         // permutations whose statistic reached the observed one, per node
    std::vector<double> flank_p;   // [node][2] flank p-values with the permutations forced, -1 = no test
};

struct Decision {
    int state, nrejc, best_i, best_j;
    double ostat;
};

struct FlankTest {
    int node, off, n1, n2, tid, m1, s0, need_perm;
    double rm1, xbar, ostat, pvalue;
};

__device__ __forceinline__ double pt(const double* a, int lo, int i) {
    return i ? a[lo + i - 1] : 0.0;
}

// ------------------------------------------------------------------------------------ prep
// Sums of R's changepoints() for weighted data, and how they are evaluated here:
//   * sum(w), sum(x*w), sum(w*xc^2) are R-level sum() calls -- R accumulates them in 80-bit long
//     double, i.e. they are (to within an ulp) the correctly rounded exact sums.  They are
//     double-double (compensated) reductions: the rounded result does not depend on the order.
//   * the prefix sums S_i = sum_{l<=i} xc_l w_l (DNAcopy's partial sums) and cumsum(w) are a
//     BLOCKED FP64 prefix sum with one fixed, documented association (oracle/cbs_oracle.c
//     `blocked_prefix` replays it operation by operation, so GPU and oracle agree bit for bit):
//       tile   = 2048 consecutive terms; thread j of 256 owns terms 8j..8j+7
//       p[j,e] = left-to-right running sum of the thread's terms          (8 dependent adds)
//       A[j]   = Kogge-Stone inclusive scan of the thread totals p[j,7] inside each warp
//                (steps 1,2,4,8,16: A[l] += A[l-o] for lanes l >= o), E[j] = A[j-1] (0 for lane 0)
//       X[w]   = W[0] + W[1] + ... + W[w-1] left to right over the warp totals W[w] = A[32w+31]
//       local  = (X[w] + E[j]) + p[j,e];   T = local of the tile's last slot (missing terms are 0.0)
//       C[k]   = T[0] + T[1] + ... + T[k-1] left to right over the node's tiles (C[0] = 0)
//       out    = C[k] + local
//     It differs from a plain left-to-right sum by rounding only (its error growth is smaller).
// A node is cut into tiles of 2048 prefix POINTS (point i = S_i, i = 0..n; point i > 0 is term i-1)
// and every tile is one CTA, so the largest arm no longer sits on a single SM:
//   sums   (tile)  : min / max of x, double-double partials of sum w and sum x*w
//   scan   (tile)  : node totals from the partials -> avg; centred data, tile-local prefix scan, tile
//                    totals, double-double partial of sum w*xc^2
//   finish (tile)  : carry C[k] from the tile totals, final S and cw, per-256-point block min / max
//                    (+ arg, + cw at the arg), partial minimum for getmncwt
//   node   (node)  : tss, total, delta, reset of the running best
static const int PREP_THREADS = 256;
static const int PREP_E = 8;                            // terms per thread
static const int PREP_TILE = PREP_THREADS * PREP_E;     // 2048 terms / points per tile
static const int PREP_HALO = HYB_KMAX + 2;              // getmncwt looks kmax+1 points ahead
__host__ __device__ __forceinline__ int prep_pad(int k) { return k + (k >> 3); }  // bank-conflict padding

// double-double accumulator: hi + lo carries the running sum to ~2^-104 relative
struct DD {
    double hi, lo;
};
__device__ __forceinline__ void dd_add(DD& a, double p) {          // a += p  (Knuth two-sum)
    const double s = __dadd_rn(a.hi, p);
    const double bb = __dsub_rn(s, a.hi);
    const double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(p, bb));
    a.hi = s;
    a.lo = __dadd_rn(a.lo, e);
}
__device__ __forceinline__ void dd_merge(DD& a, double bhi, double blo) {  // a += (bhi, blo)
    dd_add(a, bhi);
    a.lo = __dadd_rn(a.lo, blo);
}
__device__ __forceinline__ DD dd_warp_sum(DD a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double h = __shfl_xor_sync(0xffffffffu, a.hi, o);
        const double l = __shfl_xor_sync(0xffffffffu, a.lo, o);
        dd_merge(a, h, l);
    }
    return a;
}
// CTA-wide double-double sum (not rounded); result valid in thread 0
__device__ __forceinline__ DD dd_block_sum(DD a, double* s_hi, double* s_lo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = dd_warp_sum(a);
    if (lane == 0) { s_hi[warp] = a.hi; s_lo[warp] = a.lo; }
    __syncthreads();
    DD t = {0.0, 0.0};
    if (threadIdx.x == 0) {
        t.hi = s_hi[0]; t.lo = s_lo[0];
        for (int k = 1; k < PREP_THREADS / 32; ++k) dd_merge(t, s_hi[k], s_lo[k]);
    }
    __syncthreads();
    return t;
}

struct TileSums {      // per tile, written by the sums kernel
    double mn, mx, sw_hi, sw_lo, swx_hi, swx_lo;
};
struct TileScan {      // per tile, written by the scan kernel
    double Tw, Ts, tss_hi, tss_lo, dmin, pad;
};

// tile -> (node, tile index inside the node); ptile_off[a] = first tile of node a
__device__ __forceinline__ void decode_ptile(const int* __restrict__ ptile_off, int nnodes, int tile, int& a, int& k) {
    int lo_ = 0, hi_ = nnodes;
    while (hi_ - lo_ > 1) {
        const int m = (lo_ + hi_) >> 1;
        if (ptile_off[m] <= tile) lo_ = m; else hi_ = m;
    }
    a = lo_;
    k = tile - ptile_off[a];
}

__global__ void __launch_bounds__(PREP_THREADS)
cbs_prep_sums_kernel(const Node* __restrict__ nodes, const int* __restrict__ ptile_off, int nnodes,
                     const double* __restrict__ x, const double* __restrict__ w, TileSums* __restrict__ part) {
    __shared__ double s_a[32], s_b[32];
    int a, k;
    decode_ptile(ptile_off, nnodes, blockIdx.x, a, k);
    const int lo = nodes[a].lo, n = nodes[a].hi - nodes[a].lo;
    const int t0 = k * PREP_TILE, m = max(0, min(PREP_TILE, n - t0));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double mn = INFINITY, mx = -INFINITY;
    DD a_w = {0.0, 0.0}, a_wx = {0.0, 0.0};
    for (int r = tid; r < m; r += PREP_THREADS) {
        const double v = x[lo + t0 + r], wv = w[lo + t0 + r];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        dd_add(a_w, wv);
        dd_add(a_wx, __dmul_rn(v, wv));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_a[warp] = mn; s_b[warp] = mx; }
    __syncthreads();
    if (tid == 0)
        for (int q = 1; q < PREP_THREADS / 32; ++q) { mn = fmin(mn, s_a[q]); mx = fmax(mx, s_b[q]); }
    __syncthreads();
    const DD sw = dd_block_sum(a_w, s_a, s_b);
    const DD swx = dd_block_sum(a_wx, s_a, s_b);
    if (tid == 0) {
        TileSums o;
        o.mn = mn; o.mx = mx; o.sw_hi = sw.hi; o.sw_lo = sw.lo; o.swx_hi = swx.hi; o.swx_lo = swx.lo;
        part[blockIdx.x] = o;
    }
}

// node totals from the tile partials, in tile order (every CTA of the node computes the same values)
__device__ __forceinline__ void node_totals(const TileSums* __restrict__ part, int ntiles, double& range,
                                            double& sw, double& swx) {
    double mn = INFINITY, mx = -INFINITY;
    DD a = {0.0, 0.0}, b = {0.0, 0.0};
    for (int t = 0; t < ntiles; ++t) {
        const TileSums p = part[t];
        mn = fmin(mn, p.mn);
        mx = fmax(mx, p.mx);
        if (t == 0) { a.hi = p.sw_hi; a.lo = p.sw_lo; b.hi = p.swx_hi; b.lo = p.swx_lo; }
        else { dd_merge(a, p.sw_hi, p.sw_lo); dd_merge(b, p.swx_hi, p.swx_lo); }
    }
    range = mx - mn;
    sw = __dadd_rn(a.hi, a.lo);
    swx = __dadd_rn(b.hi, b.lo);
}

__global__ void __launch_bounds__(PREP_THREADS, 4)
cbs_prep_scan_kernel(Node* __restrict__ nodes, const int* __restrict__ ptile_off, int nnodes,
                     const double* __restrict__ x, const double* __restrict__ w, const TileSums* __restrict__ part1,
                     double* __restrict__ xc, double* __restrict__ y, double* __restrict__ wn,
                     double* __restrict__ Sl, double* __restrict__ cwl, TileScan* __restrict__ part3) {
    extern __shared__ __align__(16) double s_dyn[];      // [2 arrays][padded tile]
    __shared__ double s_a[32], s_b[32];
    __shared__ double s_scalar[4];
    __shared__ double s_wtot[2][PREP_THREADS / 32];
    int a, k;
    decode_ptile(ptile_off, nnodes, blockIdx.x, a, k);
    const int lo = nodes[a].lo, n = nodes[a].hi - nodes[a].lo;
    const int ntiles = ptile_off[a + 1] - ptile_off[a];
    const int t0 = k * PREP_TILE, m = max(0, min(PREP_TILE, n - t0));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* s_w = s_dyn;
    double* s_s = s_dyn + (PREP_TILE + PREP_TILE / 8);
    if (tid == 0) {
        double range, sw, swx;
        node_totals(part1 + ptile_off[a], ntiles, range, sw, swx);
        s_scalar[0] = range; s_scalar[1] = sw; s_scalar[2] = swx;
    }
    __syncthreads();
    // range check: isTRUE(all.equal(diff(range(x)), 0)) -> final (changepoints R code)
    if (s_scalar[0] <= 1.4901161193847656e-08) {
        if (tid == 0 && k == 0) { nodes[a].flat = 1; nodes[a].best_q = -1.0; }
        return;
    }
    const double sw = s_scalar[1], swx = s_scalar[2];
    const double avg = __ddiv_rn(swx, sw), rsw = sqrt(sw);
    if (tid == 0 && k == 0) { nodes[a].flat = 0; nodes[a].avg = avg; nodes[a].rsw = rsw; }

    DD a_tss = {0.0, 0.0};
    for (int r = tid; r < PREP_TILE; r += PREP_THREADS) {              // coalesced staging
        double wv = 0.0, sv = 0.0;
        if (r < m) {
            const int g = lo + t0 + r;
            wv = w[g];
            const double cc = __dsub_rn(x[g], avg);
            xc[g] = cc;
            y[g] = __dmul_rn(cc, sqrt(wv));
            wn[g] = __ddiv_rn(wv, rsw);
            sv = __dmul_rn(cc, wv);
            dd_add(a_tss, __dmul_rn(wv, __dmul_rn(cc, cc)));
        }
        s_w[prep_pad(r)] = wv;
        s_s[prep_pad(r)] = sv;
    }
    __syncthreads();
    double pw[PREP_E], ps[PREP_E];
    {
        double rw = 0.0, rs = 0.0;
        const int b0 = prep_pad(PREP_E * tid);                         // 8 consecutive padded slots
#pragma unroll
        for (int e = 0; e < PREP_E; ++e) {
            const double vw = s_w[b0 + e], vs = s_s[b0 + e];
            rw = e ? __dadd_rn(rw, vw) : vw;
            rs = e ? __dadd_rn(rs, vs) : vs;
            pw[e] = rw;
            ps[e] = rs;
        }
    }
    double Aw = pw[PREP_E - 1], As = ps[PREP_E - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {                                  // Kogge-Stone inclusive warp scan
        const double tw = __shfl_up_sync(0xffffffffu, Aw, o), ts = __shfl_up_sync(0xffffffffu, As, o);
        if (lane >= o) { Aw = __dadd_rn(Aw, tw); As = __dadd_rn(As, ts); }
    }
    double Ew = __shfl_up_sync(0xffffffffu, Aw, 1), Es = __shfl_up_sync(0xffffffffu, As, 1);
    if (lane == 0) { Ew = 0.0; Es = 0.0; }
    if (lane == 31) { s_wtot[0][warp] = Aw; s_wtot[1][warp] = As; }
    __syncthreads();
    double Xw = 0.0, Xs = 0.0;
    for (int q = 0; q < warp; ++q) {
        Xw = q ? __dadd_rn(Xw, s_wtot[0][q]) : s_wtot[0][0];
        Xs = q ? __dadd_rn(Xs, s_wtot[1][q]) : s_wtot[1][0];
    }
    const double bw_ = __dadd_rn(Xw, Ew), bs_ = __dadd_rn(Xs, Es);
    {
        const int b0 = prep_pad(PREP_E * tid);
#pragma unroll
        for (int e = 0; e < PREP_E; ++e) {
            s_w[b0 + e] = __dadd_rn(bw_, pw[e]);
            s_s[b0 + e] = __dadd_rn(bs_, ps[e]);
        }
    }
    __syncthreads();
    for (int r = tid; r < m; r += PREP_THREADS) {                       // coalesced write-back of the locals
        const int g = lo + t0 + r;
        cwl[g] = s_w[prep_pad(r)];
        Sl[g] = s_s[prep_pad(r)];
    }
    const double Tw = s_w[prep_pad(PREP_TILE - 1)], Ts = s_s[prep_pad(PREP_TILE - 1)];
    const DD tss = dd_block_sum(a_tss, s_a, s_b);
    if (tid == 0) {
        TileScan o;
        o.Tw = Tw; o.Ts = Ts; o.tss_hi = tss.hi; o.tss_lo = tss.lo; o.dmin = INFINITY; o.pad = 0.0;
        part3[blockIdx.x] = o;
    }
}

__global__ void __launch_bounds__(PREP_THREADS)
cbs_prep_finish_kernel(const Node* __restrict__ nodes, const int* __restrict__ ptile_off, int nnodes,
                       const double* __restrict__ Sl, const double* __restrict__ cwl, TileScan* __restrict__ part3,
                       double* __restrict__ S, double* __restrict__ cw, double* __restrict__ blk_min,
                       double* __restrict__ blk_max, int* __restrict__ blk_imin, int* __restrict__ blk_imax,
                       double* __restrict__ blk_cmin, double* __restrict__ blk_cmax, int hk, int nmin) {
    __shared__ double pS[PREP_TILE];
    __shared__ double pC[PREP_TILE + PREP_HALO];
    __shared__ double s_carry[4];
    __shared__ double s_red[8];
    int a, k;
    decode_ptile(ptile_off, nnodes, blockIdx.x, a, k);
    const Node* nd = &nodes[a];
    if (nd->flat) return;
    const int lo = nd->lo, n = nd->hi - nd->lo, boff = nd->blk_off;
    const double rsw = nd->rsw;
    const int t0 = k * PREP_TILE, m = max(0, min(PREP_TILE, n - t0));
    const int npts = min(PREP_TILE, n + 1 - t0);                         // points t0 .. t0+npts-1 (<= n)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        // C[k] = T[0] + ... + T[k-1], left to right; C[k+1] for the halo
        const TileScan* ts = part3 + ptile_off[a];
        double Cw = 0.0, Cs = 0.0;
        for (int t = 0; t < k; ++t) {
            Cw = t ? __dadd_rn(Cw, ts[t].Tw) : ts[0].Tw;
            Cs = t ? __dadd_rn(Cs, ts[t].Ts) : ts[0].Ts;
        }
        s_carry[0] = Cw; s_carry[1] = Cs;
        s_carry[2] = k ? __dadd_rn(Cw, ts[k].Tw) : ts[0].Tw;           // C[k+1]
    }
    __syncthreads();
    const double Cw = s_carry[0], Cs = s_carry[1], Cw1 = s_carry[2];
    if (tid == 0) {                                                      // the tile's first point is the carry
        pS[0] = k ? Cs : 0.0;
        pC[0] = k ? __ddiv_rn(Cw, rsw) : 0.0;
    }
    for (int r = tid; r < m; r += PREP_THREADS) {
        const int g = lo + t0 + r;
        // tile 0 has C = 0: 0 + local = local exactly, as the oracle computes it
        const double vS = k ? __dadd_rn(Cs, Sl[g]) : Sl[g];
        const double vC = __ddiv_rn(k ? __dadd_rn(Cw, cwl[g]) : cwl[g], rsw);
        S[g] = vS;
        cw[g] = vC;
        if (r + 1 < PREP_TILE) pS[r + 1] = vS;
        pC[r + 1] = vC;                                                  // slot PREP_TILE = first halo point
    }
    // halo for getmncwt: cw at points t0+PREP_TILE+1 .. (terms of the next tile, carried by C[k+1])
    const bool want_delta = hk > 0 && n > nmin;
    const int jw = hk + 1;
    if (want_delta) {
        for (int h = tid; h < jw; h += PREP_THREADS) {
            const int p = t0 + PREP_TILE + 1 + h;                        // point index; its term is p-1
            if (p <= n) pC[PREP_TILE + 1 + h] = __ddiv_rn(__dadd_rn(Cw1, cwl[lo + p - 1]), rsw);
        }
    }
    __syncthreads();
    // per-256-point block min / max of S with arg (ties -> smallest index) and cw at the arg
    if (warp * CBS_BLK < npts) {
        const int b = (t0 / CBS_BLK) + warp;
        double bmn = INFINITY, bmx = -INFINITY;
        int bimn = 0, bimx = 0;
        for (int q = lane; q < CBS_BLK; q += 32) {
            const int lp = warp * CBS_BLK + q;
            if (lp < npts) {
                const double v = pS[lp];
                if (v < bmn) { bmn = v; bimn = lp; }
                if (v > bmx) { bmx = v; bimx = lp; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double omn = __shfl_xor_sync(0xffffffffu, bmn, o);
            const int oimn = __shfl_xor_sync(0xffffffffu, bimn, o);
            const double omx = __shfl_xor_sync(0xffffffffu, bmx, o);
            const int oimx = __shfl_xor_sync(0xffffffffu, bimx, o);
            if (omn < bmn || (omn == bmn && oimn < bimn)) { bmn = omn; bimn = oimn; }
            if (omx > bmx || (omx == bmx && oimx < bimx)) { bmx = omx; bimx = oimx; }
        }
        if (lane == 0) {
            blk_min[boff + b] = bmn;
            blk_max[boff + b] = bmx;
            blk_imin[boff + b] = t0 + bimn;
            blk_imax[boff + b] = t0 + bimx;
            blk_cmin[boff + b] = pC[bimn];     // cw at the extreme points: the seed kernel then needs
            blk_cmax[boff + b] = pC[bimx];     // no scattered load
        }
    }
    // DNAcopy getmncwt, linear arcs starting in this tile: min over i of cw(i+jw) - cw(i)
    if (want_delta) {
        double dmin = INFINITY;
        for (int q = tid; q < npts; q += PREP_THREADS) {
            const int i = t0 + q;
            if (i + jw <= n) dmin = fmin(dmin, __dsub_rn(pC[q + jw], pC[q]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        if (lane == 0) s_red[warp] = dmin;
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < PREP_THREADS / 32; ++q) dmin = fmin(dmin, s_red[q]);
            part3[blockIdx.x].dmin = dmin;
        }
    }
}

// one warp per node: tss, total, delta (linear partials + the wrap-around arcs), reset of the best
__global__ void __launch_bounds__(128)
cbs_prep_node_kernel(Node* __restrict__ nodes, const int* __restrict__ ptile_off, int nnodes,
                     const TileScan* __restrict__ part3, const double* __restrict__ cw,
                     const double* __restrict__ blk_min, const double* __restrict__ blk_max,
                     const int* __restrict__ blk_imin, const int* __restrict__ blk_imax,
                     const double* __restrict__ blk_cmin, const double* __restrict__ blk_cmax, int al0, int hk,
                     int nmin) {
    const int a = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (a >= nnodes) return;
    Node* nd = &nodes[a];
    if (nd->flat) return;
    const int lo = nd->lo, n = nd->hi - nd->lo;
    const int ntiles = ptile_off[a + 1] - ptile_off[a];
    const TileScan* ts = part3 + ptile_off[a];
    const double total = cw[lo + n - 1];
    double delta = 0.0;
    if (hk > 0 && n > nmin) {
        const int jw = hk + 1;
        double dmin = INFINITY;
        for (int t = lane; t < ntiles; t += 32) dmin = fmin(dmin, ts[t].dmin);
        for (int i = 1 + lane; i <= jw; i += 32)
            dmin = fmin(dmin, __dadd_rn(__dsub_rn(total, pt(cw, lo, n - jw + i)), pt(cw, lo, i)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        delta = __ddiv_rn(dmin, total);
    }
    // seed arc: from the global argmin of S to its global argmax (a real arc, scored exactly) -- gives
    // the seed kernel's CTAs a running best to filter against before they take the node's lock
    const int nb = (n + 1 + CBS_BLK - 1) / CBS_BLK, boff = nd->blk_off;
    double gmn = INFINITY, gmx = -INFINITY, cmn = 0.0, cmx = 0.0;
    int imn = 0x7fffffff, imx = 0x7fffffff;
    for (int b = lane; b < nb; b += 32) {
        const double vmn = blk_min[boff + b], vmx = blk_max[boff + b];
        if (vmn < gmn) { gmn = vmn; imn = blk_imin[boff + b]; cmn = blk_cmin[boff + b]; }
        if (vmx > gmx) { gmx = vmx; imx = blk_imax[boff + b]; cmx = blk_cmax[boff + b]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double omn = __shfl_xor_sync(0xffffffffu, gmn, o), ocmn = __shfl_xor_sync(0xffffffffu, cmn, o);
        const int oimn = __shfl_xor_sync(0xffffffffu, imn, o);
        const double omx = __shfl_xor_sync(0xffffffffu, gmx, o), ocmx = __shfl_xor_sync(0xffffffffu, cmx, o);
        const int oimx = __shfl_xor_sync(0xffffffffu, imx, o);
        if (omn < gmn || (omn == gmn && oimn < imn)) { gmn = omn; imn = oimn; cmn = ocmn; }
        if (omx > gmx || (omx == gmx && oimx < imx)) { gmx = omx; imx = oimx; cmx = ocmx; }
    }
    if (lane == 0) {
        DD t = {ts[0].tss_hi, ts[0].tss_lo};
        for (int q = 1; q < ntiles; ++q) dd_merge(t, ts[q].tss_hi, ts[q].tss_lo);
        nd->tss = __dadd_rn(t.hi, t.lo);
        nd->total = total;
        nd->delta = delta;
        nd->lock = 0;
        double q = -1.0;
        int qi = 0x7fffffff, qj = 0x7fffffff;
        const bool fwd = imn < imx;
        const int i = fwd ? imn : imx, j = fwd ? imx : imn;
        if (j - i >= al0 && j - i <= n - al0) {
            const double d = fwd ? __dsub_rn(gmx, gmn) : __dsub_rn(gmn, gmx);          // S_j - S_i
            const double c = fwd ? __dsub_rn(cmx, cmn) : __dsub_rn(cmn, cmx);          // cw_j - cw_i
            const double den = __dmul_rn(c, __dsub_rn(total, c));
            q = __ddiv_rn(__dmul_rn(d, d), den);
            qi = i; qj = j;
        }
        nd->best_q = q;
        nd->best_i = qi;
        nd->best_j = qj;
    }
}

// ------------------------------------------------------------------------------------ scan
struct Cand {
    double q, qlo;
    int i, j;
};

__device__ __forceinline__ bool cand_better(double q, int i, int j, double bq, int bi, int bj) {
    return q > bq || (q == bq && (i < bi || (i == bi && j < bj)));
}

__device__ __noinline__ void cand_update(Cand& b, double num, double den, int i, int j) {
    const double q = __ddiv_rn(num, den);
    if (cand_better(q, i, j, b.q, b.i, b.j)) {
        b.q = q;
        b.i = i;
        b.j = j;
        b.qlo = q * (1.0 - 1e-12);
    }
}

// inlined form for kernels that scan one arc per iteration: the candidate stays in registers
__device__ __forceinline__ void cand_update_inl(Cand& b, double num, double den, int i, int j) {
    const double q = __ddiv_rn(num, den);
    if (cand_better(q, i, j, b.q, b.i, b.j)) {
        b.q = q;
        b.i = i;
        b.j = j;
        b.qlo = q * (1.0 - 1e-12);
    }
}

template <int NA, bool MASKED>
__device__ __forceinline__ void scan_subtiles(Cand& best, const double2* __restrict__ sj, const double* Si,
                                              const double* ci, const int* ii, int jbase, int jcount, double total,
                                              int n, int al0) {
#pragma unroll 2
    for (int jj = 0; jj < jcount; ++jj) {
        const double2 pj = sj[jj];
        const int j = jbase + jj;
#pragma unroll
        for (int r = 0; r < NA; ++r) {
            const double d = __dsub_rn(pj.x, Si[r]);
            const double c = __dsub_rn(pj.y, ci[r]);
            const double den = __dmul_rn(c, __dsub_rn(total, c));
            const double num = __dmul_rn(d, d);
            bool hit = num > __dmul_rn(best.qlo, den);
            if (MASKED) {
                const int kw = j - ii[r];
                hit = hit && (kw >= al0) && (kw <= n - al0) && (ii[r] <= n);
            }
            if (hit) cand_update(best, num, den, ii[r], j);
        }
    }
}

// The arc scan of one recursion level runs as three kernels over the same tile geometry
// (node, 1024-point start block ib = 4 bound blocks, 256-point end block jb):
//   seed  : every pair of bound blocks proposes the arcs between its extreme prefix points
//           (argmin S of one block, argmax S of the other) -- real arcs, scored exactly, so the
//           node's running best starts within a few percent of the maximum;
//   bound : the pruning test of every (start block, end block) pair against that best; survivors
//           are appended to a device-wide queue of 1024x256 tiles;
//   scan  : persistent CTAs pop tiles from the queue (each re-tests its bound against the best
//           found meanwhile) -- equal-sized work items, so no CTA is left walking a long list.
struct TileRef {
    int node, ib, jb;
    unsigned mask;
};

struct SeedRes {   // best seed candidate of one task
    double q;
    int i, j;
    long long pad;
};

// task -> (node a, start block ib, first end block jb0): a host-built table, one int4 per task
__device__ __forceinline__ void decode_task(const Node* __restrict__ nodes, const int4* __restrict__ task_tab,
                                            int task, int& a, int& ib, int& jb0, int& NJ) {
    const int4 t = task_tab[task];
    a = t.x;
    ib = t.y;
    jb0 = t.z;
    NJ = t.w;
}

// CTA-wide argmax of the threads' candidates (ties -> smallest (i,j)) and one lock-protected
// update of the node; only candidates found by this CTA carry i != INT_MAX.
__device__ __forceinline__ void cta_commit(Node* nd, double q, int qi, int qj, double* s_q, int* s_i, int* s_j) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!__syncthreads_or(qi != 0x7fffffff)) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oq = __shfl_xor_sync(0xffffffffu, q, o);
        const int oi = __shfl_xor_sync(0xffffffffu, qi, o);
        const int oj = __shfl_xor_sync(0xffffffffu, qj, o);
        if (cand_better(oq, oi, oj, q, qi, qj)) { q = oq; qi = oi; qj = oj; }
    }
    if (lane == 0) { s_q[warp] = q; s_i[warp] = qi; s_j[warp] = qj; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < 8; ++k)
            if (cand_better(s_q[k], s_i[k], s_j[k], q, qi, qj)) { q = s_q[k]; qi = s_i[k]; qj = s_j[k]; }
        // best_q only rises, so a candidate strictly below it can never win: skip the lock
        if (qi != 0x7fffffff && !(q < *((volatile double*)&nd->best_q))) {
            while (atomicCAS(&nd->lock, 0, 1) != 0) {}
            __threadfence();
            const double cq = *((volatile double*)&nd->best_q);
            const int cqi = *((volatile int*)&nd->best_i), cqj = *((volatile int*)&nd->best_j);
            if (cand_better(q, qi, qj, cq, cqi, cqj)) {
                *((volatile int*)&nd->best_i) = qi;
                *((volatile int*)&nd->best_j) = qj;
                *((volatile double*)&nd->best_q) = q;
            }
            __threadfence();
            atomicExch(&nd->lock, 0);
        }
    }
}

__global__ void __launch_bounds__(256)
cbs_seed_kernel(const Node* __restrict__ nodes, const int4* __restrict__ task_tab, SeedRes* __restrict__ res,
                const double* __restrict__ blk_min, const double* __restrict__ blk_max,
                const int* __restrict__ blk_imin, const int* __restrict__ blk_imax,
                const double* __restrict__ blk_cmin, const double* __restrict__ blk_cmax, int al0) {
    __shared__ double s_q[8];
    __shared__ int s_i[8], s_j[8];
    int a, ib, jb0, NJ;
    decode_task(nodes, task_tab, blockIdx.x, a, ib, jb0, NJ);
    const Node* nd = &nodes[a];
    if (nd->flat) { if (threadIdx.x == 0) { SeedRes r_; r_.q = -1.0; r_.i = 0x7fffffff; r_.j = 0x7fffffff; r_.pad = 0; res[blockIdx.x] = r_; } return; }
    const int n = nd->hi - nd->lo, boff = nd->blk_off;
    const double total = nd->total;
    double bq = -1.0;
    int bi_ = 0x7fffffff, bj_ = 0x7fffffff;
    const int jb = jb0 + threadIdx.x;
    if (jb < NJ) {
        // per block: (S, cw, index) at the argmin and at the argmax of S -- all coalesced loads
        const double jSmn = blk_min[boff + jb], jSmx = blk_max[boff + jb];
        const double jCmn = blk_cmin[boff + jb], jCmx = blk_cmax[boff + jb];
        const int jImn = blk_imin[boff + jb], jImx = blk_imax[boff + jb];
#pragma unroll
        for (int r = 0; r < CBS_IW; ++r) {
            const int bi = CBS_IW * ib + r;
            if (bi > jb || bi >= NJ) continue;
            const double iSmn = blk_min[boff + bi], iSmx = blk_max[boff + bi];
            const double iCmn = blk_cmin[boff + bi], iCmx = blk_cmax[boff + bi];
            const int iImn = blk_imin[boff + bi], iImx = blk_imax[boff + bi];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                // c = 0: from the start block's minimum to the end block's maximum; c = 1: max -> min
                int i = c ? iImx : iImn, j = c ? jImn : jImx;
                double Si = c ? iSmx : iSmn, Sj = c ? jSmn : jSmx, ci = c ? iCmx : iCmn, cj = c ? jCmn : jCmx;
                if (i > j) {                         // same block: order the two points
                    const int t_ = i; i = j; j = t_;
                    const double ts = Si; Si = Sj; Sj = ts;
                    const double tc = ci; ci = cj; cj = tc;
                }
                const int kw = j - i;
                if (kw < al0 || kw > n - al0) continue;
                const double d = __dsub_rn(Sj, Si);
                const double cc = __dsub_rn(cj, ci);
                const double den = __dmul_rn(cc, __dsub_rn(total, cc));
                const double num = __dmul_rn(d, d);
                if (!(num >= __dmul_rn(bq * (1.0 - 1e-12), den)) && bq >= 0.0) continue;
                const double q = __ddiv_rn(num, den);
                if (cand_better(q, i, j, bq, bi_, bj_)) { bq = q; bi_ = i; bj_ = j; }
            }
        }
    }
    // CTA argmax (ties -> smallest (i,j)); the per-task results are reduced per node by
    // cbs_seed_reduce_kernel -- no lock, no atomics
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oq = __shfl_xor_sync(0xffffffffu, bq, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi_, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj_, o);
        if (cand_better(oq, oi, oj, bq, bi_, bj_)) { bq = oq; bi_ = oi; bj_ = oj; }
    }
    if (lane == 0) { s_q[warp] = bq; s_i[warp] = bi_; s_j[warp] = bj_; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k)
            if (cand_better(s_q[k], s_i[k], s_j[k], bq, bi_, bj_)) { bq = s_q[k]; bi_ = s_i[k]; bj_ = s_j[k]; }
        SeedRes r_;
        r_.q = bq; r_.i = bi_; r_.j = bj_; r_.pad = 0;
        res[blockIdx.x] = r_;
    }
}

// one warp per node: best of the node's per-task seed candidates and of the arc prep_node left there
__global__ void __launch_bounds__(128)
cbs_seed_reduce_kernel(Node* __restrict__ nodes, int nnodes, const int* __restrict__ task_off,
                       const SeedRes* __restrict__ res) {
    const int a = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (a >= nnodes) return;
    Node* nd = &nodes[a];
    if (nd->flat) return;
    double q = -1.0;
    int qi = 0x7fffffff, qj = 0x7fffffff;
    for (int t = task_off[a] + lane; t < task_off[a + 1]; t += 32) {
        const SeedRes r_ = res[t];
        if (cand_better(r_.q, r_.i, r_.j, q, qi, qj)) { q = r_.q; qi = r_.i; qj = r_.j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oq = __shfl_xor_sync(0xffffffffu, q, o);
        const int oi = __shfl_xor_sync(0xffffffffu, qi, o);
        const int oj = __shfl_xor_sync(0xffffffffu, qj, o);
        if (cand_better(oq, oi, oj, q, qi, qj)) { q = oq; qi = oi; qj = oj; }
    }
    if (lane == 0 && qi != 0x7fffffff && cand_better(q, qi, qj, nd->best_q, nd->best_i, nd->best_j)) {
        nd->best_q = q;
        nd->best_i = qi;
        nd->best_j = qj;
    }
}


// upper bound test of one (start block bi, end block jb) pair, bi < jb
__device__ __forceinline__ bool tile_needed(const double* __restrict__ cw, const double* __restrict__ blk_min,
                                            const double* __restrict__ blk_max, int lo, int n, int boff, int bi,
                                            int jb, double total, double gq) {
    const int jbase = jb * CBS_BLK, jend = min(jbase + CBS_BLK - 1, n);
    const int ibase = bi * CBS_BLK, iend = min(ibase + CBS_BLK - 1, n);
    const double d1 = blk_max[boff + jb] - blk_min[boff + bi];
    const double d2 = blk_max[boff + bi] - blk_min[boff + jb];
    const double dm = fmax(fabs(d1), fabs(d2));
    const double cmin = pt(cw, lo, jbase) - pt(cw, lo, iend);
    const double cmax = pt(cw, lo, jend) - pt(cw, lo, ibase);
    const double fmin_ = fmin(cmin * (total - cmin), cmax * (total - cmax));
    if (fmin_ > 0.0) return !((dm * dm) / fmin_ * (1.0 + 1e-9) < gq);
    return true;
}

// Narrow arcs: every (i, j) whose 256-point blocks are equal or adjacent.  At 256-point granularity
// these tiles can never be pruned (the lightest arc of the pair is a single bin), and they used to
// be most of the scan; here one CTA owns one start block, keeps its 512 candidate end points in
// shared memory and repeats the bound test on 32x32 sub-blocks, warp-uniformly (warp <-> 32 starts).
static const int BAND_SUB = 32;
static const int BAND_PTS = 2 * CBS_BLK;

template <bool PRUNE>
__global__ void __launch_bounds__(256)
cbs_band_kernel(Node* __restrict__ nodes, int nnodes,
                const double* __restrict__ S, const double* __restrict__ cw, int al0,
                unsigned long long* __restrict__ stats) {
    __shared__ double2 sp[BAND_PTS];
    __shared__ double s_mn[BAND_PTS / BAND_SUB], s_mx[BAND_PTS / BAND_SUB];
    __shared__ double s_q[8];
    __shared__ int s_i[8], s_j[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int a = 0, hi_ = nnodes;
    while (hi_ - a > 1) {
        const int m = (a + hi_) >> 1;
        if (nodes[m].blk_off <= (int)blockIdx.x) a = m; else hi_ = m;
    }
    Node* nd = &nodes[a];
    if (nd->flat) return;
    const int lo = nd->lo, n = nd->hi - nd->lo;
    const int bi = blockIdx.x - nd->blk_off;
    const int ibase = bi * CBS_BLK;
    const int npts = min(BAND_PTS, n + 1 - ibase);   // points ibase .. ibase+npts-1 (<= n)
    const double total = nd->total;
    for (int k = tid; k < BAND_PTS; k += 256)
        sp[k] = (k < npts) ? make_double2(pt(S, lo, ibase + k), pt(cw, lo, ibase + k)) : make_double2(0.0, 0.0);
    __syncthreads();
    const int nsub = (npts + BAND_SUB - 1) / BAND_SUB;
    for (int sb = warp; sb < nsub; sb += 8) {       // min/max of S per 32-point sub-block
        const int k = sb * BAND_SUB + lane;
        double vmn = (k < npts) ? sp[k].x : INFINITY, vmx = (k < npts) ? sp[k].x : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmn = fmin(vmn, __shfl_xor_sync(0xffffffffu, vmn, o));
            vmx = fmax(vmx, __shfl_xor_sync(0xffffffffu, vmx, o));
        }
        if (lane == 0) { s_mn[sb] = vmn; s_mx[sb] = vmx; }
    }
    __syncthreads();
    const double gq = *((volatile double*)&nd->best_q);
    Cand best;
    best.q = gq;
    best.qlo = (gq < 0.0) ? -1.0 : gq * (1.0 - 1e-12);
    best.i = 0x7fffffff;
    best.j = 0x7fffffff;
    unsigned my_pairs = 0;
    const int il = warp * BAND_SUB + lane;          // this lane's start point (local)
    if (warp * BAND_SUB < npts && warp * BAND_SUB < CBS_BLK) {
        const int i = ibase + il;
        const bool iok = il < npts;
        // lanes without a start point carry NaN so that their comparisons are all false
        const double Si = iok ? sp[il].x : __longlong_as_double(0x7ff8000000000000ll), ci = iok ? sp[il].y : 0.0;
        const unsigned kw_span = (unsigned)(n - 2 * al0);   // al0 <= kw <= n - al0  <=>  kw - al0 <= n - 2 al0
        const int wend = min(warp * BAND_SUB + BAND_SUB - 1, npts - 1);
        // weight of the lightest al0-wide window that starts in this warp's sub-block
        double wmin = (il + al0 < npts) ? sp[il + al0].y - ci : INFINITY;
        if (PRUNE) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        }
        // lane <-> end sub-block: one bound test each, then walk the survivors (warp-uniform)
        unsigned need = 0;
        {
            const int sb = lane;
            bool keep = sb >= warp && sb < nsub;
            if (PRUNE && keep && gq >= 0.0) {
                const int j0 = sb * BAND_SUB, j1 = min(j0 + BAND_SUB - 1, npts - 1);
                const double d1 = s_mx[sb] - s_mn[warp], d2 = s_mx[warp] - s_mn[sb];
                const double dm = fmax(fabs(d1), fabs(d2));
                // lightest arc: for the own and the next sub-block every admissible arc holds an
                // al0-wide window that starts in this warp's sub-block (the weights are >= 0)
                const double cmin = (sb > warp + 1) ? sp[j0].y - sp[wend].y : wmin;
                const double cmax = sp[j1].y - sp[warp * BAND_SUB].y;
                const double fmin_ = fmin(cmin * (total - cmin), cmax * (total - cmax));
                if (fmin_ > 0.0 && (dm * dm) / fmin_ * (1.0 + 1e-9) < gq) keep = false;
            }
            need = __ballot_sync(0xffffffffu, keep);
        }
        while (need) {
            const int sb = __ffs(need) - 1;
            need &= need - 1;
            const int j0 = sb * BAND_SUB, j1 = min(j0 + BAND_SUB - 1, npts - 1);
            ++my_pairs;
            for (int jl = j0; jl <= j1; ++jl) {
                const double2 pj = sp[jl];
                const int kw = jl - il;
                const double d = __dsub_rn(pj.x, Si);
                const double c = __dsub_rn(pj.y, ci);
                const double den = __dmul_rn(c, __dsub_rn(total, c));
                const double num = __dmul_rn(d, d);
                if ((unsigned)(kw - al0) <= kw_span && num > __dmul_rn(best.qlo, den))
                    cand_update_inl(best, num, den, i, ibase + jl);
            }
        }
    }
    if (stats && lane == 0 && my_pairs) atomicAdd(&stats[2], (unsigned long long)my_pairs);
    cta_commit(nd, best.q, best.i, best.j, s_q, s_i, s_j);
}

template <bool PRUNE>
__global__ void __launch_bounds__(256)
cbs_bound_kernel(const Node* __restrict__ nodes, const int4* __restrict__ task_tab,
                 const double* __restrict__ cw, const double* __restrict__ blk_min,
                 const double* __restrict__ blk_max, TileRef* __restrict__ queue, int* __restrict__ qcount,
                 unsigned long long* __restrict__ stats) {
    __shared__ int s_warp_cnt[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int a, ib, jb0, NJ;
    decode_task(nodes, task_tab, blockIdx.x, a, ib, jb0, NJ);
    const Node* nd = &nodes[a];
    if (nd->flat) return;
    const int lo = nd->lo, n = nd->hi - nd->lo, boff = nd->blk_off;
    const double total = nd->total, gq = nd->best_q;
    unsigned mask = 0;
    u32 checked = 0;
    const int jb = jb0 + tid;
    if (jb < NJ) {
#pragma unroll
        for (int r = 0; r < CBS_IW; ++r) {
            const int bi = CBS_IW * ib + r;
            if (bi * CBS_BLK > n || bi + 1 >= jb) continue;  // bi >= jb-1: the band kernel's arcs
            bool need = true;
            if (PRUNE && gq >= 0.0) need = tile_needed(cw, blk_min, blk_max, lo, n, boff, bi, jb, total, gq);
            ++checked;
            if (need) mask |= 1u << r;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, mask != 0);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int k = 0; k < 8; ++k) tot += s_warp_cnt[k];
        s_base = tot ? atomicAdd(qcount, tot) : 0;
    }
    __syncthreads();
    if (mask) {
        int base = s_base;
        for (int k = 0; k < warp; ++k) base += s_warp_cnt[k];
        TileRef e;
        e.node = a; e.ib = ib; e.jb = jb; e.mask = mask;
        queue[base + __popc(bal & ((1u << lane) - 1u))] = e;
    }
    if (stats) {
        checked = warp_sum_u32(checked);
        if (lane == 0 && checked) atomicAdd(&stats[1], (unsigned long long)checked);
    }
}

template <bool PRUNE>
__global__ void __launch_bounds__(256)
cbs_scan_kernel(Node* __restrict__ nodes, const TileRef* __restrict__ queue, const int* __restrict__ qcount,
                int* __restrict__ pop_counter, const double* __restrict__ S, const double* __restrict__ cw,
                const double* __restrict__ blk_min, const double* __restrict__ blk_max, int al0,
                unsigned long long* __restrict__ stats) {
    __shared__ double2 sj[CBS_BLK];
    __shared__ int s_idx;
    __shared__ unsigned s_mask;
    __shared__ double s_q[8];
    __shared__ int s_i[8], s_j[8];
    const int tid = threadIdx.x;
    const int nq = *qcount;
    unsigned long long my_scanned = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(pop_counter, 1);
        __syncthreads();
        const int idx = s_idx;
        if (idx >= nq) break;
        const TileRef e = queue[idx];
        Node* nd = &nodes[e.node];
        const int lo = nd->lo, n = nd->hi - nd->lo, boff = nd->blk_off;
        const int P = n + 1;
        const int ib = e.ib, jb = e.jb;
        const double total = nd->total;
        const double gq = *((volatile double*)&nd->best_q);
        unsigned mask = e.mask;
        if (PRUNE) {
            // the best may have risen since the bound kernel ran: re-test (warp 0, one lane per start block)
            if (tid < 32) {
                bool keep = false;
                if (tid < CBS_IW && (mask & (1u << tid))) {
                    const int bi = CBS_IW * ib + tid;
                    keep = !(bi < jb && gq >= 0.0) || tile_needed(cw, blk_min, blk_max, lo, n, boff, bi, jb, total, gq);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (tid == 0) s_mask = m;
            }
            __syncthreads();
            mask = s_mask;
            if (mask == 0) continue;
        }
        const int jbase = jb * CBS_BLK;
        const int jcount = min(CBS_BLK, P - jbase);
        const int jend = jbase + jcount - 1;
        const int na = __popc(mask);
        bool masked = false;
#pragma unroll
        for (int r = 0; r < CBS_IW; ++r) {
            if (!(mask & (1u << r))) continue;
            const int ibase = (CBS_IW * ib + r) * CBS_BLK;
            const int iend = min(ibase + CBS_BLK - 1, n);
            if (!(jbase - iend >= al0 && jend - ibase <= n - al0 && ibase + CBS_BLK - 1 <= n)) masked = true;
        }
        my_scanned += (tid == 0) ? na : 0;
        if (tid < jcount) sj[tid] = make_double2(pt(S, lo, jbase + tid), pt(cw, lo, jbase + tid));
        double Si[CBS_IW], ci[CBS_IW];
        int ii[CBS_IW];
#pragma unroll
        for (int r = 0; r < CBS_IW; ++r) {
            const int slot = (int)__fns(mask, 0, (r < na) ? r + 1 : 1);
            const int i = (CBS_IW * ib + slot) * CBS_BLK + tid;
            ii[r] = i;
            const bool ok = i <= n;
            Si[r] = ok ? pt(S, lo, i) : 0.0;
            ci[r] = ok ? pt(cw, lo, i) : 0.0;
        }
        Cand best;
        best.q = gq;
        best.qlo = (gq < 0.0) ? -1.0 : gq * (1.0 - 1e-12);
        best.i = 0x7fffffff;
        best.j = 0x7fffffff;
        __syncthreads();
        if (masked) {
            switch (na) {
                case 1: scan_subtiles<1, true>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
                case 2: scan_subtiles<2, true>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
                case 3: scan_subtiles<3, true>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
                default: scan_subtiles<4, true>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
            }
        } else {
            switch (na) {
                case 1: scan_subtiles<1, false>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
                case 2: scan_subtiles<2, false>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
                case 3: scan_subtiles<3, false>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
                default: scan_subtiles<4, false>(best, sj, Si, ci, ii, jbase, jcount, total, n, al0); break;
            }
        }
        cta_commit(nd, best.q, best.i, best.j, s_q, s_i, s_j);
    }
    if (stats && tid == 0 && my_scanned) atomicAdd(&stats[0], my_scanned);
}

// ------------------------------------------------------------------------------------ decide
__device__ __forceinline__ double pnorm_std(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

__device__ double nu_dev(double x, double tol) {
    double lnu1;
    if (x > 0.01) {
        lnu1 = log(2.0) - 2.0 * log(x);
        double lnu0 = lnu1, dk = 0.0;
        int k = 2;
        for (int i = 0; i < k; ++i) {
            dk += 1.0;
            lnu1 -= 2.0 * pnorm_std(-x * sqrt(dk) / 2.0) / dk;
        }
        while (fabs((lnu1 - lnu0) / lnu1) > tol) {
            lnu0 = lnu1;
            for (int i = 0; i < k; ++i) {
                dk += 1.0;
                lnu1 -= 2.0 * pnorm_std(-x * sqrt(dk) / 2.0) / dk;
            }
            k *= 2;
        }
    } else {
        lnu1 = -0.583 * x;
    }
    return exp(lnu1);
}

__device__ __forceinline__ double it1tsq_dev(double x, double a) {
    double yy = x + a - 0.5;
    double r = (8.0 * yy) / (1.0 - 4.0 * yy * yy) + 2.0 * log((1.0 + 2.0 * yy) / (1.0 - 2.0 * yy));
    yy = x - 0.5;
    r = r - (8.0 * yy) / (1.0 - 4.0 * yy * yy) - 2.0 * log((1.0 + 2.0 * yy) / (1.0 - 2.0 * yy));
    return r;
}

__device__ __forceinline__ double t2_of(double bss, double tss, int n) {
    if (tss <= bss + 0.0001) tss = bss + 1.0;
    return bss / ((tss - bss) / ((double)n - 2.0));
}

enum { ST_NEED_TAILP = 4 };

// thread per node: observed T^2 and the decisions that need no p-value
__global__ void cbs_stat_kernel(Node* __restrict__ nodes, int nnodes, double alpha, int nperm, int nmin,
                                int hybrid_method) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnodes) return;
    Node* nd = &nodes[idx];
    const int n = nd->hi - nd->lo;
    int state = ST_FINAL, nrejc = 0;
    double ostat = 0.0, ostat1 = 0.0;
    if (!nd->flat && nd->best_q >= 0.0) {
        const double t2 = t2_of(nd->best_q, nd->tss, n);
        ostat1 = sqrt(t2);
        ostat = t2 * 0.99999;
        if (ostat1 > 0.1 || (hybrid_method & 2)) {
            const int bi = nd->best_i, bj = nd->best_j;
            const int l = min(bj - bi, n - bj + bi);
            const bool diag = (hybrid_method & 2) != 0;   // diagnostics: every hybrid node gets its tail probability
            if (ostat1 >= 7.0 && l >= 10 && !diag) state = ST_SPLIT;
            else if ((hybrid_method & 1) && n > nmin) state = ST_NEED_TAILP;
            else { state = ST_PERM_EXACT; nrejc = (int)(alpha * (double)nperm); }
        }
    }
    nd->state = state;
    nd->nrejc = nrejc;
    nd->ostat = ostat;
    nd->ostat1 = ostat1;
}

// Cheap two-sided bracket of DNAcopy's tailp that settles almost every node without the series.
// The Siegmund-Yakir closed form  nu_sy(x) = (2/x)(Phi(x/2)-1/2) / ((x/2)Phi(x/2)+phi(x/2))  lies
// below the reference's truncated series value by at most 2.2 % for every x > 0.01 (it is exact to
// 1e-15 for x > 12; tests/test_oracle_pins.py sweeps the bracket), so with p_sy = tailp computed
// from nu_sy:  p_sy (1-1e-6) <= p_reference <= 1.05 p_sy.  One warp per node.  A node whose bracket
// straddles alpha (or an nrejc step alpha - m/nperm) keeps ST_NEED_TAILP and gets the exact series.
__global__ void __launch_bounds__(128)
cbs_tailp_bounds_kernel(Node* __restrict__ nodes, int nnodes, double alpha, int nperm, int ngrid) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * 4 + warp;
    if (idx >= nnodes) return;
    Node* nd = &nodes[idx];
    if (nd->state != ST_NEED_TAILP) return;
    const int n = nd->hi - nd->lo;
    const double b = nd->ostat1, delta = nd->delta;
    const double dincr = (0.5 - delta) / (double)ngrid;
    const double bsqrtm = b / sqrt((double)n);
    double p = 0.0;
    for (int g = lane; g < ngrid; g += 32) {
        const double tl = 0.5 - dincr * (double)(g + 1), tt = 0.5 - dincr * ((double)g + 0.5);
        const double x = bsqrtm / sqrt(tt * (1.0 - tt));
        double nux;
        if (x > 0.01) {
            const double h = 0.5 * x;
            const double Pc = 0.5 * erf(h * 0.70710678118654752440);  // Phi(h) - 1/2
            const double ph = 0.3989422804014327 * exp(-0.5 * h * h);
            nux = ((2.0 / x) * Pc) / (h * (0.5 + Pc) + ph);
        } else {
            nux = exp(-0.583 * x);
        }
        p += (nux * nux) * it1tsq_dev(tl, dincr);
    }
    p = warp_sum(p);
    if (lane == 0) {
        p = 2.0 * (9.973557e-2 * (b * b * b) * exp(-(b * b) / 2.0) * p);
        const double plo = p * (1.0 - 1e-6), phi = p * 1.05;
        if (plo > alpha) {
            nd->state = ST_FINAL;
        } else if (phi <= alpha) {
            const int r_hi = (int)((alpha - plo) * (double)nperm), r_lo = (int)((alpha - phi) * (double)nperm);
            if (r_hi == r_lo) { nd->state = ST_PERM_HYBRID; nd->nrejc = r_lo; }
        }
    }
}

// DNAcopy tailp: one CTA per (grid point, node).  nu(x)'s series is summed in the reference's
// doubling blocks (2,2,4,8,...), each block split over the 128 threads and tree-reduced; the
// relative-change stopping rule is evaluated between blocks exactly as in the serial code.
__global__ void __launch_bounds__(128)
cbs_tailp_terms_kernel(const Node* __restrict__ nodes, double* __restrict__ terms, int kmax, int ngrid, double tol) {
    __shared__ double s_part[4];
    __shared__ double s_sum;
    const Node* nd = &nodes[blockIdx.y];
    if (nd->state != ST_NEED_TAILP) return;
    const int n = nd->hi - nd->lo;
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double b = nd->ostat1;
    const double delta = nd->delta;
    const double dincr = (0.5 - delta) / (double)ngrid;
    const double bsqrtm = b / sqrt((double)n);
    double tl = 0.5 - dincr, tt = 0.5 - 0.5 * dincr;
    for (int s = 0; s < g; ++s) { tl -= dincr; tt -= dincr; }  // same roundings as the serial loop
    const double x = bsqrtm / sqrt(tt * (1.0 - tt));
    double lnu1;
    if (x > 0.01) {
        lnu1 = log(2.0) - 2.0 * log(x);
        double lnu0 = lnu1;
        long long done = 0;
        int k = 2;
        bool first = true;
        for (;;) {
            double part = 0.0;
            for (long long i = done + 1 + tid; i <= done + k; i += 128) {
                const double dk = (double)i;
                part += 2.0 * pnorm_std(-x * sqrt(dk) / 2.0) / dk;
            }
            part = warp_sum(part);
            if (lane == 0) s_part[warp] = part;
            __syncthreads();
            if (tid == 0) s_sum = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
            __syncthreads();
            lnu1 -= s_sum;
            done += k;
            if (!first) k *= 2;
            first = false;
            if (!(fabs((lnu1 - lnu0) / lnu1) > tol)) break;
            lnu0 = lnu1;
            __syncthreads();
        }
    } else {
        lnu1 = -0.583 * x;
    }
    if (tid == 0) {
        const double nux = exp(lnu1);
        terms[(size_t)blockIdx.y * ngrid + g] = (nux * nux) * it1tsq_dev(tl, dincr);
    }
}

// thread per node: finish the tail probability and the hybrid decision
__global__ void cbs_decide_kernel(Node* __restrict__ nodes, int nnodes, Decision* __restrict__ dec,
                                  const double* __restrict__ terms, double alpha, int nperm, int ngrid) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnodes) return;
    Node* nd = &nodes[idx];
    int state = nd->state, nrejc = nd->nrejc;
    if (state == ST_NEED_TAILP) {
        const double b = nd->ostat1;
        double p = 0.0;
        for (int g = 0; g < ngrid; ++g) p += terms[(size_t)idx * ngrid + g];
        p = 9.973557e-2 * (b * b * b) * exp(-(b * b) / 2.0) * p;
        p = 2.0 * p;
        nd->tailp = p;
        if (p > alpha) state = ST_FINAL;
        else { state = ST_PERM_HYBRID; nrejc = (int)((alpha - p) * (double)nperm); }
        nd->state = state;
        nd->nrejc = nrejc;
    }
    Decision d;
    d.state = state;
    d.nrejc = nrejc;
    d.best_i = nd->best_i;
    d.best_j = nd->best_j;
    d.ostat = nd->ostat;
    dec[idx] = d;
}

// ------------------------------------------------------------------------------------ permutations
// Pruning limits for the hybrid permutation statistic of one node (DNAcopy keeps the same kind of table,
// mncwt, for hwtmaxp).  A permutation only has to answer "does any short arc reach the observed
// statistic?", i.e. is  acc^2 / den(s,k) >= q_thr  for some arc, with q_thr the smallest BSS whose T^2
// reaches ostat.  The weights are not permuted, so den(s,k) = c (W - c), c = sum of the arc's k
// normalised weights, is the same in every permutation: lim[k] = q_thr * min_s den(s,k) * (1 - 1e-12)
// lets the permutation kernel drop an arc after one add, one multiply and one compare.
// One CTA per node; cacc is accumulated exactly as the permutation kernel accumulates it.
static const int HYB_LIM = HYB_KMAX + 1;

__global__ void __launch_bounds__(256)
cbs_perm_bounds_kernel(const Node* __restrict__ nodes, const int* __restrict__ list, const double* __restrict__ wn,
                       int al0, int kmax, unsigned long long* __restrict__ dmin_bits) {
    // grid (tile of HYB_TP arc starts, list index): min over the tile's starts of den(s, k) for every k,
    // merged into dmin_bits[node][k] with atomicMin on the bit pattern (den > 0, so the order is numeric)
    __shared__ double s_w[HYB_TP + HYB_KMAX];
    __shared__ double s_min[8][HYB_LIM];
    const int node = list[blockIdx.y];
    const Node nd = nodes[node];
    const int n = nd.hi - nd.lo, lo = nd.lo;
    const int s0 = blockIdx.x * HYB_TP;
    if (s0 >= n) return;
    const double total = nd.total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = min(HYB_TP, n - s0);
    const int need = count + kmax - 1;
    for (int k = threadIdx.x; k < need; k += 256) {
        int pos = s0 + k;
        if (pos >= n) pos -= n;
        s_w[k] = wn[lo + pos];
    }
    __syncthreads();
    double dmin[HYB_LIM];
#pragma unroll 1
    for (int k = 0; k <= kmax; ++k) dmin[k] = INFINITY;
    for (int s = threadIdx.x; s < count; s += 256) {
        double cacc = 0.0;
#pragma unroll 1
        for (int k = 1; k <= kmax; ++k) {
            cacc = __dadd_rn(cacc, s_w[s + k - 1]);
            if (k >= al0) dmin[k] = fmin(dmin[k], __dmul_rn(cacc, __dsub_rn(total, cacc)));
        }
    }
#pragma unroll 1
    for (int k = al0; k <= kmax; ++k) {
        double v = dmin[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) s_min[warp][k] = v;
    }
    __syncthreads();
    for (int k = al0 + threadIdx.x; k <= kmax; k += 256) {
        double v = s_min[0][k];
        for (int w_ = 1; w_ < 8; ++w_) v = fmin(v, s_min[w_][k]);
        if (v < INFINITY) atomicMin(&dmin_bits[(size_t)node * HYB_LIM + k], (unsigned long long)__double_as_longlong(v));
    }
}

// hybrid: circular arcs of width al0..kmax over the permuted terms; grid (pos tile, perm, list idx)
// KMAX_T / AL0_T > 0: widths fixed at compile time (DNAcopy's defaults kmax = 25, min.width = 2) so that the
// arc loop unrolls into LDS + DADD + DMUL + DSETP per arc; 0 = run-time widths.
template <int KMAX_T, int AL0_T>
__global__ void __launch_bounds__(256, 3)
cbs_perm_hybrid_kernel(const Node* __restrict__ nodes, const int* __restrict__ list, int p0,
                       const double* __restrict__ y, const double* __restrict__ w_sqrt,
                       const double* __restrict__ wn, const unsigned long long* __restrict__ dmin_bits, u64 seed,
                       int al0, int kmax, unsigned long long* __restrict__ pmax, int chunk) {
    __shared__ double s_t[HYB_TP + HYB_KMAX];
    __shared__ double s_lim[HYB_LIM];
    __shared__ double s_best[8];
    const int node = list[blockIdx.z];
    const Node nd = nodes[node];
    const int n = nd.hi - nd.lo, lo = nd.lo;
    const int s0 = blockIdx.x * HYB_TP;
    if (s0 >= n) return;
    const int perm = p0 + blockIdx.y;
    Prp prp;
    prp.init(seed, (u32)(nd.lo - nd.base), (u32)(nd.hi - nd.base), 0u, (u32)perm, (u32)n);
    const int count = min(HYB_TP, n - s0);
    const int need = count + kmax - 1;
    for (int k = threadIdx.x; k < need; k += 256) {
        int pos = s0 + k;
        if (pos >= n) pos -= n;
        s_t[k] = __dmul_rn(y[lo + prp((u32)pos)], w_sqrt[lo + pos]);
    }
    if (threadIdx.x <= kmax) {
        // smallest BSS whose T^2 reaches ostat: bss (n-2) / (tss - bss) >= ostat (t2_of is increasing there);
        // no pruning when tss is so small that t2_of's guard (tss <= bss + 1e-4) could come into play
        const double nm2 = (double)n - 2.0;
        const double qthr = nd.ostat * nd.tss / (nm2 + nd.ostat) * (1.0 - 1e-9);
        const bool safe = nd.tss * nm2 / (nm2 + nd.ostat) > 1e-3 && nd.ostat > 0.0 && threadIdx.x >= al0;
        const double dm = __longlong_as_double((long long)dmin_bits[(size_t)node * HYB_LIM + threadIdx.x]);
        s_lim[threadIdx.x] = safe ? qthr * dm * (1.0 - 1e-12) : 0.0;
    }
    __syncthreads();
    const double total = nd.total;
    double best = 0.0, bestlo = 0.0;
    // rare path: this arc could reach the observed statistic -> score it exactly
    auto score = [&](int s, int k, double num) {
        double cacc = 0.0;
        int pos = s0 + s;                                  // the weights are read from global memory here only
        for (int j = 0; j < k; ++j) {
            if (pos >= n) pos -= n;
            cacc = __dadd_rn(cacc, wn[lo + pos]);
            ++pos;
        }
        const double den = __dmul_rn(cacc, __dsub_rn(total, cacc));
        if (num > __dmul_rn(bestlo, den)) {
            const double q = __ddiv_rn(num, den);
            if (q > best) { best = q; bestlo = q * (1.0 - 1e-12); }
        }
    };
    if (KMAX_T > 0) {
        for (int s = threadIdx.x; s < count; s += 256) {
            double acc = 0.0;
#pragma unroll
            for (int k = 1; k <= KMAX_T; ++k) {
                acc = __dadd_rn(acc, s_t[s + k - 1]);
                if (k >= AL0_T) {
                    const double num = __dmul_rn(acc, acc);
                    if (num >= s_lim[k]) score(s, k, num);
                }
            }
        }
    } else {
        for (int s = threadIdx.x; s < count; s += 256) {
            double acc = 0.0;
            for (int k = 1; k <= kmax; ++k) {
                acc = __dadd_rn(acc, s_t[s + k - 1]);
                if (k >= al0) {
                    const double num = __dmul_rn(acc, acc);
                    if (num >= s_lim[k]) score(s, k, num);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) best = fmax(best, s_best[k]);
        if (best > 0.0)
            atomicMax(&pmax[(size_t)blockIdx.z * chunk + blockIdx.y], (unsigned long long)__double_as_longlong(best));
    }
}

// exact: all arcs of the permuted node (n <= EXACT_NMAX); one warp per (node, perm)
__global__ void __launch_bounds__(128)
cbs_perm_exact_kernel(const Node* __restrict__ nodes, const int* __restrict__ list, int p0,
                      const double* __restrict__ y, const double* __restrict__ w_sqrt,
                      const double* __restrict__ cw, u64 seed, int al0, unsigned long long* __restrict__ pmax,
                      int chunk) {
    __shared__ double s_S[4][EXACT_NMAX + 1];
    __shared__ double s_cw[EXACT_NMAX + 1];
    const Node nd = nodes[list[blockIdx.y]];
    const int n = nd.hi - nd.lo, lo = nd.lo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k <= n; k += 128) s_cw[k] = pt(cw, lo, k);
    const int pidx = blockIdx.x * 4 + warp;
    const bool live = pidx < chunk;
    double* Sp = s_S[warp];
    if (live) {
        Prp prp;
        prp.init(seed, (u32)(nd.lo - nd.base), (u32)(nd.hi - nd.base), 0u, (u32)(p0 + pidx), (u32)n);
        for (int k = lane; k < n; k += 32) Sp[k + 1] = __dmul_rn(y[lo + prp((u32)k)], w_sqrt[lo + k]);
        __syncwarp();
        if (lane == 0) {
            double run = 0.0;
            Sp[0] = 0.0;
            for (int k = 1; k <= n; ++k) {
                run = __dadd_rn(run, Sp[k]);
                Sp[k] = run;
            }
        }
    }
    __syncthreads();
    if (!live) return;
    const double total = nd.total;
    double best = 0.0, bestlo = 0.0;
    for (int i = lane; i <= n - al0; i += 32) {
        const double Si = Sp[i], ci = s_cw[i];
        const int jhi = min(n, i + (n - al0));
        for (int j = i + al0; j <= jhi; ++j) {
            const double d = __dsub_rn(Sp[j], Si);
            const double c = __dsub_rn(s_cw[j], ci);
            const double den = __dmul_rn(c, __dsub_rn(total, c));
            const double num = __dmul_rn(d, d);
            if (num > __dmul_rn(bestlo, den)) {
                const double q = __ddiv_rn(num, den);
                if (q > best) { best = q; bestlo = q * (1.0 - 1e-12); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) pmax[(size_t)blockIdx.y * chunk + pidx] = (unsigned long long)__double_as_longlong(best);
}

// exceedance flags: ostat <= T^2(pmax)
__global__ void cbs_perm_flags_kernel(const Node* __restrict__ nodes, const int* __restrict__ list, int nlist,
                                      const unsigned long long* __restrict__ pmax, int chunk,
                                      unsigned char* __restrict__ flags) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nlist * chunk) return;
    const Node* nd = &nodes[list[k / chunk]];
    const double pb = __longlong_as_double((long long)pmax[k]);
    const double pstat = t2_of(pb, nd->tss, nd->hi - nd->lo);
    flags[k] = (nd->ostat <= pstat) ? 1 : 0;
}

// ------------------------------------------------------------------------------------ flank tests
// wtpermp set-up: every test's rows are split over FLANK_SPLIT CTAs (partial sums in a fixed order),
// then one thread per test combines the partials left to right -- deterministic, thresholds only
static const int FLANK_SPLIT = 16;

__global__ void __launch_bounds__(256)
cbs_flank_part_kernel(const Node* __restrict__ nodes, const FlankTest* __restrict__ tests,
                      const double* __restrict__ xc, const double* __restrict__ w, double* __restrict__ part) {
    __shared__ double s_part[5][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x, sp = blockIdx.y;
    const FlankTest ft = tests[idx];
    const Node* nd = &nodes[ft.node];
    const int g0 = nd->lo + ft.off;
    const int n1 = ft.n1, n = ft.n1 + ft.n2;
    const int chunk = (n + FLANK_SPLIT - 1) / FLANK_SPLIT;
    const int k0 = sp * chunk, k1 = min(n, k0 + chunk);
    double xs1 = 0, xs2 = 0, tss = 0, rn1 = 0, rn2 = 0;
    for (int k = k0 + threadIdx.x; k < k1; k += 256) {
        const double xv = xc[g0 + k], wv = w[g0 + k];
        const double wx = wv * xv;
        tss += wv * (xv * xv);
        if (k < n1) { xs1 += wx; rn1 += wv; } else { xs2 += wx; rn2 += wv; }
    }
    xs1 = warp_sum(xs1); xs2 = warp_sum(xs2); tss = warp_sum(tss); rn1 = warp_sum(rn1); rn2 = warp_sum(rn2);
    if (lane == 0) { s_part[0][warp] = xs1; s_part[1][warp] = xs2; s_part[2][warp] = tss; s_part[3][warp] = rn1; s_part[4][warp] = rn2; }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0.0;
        for (int k = 0; k < 8; ++k) v += s_part[threadIdx.x][k];
        part[((size_t)idx * FLANK_SPLIT + sp) * 5 + threadIdx.x] = v;
    }
}

__global__ void cbs_flank_prep_kernel(FlankTest* __restrict__ tests, int ntests, const double* __restrict__ part,
                                      int force_perm) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ntests) return;
    FlankTest* o = &tests[idx];
    const int n1 = o->n1, n2 = o->n2, n = n1 + n2;
    if (n1 == 1 || n2 == 1) { o->pvalue = 1.0; o->need_perm = 0; return; }
    double xs1 = 0, xs2 = 0, tss = 0, rn1 = 0, rn2 = 0;
    for (int sp = 0; sp < FLANK_SPLIT; ++sp) {
        const double* q = part + ((size_t)idx * FLANK_SPLIT + sp) * 5;
        xs1 += q[0]; xs2 += q[1]; tss += q[2]; rn1 += q[3]; rn2 += q[4];
    }
    const double rn = rn1 + rn2;
    const double xbar = (xs1 + xs2) / rn;
    tss = tss - rn * (xbar * xbar);
    int m1, s0;
    double rm1, odiff;
    if (n1 <= n2) { m1 = n1; s0 = 0; rm1 = rn1; odiff = fabs(xs1 / rn1 - xbar); }
    else { m1 = n2; s0 = n1; rm1 = rn2; odiff = fabs(xs2 / rn2 - xbar); }
    double tstat = rn * rm1 * (odiff * odiff) / (rn - rm1);
    tstat = tstat / ((tss - tstat) / ((double)n - 2.0));
    o->m1 = m1; o->s0 = s0; o->rm1 = rm1; o->xbar = xbar; o->ostat = 0.99999 * odiff;
    if (tstat > 25.0 && m1 >= 10 && !force_perm) { o->pvalue = 0.0; o->need_perm = 0; }
    else { o->pvalue = -1.0; o->need_perm = 1; }
}

// one warp per (test, perm): pstat = |sum_{short side} y[pi(k)] rw[k] / rm1 - xbar|
__global__ void __launch_bounds__(128)
cbs_flank_perm_kernel(const Node* __restrict__ nodes, const FlankTest* __restrict__ tests,
                      const int* __restrict__ list, int nperm, const double* __restrict__ xc,
                      const double* __restrict__ w_sqrt, u64 seed, int* __restrict__ nrej) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int perm = blockIdx.x * 4 + warp;
    if (perm >= nperm) return;
    const FlankTest ft = tests[list[blockIdx.y]];
    const Node* nd = &nodes[ft.node];
    const int g0 = nd->lo + ft.off;
    const int n = ft.n1 + ft.n2;
    Prp prp;
    prp.init(seed, (u32)(nd->lo - nd->base), (u32)(nd->hi - nd->base), (u32)ft.tid, (u32)perm, (u32)n);
    double xs = 0.0;
    for (int k = lane; k < ft.m1; k += 32) {
        const int src = g0 + (int)prp((u32)(ft.s0 + k));
        xs += (xc[src] * w_sqrt[src]) * w_sqrt[g0 + ft.s0 + k];
    }
    xs = warp_sum(xs);
    if (lane == 0) {
        const double pstat = fabs(xs / ft.rm1 - ft.xbar);
        if (ft.ostat <= pstat) atomicAdd(&nrej[blockIdx.y], 1);
    }
}

// ------------------------------------------------------------------------------------ misc kernels
__global__ void sqrt_kernel(const double* __restrict__ w, double* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sqrt(w[i]);
}

// weighted.mean(log2, weight) per segment (cbs.py:70-71); one CTA per segment
__global__ void __launch_bounds__(1024)
seg_mean_kernel(const double* __restrict__ x, const double* __restrict__ w, const long long* __restrict__ lo,
                const long long* __restrict__ hi, int nseg, double* __restrict__ mean) {
    __shared__ double s_a[32], s_b[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    double sw = 0.0, swx = 0.0;
    for (long long k = lo[s] + threadIdx.x; k < hi[s]; k += 1024) {
        sw += w[k];
        swx += x[k] * w[k];
    }
    sw = warp_sum(sw);
    swx = warp_sum(swx);
    if (lane == 0) { s_a[warp] = sw; s_b[warp] = swx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; ++k) { sw += s_a[k]; swx += s_b[k]; }
        mean[s] = swx / sw;
    }
}

// ------------------------------------------------------------------------------------ host side
// DNAcopy getbdry (sequential stopping boundary), restated with an exact draw-by-draw
// hypergeometric recursion instead of phyper()/closed-form pexceed.
static void etabdry_host(int nperm, double eta0, int j, int* ibdry) {
    std::vector<double> dist(j + 1, 0.0);
    dist[0] = 1.0;
    int k = 0;
    for (int i = 1; i <= nperm && k < j; ++i) {
        const double rem = (double)(nperm - i + 1);
        for (int xx = std::min(i, j); xx >= 0; --xx) {
            const double stay = dist[xx] * (1.0 - (double)(j - xx) / rem);
            const double up = xx > 0 ? dist[xx - 1] * ((double)(j - (xx - 1)) / rem) : 0.0;
            dist[xx] = stay + up;
        }
        double cdf = 0.0;
        for (int xx = 0; xx <= k; ++xx) cdf += dist[xx];
        if (cdf <= eta0) ibdry[k++] = i;
    }
    for (; k < j; ++k) ibdry[k] = nperm;
}

static double pexceed_host(int nperm, int j, const int* ibdry) {
    std::vector<double> alive(j + 1, 0.0);
    alive[0] = 1.0;
    double crossed = 0.0;
    int kb = 0;
    for (int i = 1; i <= nperm; ++i) {
        const double rem = (double)(nperm - i + 1);
        for (int xx = std::min(i, j); xx >= 0; --xx) {
            const double stay = alive[xx] * (1.0 - (double)(j - xx) / rem);
            const double up = xx > 0 ? alive[xx - 1] * ((double)(j - (xx - 1)) / rem) : 0.0;
            alive[xx] = stay + up;
        }
        while (kb < j && ibdry[kb] == i) {
            for (int xx = 0; xx <= kb; ++xx) { crossed += alive[xx]; alive[xx] = 0.0; }
            ++kb;
        }
    }
    return crossed;
}

void cbs_getbdry(double eta, int nperm, int max_ones, std::vector<int>& out) {
    out.assign((size_t)max_ones * (max_ones + 1) / 2, 0);
    out[0] = nperm - (int)((double)nperm * eta);
    double eta0 = eta;
    int l = 1;
    for (int j = 2; j <= max_ones; ++j) {
        int* b = out.data() + l;
        double etahi = eta0 * 1.1, etalo = eta0 * 0.25;
        etabdry_host(nperm, etahi, j, b);
        double phi = pexceed_host(nperm, j, b);
        etabdry_host(nperm, etalo, j, b);
        double plo = pexceed_host(nperm, j, b);
        while ((etahi - etalo) / etalo > 1e-2) {
            eta0 = etalo + (etahi - etalo) * (eta - plo) / (phi - plo);
            etabdry_host(nperm, eta0, j, b);
            const double pex = pexceed_host(nperm, j, b);
            if (pex > eta) { etahi = eta0; phi = pex; } else { etalo = eta0; plo = pex; }
        }
        l += j;
    }
}

// Pinned staging memory for the small per-level host<->device copies (node tables, decisions, flags):
// copies from pageable std::vector storage go through the driver's own staging buffer and cost ~10 us
// each, which adds up over ~10 copies per recursion level.  Chunks persist per thread across calls.
struct PinnedArena {
    struct Chunk { char* p; size_t cap; };
    std::vector<Chunk> chunks;
    size_t cur = 0, used = 0;
    void rewind() { cur = 0; used = 0; }
    void* get(size_t bytes) {
        bytes = (bytes + 63) & ~(size_t)63;
        while (cur < chunks.size() && used + bytes > chunks[cur].cap) { ++cur; used = 0; }
        if (cur == chunks.size()) {
            Chunk c;
            c.cap = std::max(bytes, (size_t)1 << 20);
            if (cudaHostAlloc((void**)&c.p, c.cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
            chunks.push_back(c);
            used = 0;
        }
        void* r = chunks[cur].p + used;
        used += bytes;
        return r;
    }
    template <typename T>
    T* put(const T* src, size_t count) {          // copy into pinned memory, return the pinned pointer
        T* d = (T*)get(count * sizeof(T));
        if (d && count) memcpy(d, src, count * sizeof(T));
        return d;
    }
};
static thread_local PinnedArena g_pinned;

#define CNVK_PIN_CHECK(ptr)                                                     \
    do {                                                                        \
        if (!(ptr)) { set_last_error("cbs_segment: pinned staging allocation failed"); return (int)cudaErrorMemoryAllocation; } \
    } while (0)

struct HostNode {
    int lo, hi, base, arm;
};

// tasks (CTAs of the seed / bound kernels) and 1024x256 tiles of a node of n bins
static void tiles_for(int n, int node, std::vector<int4>* table, long long* tasks, long long* tiles) {
    const int P = n + 1;
    const int NJ = (P + CBS_BLK - 1) / CBS_BLK, NI = (P + CBS_IBLK - 1) / CBS_IBLK;
    long long t = 0, q = 0;
    for (int ib = 0; ib < NI; ++ib) {
        const int m = std::max(0, NJ - CBS_IW * ib);
        const int chunks = (m + 255) / 256;
        for (int c = 0; c < chunks; ++c) table->push_back(make_int4(node, ib, CBS_IW * ib + 256 * c, NJ));
        t += chunks;
        q += m;
    }
    *tasks = t;
    *tiles = q;
}

int cbs_segment(const double* d_x, const double* d_w, const long long* h_off, int n_arms, const CbsParams& P_in,
                std::vector<CbsSeg>& segs, CbsStats* stats, cudaStream_t stream, CbsDiag* diag) {
    // diagnostics (cnvk_cbs_node_diag): every interval is ONE node; one level only, the tail series is always
    // summed, every node runs all nperm permutations (no stopping rule) and both flank tests with permutations
    CbsParams P = P_in;
    if (diag) P.prune |= 2;
    segs.clear();
    CbsStats st;
    memset(&st, 0, sizeof(st));
    if (n_arms <= 0) { if (stats) *stats = st; return 0; }
    const long long N = h_off[n_arms];
    if (N >= (1ll << 31)) { set_last_error("cbs_segment: too many bins (%lld)", N); return -2; }
    if (P.kmax > HYB_KMAX || P.nmin > EXACT_NMAX || P.min_width < 2 || P.kmax < P.min_width || P.nperm < 1) {
        set_last_error("cbs_segment: unsupported parameters (kmax<=%d, nmin<=%d, min_width>=2)", HYB_KMAX, EXACT_NMAX);
        return -2;
    }
    if (P.nmin < 2 * P.kmax) { set_last_error("cbs_segment: need nmin >= 2*kmax"); return -2; }
    int sm_count = 148;
    {
        int dev = 0;
        CNVK_CHECK(cudaGetDevice(&dev));
        CNVK_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int max_ones = (int)std::floor((double)P.nperm * P.alpha) + 1;
    // the stopping boundary depends on (eta, nperm, max.ones) only and costs ~nperm * max.ones * 20 flops to
    // build: keep the last one per thread
    static thread_local std::vector<int> sbdry;
    static thread_local double sb_eta = -1.0;
    static thread_local int sb_nperm = -1, sb_ones = -1;
    if (sb_eta != P.eta || sb_nperm != P.nperm || sb_ones != max_ones) {
        cbs_getbdry(P.eta, P.nperm, max_ones, sbdry);
        sb_eta = P.eta; sb_nperm = P.nperm; sb_ones = max_ones;
    }

    Scratch scratch(stream);
    double *d_rw, *d_xc, *d_y, *d_S, *d_cw, *d_wn, *d_bmin, *d_bmax, *d_Sl, *d_cwl;
    const size_t NB = (size_t)(N / CBS_BLK) + 2 * (size_t)n_arms + 2 * (size_t)N / 4 + 16;  // >= sum ceil((n+1)/256)
    CNVK_CHECK(scratch.alloc(&d_rw, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_xc, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_y, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_S, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_cw, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_wn, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_Sl, (size_t)N));    // tile-local prefix sums (before the carries)
    CNVK_CHECK(scratch.alloc(&d_cwl, (size_t)N));
    CNVK_CHECK(scratch.alloc(&d_bmin, NB));
    CNVK_CHECK(scratch.alloc(&d_bmax, NB));
    int *d_bimin, *d_bimax;
    double *d_bcmin, *d_bcmax;
    CNVK_CHECK(scratch.alloc(&d_bimin, NB));
    CNVK_CHECK(scratch.alloc(&d_bimax, NB));
    CNVK_CHECK(scratch.alloc(&d_bcmin, NB));
    CNVK_CHECK(scratch.alloc(&d_bcmax, NB));
    if (N > 0) {
        sqrt_kernel<<<div_up(N, 256), 256, 0, stream>>>(d_w, d_rw, N);
        CNVK_LAUNCH_CHECK();
    }
    int* d_counter;  // [0] tiles queued, [1] tiles popped
    unsigned long long* d_stats;
    CNVK_CHECK(scratch.alloc(&d_counter, 2));
    CNVK_CHECK(scratch.alloc(&d_stats, 3));  // [0] 256x256 sub-tiles scanned, [1] tested, [2] 32x32 band pairs scanned
    CNVK_CHECK(cudaMemsetAsync(d_stats, 0, 3 * sizeof(unsigned long long), stream));

    std::vector<HostNode> frontier;
    std::vector<std::vector<int>> ends(n_arms);
    for (int a = 0; a < n_arms; ++a) {
        const int lo = (int)h_off[a], hi = (int)h_off[a + 1];
        if (hi <= lo) continue;
        if (hi - lo >= 2 * P.min_width) frontier.push_back({lo, hi, lo, a});
        else ends[a].push_back(hi);
    }

    // growable device buffers for per-level arrays
    size_t cap_nodes = 0, cap_queue = 0, cap_ptiles = 0;
    TileRef* d_queue = nullptr;
    TileSums* d_part1 = nullptr;
    TileScan* d_part3 = nullptr;
    int* d_ptile_off = nullptr;
    std::vector<int> h_ptile_off;
    std::vector<int4> h_task_tab;
    int4* d_task_tab = nullptr;
    SeedRes* d_seedres = nullptr;
    int* d_task_off = nullptr;
    size_t cap_tasks = 0, cap_task_off = 0;
    struct TaskGuard { int4*& p; SeedRes*& r; int*& o; cudaStream_t s; ~TaskGuard() { if (p) cudaFreeAsync(p, s); if (r) cudaFreeAsync(r, s); if (o) cudaFreeAsync(o, s); } } tguard{d_task_tab, d_seedres, d_task_off, stream};
    struct PartGuard { TileSums*& a; TileScan*& b; cudaStream_t s; ~PartGuard() { if (a) cudaFreeAsync(a, s); if (b) cudaFreeAsync(b, s); } } pguard{d_part1, d_part3, stream};
    Node* d_nodes = nullptr;
    Decision* d_dec = nullptr;
    double* d_terms = nullptr;
    struct QueueGuard { TileRef*& q; cudaStream_t s; ~QueueGuard() { if (q) cudaFreeAsync(q, s); } } qguard{d_queue, stream};
    auto free_level = [&]() {
        if (d_terms) cudaFreeAsync(d_terms, stream);
        d_terms = nullptr;
        if (d_nodes) cudaFreeAsync(d_nodes, stream);
        if (d_dec) cudaFreeAsync(d_dec, stream);
        if (d_ptile_off) cudaFreeAsync(d_ptile_off, stream);
        d_nodes = nullptr; d_dec = nullptr; d_ptile_off = nullptr;
    };
    struct Guard { decltype(free_level)& f; ~Guard() { f(); } } guard{free_level};

    std::vector<Node> h_nodes;
    std::vector<Decision> h_dec;
    std::vector<int> h_tile_off;

    while (!frontier.empty()) {
        st.levels++;
        const int nn = (int)frontier.size();
        st.nodes += nn;
        if ((size_t)nn > cap_nodes) {
            free_level();
            cap_nodes = (size_t)nn * 2;
            CNVK_CHECK(cudaMallocAsync((void**)&d_nodes, cap_nodes * sizeof(Node), stream));
            CNVK_CHECK(cudaMallocAsync((void**)&d_dec, cap_nodes * sizeof(Decision), stream));
            CNVK_CHECK(cudaMallocAsync((void**)&d_ptile_off, (cap_nodes + 1) * sizeof(int), stream));
            CNVK_CHECK(cudaMallocAsync((void**)&d_terms, cap_nodes * 100 * sizeof(double), stream));
        }
        h_nodes.assign(nn, Node());
        h_tile_off.assign(nn + 1, 0);
        h_ptile_off.assign(nn + 1, 0);
        h_task_tab.clear();
        long long blk_off = 0, tiles = 0, qtiles = 0, ptiles = 0;
        for (int k = 0; k < nn; ++k) {
            Node& nd = h_nodes[k];
            memset(&nd, 0, sizeof(Node));
            nd.lo = frontier[k].lo; nd.hi = frontier[k].hi; nd.base = frontier[k].base;
            nd.blk_off = (int)blk_off;
            nd.best_q = -1.0; nd.best_i = 0x7fffffff; nd.best_j = 0x7fffffff;
            nd.tailp = -1.0;
            const int n = nd.hi - nd.lo;
            st.prep_bins += n;
            h_ptile_off[k] = (int)ptiles;
            ptiles += (n + 1 + PREP_TILE - 1) / PREP_TILE;     // tiles of 2048 prefix points
            blk_off += (n + 1 + CBS_BLK - 1) / CBS_BLK;
            h_tile_off[k] = (int)tiles;
            long long nt_, nq_;
            tiles_for(n, k, &h_task_tab, &nt_, &nq_);
            tiles += nt_;
            qtiles += nq_;
            if (tiles >= (1ll << 31) || qtiles >= (1ll << 31)) { set_last_error("cbs_segment: tile count overflow"); return -2; }
        }
        h_tile_off[nn] = (int)tiles;
        h_ptile_off[nn] = (int)ptiles;
        if ((size_t)ptiles > cap_ptiles) {
            if (d_part1) cudaFreeAsync(d_part1, stream);
            if (d_part3) cudaFreeAsync(d_part3, stream);
            d_part1 = nullptr; d_part3 = nullptr;
            cap_ptiles = (size_t)ptiles * 2;
            CNVK_CHECK(cudaMallocAsync((void**)&d_part1, cap_ptiles * sizeof(TileSums), stream));
            CNVK_CHECK(cudaMallocAsync((void**)&d_part3, cap_ptiles * sizeof(TileScan), stream));
        }
        if ((size_t)blk_off > NB) { set_last_error("cbs_segment: block buffer too small"); return -3; }
        g_pinned.rewind();   // every copy of the previous level has completed (the level ends with a sync)
        if (h_task_tab.size() > cap_tasks) {
            if (d_task_tab) cudaFreeAsync(d_task_tab, stream);
            if (d_seedres) cudaFreeAsync(d_seedres, stream);
            d_task_tab = nullptr; d_seedres = nullptr;
            cap_tasks = h_task_tab.size() * 2;
            CNVK_CHECK(cudaMallocAsync((void**)&d_task_tab, cap_tasks * sizeof(int4), stream));
            CNVK_CHECK(cudaMallocAsync((void**)&d_seedres, cap_tasks * sizeof(SeedRes), stream));
        }
        if ((size_t)nn + 1 > cap_task_off) {
            if (d_task_off) cudaFreeAsync(d_task_off, stream);
            d_task_off = nullptr;
            cap_task_off = ((size_t)nn + 1) * 2;
            CNVK_CHECK(cudaMallocAsync((void**)&d_task_off, cap_task_off * sizeof(int), stream));
        }
        {
            int* p_to = g_pinned.put(h_tile_off.data(), (size_t)nn + 1);
            CNVK_PIN_CHECK(p_to);
            CNVK_CHECK(cudaMemcpyAsync(d_task_off, p_to, ((size_t)nn + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        if (!h_task_tab.empty()) {
            int4* p_tab = g_pinned.put(h_task_tab.data(), h_task_tab.size());
            CNVK_PIN_CHECK(p_tab);
            CNVK_CHECK(cudaMemcpyAsync(d_task_tab, p_tab, h_task_tab.size() * sizeof(int4), cudaMemcpyHostToDevice, stream));
        }
        {
            Node* p_nodes = g_pinned.put(h_nodes.data(), (size_t)nn);
            int* p_pto = g_pinned.put(h_ptile_off.data(), (size_t)nn + 1);
            CNVK_PIN_CHECK(p_nodes && p_pto);
            CNVK_CHECK(cudaMemcpyAsync(d_nodes, p_nodes, nn * sizeof(Node), cudaMemcpyHostToDevice, stream));
            CNVK_CHECK(cudaMemcpyAsync(d_ptile_off, p_pto, (nn + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        CNVK_CHECK(cudaMemsetAsync(d_counter, 0, 2 * sizeof(int), stream));
        if ((size_t)qtiles > cap_queue) {
            if (d_queue) cudaFreeAsync(d_queue, stream);
            d_queue = nullptr;
            cap_queue = (size_t)qtiles;
            CNVK_CHECK(cudaMallocAsync((void**)&d_queue, cap_queue * sizeof(TileRef), stream));
        }
        {
            ProfScope ps(PC_CBS_PREP, stream);
            const int hk = P.hybrid ? P.kmax : 0;
            cbs_prep_sums_kernel<<<(int)ptiles, PREP_THREADS, 0, stream>>>(d_nodes, d_ptile_off, nn, d_x, d_w, d_part1);
            CNVK_LAUNCH_CHECK();
            cbs_prep_scan_kernel<<<(int)ptiles, PREP_THREADS, PREP_SMEM, stream>>>(d_nodes, d_ptile_off, nn, d_x, d_w, d_part1, d_xc, d_y, d_wn, d_Sl, d_cwl, d_part3);
            CNVK_LAUNCH_CHECK();
            cbs_prep_finish_kernel<<<(int)ptiles, PREP_THREADS, 0, stream>>>(d_nodes, d_ptile_off, nn, d_Sl, d_cwl, d_part3, d_S, d_cw, d_bmin, d_bmax, d_bimin, d_bimax, d_bcmin, d_bcmax, hk, P.nmin);
            CNVK_LAUNCH_CHECK();
            cbs_prep_node_kernel<<<div_up(nn, 4), 128, 0, stream>>>(d_nodes, d_ptile_off, nn, d_part3, d_cw, d_bmin, d_bmax, d_bimin, d_bimax, d_bcmin, d_bcmax, P.min_width, hk, P.nmin);
            CNVK_LAUNCH_CHECK();
        }
        {
            if (tiles > 0) {
                const int grid = sm_count * 4;
                if (P.prune & 1) {
                    ProfScope ps(PC_CBS_SEED, stream);
                    cbs_seed_kernel<<<(int)tiles, 256, 0, stream>>>(d_nodes, d_task_tab, d_seedres, d_bmin, d_bmax, d_bimin, d_bimax, d_bcmin, d_bcmax, P.min_width);
                    CNVK_LAUNCH_CHECK();
                    cbs_seed_reduce_kernel<<<div_up(nn, 4), 128, 0, stream>>>(d_nodes, nn, d_task_off, d_seedres);
                    CNVK_LAUNCH_CHECK();
                }
                {
                    ProfScope ps(PC_CBS_BAND, stream);
                    if (P.prune & 1) cbs_band_kernel<true><<<(int)blk_off, 256, 0, stream>>>(d_nodes, nn, d_S, d_cw, P.min_width, d_stats);
                    else cbs_band_kernel<false><<<(int)blk_off, 256, 0, stream>>>(d_nodes, nn, d_S, d_cw, P.min_width, d_stats);
                    CNVK_LAUNCH_CHECK();
                }
                {
                    ProfScope ps(PC_CBS_BOUND, stream);
                    if (P.prune & 1) cbs_bound_kernel<true><<<(int)tiles, 256, 0, stream>>>(d_nodes, d_task_tab, d_cw, d_bmin, d_bmax, d_queue, d_counter, d_stats);
                    else cbs_bound_kernel<false><<<(int)tiles, 256, 0, stream>>>(d_nodes, d_task_tab, d_cw, d_bmin, d_bmax, d_queue, d_counter, d_stats);
                    CNVK_LAUNCH_CHECK();
                }
                {
                    ProfScope ps(PC_CBS_SCAN, stream);
                    if (P.prune & 1) cbs_scan_kernel<true><<<grid, 256, 0, stream>>>(d_nodes, d_queue, d_counter, d_counter + 1, d_S, d_cw, d_bmin, d_bmax, P.min_width, d_stats);
                    else cbs_scan_kernel<false><<<grid, 256, 0, stream>>>(d_nodes, d_queue, d_counter, d_counter + 1, d_S, d_cw, d_bmin, d_bmax, P.min_width, d_stats);
                    CNVK_LAUNCH_CHECK();
                }
            }
        }
        {
            ProfScope ps(PC_CBS_DECIDE, stream);
            const int ngrid = 100;
            cbs_stat_kernel<<<div_up(nn, 128), 128, 0, stream>>>(d_nodes, nn, P.alpha, P.nperm, P.nmin, (P.hybrid ? 1 : 0) | (diag ? 2 : 0));
            CNVK_LAUNCH_CHECK();
            if (P.hybrid && !(P.prune & 2)) {
                cbs_tailp_bounds_kernel<<<div_up(nn, 4), 128, 0, stream>>>(d_nodes, nn, P.alpha, P.nperm, ngrid);
                CNVK_LAUNCH_CHECK();
            }
            if (P.hybrid) {
                for (int y0 = 0; y0 < nn; y0 += 32768) {
                    const int ny = std::min(32768, nn - y0);
                    cbs_tailp_terms_kernel<<<dim3(ngrid, ny), 128, 0, stream>>>(d_nodes + y0, d_terms + (size_t)y0 * ngrid, P.kmax, ngrid, 1e-6);
                    CNVK_LAUNCH_CHECK();
                }
            }
            cbs_decide_kernel<<<div_up(nn, 128), 128, 0, stream>>>(d_nodes, nn, d_dec, d_terms, P.alpha, P.nperm, ngrid);
            CNVK_LAUNCH_CHECK();
        }
        h_dec.resize(nn);
        {
            Decision* p_dec = (Decision*)g_pinned.get((size_t)nn * sizeof(Decision));
            CNVK_PIN_CHECK(p_dec);
            CNVK_CHECK(cudaMemcpyAsync(p_dec, d_dec, nn * sizeof(Decision), cudaMemcpyDeviceToHost, stream));
            CNVK_CHECK(cudaStreamSynchronize(stream));
            memcpy(h_dec.data(), p_dec, (size_t)nn * sizeof(Decision));
        }
        std::vector<int> diag_nrej;
        if (diag) {
            diag->nodes.resize(nn);
            CNVK_CHECK(cudaMemcpyAsync(diag->nodes.data(), d_nodes, nn * sizeof(Node), cudaMemcpyDeviceToHost, stream));
            CNVK_CHECK(cudaStreamSynchronize(stream));
            diag_nrej.assign(nn, 0);
            for (int k = 0; k < nn; ++k) {
                const Node& nd = diag->nodes[k];
                if (nd.flat || nd.best_q < 0.0) { h_dec[k].state = ST_FINAL; continue; }
                h_dec[k].state = (P.hybrid && nd.hi - nd.lo > P.nmin) ? ST_PERM_HYBRID : ST_PERM_EXACT;
                h_dec[k].nrejc = P.nperm;
            }
        }

        // ---- permutation reference distributions
        for (int mode = ST_PERM_HYBRID; mode <= ST_PERM_EXACT; ++mode) {
            std::vector<int> list;
            for (int k = 0; k < nn; ++k) if (h_dec[k].state == mode) list.push_back(k);
            if (list.empty()) continue;
            st.perm_nodes += (long long)list.size();
            std::vector<int> nrej(list.size(), 0), kk(list.size()), done(list.size(), 0);
            for (size_t q = 0; q < list.size(); ++q) {
                const int nrejc = h_dec[list[q]].nrejc;
                kk[q] = nrejc * (nrejc + 1) / 2 + 1;
            }
            int p0 = 0;
            int chunk = 64;
            std::vector<int> active(list.begin(), list.end());
            std::vector<size_t> active_q(list.size());
            for (size_t q = 0; q < list.size(); ++q) active_q[q] = q;
            Scratch limscratch(stream);
            unsigned long long* d_lim = nullptr;
            if (mode == ST_PERM_HYBRID) {
                // pruning limits of every node that enters the hybrid permutation test (once per node)
                int* d_all;
                CNVK_CHECK(limscratch.alloc(&d_lim, (size_t)nn * HYB_LIM));
                CNVK_CHECK(limscratch.alloc(&d_all, list.size()));
                CNVK_CHECK(cudaMemsetAsync(d_lim, 0x7f, (size_t)nn * HYB_LIM * sizeof(unsigned long long), stream));
                int* p_all = g_pinned.put(list.data(), list.size());
                CNVK_PIN_CHECK(p_all);
                CNVK_CHECK(cudaMemcpyAsync(d_all, p_all, list.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
                ProfScope bscope(PC_CBS_PERM_HYBRID, stream);
                int maxn = 0;
                for (int k : list) maxn = std::max(maxn, frontier[k].hi - frontier[k].lo);
                for (size_t z0 = 0; z0 < list.size(); z0 += 32768) {
                    const int nz = (int)std::min<size_t>(32768, list.size() - z0);
                    cbs_perm_bounds_kernel<<<dim3(div_up(maxn, HYB_TP), nz), 256, 0, stream>>>(d_nodes, d_all + z0, d_wn, P.min_width, P.kmax, d_lim);
                    CNVK_LAUNCH_CHECK();
                }
            }
            while (!active.empty() && p0 < P.nperm) {
                chunk = std::min(chunk, P.nperm - p0);
                const int na = (int)active.size();
                Scratch ps(stream);
                int* d_list;
                unsigned long long* d_pmax;
                unsigned char* d_flags;
                CNVK_CHECK(ps.alloc(&d_list, (size_t)na));
                CNVK_CHECK(ps.alloc(&d_pmax, (size_t)na * chunk));
                CNVK_CHECK(ps.alloc(&d_flags, (size_t)na * chunk));
                {
                    int* p_list = g_pinned.put(active.data(), (size_t)na);
                    CNVK_PIN_CHECK(p_list);
                    CNVK_CHECK(cudaMemcpyAsync(d_list, p_list, na * sizeof(int), cudaMemcpyHostToDevice, stream));
                }
                CNVK_CHECK(cudaMemsetAsync(d_pmax, 0, (size_t)na * chunk * sizeof(unsigned long long), stream));
                ProfScope pscope(mode == ST_PERM_HYBRID ? PC_CBS_PERM_HYBRID : PC_CBS_PERM_EXACT, stream);
                if (mode == ST_PERM_HYBRID) {
                    int maxn = 0;
                    for (int k : active) maxn = std::max(maxn, frontier[k].hi - frontier[k].lo);
                    // grid.z is limited to 65535 and grid.y too: split the node list if needed
                    for (int z0 = 0; z0 < na; z0 += 32768) {
                        const int nz = std::min(32768, na - z0);
                        dim3 grid(div_up(maxn, HYB_TP), chunk, nz);
                        if (P.kmax == 25 && P.min_width == 2)
                            cbs_perm_hybrid_kernel<25, 2><<<grid, 256, 0, stream>>>(d_nodes, d_list + z0, p0, d_y, d_rw, d_wn, d_lim, P.seed, P.min_width, P.kmax, d_pmax + (size_t)z0 * chunk, chunk);
                        else
                            cbs_perm_hybrid_kernel<0, 0><<<grid, 256, 0, stream>>>(d_nodes, d_list + z0, p0, d_y, d_rw, d_wn, d_lim, P.seed, P.min_width, P.kmax, d_pmax + (size_t)z0 * chunk, chunk);
                        CNVK_LAUNCH_CHECK();
                    }
                    for (int k : active) st.perm_arcs += (long long)(frontier[k].hi - frontier[k].lo) * (P.kmax - P.min_width + 1) * chunk;
                } else {
                    for (int z0 = 0; z0 < na; z0 += 32768) {
                        const int nz = std::min(32768, na - z0);
                        dim3 grid(div_up(chunk, 4), nz);
                        cbs_perm_exact_kernel<<<grid, 128, 0, stream>>>(d_nodes, d_list + z0, p0, d_y, d_rw, d_cw, P.seed, P.min_width, d_pmax + (size_t)z0 * chunk, chunk);
                        CNVK_LAUNCH_CHECK();
                    }
                    for (int k : active) { const long long n = frontier[k].hi - frontier[k].lo; st.perm_arcs += n * n / 2 * chunk; }
                }
                cbs_perm_flags_kernel<<<div_up((long long)na * chunk, 256), 256, 0, stream>>>(d_nodes, d_list, na, d_pmax, chunk, d_flags);
                CNVK_LAUNCH_CHECK();
                unsigned char* flags = (unsigned char*)g_pinned.get((size_t)na * chunk);
                CNVK_PIN_CHECK(flags);
                CNVK_CHECK(cudaMemcpyAsync(flags, d_flags, (size_t)na * chunk, cudaMemcpyDeviceToHost, stream));
                CNVK_CHECK(cudaStreamSynchronize(stream));
                st.perms_run += (long long)na * chunk;
                // replay DNAcopy's sequential stopping rule on the flags
                std::vector<int> next_active;
                std::vector<size_t> next_q;
                for (int a = 0; a < na; ++a) {
                    const size_t q = active_q[a];
                    const int nrejc = h_dec[list[q]].nrejc;
                    int decided = 0;
                    for (int c = 0; c < chunk && !decided; ++c) {
                        const int np = p0 + c + 1;
                        if (flags[(size_t)a * chunk + c]) { nrej[q]++; kk[q]++; }
                        if (diag) continue;
                        if (nrej[q] > nrejc) decided = -1;
                        else if (np >= sbdry[kk[q] - 1]) decided = 1;
                    }
                    if (diag) diag_nrej[list[q]] = nrej[q];
                    if (decided == 0 && p0 + chunk >= P.nperm) decided = 1;  // survived every permutation
                    if (decided == 0) { next_active.push_back(active[a]); next_q.push_back(q); }
                    else h_dec[list[q]].state = decided > 0 ? ST_SPLIT : ST_FINAL;
                }
                active.swap(next_active);
                active_q.swap(next_q);
                p0 += chunk;
                chunk = std::min(chunk * 4, 2048);
            }
        }

        // ---- ternary-split confirmation
        std::vector<FlankTest> tests;
        std::vector<int> test_owner;
        for (int k = 0; k < nn; ++k) {
            if (h_dec[k].state != ST_SPLIT) continue;
            const int n = frontier[k].hi - frontier[k].lo, ai = h_dec[k].best_i, aj = h_dec[k].best_j;
            if (aj == n || ai == 0) continue;
            FlankTest f;
            memset(&f, 0, sizeof(f));
            f.node = k; f.off = 0; f.n1 = ai; f.n2 = aj - ai; f.tid = 1;
            tests.push_back(f); test_owner.push_back(k);
            f.off = ai; f.n1 = aj - ai; f.n2 = n - aj; f.tid = 2;
            tests.push_back(f); test_owner.push_back(k);
        }
        std::vector<double> pvals(tests.size(), 0.0);
        if (!tests.empty()) {
            const int nt = (int)tests.size();
            Scratch fs(stream);
            FlankTest* d_tests;
            CNVK_CHECK(fs.alloc(&d_tests, (size_t)nt));
            FlankTest* p_tests = g_pinned.put(tests.data(), (size_t)nt);
            CNVK_PIN_CHECK(p_tests);
            CNVK_CHECK(cudaMemcpyAsync(d_tests, p_tests, nt * sizeof(FlankTest), cudaMemcpyHostToDevice, stream));
            ProfScope fscope(PC_CBS_FLANK, stream);
            double* d_part;
            CNVK_CHECK(fs.alloc(&d_part, (size_t)nt * FLANK_SPLIT * 5));
            for (int z0 = 0; z0 < nt; z0 += 32768) {   // grid.x is the test index
                const int nz = std::min(32768, nt - z0);
                cbs_flank_part_kernel<<<dim3(nz, FLANK_SPLIT), 256, 0, stream>>>(d_nodes, d_tests + z0, d_xc, d_w, d_part + (size_t)z0 * FLANK_SPLIT * 5);
                CNVK_LAUNCH_CHECK();
            }
            cbs_flank_prep_kernel<<<div_up(nt, 128), 128, 0, stream>>>(d_tests, nt, d_part, diag ? 1 : 0);
            CNVK_LAUNCH_CHECK();
            CNVK_CHECK(cudaMemcpyAsync(p_tests, d_tests, nt * sizeof(FlankTest), cudaMemcpyDeviceToHost, stream));
            CNVK_CHECK(cudaStreamSynchronize(stream));
            memcpy(tests.data(), p_tests, (size_t)nt * sizeof(FlankTest));
            std::vector<int> plist;
            for (int t = 0; t < nt; ++t) if (tests[t].need_perm) plist.push_back(t);
            if (!plist.empty()) {
                const int np_ = (int)plist.size();
                st.flank_perm_tests += np_;
                int *d_plist, *d_nrej;
                CNVK_CHECK(fs.alloc(&d_plist, (size_t)np_));
                CNVK_CHECK(fs.alloc(&d_nrej, (size_t)np_));
                CNVK_CHECK(cudaMemcpyAsync(d_plist, plist.data(), np_ * sizeof(int), cudaMemcpyHostToDevice, stream));
                CNVK_CHECK(cudaMemsetAsync(d_nrej, 0, np_ * sizeof(int), stream));
                for (int z0 = 0; z0 < np_; z0 += 32768) {
                    const int nz = std::min(32768, np_ - z0);
                    dim3 grid(div_up(P.nperm, 4), nz);
                    cbs_flank_perm_kernel<<<grid, 128, 0, stream>>>(d_nodes, d_tests, d_plist + z0, P.nperm, d_xc, d_rw, P.seed, d_nrej + z0);
                    CNVK_LAUNCH_CHECK();
                }
                std::vector<int> h_nrej(np_);
                CNVK_CHECK(cudaMemcpyAsync(h_nrej.data(), d_nrej, np_ * sizeof(int), cudaMemcpyDeviceToHost, stream));
                CNVK_CHECK(cudaStreamSynchronize(stream));
                for (int q = 0; q < np_; ++q) tests[plist[q]].pvalue = (double)h_nrej[q] / (double)P.nperm;
            }
            for (int t = 0; t < nt; ++t) pvals[t] = tests[t].pvalue;
        }

        if (diag) {
            diag->nrej = diag_nrej;
            diag->flank_p.assign((size_t)2 * nn, -1.0);
            for (size_t t = 0; t < tests.size(); ++t) diag->flank_p[(size_t)2 * test_owner[t] + (tests[t].tid - 1)] = pvals[t];
            if (stats) *stats = st;
            return 0;
        }
        // ---- next frontier
        std::vector<HostNode> next;
        size_t tcur = 0;
        for (int k = 0; k < nn; ++k) {
            const HostNode& hn = frontier[k];
            const int n = hn.hi - hn.lo;
            int ncpt = 0, icpt[2] = {0, 0};
            if (h_dec[k].state == ST_SPLIT) {
                const int ai = h_dec[k].best_i, aj = h_dec[k].best_j;
                if (aj == n) { ncpt = 1; icpt[0] = ai; }
                else if (ai == 0) { ncpt = 1; icpt[0] = aj; }
                else {
                    if (pvals[tcur] <= P.alpha) { ncpt = 1; icpt[0] = ai; }
                    if (pvals[tcur + 1] <= P.alpha) { icpt[ncpt] = aj; ++ncpt; }
                    if (ncpt == 0) { ncpt = 2; icpt[0] = ai; icpt[1] = aj; }
                    tcur += 2;
                }
            }
            if (ncpt == 0) { ends[hn.arm].push_back(hn.hi); continue; }
            int cuts[4] = {hn.lo, hn.lo + icpt[0], ncpt == 2 ? hn.lo + icpt[1] : hn.hi, hn.hi};
            const int npieces = ncpt + 1;
            for (int c = 0; c < npieces; ++c) {
                const int a0 = cuts[c], a1 = (c == npieces - 1) ? hn.hi : cuts[c + 1];
                if (a1 - a0 >= 2 * P.min_width) next.push_back({a0, a1, hn.base, hn.arm});
                else ends[hn.arm].push_back(a1);
            }
        }
        frontier.swap(next);
    }

    // ---- segment table + weighted means
    std::vector<long long> slo, shi;
    for (int a = 0; a < n_arms; ++a) {
        std::sort(ends[a].begin(), ends[a].end());
        long long prev = h_off[a];
        for (int e : ends[a]) {
            CbsSeg s;
            s.arm = a; s.lo = prev; s.hi = e; s.mean = 0.0;
            segs.push_back(s);
            slo.push_back(prev); shi.push_back(e);
            prev = e;
        }
    }
    if (!segs.empty()) {
        const int ns = (int)segs.size();
        Scratch ms(stream);
        long long *d_lo, *d_hi;
        double* d_mean;
        CNVK_CHECK(ms.alloc(&d_lo, (size_t)ns));
        CNVK_CHECK(ms.alloc(&d_hi, (size_t)ns));
        CNVK_CHECK(ms.alloc(&d_mean, (size_t)ns));
        CNVK_CHECK(cudaMemcpyAsync(d_lo, slo.data(), ns * sizeof(long long), cudaMemcpyHostToDevice, stream));
        CNVK_CHECK(cudaMemcpyAsync(d_hi, shi.data(), ns * sizeof(long long), cudaMemcpyHostToDevice, stream));
        {
            ProfScope ps(PC_CBS_MEANS, stream);
            seg_mean_kernel<<<ns, 1024, 0, stream>>>(d_x, d_w, d_lo, d_hi, ns, d_mean);
            CNVK_LAUNCH_CHECK();
        }
        std::vector<double> means(ns);
        CNVK_CHECK(cudaMemcpyAsync(means.data(), d_mean, ns * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CNVK_CHECK(cudaStreamSynchronize(stream));
        for (int s = 0; s < ns; ++s) segs[s].mean = means[s];
    }
    {
        unsigned long long h_stats[3];
        CNVK_CHECK(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, stream));
        CNVK_CHECK(cudaStreamSynchronize(stream));
        st.obs_subtiles_scanned = (long long)h_stats[0];
        st.obs_subtiles_total = (long long)h_stats[1];
        st.obs_arcs = st.obs_subtiles_scanned * (long long)CBS_BLK * CBS_BLK + (long long)h_stats[2] * BAND_SUB * BAND_SUB;
    }
    if (stats) *stats = st;
    return 0;
}

// One level over the given intervals, nothing decided early: the numbers the statistical tests compare with
// true-shuffle Monte-Carlo estimates (tests/test_cbs_stats_gpu.py; oracle twin: cbso_node_diag).
// out6[k] = {sqrt(T^2), tailp | -1, delta | -1, 0.99999 T^2, flank p 1 | -1, flank p 2 | -1},
// iout5[k] = {arc i, arc j, exceedances among nperm permutations, hybrid?, constant data?}.
int cbs_node_diag(const double* d_x, const double* d_w, const long long* h_off, int n_nodes, const CbsParams& P,
                  double* out6, int* iout5, cudaStream_t stream) {
    for (int k = 0; k < n_nodes; ++k)
        if (h_off[k + 1] - h_off[k] < 2 * P.min_width) { set_last_error("cbs_node_diag: node %d is shorter than 2*min_width", k); return -2; }
    CbsDiag diag;
    std::vector<CbsSeg> segs;
    CbsStats st;
    const int rc = cbs_segment(d_x, d_w, h_off, n_nodes, P, segs, &st, stream, &diag);
    if (rc) return rc;
    for (int k = 0; k < n_nodes; ++k) {
        const Node& nd = diag.nodes[k];
        double* o = out6 + (size_t)6 * k;
        int* io = iout5 + (size_t)5 * k;
        const bool live = !nd.flat && nd.best_q >= 0.0;
        const bool hyb = P.hybrid && (nd.hi - nd.lo) > P.nmin;
        o[0] = live ? nd.ostat1 : -1.0;
        o[1] = (live && hyb) ? nd.tailp : -1.0;
        o[2] = (live && hyb) ? nd.delta : -1.0;
        o[3] = live ? nd.ostat : -1.0;
        o[4] = diag.flank_p[(size_t)2 * k];
        o[5] = diag.flank_p[(size_t)2 * k + 1];
        io[0] = live ? nd.best_i : 0;
        io[1] = live ? nd.best_j : 0;
        io[2] = diag.nrej[k];
        io[3] = hyb ? 1 : 0;
        io[4] = nd.flat ? 1 : 0;
    }
    return 0;
}

}  // namespace cnvk
