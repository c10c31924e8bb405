// extern "C" surface of libcnvkit_b200.so -- see include/cnvkit_b200.h for the contract.
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/cnvkit_b200.h"
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {
std::atomic<unsigned long long> g_launch_count{0};
int g_prof_enabled = 0;
int g_nvtx_enabled = getenv("CNVK_NVTX") ? atoi(getenv("CNVK_NVTX")) : 0;
struct ProfPending { int cls; cudaEvent_t a, b; };
static std::vector<ProfPending> g_prof_pending;
static std::mutex g_prof_mutex;   // host threads of a cohort run share the profiler
static double g_prof_ms[PC_COUNT];
static unsigned long long g_prof_n[PC_COUNT];
void prof_begin(int cls, cudaStream_t stream, cudaEvent_t* start) {
    if (cudaEventCreate(start) != cudaSuccess) { *start = nullptr; return; }
    cudaEventRecord(*start, stream);
}
void prof_end(int cls, cudaStream_t stream, cudaEvent_t start) {
    cudaEvent_t stop;
    if (cudaEventCreate(&stop) != cudaSuccess) { cudaEventDestroy(start); return; }
    cudaEventRecord(stop, stream);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_pending.push_back({cls, start, stop});
}
static void prof_collect() {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& p : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof_ms[p.cls] += ms;
            g_prof_n[p.cls] += 1;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_prof_pending.clear();
}
static const char* const kProfNames[PC_COUNT] = {
    "radix_sort", "wavelet_build", "window_query", "cbs_prep", "cbs_scan", "cbs_seed", "cbs_band",
    "cbs_bound", "cbs_means", "cbs_decide", "cbs_perm_hybrid",
    "cbs_perm_exact", "cbs_flank", "fix", "haar", "biweight", "other"};
void prof_nvtx_push(int cls) { nvtxRangePushA((cls >= 0 && cls < PC_COUNT) ? kProfNames[cls] : "cnvk"); }
void prof_nvtx_pop() { nvtxRangePop(); }
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace cnvk

using namespace cnvk;

// Keep freed scratch in the stream-ordered pool across synchronisations (the default release
// threshold of 0 hands it back to the driver at every sync, which costs milliseconds per level).
static int ensure_init() {
    static thread_local int done_for = -1;
    int dev = 0;
    CNVK_CHECK(cudaGetDevice(&dev));
    if (done_for != dev) {
        cudaMemPool_t pool;
        CNVK_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long thr = ~0ull;
        CNVK_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        done_for = dev;
    }
    return 0;
}

// parked scratch slabs (common.cuh); the vector is leaked on purpose: thread-local destructors may run after statics
static std::mutex* g_slab_mu = new std::mutex;
static std::vector<cnvk::ScratchSlab>* g_slabs = new std::vector<cnvk::ScratchSlab>;
namespace cnvk {
bool scratch_stack_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CNVK_SCRATCH_STACK");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
void scratch_slab_park(const ScratchSlab& s) {
    std::lock_guard<std::mutex> lock(*g_slab_mu);
    g_slabs->push_back(s);
}
bool scratch_slab_adopt(int device, ScratchSlab* s) {
    std::lock_guard<std::mutex> lock(*g_slab_mu);
    int best = -1;
    for (size_t i = 0; i < g_slabs->size(); ++i)
        if ((*g_slabs)[i].device == device && (best < 0 || (*g_slabs)[i].cap > (*g_slabs)[best].cap)) best = (int)i;
    if (best < 0) return false;
    *s = (*g_slabs)[best];
    g_slabs->erase(g_slabs->begin() + best);
    return true;
}
}  // namespace cnvk

#define RC0(expr)              \
    do {                       \
        int _rc = (expr);      \
        if (_rc) return _rc;   \
    } while (0)

// Host-buffer helper: stage inputs, run, fetch outputs on a private stream.
struct HostCall {
    cudaStream_t stream;
    Scratch* scratch;
    HostCall() : stream(nullptr), scratch(nullptr) {}
    int begin() {
        RC0(ensure_init());
        CNVK_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        scratch = new Scratch(stream);
        return 0;
    }
    template <typename T>
    int upload(const T* h, size_t count, T** d) {
        CNVK_CHECK(scratch->alloc(d, count));
        if (count) CNVK_CHECK(cudaMemcpyAsync(*d, h, count * sizeof(T), cudaMemcpyHostToDevice, stream));
        return 0;
    }
    template <typename T>
    int download(T* h, const T* d, size_t count) {
        if (count) CNVK_CHECK(cudaMemcpyAsync(h, d, count * sizeof(T), cudaMemcpyDeviceToHost, stream));
        return 0;
    }
    int finish() {
        CNVK_CHECK(cudaStreamSynchronize(stream));
        return 0;
    }
    ~HostCall() {
        delete scratch;
        if (stream) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
    }
};

#define RC(expr)               \
    do {                       \
        int _rc = (expr);      \
        if (_rc) return _rc;   \
    } while (0)

extern "C" {

const char* cnvk_last_error(void) { return g_err; }
int cnvk_abi_version(void) { return 3; }
unsigned long long cnvk_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

void cnvk_prof_enable(int on) { g_prof_enabled = on; }
void cnvk_nvtx_enable(int on) { g_nvtx_enabled = on; }
void cnvk_prof_reset(void) {
    prof_collect();
    for (int k = 0; k < PC_COUNT; ++k) { g_prof_ms[k] = 0.0; g_prof_n[k] = 0; }
}
int cnvk_prof_count(void) { return PC_COUNT; }
const char* cnvk_prof_name(int cls) { return (cls >= 0 && cls < PC_COUNT) ? kProfNames[cls] : ""; }
int cnvk_prof_get(double* ms, unsigned long long* scopes, int n) {
    prof_collect();
    for (int k = 0; k < n && k < PC_COUNT; ++k) { ms[k] = g_prof_ms[k]; scopes[k] = g_prof_n[k]; }
    return 0;
}

int cnvk_fp64_peak(int iters, int reps, double* tflops) { return fp64_peak(iters, reps, tflops, nullptr); }

int cnvk_device_info(int device, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes) {
    cudaDeviceProp prop;
    CNVK_CHECK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    size_t f = 0, t = 0;
    CNVK_CHECK(cudaSetDevice(device));
    CNVK_CHECK(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return 0;
}

int cnvk_rolling_median_dev(const double* d_x, int64_t n, int64_t wing, double* d_out, void* stream) {
    RC(ensure_init());
    return rolling_window_stat(d_x, n, wing, 0, 0.5, 1, d_out, (cudaStream_t)stream);
}

int cnvk_rolling_quantile_dev(const double* d_x, int64_t n, int64_t wing, double q, double* d_out, void* stream) {
    RC(ensure_init());
    return rolling_window_stat(d_x, n, wing, 1, q, 2, d_out, (cudaStream_t)stream);
}

static int rolling_host(const double* x, int64_t n, int64_t wing, int mode, double q, int minp, double* out) {
    HostCall hc;
    RC(hc.begin());
    double *dx, *dout;
    RC(hc.upload(x, (size_t)n, &dx));
    CNVK_CHECK(hc.scratch->alloc(&dout, (size_t)n));
    RC(rolling_window_stat(dx, n, wing, mode, q, minp, dout, hc.stream));
    RC(hc.download(out, dout, (size_t)n));
    return hc.finish();
}

int cnvk_rolling_median(const double* x, int64_t n, int64_t wing, double* out) {
    return rolling_host(x, n, wing, 0, 0.5, 1, out);
}

int cnvk_rolling_quantile(const double* x, int64_t n, int64_t wing, double q, double* out) {
    return rolling_host(x, n, wing, 1, q, 2, out);
}

int cnvk_argsort_f64_dev(const double* d_x, int64_t n, uint32_t* d_order, void* stream) {
    RC(ensure_init());
    return argsort_f64(d_x, n, d_order, (cudaStream_t)stream);
}

int cnvk_fix_mask_dev(const double* d_sample_log2, const double* d_ref_log2, const double* d_ref_spread,
                      const double* d_ref_depth, const double* d_ref_gc, int64_t n, uint8_t* d_keep, void* stream) {
    RC(ensure_init());
    return fix_mask(d_sample_log2, d_ref_log2, d_ref_spread, d_ref_depth, d_ref_gc, n, d_keep, (cudaStream_t)stream);
}

int cnvk_center_all_dev(double* d_log2, const double* d_depth, const uint8_t* d_sel, int64_t n,
                        const int64_t* d_chrom_offsets, int32_t nchrom, int32_t skip_low, double* d_shift_out,
                        void* stream) {
    RC(ensure_init());
    return center_all(d_log2, d_depth, d_sel, n, (const long long*)d_chrom_offsets, nchrom, skip_low, d_shift_out,
                      (cudaStream_t)stream);
}

int cnvk_center_by_window_dev(double* d_log2, const double* d_key, const uint32_t* d_perm, int64_t n, int64_t wing,
                              void* stream) {
    RC(ensure_init());
    return center_by_window(d_log2, d_key, d_perm, n, wing, (cudaStream_t)stream);
}

int cnvk_center_by_window(double* log2, const double* key, const uint32_t* perm, int64_t n, int64_t wing) {
    HostCall hc;
    RC(hc.begin());
    double *dl, *dk;
    uint32_t* dp;
    RC(hc.upload(log2, (size_t)n, &dl));
    RC(hc.upload(key, (size_t)n, &dk));
    RC(hc.upload(perm, (size_t)n, &dp));
    RC(center_by_window(dl, dk, dp, n, wing, hc.stream));
    RC(hc.download(log2, dl, (size_t)n));
    return hc.finish();
}

int cnvk_edge_bias_dev(const int64_t* d_start, const int64_t* d_end, const int32_t* d_chrom_id, int64_t n,
                       int64_t margin, double* d_out, void* stream) {
    RC(ensure_init());
    return edge_bias((const long long*)d_start, (const long long*)d_end, d_chrom_id, n, margin, d_out,
                     (cudaStream_t)stream);
}

int cnvk_fix_group_dev(double* d_log2, const double* d_depth, const uint8_t* d_auto_sel, int64_t n,
                       const int64_t* d_chrom_offsets, int32_t nchrom, const int32_t* d_chrom_id,
                       const int64_t* d_start, const int64_t* d_end, const double* d_gc, const double* d_rmask,
                       const double* d_edge_key, int32_t do_edge, int32_t skip_low, int32_t do_center, int64_t wing,
                       const uint32_t* d_perm, int64_t margin, const char* order, int32_t* skipped, void* stream) {
    RC(ensure_init());
    int sk = 0;
    int rc = fix_group(d_log2, d_depth, d_auto_sel, n, (const long long*)d_chrom_offsets, nchrom, d_chrom_id,
                       (const long long*)d_start, (const long long*)d_end, d_gc, d_rmask, d_edge_key, do_edge,
                       skip_low, do_center, wing, d_perm, margin, order ? order : "ger", nullptr, nullptr, 0, &sk,
                       (cudaStream_t)stream);
    if (skipped) *skipped = sk;
    return rc;
}

int cnvk_trait_order_dev(const double* d_key, const uint32_t* d_perm, int64_t n, uint32_t* d_order, void* stream) {
    RC(ensure_init());
    return trait_order(d_key, d_perm, n, d_order, (cudaStream_t)stream);
}

int cnvk_center_by_order_dev(double* d_log2, const uint32_t* d_order, int64_t n, int64_t wing, void* stream) {
    RC(ensure_init());
    return center_by_order(d_log2, d_order, n, wing, (cudaStream_t)stream);
}

int cnvk_fix_group_ordered_dev(double* d_log2, const double* d_depth, const uint8_t* d_auto_sel, int64_t n,
                               const int64_t* d_chrom_offsets, int32_t nchrom, const uint32_t* d_order0,
                               const uint32_t* d_order1, const uint32_t* d_order2, int32_t skip_low, int32_t do_center,
                               int64_t wing, int32_t* skipped, void* stream) {
    RC(ensure_init());
    int sk = 0;
    const uint32_t* orders[3] = {d_order0, d_order1, d_order2};
    int rc = fix_group_ordered(d_log2, d_depth, d_auto_sel, n, (const long long*)d_chrom_offsets, nchrom, orders, 3,
                               skip_low, do_center, wing, &sk, (cudaStream_t)stream);
    if (skipped) *skipped = sk;
    return rc;
}

int cnvk_bias_correct_logr_dev(double* d_log2, const double* d_depth, const uint8_t* d_auto_sel, int64_t n,
                               const int64_t* d_chrom_offsets, int32_t nchrom, const double* d_gc,
                               const double* d_rmask, const double* d_edge_key, int32_t skip_low, int64_t wing,
                               const uint32_t* d_perm, const double* d_flat_logr, const uint8_t* d_sex_chrom,
                               int32_t is_xx, int32_t* skipped, void* stream) {
    RC(ensure_init());
    int sk = 0;
    int rc = fix_group(d_log2, d_depth, d_auto_sel, n, (const long long*)d_chrom_offsets, nchrom, nullptr, nullptr,
                       nullptr, d_gc, d_rmask, d_edge_key, d_edge_key != nullptr, skip_low, 1, wing, d_perm, 0, "gre",
                       d_flat_logr, d_sex_chrom, is_xx, &sk, (cudaStream_t)stream);
    if (skipped) *skipped = sk;
    return rc;
}

int cnvk_fix_finish_dev(double* d_log2, const double* d_depth, const double* d_ref_log2, const double* d_ref_spread,
                        const int64_t* d_start, const int64_t* d_end, const int32_t* d_chrom_id,
                        const uint8_t* d_is_anti, const uint8_t* d_auto_sel, int64_t n,
                        const int64_t* d_chrom_offsets, int32_t nchrom, double* d_weight, double* d_info,
                        void* stream) {
    RC(ensure_init());
    return fix_finish(d_log2, d_depth, d_ref_log2, d_ref_spread, (const long long*)d_start, (const long long*)d_end,
                      d_chrom_id, d_is_anti, d_auto_sel, n, (const long long*)d_chrom_offsets, nchrom, d_weight,
                      d_info, (cudaStream_t)stream);
}

int cnvk_segmented_median_dev(const double* d_x, const uint8_t* d_use, const int64_t* d_seg_offsets, int32_t nseg,
                              double* d_med, int32_t* d_cnt, void* stream) {
    RC(ensure_init());
    return segmented_median(d_x, d_use, (const long long*)d_seg_offsets, nseg, d_med, d_cnt, (cudaStream_t)stream);
}

int cnvk_biweight(const double* x, int64_t n, const double* initial, double* out) {
    HostCall hc;
    RC(hc.begin());
    double *dx, *dsorted, *dout, *dinit = nullptr;
    u64 *k0, *k1;
    u32 *v0, *v1, *nvalid;
    const size_t m = (size_t)(n > 0 ? n : 1);
    RC(hc.upload(x, (size_t)n, &dx));
    CNVK_CHECK(hc.scratch->alloc(&dsorted, m));
    CNVK_CHECK(hc.scratch->alloc(&dout, 2));
    CNVK_CHECK(hc.scratch->alloc(&k0, m));
    CNVK_CHECK(hc.scratch->alloc(&k1, m));
    CNVK_CHECK(hc.scratch->alloc(&v0, m));
    CNVK_CHECK(hc.scratch->alloc(&v1, m));
    CNVK_CHECK(hc.scratch->alloc(&nvalid, 1));
    if (initial) RC(hc.upload(initial, 1, &dinit));
    RC(sort_values_f64(dx, n, k0, k1, v0, v1, dsorted, nvalid, hc.stream));
    RC(biweight_of_sorted(dsorted, nvalid, 0, dinit, dout, hc.stream));
    RC(hc.download(out, dout, 2));
    return hc.finish();
}

int cnvk_biweight_columns_dev(const double* d_mat, int32_t rows, int64_t ncols, int32_t mode, double* d_loc,
                              double* d_var, void* stream) {
    RC(ensure_init());
    return biweight_columns(d_mat, rows, ncols, mode, d_loc, d_var, (cudaStream_t)stream);
}

int cnvk_biweight_columns(const double* mat, int32_t rows, int64_t ncols, int32_t mode, double* loc, double* var) {
    HostCall hc;
    RC(hc.begin());
    double *dm, *dl, *dv;
    RC(hc.upload(mat, (size_t)rows * (size_t)ncols, &dm));
    CNVK_CHECK(hc.scratch->alloc(&dl, (size_t)ncols));
    CNVK_CHECK(hc.scratch->alloc(&dv, (size_t)ncols));
    RC(biweight_columns(dm, rows, ncols, mode, dl, dv, hc.stream));
    if (mode & 1) RC(hc.download(loc, dl, (size_t)ncols));
    if (mode & 2) RC(hc.download(var, dv, (size_t)ncols));
    return hc.finish();
}

int cnvk_fasta_gc_lo_dev(const uint8_t* d_bytes, int64_t nbytes, const int64_t* d_range_lo,
                         const int64_t* d_range_hi, int64_t n_ranges, double* d_gc, double* d_rmask,
                         void* stream) {
    RC(ensure_init());
    return fasta_gc_lo(d_bytes, nbytes, (const long long*)d_range_lo, (const long long*)d_range_hi, n_ranges, d_gc,
                       d_rmask, (cudaStream_t)stream);
}

int cnvk_fasta_gc_lo(const uint8_t* bytes, int64_t nbytes, const int64_t* range_lo, const int64_t* range_hi,
                     int64_t n_ranges, double* gc, double* rmask) {
    HostCall hc;
    RC(hc.begin());
    unsigned char* db;
    long long *dl, *dh;
    double *dg, *dr;
    RC(hc.upload((const unsigned char*)bytes, (size_t)nbytes, &db));
    RC(hc.upload((const long long*)range_lo, (size_t)n_ranges, &dl));
    RC(hc.upload((const long long*)range_hi, (size_t)n_ranges, &dh));
    CNVK_CHECK(hc.scratch->alloc(&dg, (size_t)n_ranges));
    CNVK_CHECK(hc.scratch->alloc(&dr, (size_t)n_ranges));
    RC(fasta_gc_lo(db, nbytes, dl, dh, n_ranges, dg, dr, hc.stream));
    RC(hc.download(gc, dg, (size_t)n_ranges));
    RC(hc.download(rmask, dr, (size_t)n_ranges));
    return hc.finish();
}

static void boot_pcg(const cnvk_boot_params* p, u64 pcg[4]) {
    pcg[0] = p->pcg_state[0]; pcg[1] = p->pcg_state[1]; pcg[2] = p->pcg_inc[0]; pcg[3] = p->pcg_inc[1];
    if (!(pcg[0] | pcg[1] | pcg[2] | pcg[3])) {   // numpy.random.PCG64(0xA5EED).state
        pcg[0] = 0x52637b81f59d8cc6ull; pcg[1] = 0x00b11abc6a6a0481ull;
        pcg[2] = 0x392a8421e3a06208ull; pcg[3] = 0x808513643339beb7ull;
    }
}

int cnvk_bootstrap_ci_dev(const double* d_values, const double* d_weights, const int64_t* seg_lo,
                          const int64_t* seg_hi, int32_t n_segs, const cnvk_boot_params* params,
                          double* d_ci_lo, double* d_ci_hi, void* stream) {
    RC(ensure_init());
    u64 pcg[4];
    boot_pcg(params, pcg);
    return bootstrap_ci(d_values, d_weights, (const long long*)seg_lo, (const long long*)seg_hi, n_segs, params->alpha,
                        params->bootstraps, params->smooth_upto, pcg, params->noise_seed, d_ci_lo, d_ci_hi,
                        (cudaStream_t)stream);
}

int cnvk_bootstrap_ci(const double* values, const double* weights, int64_t n, const int64_t* seg_lo,
                      const int64_t* seg_hi, int32_t n_segs, const cnvk_boot_params* params, double* ci_lo,
                      double* ci_hi) {
    for (int32_t s = 0; s < n_segs; ++s) {
        if (seg_lo[s] < 0 || seg_hi[s] > n || seg_hi[s] < seg_lo[s]) {
            set_last_error("cnvk_bootstrap_ci: segment %d spans rows [%lld, %lld) of %lld", (int)s,
                           (long long)seg_lo[s], (long long)seg_hi[s], (long long)n);
            return -2;
        }
    }
    HostCall hc;
    RC(hc.begin());
    double *dv, *dw, *dl, *dh;
    RC(hc.upload(values, (size_t)n, &dv));
    RC(hc.upload(weights, (size_t)n, &dw));
    CNVK_CHECK(hc.scratch->alloc(&dl, (size_t)n_segs));
    CNVK_CHECK(hc.scratch->alloc(&dh, (size_t)n_segs));
    u64 pcg[4];
    boot_pcg(params, pcg);
    RC(bootstrap_ci(dv, dw, (const long long*)seg_lo, (const long long*)seg_hi, n_segs, params->alpha,
                    params->bootstraps, params->smooth_upto, pcg, params->noise_seed, dl, dh, hc.stream));
    RC(hc.download(ci_lo, dl, (size_t)n_segs));
    RC(hc.download(ci_hi, dh, (size_t)n_segs));
    return hc.finish();
}

static int cbs_common(const double* d_x, const double* d_w, const int64_t* arm_offsets, int32_t n_arms,
                      const cnvk_cbs_params* p, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo,
                      int64_t* seg_hi, double* seg_mean, int64_t* n_segs, cnvk_cbs_stats* stats,
                      cudaStream_t stream) {
    CbsParams P;
    P.alpha = p->alpha; P.nperm = p->nperm; P.min_width = p->min_width; P.kmax = p->kmax; P.nmin = p->nmin;
    P.eta = p->eta; P.hybrid = p->hybrid; P.prune = p->prune; P.seed = p->seed;
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1);
    std::vector<CbsSeg> segs;
    CbsStats st;
    RC(cbs_segment(d_x, d_w, off.data(), n_arms, P, segs, &st, stream));
    if ((int64_t)segs.size() > capacity) {
        set_last_error("cnvk_cbs_segment: %zu segments exceed capacity %lld", segs.size(), (long long)capacity);
        return -4;
    }
    for (size_t s = 0; s < segs.size(); ++s) {
        seg_arm[s] = segs[s].arm; seg_lo[s] = segs[s].lo; seg_hi[s] = segs[s].hi; seg_mean[s] = segs[s].mean;
    }
    *n_segs = (int64_t)segs.size();
    if (stats) {
        stats->nodes = st.nodes; stats->levels = st.levels; stats->obs_subtiles_total = st.obs_subtiles_total;
        stats->obs_subtiles_scanned = st.obs_subtiles_scanned; stats->obs_arcs = st.obs_arcs;
        stats->perm_arcs = st.perm_arcs; stats->perm_nodes = st.perm_nodes; stats->perms_run = st.perms_run;
        stats->flank_perm_tests = st.flank_perm_tests;
        stats->prep_bins = st.prep_bins;
    }
    return 0;
}

int cnvk_cbs_segment_dev(const double* d_log2, const double* d_weight, const int64_t* arm_offsets, int32_t n_arms,
                         const cnvk_cbs_params* params, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo,
                         int64_t* seg_hi, double* seg_mean, int64_t* n_segs, cnvk_cbs_stats* stats, void* stream) {
    RC(ensure_init());
    return cbs_common(d_log2, d_weight, arm_offsets, n_arms, params, capacity, seg_arm, seg_lo, seg_hi, seg_mean,
                      n_segs, stats, (cudaStream_t)stream);
}

int cnvk_cbs_segment(const double* log2, const double* weight, const int64_t* arm_offsets, int32_t n_arms,
                     const cnvk_cbs_params* params, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo,
                     int64_t* seg_hi, double* seg_mean, int64_t* n_segs, cnvk_cbs_stats* stats) {
    HostCall hc;
    RC(hc.begin());
    const size_t N = n_arms > 0 ? (size_t)arm_offsets[n_arms] : 0;
    double *dx, *dw;
    RC(hc.upload(log2, N, &dx));
    RC(hc.upload(weight, N, &dw));
    RC(cbs_common(dx, dw, arm_offsets, n_arms, params, capacity, seg_arm, seg_lo, seg_hi, seg_mean, n_segs, stats,
                  hc.stream));
    return hc.finish();
}

int cnvk_cbs_input_filter_dev(const double* d_log2, const double* d_weight, const uint8_t* d_outlier, const int64_t* d_arm_id,
                              const int64_t* arm_offsets, int32_t n_arms, int64_t n, double* d_log2_out, double* d_weight_out,
                              int64_t* d_src_rows, int64_t* unit_offsets, uint32_t* flags, void* stream) {
    RC(ensure_init());
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1), units((size_t)n_arms + 1, 0);
    unsigned int f = 0;
    RC(cbs_input_filter(d_log2, d_weight, d_outlier, (const long long*)d_arm_id, off.data(), n_arms, n, d_log2_out,
                        d_weight_out, (long long*)d_src_rows, units.data(), &f, (cudaStream_t)stream));
    for (int a = 0; a <= n_arms; ++a) unit_offsets[a] = units[a];
    *flags = f;
    return 0;
}

int cnvk_coverage_finalize_dev(const double* d_basecount, const int64_t* d_start, const int64_t* d_end, int64_t n,
                               double null_log2, double* d_depth, double* d_log2, void* stream) {
    RC(ensure_init());
    return coverage_finalize(d_basecount, (const long long*)d_start, (const long long*)d_end, n, null_log2, d_depth,
                             d_log2, (cudaStream_t)stream);
}

int cnvk_coverage_finalize(const double* basecount, const int64_t* start, const int64_t* end, int64_t n,
                           double null_log2, double* depth, double* log2) {
    HostCall hc;
    RC(hc.begin());
    double *db, *dd, *dl;
    long long *ds, *de;
    RC(hc.upload(basecount, (size_t)n, &db));
    RC(hc.upload((const long long*)start, (size_t)n, &ds));
    RC(hc.upload((const long long*)end, (size_t)n, &de));
    CNVK_CHECK(hc.scratch->alloc(&dd, (size_t)n));
    CNVK_CHECK(hc.scratch->alloc(&dl, (size_t)n));
    RC(coverage_finalize(db, ds, de, n, null_log2, dd, dl, hc.stream));
    RC(hc.download(depth, dd, (size_t)n));
    RC(hc.download(log2, dl, (size_t)n));
    return hc.finish();
}

int cnvk_smooth_cna_dev(const double* d_log2, const int64_t* arm_offsets, int32_t n_arms, int32_t smooth_region,
                        double outlier_sd_scale, double smooth_sd_scale, double trim, double* d_out, double* d_sd,
                        void* stream) {
    RC(ensure_init());
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1);
    return smooth_cna(d_log2, off.data(), n_arms, smooth_region, outlier_sd_scale, smooth_sd_scale, trim, d_out, d_sd,
                      (cudaStream_t)stream);
}

int cnvk_smooth_cna(const double* log2, const int64_t* arm_offsets, int32_t n_arms, int32_t smooth_region,
                    double outlier_sd_scale, double smooth_sd_scale, double trim, double* out, double* sd) {
    HostCall hc;
    RC(hc.begin());
    const size_t N = n_arms > 0 ? (size_t)arm_offsets[n_arms] : 0;
    double *dx, *dout, *dsd;
    RC(hc.upload(log2, N, &dx));
    CNVK_CHECK(hc.scratch->alloc(&dout, N));
    CNVK_CHECK(hc.scratch->alloc(&dsd, (size_t)std::max(n_arms, 1)));
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1);
    RC(smooth_cna(dx, off.data(), n_arms, smooth_region, outlier_sd_scale, smooth_sd_scale, trim, dout, dsd, hc.stream));
    RC(hc.download(out, dout, N));
    if (sd) RC(hc.download(sd, dsd, (size_t)n_arms));
    return hc.finish();
}

int cnvk_cbs_node_diag(const double* log2, const double* weight, const int64_t* node_offsets, int32_t n_nodes,
                       const cnvk_cbs_params* p, double* out6, int32_t* iout5) {
    HostCall hc;
    RC(hc.begin());
    const size_t N = n_nodes > 0 ? (size_t)node_offsets[n_nodes] : 0;
    double *dx, *dw;
    RC(hc.upload(log2, N, &dx));
    RC(hc.upload(weight, N, &dw));
    CbsParams P;
    P.alpha = p->alpha; P.nperm = p->nperm; P.min_width = p->min_width; P.kmax = p->kmax; P.nmin = p->nmin;
    P.eta = p->eta; P.hybrid = p->hybrid; P.prune = p->prune; P.seed = p->seed;
    std::vector<long long> off(node_offsets, node_offsets + n_nodes + 1);
    RC(cbs_node_diag(dx, dw, off.data(), n_nodes, P, out6, (int*)iout5, hc.stream));
    return hc.finish();
}

int64_t cnvk_panel_sizeof(void) { return (int64_t)sizeof(cnvk_panel); }

int cnvk_panel_prepare_dev(const cnvk_panel* panel, int32_t n_samples, const double* const* d_in, double* const* d_out,
                           int32_t want_cbs, double skip_outliers, double* const* d_cbs, int64_t* const* d_src,
                           int64_t* unit_offsets, uint32_t* status, void* stream) {
    RC(ensure_init());
    if (!panel || !d_in || !d_out || !status || (want_cbs && (!d_cbs || !d_src || !unit_offsets)))
    {
        set_last_error("cnvk_panel_prepare_dev: null argument");
        return -2;
    }
    return panel_prepare(panel, n_samples, d_in, d_out, want_cbs, skip_outliers, d_cbs, (long long* const*)d_src,
                         (long long*)unit_offsets, status, (cudaStream_t)stream);
}

int cnvk_panel_transfer_dev(const cnvk_panel* panel, int32_t n_samples, int32_t n_seg, const int64_t* seg,
                            const int64_t* const* d_src, const double* const* d_depth, const double* const* d_weight,
                            int64_t* bounds, double* sums, void* stream) {
    RC(ensure_init());
    if (!panel || (n_seg > 0 && (!seg || !d_src || !d_depth || !d_weight || !bounds || !sums)))
    {
        set_last_error("cnvk_panel_transfer_dev: null argument");
        return -2;
    }
    return panel_transfer(panel, n_samples, n_seg, (const long long*)seg, (const long long* const*)d_src, d_depth, d_weight,
                          (long long*)bounds, sums, (cudaStream_t)stream);
}

int cnvk_outlier_mask_dev(const double* d_log2, const int64_t* arm_offsets, int32_t n_arms, double q, double factor,
                          uint8_t* d_mask, void* stream) {
    RC(ensure_init());
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1);
    return outlier_mask(d_log2, off.data(), n_arms, q, factor, d_mask, (cudaStream_t)stream);
}

int cnvk_round_sig_dev(const double* d_x, int64_t n, int32_t digits, double* d_out, uint64_t* unhandled, void* stream) {
    RC(ensure_init());
    unsigned long long u = 0;
    int rc = round_sig(d_x, n, digits, d_out, &u, (cudaStream_t)stream);
    if (unhandled) *unhandled = u;
    return rc;
}

int cnvk_segment_bin_sums_dev(const double* d_depth, const double* d_weight, const int64_t* d_lo, const int64_t* d_hi,
                              int32_t nseg, double* d_seg_weight, double* d_seg_depth, void* stream) {
    RC(ensure_init());
    return segment_bin_sums(d_depth, d_weight, (const long long*)d_lo, (const long long*)d_hi, nullptr, nullptr, nseg,
                            d_seg_weight, d_seg_depth, (cudaStream_t)stream);
}

int cnvk_segment_bin_sums_masked_dev(const double* d_depth, const double* d_weight, const int64_t* d_lo,
                                     const int64_t* d_hi, const int64_t* d_bin_end, const int64_t* d_seg_start,
                                     int32_t nseg, double* d_seg_weight, double* d_seg_depth, void* stream) {
    RC(ensure_init());
    return segment_bin_sums(d_depth, d_weight, (const long long*)d_lo, (const long long*)d_hi,
                            (const long long*)d_bin_end, (const long long*)d_seg_start, nseg, d_seg_weight,
                            d_seg_depth, (cudaStream_t)stream);
}

static int haar_common(const double* d_x, const double* d_w, const int64_t* arm_offsets, int32_t n_arms, double q,
                       int64_t capacity, int32_t* seg_arm, int64_t* seg_lo, int64_t* seg_hi, double* seg_mean,
                       int64_t* n_segs, cudaStream_t stream, const double* sigma_override = nullptr) {
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1);
    std::vector<CbsSeg> segs;
    RC(haar_segment(d_x, d_w, off.data(), n_arms, q, segs, stream, sigma_override));
    if ((int64_t)segs.size() > capacity) {
        set_last_error("cnvk_haar_segment: %zu segments exceed capacity %lld", segs.size(), (long long)capacity);
        return -4;
    }
    for (size_t s = 0; s < segs.size(); ++s) {
        seg_arm[s] = segs[s].arm; seg_lo[s] = segs[s].lo; seg_hi[s] = segs[s].hi; seg_mean[s] = segs[s].mean;
    }
    *n_segs = (int64_t)segs.size();
    return 0;
}

int cnvk_haar_segment_dev(const double* d_log2, const double* d_weight, const int64_t* arm_offsets, int32_t n_arms,
                          double fdr_q, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo, int64_t* seg_hi,
                          double* seg_mean, int64_t* n_segs, void* stream) {
    RC(ensure_init());
    return haar_common(d_log2, d_weight, arm_offsets, n_arms, fdr_q, capacity, seg_arm, seg_lo, seg_hi, seg_mean,
                       n_segs, (cudaStream_t)stream);
}

int cnvk_haar_segment(const double* log2, const double* weight, const int64_t* arm_offsets, int32_t n_arms,
                      double fdr_q, int64_t capacity, int32_t* seg_arm, int64_t* seg_lo, int64_t* seg_hi,
                      double* seg_mean, int64_t* n_segs) {
    HostCall hc;
    RC(hc.begin());
    const size_t N = n_arms > 0 ? (size_t)arm_offsets[n_arms] : 0;
    double *dx, *dw = nullptr;
    RC(hc.upload(log2, N, &dx));
    if (weight) RC(hc.upload(weight, N, &dw));
    RC(haar_common(dx, dw, arm_offsets, n_arms, fdr_q, capacity, seg_arm, seg_lo, seg_hi, seg_mean, n_segs, hc.stream));
    return hc.finish();
}

int cnvk_haar_segment_sigma(const double* signal, const double* weight, const int64_t* arm_offsets, int32_t n_arms,
                            double fdr_q, const double* sigma_override, int64_t capacity, int32_t* seg_arm,
                            int64_t* seg_lo, int64_t* seg_hi, double* seg_mean, int64_t* n_segs) {
    HostCall hc;
    RC(hc.begin());
    const size_t N = n_arms > 0 ? (size_t)arm_offsets[n_arms] : 0;
    double *dx, *dw = nullptr;
    RC(hc.upload(signal, N, &dx));
    if (weight) RC(hc.upload(weight, N, &dw));
    RC(haar_common(dx, dw, arm_offsets, n_arms, fdr_q, capacity, seg_arm, seg_lo, seg_hi, seg_mean, n_segs, hc.stream,
                   sigma_override));
    return hc.finish();
}

int cnvk_haar_sigma(const double* signal, const int64_t* arm_offsets, int32_t n_arms, double* sigma) {
    HostCall hc;
    RC(hc.begin());
    const size_t N = n_arms > 0 ? (size_t)arm_offsets[n_arms] : 0;
    double* dx;
    RC(hc.upload(signal, N, &dx));
    std::vector<long long> off(arm_offsets, arm_offsets + n_arms + 1);
    RC(haar_sigma(dx, off.data(), n_arms, sigma, hc.stream));
    return hc.finish();
}

int cnvk_haar_segment_means(const double* signal, const double* weight, int64_t n, const int64_t* seg_lo,
                            const int64_t* seg_hi, int32_t n_segs, double* seg_mean) {
    HostCall hc;
    RC(hc.begin());
    double *dx, *dw = nullptr;
    RC(hc.upload(signal, (size_t)n, &dx));
    if (weight) RC(hc.upload(weight, (size_t)n, &dw));
    RC(haar_segment_means(dx, dw, (const long long*)seg_lo, (const long long*)seg_hi, n_segs, seg_mean, hc.stream));
    return hc.finish();
}

int cnvk_cbs_getbdry(double eta, int32_t nperm, int32_t max_ones, int32_t* out) {
    std::vector<int> b;
    cbs_getbdry(eta, nperm, max_ones, b);
    for (size_t k = 0; k < b.size(); ++k) out[k] = b[k];
    return 0;
}

}  // extern "C"
