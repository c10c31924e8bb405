// extern "C" surface of libcnvkit_b200.so -- see include/cnvkit_b200.h for the contract.
#include <stdarg.h>

#include "../../include/cnvkit_b200.h"
#include "ops.cuh"

namespace cnvk {
unsigned long long g_launch_count = 0;
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace cnvk

using namespace cnvk;

// Host-buffer helper: stage inputs, run, fetch outputs on a private stream.
struct HostCall {
    cudaStream_t stream;
    Scratch* scratch;
    HostCall() : stream(nullptr), scratch(nullptr) {}
    int begin() {
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
int cnvk_abi_version(void) { return 1; }
unsigned long long cnvk_launch_count(void) { return g_launch_count; }

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
    return rolling_window_stat(d_x, n, wing, 0, 0.5, 1, d_out, (cudaStream_t)stream);
}

int cnvk_rolling_quantile_dev(const double* d_x, int64_t n, int64_t wing, double q, double* d_out, void* stream) {
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
    return argsort_f64(d_x, n, d_order, (cudaStream_t)stream);
}

}  // extern "C"
