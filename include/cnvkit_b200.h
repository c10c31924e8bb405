/* cnvkit_b200 -- C ABI of the B200-native CNVkit hot path (libcnvkit_b200.so).
 *
 * The reference (etal/cnvkit) is pure Python and has no FFI layer; the seam it offers is a set
 * of Python callables.  Every entry point below replaces the numeric body of one of those
 * callables (file:line cited per function, relative to the reference tree) and is what a
 * ctypes/cffi stub inside the reference would bind -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / pandas types.
 *   - `*_dev` functions take DEVICE pointers and a cudaStream_t passed as void* (0 = legacy
 *     default stream); work is enqueued on that stream and the call returns without
 *     synchronising unless documented.  Functions without the suffix take HOST pointers, do
 *     their own H2D/D2H copies and return after the result is in host memory.
 *   - float columns are float64, coordinate/count columns int64, row order is the caller's.
 *   - return value 0 = success; >0 = cudaError_t; <0 = argument/shape error.  The message of
 *     the last failure on this thread is available from cnvk_last_error().
 *   - there is NO CPU fallback: every function needs a CUDA device and fails otherwise.
 */
#ifndef CNVKIT_B200_H
#define CNVKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ----------------------------------------------------------------------- */
const char* cnvk_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int cnvk_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench "gpu_launches"). */
unsigned long long cnvk_launch_count(void);
/* Device properties the host layer wants without importing torch: SM count, free/total bytes. */
int cnvk_device_info(int device, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes);

/* ---- smoothing --------------------------------------------------------------------- */
/* cnvlib/smoothing.py:77-87 rolling_median  (pandas rolling(2*wing+1, 1, center=True).median()
 * over the mirror padding of smoothing.py:72-74).  `wing` is the integer half-window the host
 * derives with smoothing._width2wing (smoothing.py:53-69); requires n >= 2, 1 <= wing <= n-1.
 * NaN inputs are skipped as pandas does (even valid counts average the two middle values;
 * an all-NaN window yields NaN). */
int cnvk_rolling_median_dev(const double* d_x, int64_t n, int64_t wing, double* d_out, void* stream);
int cnvk_rolling_median(const double* x, int64_t n, int64_t wing, double* out);

/* cnvlib/smoothing.py:135-141 rolling_quantile (pandas rolling(2*wing+1, 2, center=True)
 * .quantile(q), linear interpolation) over the same mirror padding. */
int cnvk_rolling_quantile_dev(const double* d_x, int64_t n, int64_t wing, double q, double* d_out,
                              void* stream);
int cnvk_rolling_quantile(const double* x, int64_t n, int64_t wing, double q, double* out);

/* ---- sorting ------------------------------------------------------------------------ */
/* np.argsort(x, kind="mergesort") as used by cnvlib/fix.py:414 -- stable, NaN last. */
int cnvk_argsort_f64_dev(const double* d_x, int64_t n, uint32_t* d_order, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNVKIT_B200_H */
