// Device-wide primitives written for this library: stable LSD radix sort of
// (u64 key, u32 value) pairs and a u32 exclusive scan.  All launches go to the
// caller's stream; temporary storage comes from the stream-ordered pool.
#pragma once
#include "common.cuh"

namespace cnvk {

// Sort pairs by key bits [begin_bit, end_bit) (multiples of 8), stable.
// keys/vals are ping-ponged with the *_alt buffers; on return *out_keys/*out_vals
// point at whichever buffer holds the result.
int radix_sort_pairs(u64* keys, u64* keys_alt, u32* vals, u32* vals_alt, long long n, int begin_bit,
                     int end_bit, u64** out_keys, u32** out_vals, cudaStream_t stream);

// Stable argsort of float64 keys with numpy semantics (NaN last, -0.0 == 0.0).
// d_order[n] receives the permutation; d_sorted_keys (optional) the ordered u64 keys.
int argsort_f64(const double* d_x, long long n, u32* d_order, cudaStream_t stream);

// keys[i] = orderable(x[i]), idx[i] = i
int make_sort_keys(const double* d_x, long long n, u64* d_keys, u32* d_idx, cudaStream_t stream);

// ascending copy of d_x with NaNs last (scratch k0,k1,v0,v1 of n elements each)
int sort_values_f64(const double* d_x, long long n, u64* k0, u64* k1, u32* v0, u32* v1, double* d_sorted,
                    u32* d_nvalid, cudaStream_t stream);

// out[i] = sum_{j<i} in[j]; total written to *d_total if non-null. In-place allowed.
int exclusive_scan_u32(const u32* d_in, u32* d_out, long long n, u32* d_total, cudaStream_t stream);

}  // namespace cnvk
