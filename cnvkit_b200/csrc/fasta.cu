// GC and soft-masked (lowercase = RepeatMasker) fraction of genomic intervals, counted straight from
// the bytes of a FASTA file resident in HBM: cnvlib/reference.py:859-880 get_fasta_stats /
// calculate_gc_lo.  The reference slices every bin's sequence out with pyfaidx and runs eight
// str.count passes over it; only the letters a/c/g/t in either case are counted, so line breaks
// (and N, IUPAC codes, anything else) drop out by themselves and a bin is simply a byte range of the
// file -- [offset(start), offset(end)) under the .fai line geometry, computed by the host.
//   frac_gc = (#g + #c + #G + #C) / #acgtACGT,  frac_lo = (#a + #c + #g + #t) / #acgtACGT,  0/0 -> 0.
// One warp per range; 16-byte vector loads on the aligned interior, byte loads on the ragged ends;
// every byte is read once (HBM bound, 1 B/base).
#include "ops.cuh"
#include "prof.cuh"

namespace cnvk {

__device__ __forceinline__ void classify_word(u32 w, int& n_gc, int& n_tot, int& n_lo) {
    const u32 x = w | 0x20202020u;       // fold case: only 'A'/'a' map to 'a', etc.
    const u32 at = __vcmpeq4(x, 0x61616161u) | __vcmpeq4(x, 0x74747474u);
    const u32 gc = __vcmpeq4(x, 0x67676767u) | __vcmpeq4(x, 0x63636363u);
    const u32 any = at | gc;
    const u32 low = __vcmpeq4(w & 0x20202020u, 0x20202020u);
    n_gc += __popc(gc) >> 3;
    n_tot += __popc(any) >> 3;
    n_lo += __popc(any & low) >> 3;
}

__device__ __forceinline__ void classify_byte(unsigned char c, int& n_gc, int& n_tot, int& n_lo) {
    const unsigned char x = c | 0x20;
    const bool gc = (x == 'g') | (x == 'c');
    const bool any = gc | (x == 'a') | (x == 't');
    n_gc += gc;
    n_tot += any;
    n_lo += any & ((c & 0x20) != 0);
}

__global__ void __launch_bounds__(256)
fasta_gc_lo_kernel(const unsigned char* __restrict__ bytes, long long nbytes, const long long* __restrict__ r_lo,
                   const long long* __restrict__ r_hi, long long n, double* __restrict__ gc_out,
                   double* __restrict__ lo_out) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    long long lo = r_lo[r], hi = r_hi[r];
    lo = max(0ll, min(lo, nbytes));
    hi = max(lo, min(hi, nbytes));
    int n_gc = 0, n_tot = 0, n_lo = 0;
    const unsigned long long a_lo = (unsigned long long)(bytes + lo), a_hi = (unsigned long long)(bytes + hi);
    const unsigned long long v0 = (a_lo + 15ull) & ~15ull, v1 = a_hi & ~15ull;
    if (v0 >= v1) {
        for (unsigned long long p = a_lo + lane; p < a_hi; p += 32) classify_byte(*(const unsigned char*)p, n_gc, n_tot, n_lo);
    } else {
        if (lane < 16) {
            const unsigned long long p = a_lo + lane;
            if (p < v0) classify_byte(*(const unsigned char*)p, n_gc, n_tot, n_lo);
        } else {
            const unsigned long long p = v1 + (lane - 16);
            if (p < a_hi) classify_byte(*(const unsigned char*)p, n_gc, n_tot, n_lo);
        }
        for (unsigned long long v = v0 + 16ull * lane; v < v1; v += 16ull * 32) {
            const uint4 q = __ldg((const uint4*)v);
            classify_word(q.x, n_gc, n_tot, n_lo);
            classify_word(q.y, n_gc, n_tot, n_lo);
            classify_word(q.z, n_gc, n_tot, n_lo);
            classify_word(q.w, n_gc, n_tot, n_lo);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_gc += __shfl_xor_sync(0xffffffffu, n_gc, o);
        n_tot += __shfl_xor_sync(0xffffffffu, n_tot, o);
        n_lo += __shfl_xor_sync(0xffffffffu, n_lo, o);
    }
    if (lane == 0) {
        gc_out[r] = n_tot ? __ddiv_rn((double)n_gc, (double)n_tot) : 0.0;
        lo_out[r] = n_tot ? __ddiv_rn((double)n_lo, (double)n_tot) : 0.0;
    }
}

int fasta_gc_lo(const unsigned char* d_bytes, long long nbytes, const long long* d_lo, const long long* d_hi,
                long long n, double* d_gc, double* d_rm, cudaStream_t stream) {
    if (n <= 0) return 0;
    ProfScope prof(PC_OTHER, stream);
    fasta_gc_lo_kernel<<<div_up(n, 8), 256, 0, stream>>>(d_bytes, nbytes, d_lo, d_hi, n, d_gc, d_rm);
    CNVK_LAUNCH_CHECK();
    return 0;
}

}  // namespace cnvk
