// Counter-based permutations for the CBS permutation tests.
//
// DNAcopy shuffles with R's Mersenne-Twister stream (cbs.py:49 set.seed(0xA5EED), Fortran
// xperm/wxperm Fisher-Yates); a sequential stream cannot be evaluated tile-parallel on a GPU and
// would make results depend on the order nodes are visited.  Instead permutation `perm` of test
// `tid` on the interval [lo,hi) of an arm is the keyed bijection below, evaluated independently
// per element: Philox4x32-10(counter=(perm,tid,lo,hi), key=seed) yields four round keys (four more
// are derived from them) for an alternating Feistel network over ceil(log2 n) (>= 8) bits,
// cycle-walked into [0,n).  Rounds by domain width: a balanced Feistel network over 2a bits is
// pairwise-uniform only up to O(2^-a) after 4 rounds (positions that share the right half keep their
// left-half difference with probability 2^-b + 2^-a instead of 2^-a) -- measurable for the narrow
// domains of the exact test (n <= 256) and gone after two to four more rounds, so domains of
// <= 10 / <= 14 / more bits get 8 / 6 / 4 rounds (tests/test_cbs_stats.py checks pair statistics).
#pragma once
#include "common.cuh"

namespace cnvk {

__host__ __device__ __forceinline__ void philox_round(u32 c[4], const u32 k[2]) {
    const u64 p0 = (u64)0xD2511F53u * c[0];
    const u64 p1 = (u64)0xCD9E8D57u * c[2];
    const u32 hi0 = (u32)(p0 >> 32), lo0 = (u32)p0;
    const u32 hi1 = (u32)(p1 >> 32), lo1 = (u32)p1;
    const u32 n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ u32 mix32(u32 h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

struct Prp {
    u32 k[8];
    int bits_a, bits_b, rounds;
    u32 n;

    __host__ __device__ __forceinline__ void init(u64 seed, u32 lo, u32 hi, u32 tid, u32 perm, u32 n_) {
        u32 c[4] = {perm, tid, lo, hi};
        u32 key[2] = {(u32)seed, (u32)(seed >> 32)};
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            if (r) { key[0] += 0x9E3779B9u; key[1] += 0xBB67AE85u; }
            philox_round(c, key);
        }
        k[0] = c[0]; k[1] = c[1]; k[2] = c[2]; k[3] = c[3];
#pragma unroll
        for (int r = 0; r < 4; ++r) k[4 + r] = mix32(k[r] + 0x9E3779B9u * (u32)(r + 1));
        int bits = 8;
        while (bits < 32 && (1ull << bits) < n_) ++bits;
        bits_a = bits / 2;
        bits_b = bits - bits / 2;
        rounds = bits <= 10 ? 8 : (bits <= 14 ? 6 : 4);
        n = n_;
    }

    // one application of the network (no cycle walking): the result may lie outside [0, n)
    __host__ __device__ __forceinline__ u32 once(u32 v) const {
        const int a = bits_a, b = bits_b;
        const u32 ma = (1u << a) - 1u, mb = (1u << b) - 1u;
        u32 L = v >> b, R = v & mb;
        u32 t = (L ^ mix32(R ^ k[0])) & ma; L = R; R = t;
        t = (L ^ mix32(R ^ k[1])) & mb; L = R; R = t;
        t = (L ^ mix32(R ^ k[2])) & ma; L = R; R = t;
        t = (L ^ mix32(R ^ k[3])) & mb; L = R; R = t;
        if (rounds > 4) {
            t = (L ^ mix32(R ^ k[4])) & ma; L = R; R = t;
            t = (L ^ mix32(R ^ k[5])) & mb; L = R; R = t;
            if (rounds > 6) {
                t = (L ^ mix32(R ^ k[6])) & ma; L = R; R = t;
                t = (L ^ mix32(R ^ k[7])) & mb; L = R; R = t;
            }
        }
        return (L << b) | R;
    }
    // cycle walking of a value produced by once(): same result as operator() on the original input
    __host__ __device__ __forceinline__ u32 walk(u32 v) const {
        while (v >= n) v = once(v);
        return v;
    }

    __host__ __device__ __forceinline__ u32 operator()(u32 v) const {
        const int a = bits_a, b = bits_b;
        const u32 ma = (1u << a) - 1u, mb = (1u << b) - 1u;
        do {
            u32 L = v >> b, R = v & mb;
            u32 t = (L ^ mix32(R ^ k[0])) & ma; L = R; R = t;
            t = (L ^ mix32(R ^ k[1])) & mb; L = R; R = t;
            t = (L ^ mix32(R ^ k[2])) & ma; L = R; R = t;
            t = (L ^ mix32(R ^ k[3])) & mb; L = R; R = t;
            if (rounds > 4) {
                t = (L ^ mix32(R ^ k[4])) & ma; L = R; R = t;
                t = (L ^ mix32(R ^ k[5])) & mb; L = R; R = t;
                if (rounds > 6) {
                    t = (L ^ mix32(R ^ k[6])) & ma; L = R; R = t;
                    t = (L ^ mix32(R ^ k[7])) & mb; L = R; R = t;
                }
            }
            v = (L << b) | R;
        } while (v >= n);
        return v;
    }
};

}  // namespace cnvk
