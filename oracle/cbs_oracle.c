/* TEST INFRASTRUCTURE ONLY -- CPU restatement of DNAcopy's weighted circular binary
 * segmentation as CNVkit invokes it.  Never linked into or called by the product.
 *
 * ======================= PARITY UNPINNED ==========================================
 * The arithmetic lives in a third-party dependency that is NOT in /root/reference:
 * Bioconductor `DNAcopy` (version unpinned by the reference: conda-env.yml:15,
 * environment-dev.yml:20, devtools/conda-recipe/meta.yaml:20), reached through an Rscript
 * subprocess (cnvlib/segmentation/__init__.py:292-322, script cnvlib/segmentation/cbs.py:1-78).
 * Neither R nor DNAcopy (nor their sources) exist in this container, and the reference's own
 * tests hold no CBS golden vectors (test/test_r.py asserts structure only).  This file
 * therefore restates the *published* algorithm -- Olshen et al. 2004, Venkatraman & Olshen
 * 2007, and DNAcopy's R/Fortran control flow (segment -> changepoints -> wfindcpt ->
 * wtmaxo / getmncwt / hwtmaxp / wtmaxp / tailp / wtpermp / getbdry) -- anchored on the reference's call
 * site: segment(CNA(log2, chrom, start, "logratio", presorted=T), weights=w, alpha=threshold)
 * with every other argument at DNAcopy's default (nperm=10000, p.method="hybrid",
 * min.width=2, kmax=25, nmin=200, eta=0.05, undo.splits="none").
 *
 * Deliberate, documented deviations from a bit-faithful DNAcopy:
 *  (1) RNG.  R's Mersenne-Twister stream cannot be reproduced; the north star only asks that
 *      permutation p-values agree within Monte-Carlo error.  Permutations here are
 *      counter-based: permutation p of the test `tid` on node [lo,hi) is the bijection
 *      PRP(seed, lo, hi, tid, p) on [0,n) (Philox4x32-10 round keys driving an alternating
 *      Feistel network -- 8 / 6 / 4 rounds for domains of <= 10 / <= 14 / more bits -- with cycle walking).  Results are therefore independent
 *      of the order in which nodes are visited -- which is what lets the GPU process the
 *      recursion level-synchronously and still agree with this file exactly.
 *  (2) R computes sum()/cumsum() in 80-bit extended precision.  Here the three R-level totals
 *      (sum w, sum x*w, sum w*xc^2) are compensated double-double sums rounded once -- the
 *      correctly rounded exact sum, which is what R's long-double accumulation yields to within an
 *      ulp -- while cumsum(w) and DNAcopy's partial sums S_i are a blocked FP64 prefix sum with one
 *      fixed association (`blocked_prefix` below, replayed operation by operation by the CUDA prep
 *      kernel); it equals the left-to-right Fortran loop up to rounding.
 *  (3) wtmaxo's block-pruned search is restated as an exhaustive scan over all arcs; the
 *      maximum is the same, ties resolve to the smallest (i, j).
 *  (4) hwtmaxp (hybrid permutation statistic) is restated as windowed sums over circular
 *      arcs of width min.width..kmax, accumulated element by element from the arc start.
 *  (5) wtmaxp (exact permutation statistic, nodes of <= nmin probes): the partial sums of the permuted
 *      terms use one fixed blocked association (32 runs of ceil(n/32) terms, run totals accumulated left
 *      to right; node_perm_stat_buf) instead of the Fortran left-to-right loop -- last-place differences,
 *      chosen so that a warp can produce the same bits with one run per lane.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double alpha;
    int nperm;
    int min_width;
    int kmax;
    int nmin;
    double eta;
    int hybrid; /* p.method == "hybrid" */
    int ngrid;
    double tol;
    uint64_t seed;
    int prune; /* 1: block-pruned search for the observed maximum (same result as the exhaustive scan) */
} cbso_params;

typedef struct {
    long long nodes;          /* intervals tested */
    long long obs_arcs;       /* arcs evaluated for observed statistics */
    long long perm_arcs;      /* arcs evaluated inside permutation statistics */
    long long perm_tests;     /* nodes that needed the permutation reference distribution */
    long long perms_run;      /* permutations actually generated (early stopping) */
    long long tperm_tests;    /* ternary-split flank tests that ran permutations */
} cbso_stats;

/* ------------------------------------------------------------------ RNG / permutation */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        philox_round(c, k);
    }
    memcpy(out, c, sizeof(c));
}

static inline uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

typedef struct { uint32_t k[8]; int bits_a, bits_b, rounds; uint32_t n; } prp_t;

/* Round count by domain width: a balanced Feistel network over 2a bits is only pairwise-uniform up to
 * O(2^-a) after 4 rounds (two positions that share the right half land on values whose left halves differ
 * by their own left-half difference with probability 2^-b + 2^-a instead of 2^-a) -- measurable for the
 * narrow domains of the exact test (n <= 256: a = b = 4) and gone after two to four more rounds
 * (tests/test_cbs_stats.py: pair statistics of 4e5 permutations).  Wide domains keep 4 rounds. */
static inline int prp_rounds(int bits) { return bits <= 10 ? 8 : (bits <= 14 ? 6 : 4); }

static void prp_init(prp_t* P, uint64_t seed, uint32_t lo, uint32_t hi, uint32_t tid, uint32_t perm, uint32_t n) {
    const uint32_t ctr[4] = {perm, tid, lo, hi};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key, P->k);
    for (int r = 0; r < 4; ++r) P->k[4 + r] = mix32(P->k[r] + 0x9E3779B9u * (uint32_t)(r + 1));
    int bits = 8;
    while (bits < 32 && (1ull << bits) < n) ++bits;
    P->bits_a = bits / 2;          /* left half  */
    P->bits_b = bits - bits / 2;   /* right half */
    P->rounds = prp_rounds(bits);
    P->n = n;
}

/* alternating Feistel on bits_a+bits_b bits (4, 6 or 8 rounds), cycle-walked into [0, n). */
static inline uint32_t prp_eval(const prp_t* P, uint32_t v) {
    const int a = P->bits_a, b = P->bits_b;
    const uint32_t ma = (1u << a) - 1u, mb = (1u << b) - 1u;
    do {
        uint32_t L = v >> b, R = v & mb;           /* L: a bits, R: b bits */
        for (int r = 0; r < P->rounds; r += 2) {
            /* (L:a, R:b) -> (R:b, L ^ F(R):a) -> (L:a, R:b) */
            uint32_t t = (L ^ mix32(R ^ P->k[r])) & ma; L = R; R = t;
            t = (L ^ mix32(R ^ P->k[r + 1])) & mb; L = R; R = t;
        }
        v = (L << b) | R;
    } while (v >= P->n);
    return v;
}

uint32_t cbso_prp(uint64_t seed, uint32_t lo, uint32_t hi, uint32_t tid, uint32_t perm, uint32_t n, uint32_t t) {
    prp_t P;
    prp_init(&P, seed, lo, hi, tid, perm, n);
    return prp_eval(&P, t);
}

/* permutations perm0 .. perm0+nperms-1 of the test (lo, hi, tid), row-major out[nperms][n] */
void cbso_prp_fill(uint64_t seed, uint32_t lo, uint32_t hi, uint32_t tid, uint32_t perm0, uint32_t nperms,
                   uint32_t n, uint32_t* out) {
    for (uint32_t p = 0; p < nperms; ++p) {
        prp_t P;
        prp_init(&P, seed, lo, hi, tid, perm0 + p, n);
        for (uint32_t t = 0; t < n; ++t) out[(size_t)p * n + t] = prp_eval(&P, t);
    }
}

/* the same for selected positions only: out[nperms][npos] */
void cbso_prp_at(uint64_t seed, uint32_t lo, uint32_t hi, uint32_t tid, uint32_t perm0, uint32_t nperms,
                 uint32_t n, const uint32_t* pos, uint32_t npos, uint32_t* out) {
    for (uint32_t p = 0; p < nperms; ++p) {
        prp_t P;
        prp_init(&P, seed, lo, hi, tid, perm0 + p, n);
        for (uint32_t t = 0; t < npos; ++t) out[(size_t)p * npos + t] = prp_eval(&P, pos[t]);
    }
}

/* ------------------------------------------------------------------ tail probability */
static double pnorm_std(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* Siegmund's nu(x): DNAcopy tailprobs.f `nu` */
static double nu_fn(double x, double tol) {
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

/* integral of 1/(t(1-t))^2 from x to x+a: DNAcopy tailprobs.f `it1tsq` */
static double it1tsq(double x, double a) {
    double y = x + a - 0.5;
    double r = (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    y = x - 0.5;
    r = r - (8.0 * y) / (1.0 - 4.0 * y * y) - 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    return r;
}

/* DNAcopy tailprobs.f `tailp`: P(max over arcs with fraction in [delta, 1/2] > b), two-sided */
double cbso_tailp(double b, double delta, int m, int ngrid, double tol) {
    const double dincr = (0.5 - delta) / (double)ngrid;
    const double bsqrtm = b / sqrt((double)m);
    double tl = 0.5 - dincr, t = 0.5 - 0.5 * dincr, p = 0.0;
    for (int i = 0; i < ngrid; ++i) {
        const double x = bsqrtm / sqrt(t * (1.0 - t));
        const double nux = nu_fn(x, tol);
        p += (nux * nux) * it1tsq(tl, dincr);
        tl -= dincr;
        t -= dincr;
    }
    p = 9.973557e-2 * (b * b * b) * exp(-(b * b) / 2.0) * p;
    return 2.0 * p;
}

/* ------------------------------------------------------------------ sequential boundary */
/* DNAcopy getbdry.f `etabdry`: for k = 0..j-1 the first i with P(X_i <= k) <= eta0, where X_i is
 * the number of the j "ones" that fall among the first i of nperm permutations (hypergeometric;
 * the distribution is advanced draw by draw instead of calling phyper at every i). */
static void etabdry(int nperm, double eta0, int j, int* ibdry) {
    double* dist = (double*)calloc((size_t)j + 1, sizeof(double));
    dist[0] = 1.0;
    int k = 0;
    for (int i = 1; i <= nperm && k < j; ++i) {
        for (int x = (i < j ? i : j); x >= 0; --x) {
            const double stay = dist[x] * (1.0 - (double)(j - x) / (double)(nperm - i + 1));
            const double up = (x > 0) ? dist[x - 1] * ((double)(j - (x - 1)) / (double)(nperm - i + 1)) : 0.0;
            dist[x] = stay + up;
        }
        double cdf = 0.0;
        for (int x = 0; x <= k; ++x) cdf += dist[x];
        if (cdf <= eta0) {
            ibdry[k] = i;
            ++k;
        }
    }
    for (; k < j; ++k) ibdry[k] = nperm; /* unreachable for sane eta0 */
    free(dist);
}

/* P(the count path touches the boundary) when exactly j ones are placed uniformly among
 * nperm permutations: first-crossing DP (equals getbdry.f `pexceed`'s closed forms). */
static double pexceed(int nperm, int j, const int* ibdry) {
    double* alive = (double*)calloc((size_t)j + 1, sizeof(double));
    alive[0] = 1.0;
    double crossed = 0.0;
    int kb = 0; /* next boundary index to check */
    for (int i = 1; i <= nperm; ++i) {
        /* draw i: is it a one?  P = (j - x)/(nperm - i + 1) given x ones so far */
        for (int x = (i < j ? i : j); x >= 0; --x) {
            const double stay = alive[x] * (1.0 - (double)(j - x) / (double)(nperm - i + 1));
            const double up = (x > 0) ? alive[x - 1] * ((double)(j - (x - 1)) / (double)(nperm - i + 1)) : 0.0;
            alive[x] = stay + up;
        }
        while (kb < j && ibdry[kb] == i) {
            for (int x = 0; x <= kb; ++x) { crossed += alive[x]; alive[x] = 0.0; }
            ++kb;
        }
    }
    free(alive);
    return crossed;
}

/* DNAcopy getbdry: out has max_ones*(max_ones+1)/2 entries */
int cbso_getbdry(double eta, int nperm, int max_ones, int* out) {
    const double tol = 1e-2;
    int l = 0;
    out[0] = nperm - (int)((double)nperm * eta);
    double eta0 = eta;
    l = 1;
    for (int j = 2; j <= max_ones; ++j) {
        int* b = out + l;
        double etahi = eta0 * 1.1, etalo = eta0 * 0.25, phi, plo;
        etabdry(nperm, etahi, j, b);
        phi = pexceed(nperm, j, b);
        etabdry(nperm, etalo, j, b);
        plo = pexceed(nperm, j, b);
        while ((etahi - etalo) / etalo > tol) {
            eta0 = etalo + (etahi - etalo) * (eta - plo) / (phi - plo);
            etabdry(nperm, eta0, j, b);
            const double pex = pexceed(nperm, j, b);
            if (pex > eta) { etahi = eta0; phi = pex; } else { etalo = eta0; plo = pex; }
        }
        l += j;
    }
    return l;
}

/* ------------------------------------------------------------------ exact-ish sums */
typedef struct { double hi, lo; } dd_t;
static inline void dd_add(dd_t* a, double p) { /* Knuth two-sum; lo collects the rounding errors */
    const double s = a->hi + p;
    const double bb = s - a->hi;
    const double e = (a->hi - (s - bb)) + (p - bb);
    a->hi = s;
    a->lo += e;
}

/* Blocked FP64 prefix sum with the fixed association the CUDA prep kernels use (csrc/cbs.cu):
 * tiles of 2048 terms; "thread" j of 256 owns terms 8j..8j+7 and sums them left to right (p); the 256
 * thread totals are scanned per "warp" of 32 with the Kogge-Stone steps 1,2,4,8,16 (A; E = exclusive);
 * warp totals are accumulated left to right (X); local = (X[warp] + E[thread]) + p[thread, e]; the tile
 * total T is the local value of the tile's last slot (missing terms of a short tile count as 0.0); the
 * carry of tile k is C[k] = T[0] + ... + T[k-1] left to right; out = C[k] + local (tile 0: out = local).
 * Same value as a left-to-right sum up to rounding (smaller error growth); what matters is that both
 * sides do exactly these operations. */
static void blocked_prefix(const double* t, int n, double* out) {
    enum { PE = 8, NT = 256, TILE = PE * NT };
    static _Thread_local double p[TILE], A[NT], tmp[32];
    double carry = 0.0;
    for (int t0 = 0, tile = 0; t0 < n; t0 += TILE, ++tile) {
        const int m = (n - t0 < TILE) ? n - t0 : TILE;
        for (int j = 0; j < NT; ++j) {
            double r = 0.0;
            for (int e = 0; e < PE; ++e) {
                const int k = PE * j + e;
                const double v = (k < m) ? t[t0 + k] : 0.0;
                r = e ? r + v : v;
                p[k] = r;
            }
            A[j] = r;
        }
        for (int w = 0; w < NT / 32; ++w)
            for (int o = 1; o < 32; o <<= 1) {
                for (int l = 0; l < 32; ++l) tmp[l] = A[32 * w + l];
                for (int l = o; l < 32; ++l) A[32 * w + l] = tmp[l] + tmp[l - o];
            }
        double T = 0.0;
        for (int w = 0; w < NT / 32; ++w) {
            double X = 0.0;
            for (int k = 0; k < w; ++k) X = k ? X + A[32 * k + 31] : A[31];
            for (int l = 0; l < 32; ++l) {
                const int j = 32 * w + l;
                const double E = l ? A[j - 1] : 0.0;
                const double base = X + E;
                for (int e = 0; e < PE; ++e) {
                    const int k = PE * j + e;
                    const double loc = base + p[k];
                    if (k < m) out[t0 + k] = tile ? carry + loc : loc;
                    if (k == TILE - 1) T = loc;
                }
            }
        }
        carry = tile ? carry + T : T;
    }
}

/* ------------------------------------------------------------------ node statistics */
typedef struct {
    int n;
    double* xc;   /* centred data            */
    double* w;    /* weights                 */
    double* rw;   /* sqrt(w)                 */
    double* y;    /* xc * rw (exchangeable)  */
    double* cw;   /* cumsum(w)/sqrt(sum w), cw[0..n-1]; cwp(i)= i? cw[i-1] : 0 */
    double* S;    /* prefix sums, S[0..n]    */
    double* t;    /* scratch terms           */
    double tss;
    double total; /* cw[n-1] */
} node_t;

static void node_alloc(node_t* N, int n) {
    N->n = n;
    N->xc = (double*)malloc(sizeof(double) * n);
    N->w = (double*)malloc(sizeof(double) * n);
    N->rw = (double*)malloc(sizeof(double) * n);
    N->y = (double*)malloc(sizeof(double) * n);
    N->cw = (double*)malloc(sizeof(double) * n);
    N->S = (double*)malloc(sizeof(double) * (n + 1));
    N->t = (double*)malloc(sizeof(double) * n);
}
static void node_free(node_t* N) {
    free(N->xc); free(N->w); free(N->rw); free(N->y); free(N->cw); free(N->S); free(N->t);
}

static inline double cwp(const node_t* N, int i) { return i ? N->cw[i - 1] : 0.0; }

/* max over linear arcs (i<j), al0 <= j-i <= n-al0, of (S_j-S_i)^2 / (c (total-c)), c = cwp_j-cwp_i.
 * First maximum in (i, j) lexicographic order. */
static double max_arc_full(const node_t* N, const double* S, int al0, int* bi, int* bj, long long* arcs) {
    const int n = N->n;
    double best = -1.0, bestlo = -1.0;
    int ai = 0, aj = 0;
    for (int i = 0; i <= n - al0; ++i) {
        const double Si = S[i], ci = cwp(N, i);
        int jhi = i + (n - al0);
        if (jhi > n) jhi = n;
        for (int j = i + al0; j <= jhi; ++j) {
            const double d = S[j] - Si;
            const double c = cwp(N, j) - ci;
            const double den = c * (N->total - c);
            const double num = d * d;
            if (num > bestlo * den) { /* conservative prefilter; exact test below */
                const double q = num / den;
                if (q > best) { best = q; bestlo = q * (1.0 - 1e-12); ai = i; aj = j; }
            }
        }
        if (arcs) *arcs += (jhi >= i + al0) ? (jhi - (i + al0) + 1) : 0;
    }
    *bi = ai; *bj = aj;
    return best < 0.0 ? 0.0 : best;
}

/* Same maximum (and the same (i, j) on ties) as max_arc_full, found the way DNAcopy's wtmaxo does
 * it: prefix points are grouped in blocks, the statistic of a block pair is bounded from the block
 * minima / maxima of S and the extreme arc weights, and pairs that cannot reach the current best
 * are skipped.  Used for the CPU baseline timing; tests check it against the exhaustive scan. */
static double max_arc_pruned(const node_t* N, const double* S, int al0, int* bi, int* bj, long long* arcs) {
    enum { B = 32 };
    const int n = N->n, P = n + 1, nb = (P + B - 1) / B;
    double* mn = (double*)malloc(sizeof(double) * nb);
    double* mx = (double*)malloc(sizeof(double) * nb);
    int* imn = (int*)malloc(sizeof(int) * nb);
    int* imx = (int*)malloc(sizeof(int) * nb);
    for (int b = 0; b < nb; ++b) {
        const int p0 = b * B, p1 = (p0 + B - 1 < n) ? p0 + B - 1 : n;
        mn[b] = mx[b] = S[p0]; imn[b] = imx[b] = p0;
        for (int p = p0 + 1; p <= p1; ++p) {
            if (S[p] < mn[b]) { mn[b] = S[p]; imn[b] = p; }
            if (S[p] > mx[b]) { mx[b] = S[p]; imx[b] = p; }
        }
    }
    double best = -1.0;
    int ai = 0x7fffffff, aj = 0x7fffffff;
#define CBSO_TRY(I, J)                                                                         \
    do {                                                                                       \
        const int i_ = (I), j_ = (J);                                                          \
        const int kw_ = j_ - i_;                                                               \
        if (kw_ >= al0 && kw_ <= n - al0) {                                                    \
            const double d_ = S[j_] - S[i_];                                                   \
            const double c_ = cwp(N, j_) - cwp(N, i_);                                         \
            const double q_ = (d_ * d_) / (c_ * (N->total - c_));                              \
            if (q_ > best || (q_ == best && (i_ < ai || (i_ == ai && j_ < aj)))) {            \
                best = q_; ai = i_; aj = j_;                                                   \
            }                                                                                  \
        }                                                                                      \
    } while (0)
    /* seed: arcs between block extremes */
    for (int a = 0; a < nb; ++a)
        for (int b = a; b < nb; ++b) {
            int i = imn[a], j = imx[b];
            if (i > j) { const int t = i; i = j; j = t; }
            CBSO_TRY(i, j);
            i = imx[a]; j = imn[b];
            if (i > j) { const int t = i; i = j; j = t; }
            CBSO_TRY(i, j);
        }
    long long scanned = 0;
    for (int a = 0; a < nb; ++a) {
        const int i0 = a * B, i1 = (i0 + B - 1 < n) ? i0 + B - 1 : n;
        for (int b = a; b < nb; ++b) {
            const int j0 = b * B, j1 = (j0 + B - 1 < n) ? j0 + B - 1 : n;
            if (b > a + 1 && best >= 0.0) {
                const double d1 = fabs(mx[b] - mn[a]), d2 = fabs(mx[a] - mn[b]);
                const double dm = d1 > d2 ? d1 : d2;
                const double cmin = cwp(N, j0) - cwp(N, i1), cmax = cwp(N, j1) - cwp(N, i0);
                const double f1 = cmin * (N->total - cmin), f2 = cmax * (N->total - cmax);
                const double fm = f1 < f2 ? f1 : f2;
                if (fm > 0.0 && (dm * dm) / fm * (1.0 + 1e-9) < best) continue;
            }
            for (int i = i0; i <= i1; ++i)
                for (int j = (j0 > i + al0 ? j0 : i + al0); j <= j1; ++j) CBSO_TRY(i, j);
            scanned += (long long)(i1 - i0 + 1) * (j1 - j0 + 1);
        }
    }
#undef CBSO_TRY
    if (arcs) *arcs += scanned;
    free(mn); free(mx); free(imn); free(imx);
    if (best < 0.0) { *bi = 0; *bj = 0; return 0.0; }
    *bi = ai; *bj = aj;
    return best;
}

static double t2_of(double bss, double tss, int n) {
    if (tss <= bss + 0.0001) tss = bss + 1.0;
    return bss / ((tss - bss) / ((double)n - 2.0));
}

/* hybrid permutation statistic: circular arcs of width al0..kmax over terms t[] */
static double perm_stat_hybrid(const node_t* N, const double* wn, const double* t, int al0, int kmax, long long* arcs) {
    const int n = N->n;
    double best = 0.0;
    for (int s = 0; s < n; ++s) {
        double acc = 0.0, cacc = 0.0;
        int pos = s;
        for (int k = 1; k <= kmax; ++k) {
            acc += t[pos];
            cacc += wn[pos];
            if (++pos == n) pos = 0;
            if (k >= al0) {
                const double q = (acc * acc) / (cacc * (N->total - cacc));
                if (q > best) best = q;
            }
        }
    }
    if (arcs) *arcs += (long long)n * (kmax - al0 + 1);
    return best;
}

/* ------------------------------------------------------------------ ternary-split flank test */
static double wtpermp_impl(const double* x, const double* w, const double* rw, int n1, int n2, const cbso_params* P,
                           uint32_t lo, uint32_t hi, uint32_t tid, cbso_stats* st, int force) {
    const int n = n1 + n2;
    if (n1 == 1 || n2 == 1) return 1.0;
    double xsum1 = 0, xsum2 = 0, tss = 0, rn1 = 0, rn2 = 0;
    for (int i = 0; i < n1; ++i) { xsum1 += w[i] * x[i]; tss += w[i] * (x[i] * x[i]); rn1 += w[i]; }
    for (int i = n1; i < n; ++i) { xsum2 += w[i] * x[i]; tss += w[i] * (x[i] * x[i]); rn2 += w[i]; }
    const double rn = rn1 + rn2;
    const double xbar = (xsum1 + xsum2) / rn;
    tss = tss - rn * (xbar * xbar);
    int m1, s0;
    double rm1, odiff;
    if (n1 <= n2) { m1 = n1; s0 = 0; rm1 = rn1; odiff = fabs(xsum1 / rn1 - xbar); }
    else { m1 = n2; s0 = n1; rm1 = rn2; odiff = fabs(xsum2 / rn2 - xbar); }
    const double ostat = 0.99999 * odiff;
    double tstat = rn * rm1 * (odiff * odiff) / (rn - rm1);
    tstat = tstat / ((tss - tstat) / ((double)n - 2.0));
    if (tstat > 25.0 && m1 >= 10 && !force) return 0.0;
    if (st) st->tperm_tests++;
    double* y = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) y[i] = x[i] * rw[i];
    int nrej = 0;
    for (int p = 0; p < P->nperm; ++p) {
        prp_t prp;
        prp_init(&prp, P->seed, lo, hi, tid, (uint32_t)p, (uint32_t)n);
        double xs = 0.0;
        for (int k = 0; k < m1; ++k) xs += y[prp_eval(&prp, (uint32_t)(s0 + k))] * rw[s0 + k];
        const double pstat = fabs(xs / rm1 - xbar);
        if (ostat <= pstat) ++nrej;
    }
    free(y);
    return (double)nrej / (double)P->nperm;
}
static double wtpermp(const double* x, const double* w, const double* rw, int n1, int n2, const cbso_params* P,
                      uint32_t lo, uint32_t hi, uint32_t tid, cbso_stats* st) {
    return wtpermp_impl(x, w, rw, n1, n2, P, lo, hi, tid, st, 0);
}
static double wtpermp_forced(const double* x, const double* w, const double* rw, int n1, int n2, const cbso_params* P,
                             uint32_t lo, uint32_t hi, uint32_t tid) {
    return wtpermp_impl(x, w, rw, n1, n2, P, lo, hi, tid, NULL, 1);
}

/* ------------------------------------------------------------------ wfindcpt */
/* Node set-up shared by find_cpt and the diagnostic hook: centring, R-level totals, prefix sums.
 * Returns 0 when the data are constant (changepoints(): no change point), else 1 with N allocated. */
static int node_build(const double* x, const double* w, int n, node_t* Nout, double* rsw_out) {
    /* isTRUE(all.equal(diff(range(x)), 0)) -> no change point */
    double mn = x[0], mx = x[0];
    for (int i = 1; i < n; ++i) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
    if (mx - mn <= 1.4901161193847656e-08) return 0;

    node_t N;
    node_alloc(&N, n);
    /* sum(w), sum(x*w), sum(w*xc^2) are R-level sum() calls (80-bit accumulators in R): evaluated as
     * compensated double-double sums, i.e. the correctly rounded exact sum, independent of order */
    dd_t a_w = {0.0, 0.0}, a_wx = {0.0, 0.0}, a_tss = {0.0, 0.0};
    for (int i = 0; i < n; ++i) { dd_add(&a_w, w[i]); dd_add(&a_wx, x[i] * w[i]); }
    const double sw = a_w.hi + a_w.lo, swx = a_wx.hi + a_wx.lo;
    const double avg = swx / sw;
    const double rsw = sqrt(sw);
    N.S[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        N.w[i] = w[i];
        N.xc[i] = x[i] - avg;
        N.rw[i] = sqrt(w[i]);
        N.y[i] = N.xc[i] * N.rw[i];
        dd_add(&a_tss, w[i] * (N.xc[i] * N.xc[i]));
        N.t[i] = N.xc[i] * w[i];
    }
    blocked_prefix(N.w, n, N.cw);              /* cumsum(w) ... */
    for (int i = 0; i < n; ++i) N.cw[i] = N.cw[i] / rsw;   /* ... / sqrt(sum w) */
    blocked_prefix(N.t, n, N.S + 1);           /* partial sums S_1..S_n of xc*w */
    N.tss = a_tss.hi + a_tss.lo;
    N.total = N.cw[n - 1];
    *Nout = N;
    *rsw_out = rsw;
    return 1;
}

/* DNAcopy getmncwt (weighted data): delta = lightest circular arc of kmax+1 consecutive probes as a
 * fraction of the total weight (R's (kmax+1)/n is overwritten by it) */
static double node_delta(const node_t* N, int kmax) {
    const int n = N->n, jw = kmax + 1;
    double delta = N->cw[jw - 1];
    for (int i = 1; i <= n - jw; ++i) { const double a = N->cw[i + jw - 1] - N->cw[i - 1]; if (a < delta) delta = a; }
    for (int i = 1; i <= jw; ++i) { const double a = N->total - N->cw[n - jw + i - 1] + N->cw[i - 1]; if (a < delta) delta = a; }
    return delta / N->total;
}

/* permutation statistic T^2 of permutation `perm` of the node [lo,hi): hybrid (short circular arcs)
 * or exact (all arcs) */
static double node_perm_stat_buf(const node_t* N, double* t, const double* wn, const cbso_params* P, int hybrid,
                                 int lo, int hi, int perm, cbso_stats* st) {
    const int n = N->n, al0 = P->min_width;
    prp_t prp;
    prp_init(&prp, P->seed, (uint32_t)lo, (uint32_t)hi, 0u, (uint32_t)perm, (uint32_t)n);
    for (int i = 0; i < n; ++i) t[i] = N->y[prp_eval(&prp, (uint32_t)i)] * N->rw[i];
    double pbss;
    if (hybrid) {
        pbss = perm_stat_hybrid(N, wn, t, al0, P->kmax, st ? &st->perm_arcs : NULL);
    } else {
        double* Sp = (double*)malloc(sizeof(double) * (n + 1));
        /* Prefix sums of the permuted terms in a FIXED blocked association (32 runs of ceil(n/32) terms: each run is
         * summed left to right from zero, the run totals are accumulated left to right, prefix = offset + local
         * prefix) -- the association a warp uses (one run per lane, csrc/cbs.cu cbs_perm_exact_kernel), so that
         * both sides produce the same bits.  A strictly sequential sum differs in the last place only. */
        {
            const int c = (n + 31) / 32;
            double tot[32];
            Sp[0] = 0.0;
            for (int l = 0; l < 32; ++l) {
                const int k0 = l * c, k1 = (k0 + c < n) ? k0 + c : n;
                double run = 0.0;
                for (int k = k0; k < k1; ++k) { run = run + t[k]; Sp[k + 1] = run; }
                tot[l] = run;
            }
            double off = 0.0;
            for (int l = 0; l < 32; ++l) {
                const int k0 = l * c, k1 = (k0 + c < n) ? k0 + c : n;
                for (int k = k0; k < k1; ++k) Sp[k + 1] = off + Sp[k + 1];
                off = off + tot[l];
            }
        }
        int pi, pj;
        pbss = max_arc_full(N, Sp, al0, &pi, &pj, st ? &st->perm_arcs : NULL);
        free(Sp);
    }
    if (st) st->perms_run++;
    return t2_of(pbss, N->tss, n);
}
static double node_perm_stat(node_t* N, const double* wn, const cbso_params* P, int hybrid, int lo, int hi, int perm,
                             cbso_stats* st) {
    return node_perm_stat_buf(N, N->t, wn, P, hybrid, lo, hi, perm, st);
}

/* Tests the interval [lo,hi) of the arm; returns ncpt (0,1,2) and icpt relative to lo. */
static int find_cpt(const double* x_arm, const double* w_arm, int lo, int hi, const cbso_params* P,
                    const int* sbdry, int icpt[2], cbso_stats* st, double* out_ostat1, double* out_delta) {
    const int n = hi - lo;
    const double* x = x_arm + lo;
    const double* w = w_arm + lo;
    if (out_ostat1) *out_ostat1 = 0.0;
    if (st) st->nodes++;
    node_t N;
    double rsw;
    if (!node_build(x, w, n, &N, &rsw)) return 0;
    const double tss = N.tss;

    int ncpt = 0, ai, aj;
    const int al0 = P->min_width;
    const double bss = P->prune ? max_arc_pruned(&N, N.S, al0, &ai, &aj, st ? &st->obs_arcs : NULL)
                                : max_arc_full(&N, N.S, al0, &ai, &aj, st ? &st->obs_arcs : NULL);
    const double t2 = t2_of(bss, tss, n);
    const double ostat1 = sqrt(t2);
    const double ostat = t2 * 0.99999;
    if (out_ostat1) *out_ostat1 = ostat1;
    int split = 0;
    if (ostat1 <= 0.1) goto done;
    {
        const int l = (aj - ai < n - aj + ai) ? (aj - ai) : (n - aj + ai);
        if (ostat1 >= 7.0 && l >= 10) { split = 1; }
    }
    if (!split) {
        const int hybrid = P->hybrid && (n > P->nmin);
        int nrejc, nrej = 0, k;
        double* wn = NULL;
        if (hybrid) {
            const double delta = node_delta(&N, P->kmax);
            if (out_delta) *out_delta = delta;
            const double pval1 = cbso_tailp(ostat1, delta, n, P->ngrid, P->tol);
            if (pval1 > P->alpha) goto done;
            nrejc = (int)((P->alpha - pval1) * (double)P->nperm);
            wn = (double*)malloc(sizeof(double) * n);
            for (int i = 0; i < n; ++i) wn[i] = w[i] / rsw;
        } else {
            nrejc = (int)(P->alpha * (double)P->nperm);
        }
        k = nrejc * (nrejc + 1) / 2 + 1;
        if (st) st->perm_tests++;
        split = 1; /* surviving all permutations accepts the split */
        /* Permutations are counter-based, so a block of them can be scored in any order (in parallel for
         * large nodes -- the CPU baseline's answer to a node that survives thousands of permutations) and
         * DNAcopy's sequential stopping rule replayed over the block afterwards: same decisions, and
         * perms_run counts the permutations up to the stopping point only. */
        const int par = (long long)n * (hybrid ? 24 : n / 2) >= 200000;
        int decided = 0;
        for (int p0 = 0, blk = 16; p0 < P->nperm && !decided; p0 += blk, blk = blk < 1024 ? blk * 4 : blk) {
            const int cnt = (P->nperm - p0 < blk) ? P->nperm - p0 : blk;
            double* ps = (double*)malloc(sizeof(double) * (size_t)cnt);
            long long arcs_blk = 0;
            if (par && cnt > 1) {
#pragma omp parallel
                {
                    double* tbuf = (double*)malloc(sizeof(double) * (size_t)n);
                    cbso_stats lst;
                    memset(&lst, 0, sizeof(lst));
#pragma omp for schedule(dynamic, 4)
                    for (int c = 0; c < cnt; ++c) ps[c] = node_perm_stat_buf(&N, tbuf, wn, P, hybrid, lo, hi, p0 + c, &lst);
#pragma omp atomic
                    arcs_blk += lst.perm_arcs;
                    free(tbuf);
                }
            } else {
                cbso_stats lst;
                memset(&lst, 0, sizeof(lst));
                for (int c = 0; c < cnt; ++c) ps[c] = node_perm_stat_buf(&N, N.t, wn, P, hybrid, lo, hi, p0 + c, &lst);
                arcs_blk = lst.perm_arcs;
            }
            int used = cnt;
            for (int c = 0; c < cnt; ++c) {
                const int np = p0 + c + 1;
                if (ostat <= ps[c]) { ++nrej; ++k; }
                if (nrej > nrejc) { split = 0; decided = 1; used = c + 1; break; }
                if (np >= sbdry[k - 1]) { split = 1; decided = 1; used = c + 1; break; }
            }
            if (st) { st->perms_run += used; st->perm_arcs += arcs_blk / cnt * used; }
            free(ps);
        }
        free(wn);
        if (!split) goto done;
    }
    /* ternary split bookkeeping (changepoints.f label 200) */
    if (aj == n) { ncpt = 1; icpt[0] = ai; }
    else if (ai == 0) { ncpt = 1; icpt[0] = aj; }
    else {
        double tp = wtpermp(N.xc, N.w, N.rw, ai, aj - ai, P, (uint32_t)lo, (uint32_t)hi, 1u, st);
        if (tp <= P->alpha) { ncpt = 1; icpt[0] = ai; }
        tp = wtpermp(N.xc + ai, N.w + ai, N.rw + ai, aj - ai, n - aj, P, (uint32_t)lo, (uint32_t)hi, 2u, st);
        if (tp <= P->alpha) { icpt[ncpt] = aj; ++ncpt; }
        if (ncpt == 0) { ncpt = 2; icpt[0] = ai; icpt[1] = aj; }
    }
done:
    node_free(&N);
    return ncpt;
}

/* One node, for unit tests: returns ncpt, fills icpt and the observed sqrt(T^2). */
int cbso_find_cpt(const double* x, const double* w, int lo, int hi, const cbso_params* P, const int* sbdry,
                  int* icpt, double* ostat1, cbso_stats* st) {
    return find_cpt(x, w, lo, hi, P, sbdry, icpt, st, ostat1, NULL);
}

/* DNAcopy changepoints(): returns number of segments; seg_end[k] = exclusive end of segment k.
 * seg_mean[k] = weighted.mean(log2, weight) as cbs.py:70-71 recomputes it. */
int cbso_segment_arm(const double* x, const double* w, int n, const cbso_params* P, const int* sbdry,
                     int* seg_end, double* seg_mean, cbso_stats* st) {
    if (n <= 0) return 0;
    int* stack = (int*)malloc(sizeof(int) * (size_t)(n + 2));
    int* ends = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    int k = 2, nend = 0;
    stack[0] = 0; stack[1] = n;
    while (k > 1) {
        const int lo = stack[k - 2], hi = stack[k - 1];
        int icpt[2] = {0, 0}, ncpt = 0;
        if (hi - lo >= 2 * P->min_width) ncpt = find_cpt(x, w, lo, hi, P, sbdry, icpt, st, NULL, NULL);
        if (ncpt == 0) { ends[nend++] = hi; --k; }
        else if (ncpt == 1) { stack[k - 1] = lo + icpt[0]; stack[k] = hi; ++k; }
        else { stack[k - 1] = lo + icpt[0]; stack[k] = lo + icpt[1]; stack[k + 1] = hi; k += 2; }
    }
    /* ends were produced right-to-left */
    for (int s = 0; s < nend; ++s) seg_end[s] = ends[nend - 1 - s];
    int a = 0;
    for (int s = 0; s < nend; ++s) {
        double sw = 0.0, swx = 0.0;
        for (int i = a; i < seg_end[s]; ++i) { swx += x[i] * w[i]; sw += w[i]; }
        seg_mean[s] = swx / sw;
        a = seg_end[s];
    }
    free(stack); free(ends);
    return nend;
}

/* All arms of a sample (OpenMP over arms).  Outputs are per-arm slabs: segment s of arm a is at
 * index arm_offsets[a] + s (an arm of n bins has at most n segments).  nseg[a] = count. */
int cbso_segment(const double* x, const double* w, const int64_t* arm_offsets, int n_arms, const cbso_params* P,
                 int64_t* seg_end, double* seg_mean, int* nseg, cbso_stats* st_out, int nthreads) {
    const int max_ones = (int)floor((double)P->nperm * P->alpha) + 1;
    const int sbn = max_ones * (max_ones + 1) / 2;
    int* sbdry = (int*)malloc(sizeof(int) * (size_t)sbn);
    cbso_getbdry(P->eta, P->nperm, max_ones, sbdry);
    cbso_stats total;
    memset(&total, 0, sizeof(total));
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    /* a single arm leaves the threads to the permutation blocks inside find_cpt */
#pragma omp parallel for schedule(dynamic, 1) if (n_arms > 1)
    for (int a = 0; a < n_arms; ++a) {
        const int64_t lo = arm_offsets[a], hi = arm_offsets[a + 1];
        const int n = (int)(hi - lo);
        cbso_stats st;
        memset(&st, 0, sizeof(st));
        int* ends = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        nseg[a] = cbso_segment_arm(x + lo, w + lo, n, P, sbdry, ends, seg_mean + lo, &st);
        for (int s = 0; s < nseg[a]; ++s) seg_end[lo + s] = lo + ends[s];
        free(ends);
#pragma omp critical
        {
            total.nodes += st.nodes; total.obs_arcs += st.obs_arcs; total.perm_arcs += st.perm_arcs;
            total.perm_tests += st.perm_tests; total.perms_run += st.perms_run; total.tperm_tests += st.tperm_tests;
        }
    }
    if (st_out) *st_out = total;
    free(sbdry);
    return 0;
}


/* Diagnostic hook for the statistical tests (tests/test_cbs_stats*.py): ONE node, no early stopping.
 * out[0] = sqrt(T^2) observed, out[1] = tailp (hybrid nodes; -1 otherwise), out[2] = getmncwt delta (hybrid),
 * out[3] = 0.99999 T^2 (the value permutations are compared with), out[4], out[5] = flank p-values of the
 * ternary split with the permutations forced (-1 when the arc touches an end of the node).
 * iout[0..1] = arc (i, j), iout[2] = permutations whose statistic reached the observed one among the
 * first nperm_run, iout[3] = 1 hybrid / 0 exact, iout[4] = 1 when the data are constant. */
int cbso_node_diag(const double* x_arm, const double* w_arm, int lo, int hi, const cbso_params* P, int nperm_run,
                   double* out, int* iout) {
    const int n = hi - lo;
    const double* x = x_arm + lo;
    const double* w = w_arm + lo;
    for (int k = 0; k < 6; ++k) out[k] = -1.0;
    for (int k = 0; k < 5; ++k) iout[k] = 0;
    node_t N;
    double rsw;
    if (!node_build(x, w, n, &N, &rsw)) { iout[4] = 1; return 0; }
    int ai, aj;
    const double bss = P->prune ? max_arc_pruned(&N, N.S, P->min_width, &ai, &aj, NULL)
                                : max_arc_full(&N, N.S, P->min_width, &ai, &aj, NULL);
    const double t2 = t2_of(bss, N.tss, n);
    out[0] = sqrt(t2);
    out[3] = t2 * 0.99999;
    iout[0] = ai; iout[1] = aj;
    const int hybrid = P->hybrid && (n > P->nmin);
    iout[3] = hybrid;
    double* wn = NULL;
    if (hybrid) {
        out[2] = node_delta(&N, P->kmax);
        out[1] = cbso_tailp(out[0], out[2], n, P->ngrid, P->tol);
        wn = (double*)malloc(sizeof(double) * n);
        for (int i = 0; i < n; ++i) wn[i] = w[i] / rsw;
    }
    int nrej = 0;
    for (int np = 0; np < nperm_run; ++np)
        if (out[3] <= node_perm_stat(&N, wn, P, hybrid, lo, hi, np, NULL)) ++nrej;
    iout[2] = nrej;
    free(wn);
    if (ai != 0 && aj != n) {
        out[4] = wtpermp_forced(N.xc, N.w, N.rw, ai, aj - ai, P, (uint32_t)lo, (uint32_t)hi, 1u);
        out[5] = wtpermp_forced(N.xc + ai, N.w + ai, N.rw + ai, aj - ai, n - aj, P, (uint32_t)lo, (uint32_t)hi, 2u);
    }
    node_free(&N);
    return 0;
}

/* exposed pieces for unit tests */
double cbso_perm_pvalue_flank(const double* xc, const double* w, int n1, int n2, const cbso_params* P,
                              uint32_t lo, uint32_t hi, uint32_t tid) {
    double* rw = (double*)malloc(sizeof(double) * (size_t)(n1 + n2));
    for (int i = 0; i < n1 + n2; ++i) rw[i] = sqrt(w[i]);
    const double p = wtpermp(xc, w, rw, n1, n2, P, lo, hi, tid, NULL);
    free(rw);
    return p;
}

/* ------------------------------------------------------------------ smooth.CNA (--smooth-cbs) */
/* DNAcopy::smooth.CNA with its defaults, as the R script applies it to one arm before segment()
 * (cnvlib/segmentation/cbs.py:51-58): smooth.region = 10, outlier.SD.scale = 4, smooth.SD.scale = 2,
 * trim = 0.025.  PARITY UNPINNED like the rest of this file (restated from the package's R / Fortran:
 * trimmed.variance, inflfact, smoothLR).
 *   trimmed.SD = sqrt( inflfact(trim) * sum( sort(|diff(x)|)[1:n.keep]^2 ) / (2 n.keep) ),
 *                n.keep = round((1 - 2 trim)(n - 1))            (R's round: half to even)
 *   a point is an outlier when it lies more than 4 trimmed.SD above (below) EVERY other point within 10 bins
 *   on either side (window clipped to the arm); it is then moved to median(window incl. itself) +- 2 trimmed.SD. */
static double qnorm_as241(double p) {   /* Wichura's AS241 PPND16, the algorithm behind R's qnorm */
    const double q = p - 0.5;
    double r, val;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                       + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                     + 133.14166789178437745) * r + 3.387132872796366608)
              / (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r
                     + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
                   + 42.313330701600911252) * r + 1.0);
        return val;
    }
    r = q < 0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r
                   + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
                 + 4.6303378461565452959) * r + 1.42343711074968357734)
              / (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r
                     + 0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r
                   + 2.05319162663775882187) * r + 1.0);
    } else {
        r -= 5.0;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r
                   + 0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r
                 + 5.4637849111641143699) * r + 6.6579046435011037772)
              / (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                     + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r
                   + 0.59983224954205144545) * r + 1.0);
    }
    return q < 0.0 ? -val : val;
}

double cbso_inflfact(double trim) {
    const double a = qnorm_as241(1.0 - trim);
    const int m = 10001;
    double s = 0.0;
    for (int i = 0; i < m - 1; ++i) {
        const double x0 = -a + (2.0 * a) * (double)i / (double)(m - 1), x1 = -a + (2.0 * a) * (double)(i + 1) / (double)(m - 1);
        const double xm = (x0 + x1) / 2.0;
        s += xm * xm * (exp(-0.5 * xm * xm) * 0.3989422804014327) / (1.0 - 2.0 * trim);
    }
    return 1.0 / (s * (2.0 * a / 10000.0));
}

static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* one arm; out may not alias x.  Returns trimmed.SD. */
double cbso_smooth_arm(const double* x, int n, int k, double outlier_scale, double smooth_scale, double trim, double* out) {
    for (int i = 0; i < n; ++i) out[i] = x[i];
    if (n < 2) return 0.0;
    double* d = (double*)malloc(sizeof(double) * (size_t)(n - 1));
    for (int i = 0; i + 1 < n; ++i) d[i] = fabs(x[i + 1] - x[i]);
    qsort(d, (size_t)(n - 1), sizeof(double), cmp_double);
    int nkeep = (int)nearbyint((1.0 - 2.0 * trim) * (double)(n - 1));   /* R round(): half to even */
    if (nkeep < 1) nkeep = 1;
    if (nkeep > n - 1) nkeep = n - 1;
    dd_t acc = {0.0, 0.0};
    for (int i = 0; i < nkeep; ++i) dd_add(&acc, d[i] * d[i]);          /* R-level sum(): correctly rounded */
    free(d);
    const double sd = sqrt(cbso_inflfact(trim) * ((acc.hi + acc.lo) / (2.0 * (double)nkeep)));
    const double oSD = outlier_scale * sd, sSD = smooth_scale * sd;
    double nb[64];
    for (int i = 0; i < n; ++i) {
        const int lo = i - k < 0 ? 0 : i - k, hi = i + k > n - 1 ? n - 1 : i + k;
        double above = 100.0 * oSD, below = 100.0 * oSD;   /* x_i minus the largest neighbour / smallest minus x_i */
        int first = 1;
        double mx = 0.0, mn = 0.0;
        for (int j = lo; j <= hi; ++j) {
            if (j == i) continue;
            if (first) { mx = mn = x[j]; first = 0; }
            else { if (x[j] > mx) mx = x[j]; if (x[j] < mn) mn = x[j]; }
        }
        if (first) continue;
        above = x[i] - mx;
        below = mn - x[i];
        if (!(above > oSD) && !(below > oSD)) continue;
        int m = 0;
        for (int j = lo; j <= hi; ++j) nb[m++] = x[j];
        qsort(nb, (size_t)m, sizeof(double), cmp_double);
        const double med = (m % 2) ? nb[m / 2] : (nb[m / 2 - 1] + nb[m / 2]) / 2.0;
        out[i] = above > oSD ? med + sSD : med - sSD;
    }
    return sd;
}

/* every arm of a batch (same arm_offsets convention as cbso_segment); sd_out[a] = trimmed.SD of arm a */
int cbso_smooth_cna(const double* x, const int64_t* arm_offsets, int n_arms, int k, double outlier_scale,
                    double smooth_scale, double trim, double* out, double* sd_out) {
    if (k > 31) return -1;
    for (int a = 0; a < n_arms; ++a) {
        const int64_t lo = arm_offsets[a], hi = arm_offsets[a + 1];
        const double sd = cbso_smooth_arm(x + lo, (int)(hi - lo), k, outlier_scale, smooth_scale, trim, out + lo);
        if (sd_out) sd_out[a] = sd;
    }
    return 0;
}
