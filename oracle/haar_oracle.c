/* TEST INFRASTRUCTURE ONLY -- CPU restatement of HaarSeg as the reference implements it
 * (cnvlib/segmentation/haar.py).  Pinned against vectors produced by the unmodified reference
 * (tests/golden/haar.npz) in tests/test_oracle_pins.py.  Never linked into the product.
 *
 *   haaro_conv         haar.py:460-558  HaarConv / _haar_conv_unweighted / _haar_conv_weighted
 *   haaro_find_peaks   haar.py:561-612  FindLocalPeaks (the sequential state machine, verbatim logic)
 *   haaro_fdr_thres    haar.py:374-402  FDRThres
 *   haaro_unify        haar.py:615-663  UnifyLevels
 *   haaro_segment      haar.py:257-371  haarSeg (+ SegmentByPeaks :405-451)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* numpy float64 add.reduce for a contiguous array (pairwise_sum, blocksize 128, 8 accumulators) */
static double np_pairwise(const double* a, long n) {
    if (n < 8) {
        double r = 0.0;
        for (long i = 0; i < n; ++i) r += a[i];
        return r;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
    }
}

void haaro_conv(const double* x, const double* w, int n, int h, double* out) {
    for (int k = 0; k < n; ++k) out[k] = 0.0;
    if (h > n) return;
    if (!w) {
        const double norm = sqrt(2.0 * h);
        double run = 0.0;
        for (int k = 1; k < n; ++k) {
            int hi = k + h - 1, lo = k - h - 1;
            if (hi >= n) hi = n - 1 - (hi - n);
            if (lo < 0) lo = -lo - 1;
            const double inc = (x[hi] + x[lo]) - 2 * x[k - 1];
            run += inc;
            out[k] = run / norm;
        }
        return;
    }
    double tw[64], tp[64];
    for (int i = 0; i < h; ++i) { tw[i] = w[i]; tp[i] = w[i] * x[i]; }
    double highW = np_pairwise(tw, h), highNN = np_pairwise(tp, h);
    double lowW = highW, lowNN = -highNN;
    const double norm = sqrt(h / 2.0);
    for (int k = 1; k < n; ++k) {
        int hi = k + h - 1, lo = k - h - 1;
        if (hi >= n) hi = n - 1 - (hi - n);
        if (lo < 0) lo = -lo - 1;
        lowNN += x[lo] * w[lo] - x[k - 1] * w[k - 1];
        highNN += x[hi] * w[hi] - x[k - 1] * w[k - 1];
        lowW += w[k - 1] - w[lo];
        highW += w[hi] - w[k - 1];
        out[k] = norm * (lowNN / lowW + highNN / highW);
    }
}

int haaro_find_peaks(const double* s, int n, int* peaks) {
    int np_ = 0, max_suspect = -1, min_suspect = -1;
    for (int k = 1; k < n - 1; ++k) {
        const double p = s[k - 1], c = s[k], q = s[k + 1];
        if (c > 0) {
            if (c > p && c > q) peaks[np_++] = k;
            else if (c > p && c == q) max_suspect = k;
            else if (c == p && c > q) { if (max_suspect >= 0) { peaks[np_++] = max_suspect; max_suspect = -1; } }
            else if (c == p && c < q) max_suspect = -1;
        } else if (c < 0) {
            if (c < p && c < q) peaks[np_++] = k;
            else if (c < p && c == q) min_suspect = k;
            else if (c == p && c < q) { if (min_suspect >= 0) { peaks[np_++] = min_suspect; min_suspect = -1; } }
            else if (c == p && c > q) min_suspect = -1;
        }
    }
    return np_;
}

static int cmp_desc(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x < y) - (x > y);
}

/* scipy.special.ndtr (cephes) */
static double ndtr(double a) {
    const double x = a * 0.70710678118654752440, z = fabs(x);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(x);
    double y = 0.5 * erfc(z);
    if (x > 0) y = 1.0 - y;
    return y;
}

double haaro_fdr_thres(const double* x, int M, double q, double sigma) {
    if (M < 2) return 0.0;
    if (sigma <= 0) return 0.0;
    double* v = (double*)malloc(sizeof(double) * M);
    for (int i = 0; i < M; ++i) v[i] = fabs(x[i]);
    qsort(v, M, sizeof(double), cmp_desc);
    int last = -1;
    for (int i = 0; i < M; ++i) {
        const double p = 2 * (1 - ndtr(v[i] / sigma));
        const double m = (double)(i + 1) / (double)M;
        if (p <= m * q) last = i;
    }
    const double T = (last >= 0) ? v[last] : nextafter(v[0], INFINITY);
    free(v);
    return T;
}

int haaro_unify(const int* base, int nb, const int* addon, int na, int window, int* out) {
    if (na == 0) { memcpy(out, base, sizeof(int) * nb); return nb; }
    if (nb == 0) { memcpy(out, addon, sizeof(int) * na); return na; }
    int n = 0, ib = 0;
    int* kept = (int*)malloc(sizeof(int) * na);
    int nk = 0;
    for (int i = 0; i < na; ++i) {
        const int a = addon[i];
        int pos = 0;                       /* searchsorted(base, a), side left */
        { int lo = 0, hi = nb; while (lo < hi) { int m = (lo + hi) / 2; if (base[m] < a) lo = m + 1; else hi = m; } pos = lo; }
        const int left_close = pos > 0 && (a - base[pos - 1]) <= window;
        const int right_close = pos < nb && (base[pos] - a) <= window;
        if (!(left_close || right_close)) kept[nk++] = a;
    }
    int ik = 0;
    while (ib < nb || ik < nk) {
        if (ik >= nk || (ib < nb && base[ib] <= kept[ik])) out[n++] = base[ib++];
        else out[n++] = kept[ik++];
    }
    free(kept);
    return n;
}

static int cmp_asc(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* haarSeg: returns number of segments; starts[k] = first bin of segment k; means[k] */
int haaro_segment(const double* x, const double* w, int n, double q, int* starts, double* means) {
    if (n <= 0) return 0;
    double* conv = (double*)malloc(sizeof(double) * n);
    double* tmp = (double*)malloc(sizeof(double) * n);
    int* peaks = (int*)malloc(sizeof(int) * (n + 1));
    int* bp = (int*)malloc(sizeof(int) * (n + 1));
    int* bp2 = (int*)malloc(sizeof(int) * (n + 1));
    int* addon = (int*)malloc(sizeof(int) * (n + 1));
    int nbp = 0;
    haaro_conv(x, NULL, n, 1, conv);
    for (int i = 0; i < n; ++i) tmp[i] = fabs(conv[i]);
    qsort(tmp, n, sizeof(double), cmp_asc);
    double med = (n % 2) ? tmp[n / 2] : (tmp[n / 2 - 1] + tmp[n / 2]) / 2.0;
    double sigma = med * 1.4826;
    if (sigma < 1e-10) sigma = 1e-10;
    for (int level = 1; level <= 5; ++level) {
        const int h = 1 << level;
        haaro_conv(x, w, n, h, conv);
        const int npk = haaro_find_peaks(conv, n, peaks);
        for (int i = 0; i < npk; ++i) tmp[i] = conv[peaks[i]];
        const double T = haaro_fdr_thres(tmp, npk, q, sigma);
        int na = 0;
        for (int i = 0; i < npk; ++i) if (fabs(conv[peaks[i]]) >= T) addon[na++] = peaks[i];
        const int nn = haaro_unify(bp, nbp, addon, na, 1 << (level - 1), bp2);
        memcpy(bp, bp2, sizeof(int) * nn);
        nbp = nn;
    }
    int nseg = nbp + 1;
    for (int s = 0; s < nseg; ++s) {
        const int a = s ? bp[s - 1] : 0, b = (s < nbp) ? bp[s] : n;
        starts[s] = a;
        double sw = 0, swx = 0, sx = 0;
        if (w) {
            int m = 0;
            for (int i = a; i < b; ++i) if (w[i] == w[i]) { tmp[m] = w[i]; conv[m] = x[i] * w[i]; ++m; }
            sw = np_pairwise(tmp, m);
            swx = np_pairwise(conv, m);
        }
        if (w && sw > 0) means[s] = swx / sw;
        else { sx = np_pairwise(x + a, b - a); means[s] = sx / (double)(b - a); }
    }
    free(conv); free(tmp); free(peaks); free(bp); free(bp2); free(addon);
    return nseg;
}
