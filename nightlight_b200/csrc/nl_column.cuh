// nl_column.cuh -- per-pixel column reducers of the stacking hot path.
//
// One pixel's samples across the N frames form a "column".  The reference reduces each column on
// the CPU with quick-select + sequential fp32 sums (internal/ops/stack/stack.go:274-918,
// internal/qsort/qsort.go:68-126, internal/stats/stats.go:246-261,569-586).  The result of the
// mean-type modes depends on the element ORDER left behind by that quick-select and by the
// swap-with-last clip loop, because mean and sigma are sequential fp32 sums over the permuted buffer.
// These routines therefore reproduce the reference's permutation and evaluation order exactly, but
// are organised for SIMT: a column lives in shared memory with a compile-time element stride S
// (S = pixels per warp tile, so lane == bank), and the quick-select is a flattened state machine
// that keeps the 32 lanes of a warp (32 different pixels) in one loop instead of nested
// data-dependent loops.
//
// Everything here is __host__ __device__ so tests can compile the very same code for the CPU
// (tests/host_emul.cpp, S = 1) and compare it against the oracle without a GPU.
// Compile with -fmad=false: Go/amd64 never contracts a*b+c.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NL_HD __host__ __device__ __forceinline__
#else
#define NL_HD inline
#endif

namespace nl {

NL_HD float nl_sqrtf(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);      // float32(math.Sqrt(float64(x))) == correctly rounded sqrtf
#else
    return sqrtf(x);
#endif
}
NL_HD float nl_divf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
NL_HD float nl_mulf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);    // never contracted into an FMA
#else
    return a * b;
#endif
}
NL_HD float nl_addf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
NL_HD float nl_subf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}

// Stacking modes and weighting modes, numbered like the reference (stack.go:33-42, 57-63).
enum StackMode { ST_MEDIAN = 0, ST_MEAN = 1, ST_SIGMA = 2, ST_WINSOR = 3, ST_MAD = 4, ST_LINFIT = 5, ST_AUTO = 6 };

// stack.go:45-55
NL_HD int auto_select_mode(int n_frames) {
    if (n_frames >= 25) return ST_LINFIT;
    if (n_frames >= 15) return ST_WINSOR;
    if (n_frames >= 6) return ST_SIGMA;
    return ST_MEAN;
}

// ---------------------------------------------------------------------------------------------
// Quick-select, exact permutation of qsort.go:94-126 (Hoare partition, pivot a[(l+r)>>1]).
// k is 1-based.  Flattened: every trip of the single loop advances the left and the right scan
// pointer of the current partition by at most one element each, or performs one swap, or closes
// the partition -- so 32 lanes working on 32 different columns stay converged.
// The two scans of a Hoare round are independent (no stores happen between them), so advancing
// them in lock step visits exactly the stop positions of the sequential code.
// ---------------------------------------------------------------------------------------------
template <int S>
NL_HD float qselect(float *a, int n, int k) {
    int left = 0, right = n - 1;
    if (left >= right) return a[left * S];
    float pivot = a[((left + right) >> 1) * S];
    int l = left, r = right;
    for (;;) {
        float al = a[l * S], ar = a[r * S];
        bool sl = al >= pivot;   // left scan stops here
        bool sr = ar <= pivot;   // right scan stops here
        if (sl && sr) {
            if (l < r) {          // swap, both scans move on
                a[l * S] = ar;
                a[r * S] = al;
                l++; r--;
            } else {              // scans crossed: partition index is r
                int offset = r - left + 1;
                if (k <= offset) right = r;
                else { left = r + 1; k -= offset; }
                if (left >= right) break;
                pivot = a[((left + right) >> 1) * S];
                l = left; r = right;
            }
        } else {
            l += sl ? 0 : 1;
            r -= sr ? 0 : 1;
        }
    }
    return a[left * S];
}

// qsort.go:68-82 QSelectMedianFloat32
template <int S>
NL_HD float qselect_median(float *a, int n) {
    int k = (n >> 1) + 1;
    float upper = qselect<S>(a, n, k);
    if (n & 1) return upper;
    float lower = a[0];
    for (int i = 1; i < k - 1; i++) lower = fmaxf(lower, a[i * S]);   // no NaNs in a column
    return nl_mulf(0.5f, nl_addf(lower, upper));
}

// stats.go:246-261 MeanStdDev: two sequential fp32 sums in buffer order, population sigma.
template <int S>
NL_HD void mean_stddev(const float *a, int n, float &mean, float &sd) {
    float s = 0.0f;
#pragma unroll 8
    for (int i = 0; i < n; i++) s = nl_addf(s, a[i * S]);
    float fn = (float)n;
    float m = nl_divf(s, fn);
    float v = 0.0f;
#pragma unroll 8
    for (int i = 0; i < n; i++) {
        float d = nl_subf(a[i * S], m);
        v = nl_addf(v, nl_mulf(d, d));
    }
    v = nl_divf(v, fn);
    mean = m;
    sd = nl_sqrtf(v);
}

// The clip loop shared by the sigma and winsor variants (stack.go:411-424, 495-514, 674-689,
// 779-798): an out-of-bounds sample is overwritten by the last one, the slice shrinks and slot j
// is tested again.  W: weights travel with the values.
template <int S, bool W>
NL_HD int clip_pass(float *g, float *gw, int cur, float lo, float hi, int &ncl, int &nch) {
    int j = 0;
    while (j < cur) {
        float v = g[j * S];
        bool low = v < lo, high = v > hi;
        if (low || high) {
            cur--;
            g[j * S] = g[cur * S];
            if (W) gw[j * S] = gw[cur * S];
            if (low) ncl++; else nch++;
        } else {
            j++;
        }
    }
    return cur;
}

// weighted mean of the survivors in buffer order (stack.go:518-524, 802-808)
template <int S>
NL_HD float weighted_mean(const float *g, const float *gw, int cur) {
    float ws = 0.0f, wsum = 0.0f;
    for (int i = 0; i < cur; i++) {
        float w = gw[i * S];
        ws = nl_addf(ws, nl_mulf(g[i * S], w));
        wsum = nl_addf(wsum, w);
    }
    return nl_divf(ws, wsum);
}

// inner winsorisation loop (stack.go:649-672, 754-777); wz is scratch of the same shape as g
template <int S>
NL_HD float winsor_sigma(const float *g, float *wz, int cur, float median, float sd) {
    for (int i = 0; i < cur; i++) wz[i * S] = g[i * S];
    for (;;) {
        float lo = nl_subf(median, nl_mulf(1.5f, sd));
        float hi = nl_addf(median, nl_mulf(1.5f, sd));
        int changed = 0;
        // clamp and first sum of MeanStdDev fused: the sum runs over the clamped values in order
        float s = 0.0f;
        for (int i = 0; i < cur; i++) {
            float v = wz[i * S];
            if (v < lo) { v = lo; changed++; wz[i * S] = v; }
            else if (v > hi) { v = hi; changed++; wz[i * S] = v; }
            s = nl_addf(s, v);
        }
        float fn = (float)cur;
        float m = nl_divf(s, fn);
        float var = 0.0f;
#pragma unroll 8
        for (int i = 0; i < cur; i++) {
            float d = nl_subf(wz[i * S], m);
            var = nl_addf(var, nl_mulf(d, d));
        }
        var = nl_divf(var, fn);
        float old = sd;
        sd = nl_mulf(1.134f, nl_sqrtf(var));
        float factor = nl_divf(fabsf(nl_subf(sd, old)), old);
        if (changed == 0 || factor <= 0.0005f) break;
    }
    return sd;
}

// In-place ascending sort of a column.  The reference sorts with its Hoare quicksort
// (qsort.go:26-32); the sorted array is unique, so any correct sort is bit-exact.  Heap sort:
// no recursion, no stack, O(n log n) for every input.
template <int S>
NL_HD void sort_column(float *a, int n) {
    if (n < 2) return;
    for (int start = (n >> 1) - 1; start >= 0; start--) {      // heapify
        int root = start;
        float v = a[root * S];
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            float c = a[child * S];
            if (child + 1 < n) { float c2 = a[(child + 1) * S]; if (c2 > c) { c = c2; child++; } }
            if (c <= v) break;
            a[root * S] = c;
            root = child;
        }
        a[root * S] = v;
    }
    for (int end = n - 1; end > 0; end--) {
        float v = a[end * S];
        a[end * S] = a[0];
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            float c = a[child * S];
            if (child + 1 < end) { float c2 = a[(child + 1) * S]; if (c2 > c) { c = c2; child++; } }
            if (c <= v) break;
            a[root * S] = c;
            root = child;
        }
        a[root * S] = v;
    }
}

// Insertion sort: used for the re-sorts of StackLinearFit, where the array is sorted except for
// the few slots the rejection pass overwrote.
template <int S>
NL_HD void insertion_sort_column(float *a, int n) {
    for (int i = 1; i < n; i++) {
        float v = a[i * S];
        int j = i - 1;
        if (a[j * S] <= v) continue;
        while (j >= 0 && a[j * S] > v) { a[(j + 1) * S] = a[j * S]; j--; }
        a[(j + 1) * S] = v;
    }
}

// MeanStdDev of xs = 0,1,..,n-1 (stats.go:246-261 applied to StackLinearFit's xs, stack.go:836-839)
NL_HD void ramp_mean_stddev(int n, float &mean, float &sd) {
    float s = 0.0f;
    for (int i = 0; i < n; i++) s = nl_addf(s, (float)i);
    float fn = (float)n;
    float m = nl_divf(s, fn);
    float v = 0.0f;
    for (int i = 0; i < n; i++) {
        float d = nl_subf((float)i, m);
        v = nl_addf(v, nl_mulf(d, d));
    }
    v = nl_divf(v, fn);
    mean = m;
    sd = nl_sqrtf(v);
}

// ---------------------------------------------------------------------------------------------
// The reducers.  g: gathered non-NaN samples in frame order, cur > 0 of them.  Scratch buffers as
// noted.  Clip counters accumulate into ncl / nch.
// ---------------------------------------------------------------------------------------------

// stack.go:372-436 StackSigma / stack.go:442-531 StackSigmaWeighted.  In the weighted variant the
// quick-select permutes the values but not the weights (stack.go:487 hands it gatheredCur only);
// the weights move in the clip loop alone.  Reproduced as is.
template <int S, bool W>
NL_HD float reduce_sigma(float *g, float *gw, int cur, float sig_lo, float sig_hi, int &ncl, int &nch) {
    for (;;) {
        float median = qselect_median<S>(g, cur);
        float mean, sd;
        mean_stddev<S>(g, cur, mean, sd);
        float lo = nl_subf(median, nl_mulf(sig_lo, sd));
        float hi = nl_addf(median, nl_mulf(sig_hi, sd));
        int before = cur;
        cur = clip_pass<S, W>(g, gw, cur, lo, hi, ncl, nch);
        if (cur == before || cur <= 1) return W ? weighted_mean<S>(g, gw, cur) : mean;
    }
}

// stack.go:611-705 StackWinsorSigma / stack.go:710-829 StackWinsorSigmaWeighted
template <int S, bool W>
NL_HD float reduce_winsor(float *g, float *gw, float *wz, int cur, float sig_lo, float sig_hi, int &ncl, int &nch) {
    for (;;) {
        float median = qselect_median<S>(g, cur);
        float mean, sd;
        mean_stddev<S>(g, cur, mean, sd);
        sd = winsor_sigma<S>(g, wz, cur, median, sd);
        float lo = nl_subf(median, nl_mulf(sig_lo, sd));
        float hi = nl_addf(median, nl_mulf(sig_hi, sd));
        int before = cur;
        cur = clip_pass<S, W>(g, gw, cur, lo, hi, ncl, nch);
        if (cur == before || cur <= 1) return W ? weighted_mean<S>(g, gw, cur) : mean;
    }
}

// stack.go:536-605 StackMADSigma (single pass; 0/0 -> NaN when everything is clipped, as in Go)
template <int S>
NL_HD float reduce_mad(float *g, float *ad, int cur, float sig_lo, float sig_hi, int &ncl, int &nch) {
    float median = qselect_median<S>(g, cur);
    for (int i = 0; i < cur; i++) ad[i * S] = fabsf(nl_subf(g[i * S], median));
    float mad = qselect_median<S>(ad, cur);
    float sd = nl_mulf(mad, 1.4826f);
    float lo = nl_subf(median, nl_mulf(sig_lo, sd));
    float hi = nl_addf(median, nl_mulf(sig_hi, sd));
    cur = clip_pass<S, false>(g, nullptr, cur, lo, hi, ncl, nch);
    float s = 0.0f;
    for (int i = 0; i < cur; i++) s = nl_addf(s, g[i * S]);
    return nl_divf(s, (float)cur);
}

// stack.go:834-918 StackLinearFit.  ramp[2*c], ramp[2*c+1] = MeanStdDev of 0..c-1 (precomputed per
// length by ramp_mean_stddev; the reference recomputes it per pixel, stats.go:570).
template <int S>
NL_HD float reduce_linfit(float *g, int cur, const float *ramp, float sig_lo, float sig_hi, int &ncl, int &nch) {
    float mean = 0.0f;
    bool first = true;
    for (;;) {
        if (first) sort_column<S>(g, cur); else insertion_sort_column<S>(g, cur);
        first = false;
        // LinearRegression(xs, ys), stats.go:569-586
        float xm = ramp[2 * cur], xsd = ramp[2 * cur + 1];
        float ysd;
        mean_stddev<S>(g, cur, mean, ysd);
        float corr = 0.0f;
        for (int i = 0; i < cur; i++)
            corr = nl_addf(corr, nl_mulf(nl_subf((float)i, xm), nl_subf(g[i * S], mean)));
        corr = nl_divf(corr, nl_mulf(nl_mulf(xsd, ysd), nl_addf((float)cur, 1.0f)));
        float slope = nl_divf(nl_mulf(corr, ysd), xsd);
        float icpt = nl_subf(mean, nl_mulf(slope, xm));
        // mean absolute residual, stack.go:878-886
        float sigma = 0.0f;
        for (int i = 0; i < cur; i++) {
            float lin = nl_addf(nl_mulf((float)i, slope), icpt);
            sigma = nl_addf(sigma, fabsf(nl_subf(g[i * S], lin)));
        }
        sigma = nl_divf(sigma, (float)cur);
        // rejection: overwrite from the front, stack.go:889-909
        int left = 0;
        float lob = nl_mulf(sig_lo, sigma), hib = nl_mulf(sig_hi, sigma);
        for (int i = 0; i < cur; i++) {
            float v = g[i * S];
            float lin = nl_addf(nl_mulf((float)i, slope), icpt);
            if (nl_subf(lin, v) > lob) { g[i * S] = g[left * S]; left++; ncl++; }
            else if (nl_subf(v, lin) > hib) { g[i * S] = g[left * S]; left++; nch++; }
        }
        if (left == 0 || cur < 3) break;
        g += left * S;
        cur -= left;
    }
    return mean;
}

}  // namespace nl
