/*
 * nl_oracle_amd64.c -- CPU restatement of the reference's amd64 numerics for the neighbours of the
 * stacking path: the AVX2 assembly kernels that every amd64 build of the reference runs when the CPU
 * reports AVX2 (cpuid dispatch: stats_amd64.go:24-45, noise_amd64.go:25-30, median3x3_amd64.go:26-32),
 * and the pure-Go definitions they replace.  TEST INFRASTRUCTURE ONLY (see nl_oracle.h).
 *
 * The SIMD kernels are restated lane by lane in scalar C: what matters for bit parity is which
 * element goes to which lane, the order of the dependent additions in every lane, where the kernels
 * fuse a multiply-add (VFMADD231PS) and the operand roles of VMINPS / VMAXPS ("second source wins"
 * on NaN and on +-0 ties).  nl_oracle_simd.c replays the same instruction sequences with real AVX2
 * instructions as a cross-check of this file.
 *
 * Parity pinning: the reference has no test for any of these kernels -> unpinned by the reference.
 */
#include "nl_oracle.h"

#include <math.h>
#include <string.h>

/* Intel SDM semantics in Go-assembler operand order `VMINPS src2, src1, dst`:
 * dst = src1 < src2 ? src1 : src2 (src2 on NaN and when both are zeros of either sign). */
static inline float minps(float src1, float src2) { return src1 < src2 ? src1 : src2; }
static inline float maxps(float src1, float src2) { return src1 > src2 ? src1 : src2; }

/* ---------------------------------------------------------------------------------------------
 * stats: calcMinMeanMax / calcVariance
 * ------------------------------------------------------------------------------------------- */

/* stats.go:264-277 calcMinMeanMaxPureGo */
void nlo_calc_min_mean_max_purego(const float *data, int64_t n, float *min, float *mean, float *max) {
    float mmin = data[0], mmax = data[0];
    double mmean = 0.0;
    for (int64_t i = 0; i < n; i++) {
        float mv = data[i];
        if (mv < mmin) mmin = mv;
        if (mv > mmax) mmax = mv;
        mmean += (double)mv;
    }
    *min = mmin;
    *mean = (float)(mmean / (double)n);
    *max = mmax;
}

/* stats.go:280-287 calcVariancePureGo */
double nlo_calc_variance_purego(const float *data, int64_t n, float mean) {
    double variance = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double diff = (double)(data[i] - mean);
        variance += diff * diff;
    }
    return variance / (double)n;
}

/* stats_amd64.s:27-92 calcMinMeanMaxAVX2: four lanes (element i -> lane i%4); min/max start from
 * the first vector and see it again in the first loop round; sums are float64 per lane; the lanes
 * are folded (0,1),(2,3) then across; mean = float32(sum / float64(len)).  The loop reads whole
 * vectors while the pointer is below the end, so len%4 != 0 reads past the slice: callers of the
 * oracle keep len a multiple of 4 for this variant. */
void nlo_calc_min_mean_max_avx2(const float *data, int64_t n, float *min, float *mean, float *max) {
    float mn[4], mx[4];
    double sum[4] = {0, 0, 0, 0};
    for (int j = 0; j < 4; j++) mn[j] = mx[j] = data[j];
    for (int64_t i = 0; i < n; i += 4)
        for (int j = 0; j < 4; j++) {
            float x = data[i + j];
            mn[j] = minps(mn[j], x);      /* VMINPS X1, X2, X2 */
            mx[j] = maxps(mx[j], x);      /* VMAXPS X1, X3, X3 */
            sum[j] += (double)x;          /* VCVTPS2PD, VADDPD */
        }
    /* :66-77: X4 = lanes (1,0,3,2); X2 = min(X2, X4); X4 = lanes (2,3,0,1); X2 = min(X2, X4); lane 0 */
    float a0 = minps(mn[0], mn[1]), a2 = minps(mn[2], mn[3]);
    *min = minps(a0, a2);
    float b0 = maxps(mx[0], mx[1]), b2 = maxps(mx[2], mx[3]);
    *max = maxps(b0, b2);
    /* :80-84 */
    double s = (sum[2] + sum[3]) + (sum[0] + sum[1]);
    *mean = (float)(s / (double)n);
}

/* stats_amd64.s:103-143 calcVarianceAVX2: d = x - mean in fp32 (SUBPS), widened, squared and added in
 * float64 (VMULPD then VADDPD, not fused), four lanes, same fold. */
double nlo_calc_variance_avx2(const float *data, int64_t n, float mean) {
    double sum[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n; i += 4)
        for (int j = 0; j < 4; j++) {
            float d = data[i + j] - mean;
            double dd = (double)d;
            double sq = dd * dd;
            sum[j] += sq;
        }
    double s = (sum[2] + sum[3]) + (sum[0] + sum[1]);
    return s / (double)n;
}

/* Stats.Min/Mean/Max/StdDev (stats.go:102-153): out = {min, mean, max, stddev}.
 * amd64 != 0 selects the AVX2 kernels, which needs n % 4 == 0. */
void nlo_stats(const float *data, int64_t n, int amd64, float out[4]) {
    double var;
    if (amd64) {
        nlo_calc_min_mean_max_avx2(data, n, &out[0], &out[1], &out[2]);
        var = nlo_calc_variance_avx2(data, n, out[1]);
    } else {
        nlo_calc_min_mean_max_purego(data, n, &out[0], &out[1], &out[2]);
        var = nlo_calc_variance_purego(data, n, out[1]);
    }
    out[3] = (float)sqrt(var);   /* stats.go:148-149 */
}

/* ---------------------------------------------------------------------------------------------
 * noise: EstimateNoise, AVX2 variant
 * ------------------------------------------------------------------------------------------- */

/* One flight of noise_amd64.s:106-164: lanes 0..5 hold the 3x3 neighbourhoods of the centre pixels
 * x0+1 .. x0+6 of the middle row; lanes 6,7 are always masked.  Four partial chains per lane:
 *   y0 = d00*1, y1 = d01*-2, y2 = d02*1, y3 = d10*-2,
 *   y0 = fma(d11,4,y0), y1 = fma(d12,-2,y1), y2 = fma(d20,1,y2), y3 = fma(d21,-2,y3), y0 = fma(d22,1,y0),
 *   y2 = y3+y2, y0 = y1+y0, y0 = y2+y0; |y0| masked; running sum += */
static void noise_flight(const float *r0, const float *r1, const float *r2, int64_t x0, const int valid[6], float acc[6]) {
    for (int l = 0; l < 6; l++) {
        if (!valid[l]) continue;   /* masked lanes add +0 to a non-negative sum: no change */
        const float *a = r0 + x0 + l, *b = r1 + x0 + l, *c = r2 + x0 + l;
        float y0 = a[0] * 1.0f, y1 = a[1] * -2.0f, y2 = a[2] * 1.0f, y3 = b[0] * -2.0f;
        y0 = fmaf(b[1], 4.0f, y0);
        y1 = fmaf(b[2], -2.0f, y1);
        y2 = fmaf(c[0], 1.0f, y2);
        y3 = fmaf(c[1], -2.0f, y3);
        y0 = fmaf(c[2], 1.0f, y0);
        y2 = y3 + y2;
        y0 = y1 + y0;
        y0 = y2 + y0;
        acc[l] = fabsf(y0) + acc[l];
    }
}

/* noise_amd64.s:75-192 estimateNoiseLineAVX2 on three rows of `width` floats.  Flights step by 6
 * columns while the 8-wide load stays inside the row; if columns remain, one more flight is placed
 * flush with the end of the row and the lanes already covered are masked off (:166-181). */
float nlo_estimate_noise_line_avx2(const float *rows3, int64_t width) {
    const float *r0 = rows3, *r1 = rows3 + width, *r2 = rows3 + 2 * width;
    float acc[6] = {0, 0, 0, 0, 0, 0};
    int valid[6] = {1, 1, 1, 1, 1, 1};
    int64_t si = 0, bp = width - 7;
    for (;;) {
        while (si < bp) {
            noise_flight(r0, r1, r2, si, valid, acc);
            si += 6;
        }
        int64_t ax = si - bp;
        if (ax >= 5) break;
        ax += 1;
        si -= ax;
        for (int l = 0; l < 6; l++) valid[l] = valid[l] && l >= ax;   /* filterMask2 read `ax` entries early */
        noise_flight(r0, r1, r2, si, valid, acc);
        si += 6;
    }
    /* :183-190 butterfly: (l, l^4), then (l, l^2), then (l, l^1); lanes 6 and 7 hold zero */
    float s0 = acc[0] + acc[4], s1 = acc[1] + acc[5], s2 = acc[2] + 0.0f, s3 = acc[3] + 0.0f;
    float t0 = s0 + s2, t1 = s1 + s3;
    return t0 + t1;
}

/* noise_amd64.go:33-43 estimateNoiseAVX2; the kernel needs width >= 8 (narrower rows make it read
 * before the row start), so narrower images take the pure-Go definition. */
float nlo_estimate_noise_amd64(const float *data, int32_t width, int32_t height) {
    if (width < 8) return nlo_estimate_noise(data, width, height);
    float sum = 0.0f;
    for (int line = 0; line < height - 2; line++) {
        float noise = nlo_estimate_noise_line_avx2(data + (size_t)line * width, width);
        sum += noise;
    }
    float factor = (float)sqrt(0.5 * M_PI) / (6.0f * (float)(width - 2) * (float)(height - 2));
    return sum * factor;
}

/* ---------------------------------------------------------------------------------------------
 * median: MedianFilter3x3 and pre.BadPixelMap
 * ------------------------------------------------------------------------------------------- */

/* The 19 steps of the median-of-9 network (median3x3.go:85-110, median3x3_amd64.s:124-213):
 * 's' = exchange (i gets the smaller, j the larger), 'x' = a[j] = max, 'n' = a[i] = min. */
static const struct { char op; signed char i, j; } net9[19] = {
    {'s', 0, 1}, {'s', 3, 4}, {'s', 6, 7}, {'s', 1, 2}, {'s', 4, 5}, {'s', 7, 8}, {'s', 0, 1}, {'s', 3, 4}, {'s', 6, 7},
    {'x', 0, 3}, {'x', 3, 6}, {'s', 1, 4}, {'n', 4, 7}, {'x', 1, 4}, {'n', 5, 8}, {'n', 2, 5}, {'s', 2, 4}, {'n', 4, 6},
    {'x', 2, 4}};

/* amd64 == 0: MedianFloat32Slice9's compare-and-swap form (NaN never moves);
 * amd64 != 0: the assembly's VMINPS/VMAXPS form, always (src1 = a[j], src2 = a[i]). */
static float median9_net(float a[9], int amd64) {
    for (int k = 0; k < 19; k++) {
        const int i = net9[k].i, j = net9[k].j;
        const float ai = a[i], aj = a[j];
        if (amd64) {
            if (net9[k].op != 'x') a[i] = minps(aj, ai);
            if (net9[k].op != 'n') a[j] = maxps(aj, ai);
        } else if (ai > aj) {
            if (net9[k].op != 'x') a[i] = aj;
            if (net9[k].op != 'n') a[j] = ai;
        }
    }
    return a[4];
}

/* median3x3.go:26-38 / median3x3_amd64.go:36-48: border rows and columns copied, interior = median of
 * the 3x3 neighbourhood in row-major gather order.  The AVX2 line kernel (width >= 8) stores the same
 * six medians per flight, so only the min/max semantics distinguish the two variants. */
void nlo_median_filter3x3(float *out, const float *data, int32_t width, int32_t height, int amd64) {
    if (width < 8) amd64 = 0;
    if (height <= 0 || width <= 0) return;
    memcpy(out, data, sizeof(float) * (size_t)width);
    for (int32_t y = 1; y < height - 1; y++) {
        const float *r = data + (size_t)y * width;
        float *o = out + (size_t)y * width;
        o[0] = r[0];
        for (int32_t x = 1; x < width - 1; x++) {
            float g[9];
            for (int dy = -1, k = 0; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) g[k++] = r[(ptrdiff_t)dy * width + x + dx];
            o[x] = median9_net(g, amd64);
        }
        o[width - 1] = r[width - 1];
    }
    memcpy(out + (size_t)(height - 1) * width, data + (size_t)(height - 1) * width, sizeof(float) * (size_t)width);
}

/* pre.BadPixelMap, badpixels.go:32-51.  tmp (len floats) receives data - median3x3(data); returns the
 * number of bad pixels (their indices, ascending, in bpm[0..min(count,cap))) and the stats of tmp as
 * {min, mean, max, stddev} (the reference's medianDiffStats). */
int64_t nlo_bad_pixel_map(const float *data, int64_t len, int32_t width, float sigma_low, float sigma_high, int amd64,
                          float *tmp, int32_t *bpm, int64_t cap, float stats[4]) {
    int32_t height = (int32_t)(len / width);
    nlo_median_filter3x3(tmp, data, width, height, amd64);
    for (int64_t i = 0; i < len; i++) tmp[i] = data[i] - tmp[i];   /* Subtract(tmp, data, tmp) */
    nlo_stats(tmp, len, amd64 && len % 4 == 0, stats);
    float threshold_low = -stats[3] * sigma_low, threshold_high = stats[3] * sigma_high;
    int64_t count = 0;
    for (int64_t i = 0; i < len; i++) {
        float t = tmp[i];
        if (t < threshold_low || t > threshold_high) {
            if (count < cap) bpm[count] = (int32_t)i;
            count++;
        }
    }
    return count;
}

/* pre.MedianFilterSparse, badpixels.go:79-85 (in place, in list order) with the radius-1.5 mask of
 * OpBadPixel.Apply (preprocess.go:188-189) */
void nlo_median_filter_sparse(float *data, int32_t len, int32_t width, const int32_t *indices, int64_t n) {
    int32_t mask[16];
    int nmask = nlo_create_mask(width, 1.5f, mask, 16);
    float buffer[16] = {0};
    for (int64_t k = 0; k < n; k++) data[indices[k]] = nlo_gather_and_median(data, len, indices[k], mask, nmask, buffer);
}

/* OpBadPixel.Apply, monochrome path (preprocess.go:180-191).  Returns the number of repaired pixels. */
int64_t nlo_op_bad_pixel(float *data, int64_t len, int32_t width, float sigma_low, float sigma_high, int amd64,
                         float *tmp, int32_t *bpm, float stats[4]) {
    if (sigma_low == 0.0f || sigma_high == 0.0f) return 0;
    int64_t n = nlo_bad_pixel_map(data, len, width, sigma_low, sigma_high, amd64, tmp, bpm, len, stats);
    nlo_median_filter_sparse(data, (int32_t)len, width, bpm, n);
    return n;
}
