/*
 * nl_oracle_simd.c -- the reference's four AVX2 kernels replayed with real AVX2/FMA instructions
 * (intrinsics), instruction for instruction in the order of the assembly:
 *   stats_amd64.s:27-92, :103-143; noise_amd64.s:75-192; median3x3_amd64.s:62-235.
 * TEST INFRASTRUCTURE ONLY, and only a cross-check OF THE ORACLE: tests/test_oracle.py compares the
 * scalar lane emulation in nl_oracle_amd64.c (which is what the GPU parity tests use) against this
 * file when the host CPU has AVX2+FMA, so that NaN / signed-zero operand roles and the fused
 * multiply-adds are confirmed by the hardware the reference runs on.  Built separately
 * (-mavx2 -mfma) into libnl_oracle_simd.so; never loaded on a CPU without AVX2.
 */
#include <immintrin.h>
#include <stdint.h>
#include <stddef.h>

int nlo_simd_available(void) { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"); }

/* Go assembler `VOP src2, src1, dst` == intrinsic op(src1, src2) */

void nlo_simd_min_mean_max(const float *data, int64_t n, float *min, float *mean, float *max) {
    const float *p = data, *end = data + n;
    __m128 mn = _mm_loadu_ps(p), mx = mn;
    __m256d sum = _mm256_setzero_pd();
    while (p < end) {
        __m128 x = _mm_loadu_ps(p);
        p += 4;
        mn = _mm_min_ps(mn, x);
        mx = _mm_max_ps(mx, x);
        sum = _mm256_add_pd(sum, _mm256_cvtps_pd(x));
    }
    __m128 t = _mm_permute_ps(mn, (1 << 0) + (0 << 2) + (3 << 4) + (2 << 6));
    mn = _mm_min_ps(mn, t);
    t = _mm_permute_ps(mn, (2 << 0) + (3 << 2) + (0 << 4) + (1 << 6));
    mn = _mm_min_ps(mn, t);
    *min = _mm_cvtss_f32(mn);
    t = _mm_permute_ps(mx, (1 << 0) + (0 << 2) + (3 << 4) + (2 << 6));
    mx = _mm_max_ps(mx, t);
    t = _mm_permute_ps(mx, (2 << 0) + (3 << 2) + (0 << 4) + (1 << 6));
    mx = _mm_max_ps(mx, t);
    *max = _mm_cvtss_f32(mx);
    __m256d s1 = _mm256_permute_pd(sum, 5);
    s1 = _mm256_add_pd(s1, sum);
    __m128d lo = _mm256_extractf128_pd(s1, 0), hi = _mm256_extractf128_pd(s1, 1);
    lo = _mm_add_pd(hi, lo);
    __m128d cnt = _mm_cvtsi64_sd(_mm_setzero_pd(), n);
    lo = _mm_div_sd(lo, cnt);
    *mean = _mm_cvtss_f32(_mm_cvtsd_ss(_mm_setzero_ps(), lo));
}

double nlo_simd_variance(const float *data, int64_t n, float mean) {
    const float *p = data, *end = data + n;
    __m128 m = _mm_set1_ps(mean);
    __m256d sum = _mm256_setzero_pd();
    while (p < end) {
        __m128 x = _mm_loadu_ps(p);
        p += 4;
        x = _mm_sub_ps(x, m);
        __m256d d = _mm256_cvtps_pd(x);
        d = _mm256_mul_pd(d, d);
        sum = _mm256_add_pd(sum, d);
    }
    __m256d s1 = _mm256_permute_pd(sum, 5);
    s1 = _mm256_add_pd(s1, sum);
    __m128d lo = _mm256_extractf128_pd(s1, 0), hi = _mm256_extractf128_pd(s1, 1);
    lo = _mm_add_pd(hi, lo);
    __m128d cnt = _mm_cvtsi64_sd(_mm_setzero_pd(), n);
    lo = _mm_div_sd(lo, cnt);
    return _mm_cvtsd_f64(lo);
}

static const int32_t shift_r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
static const int32_t shift_l[8] = {8, 0, 1, 2, 3, 4, 5, 6};
static const uint32_t abs_filter[8] = {0x7fffffffu, 0x7fffffffu, 0x7fffffffu, 0x7fffffffu, 0x7fffffffu, 0x7fffffffu, 0, 0};
static const uint32_t filter2[16] = {0, 0, 0, 0, 0, 0, 0, 0, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, 0, 0};

float nlo_simd_noise_line(const float *src, int64_t width) {
    const float *si = src, *bp = src + width - 7;
    __m256 acc = _mm256_setzero_ps();
    const __m256i sr = _mm256_loadu_si256((const __m256i *)shift_r);
    const __m256 w1 = _mm256_set1_ps(1.0f), wm2 = _mm256_set1_ps(-2.0f), w4 = _mm256_set1_ps(4.0f);
    __m256 mask = _mm256_loadu_ps((const float *)abs_filter);
    int fix = 0;
    for (;;) {
        while (si < bp || fix) {
            fix = 0;
            __m256 y0 = _mm256_loadu_ps(si), y3 = _mm256_loadu_ps(si + width), y6 = _mm256_loadu_ps(si + 2 * width);
            si += 6;
            __m256 y1 = _mm256_permutevar8x32_ps(y0, sr), y4 = _mm256_permutevar8x32_ps(y3, sr), y7 = _mm256_permutevar8x32_ps(y6, sr);
            __m256 y2 = _mm256_permutevar8x32_ps(y1, sr), y5 = _mm256_permutevar8x32_ps(y4, sr), y8 = _mm256_permutevar8x32_ps(y7, sr);
            y0 = _mm256_mul_ps(w1, y0);
            y1 = _mm256_mul_ps(wm2, y1);
            y2 = _mm256_mul_ps(w1, y2);
            y3 = _mm256_mul_ps(wm2, y3);
            y0 = _mm256_fmadd_ps(w4, y4, y0);
            y1 = _mm256_fmadd_ps(wm2, y5, y1);
            y2 = _mm256_fmadd_ps(w1, y6, y2);
            y3 = _mm256_fmadd_ps(wm2, y7, y3);
            y0 = _mm256_fmadd_ps(w1, y8, y0);
            y2 = _mm256_add_ps(y3, y2);
            y0 = _mm256_add_ps(y1, y0);
            y0 = _mm256_add_ps(y2, y0);
            y0 = _mm256_and_ps(mask, y0);
            acc = _mm256_add_ps(acc, y0);
        }
        ptrdiff_t ax = si - bp;
        if (ax >= 5) break;
        ax += 1;
        si -= ax;
        mask = _mm256_and_ps(mask, _mm256_loadu_ps((const float *)(filter2 + 8 - ax)));
        fix = 1;
    }
    __m256 t = _mm256_permute2f128_ps(acc, acc, 1);
    acc = _mm256_add_ps(acc, t);
    t = _mm256_permute_ps(acc, (2 << 0) + (3 << 2) + (0 << 4) + (1 << 6));
    acc = _mm256_add_ps(acc, t);
    t = _mm256_permute_ps(acc, (1 << 0) + (0 << 2) + (3 << 4) + (2 << 6));
    acc = _mm256_add_ps(acc, t);
    return _mm256_cvtss_f32(acc);
}

#define SWAP(i, j) { __m256 lo_ = _mm256_min_ps(a[j], a[i]); a[j] = _mm256_max_ps(a[j], a[i]); a[i] = lo_; }
#define MAXJ(i, j) { a[j] = _mm256_max_ps(a[j], a[i]); }
#define MINI(i, j) { a[i] = _mm256_min_ps(a[j], a[i]); }

/* dest and source: three rows of `width` floats; writes columns 1..width-2 of the middle row */
void nlo_simd_median_line(float *dest, const float *src, int64_t width) {
    const float *si = src, *bp = src + width - 7;
    float *di = dest + width;
    const __m256i sr = _mm256_loadu_si256((const __m256i *)shift_r), sl = _mm256_loadu_si256((const __m256i *)shift_l);
    const __m256i store = _mm256_setr_epi32(0, -1, -1, -1, -1, -1, -1, 0);
    int fix = 0;
    for (;;) {
        while (si < bp || fix) {
            fix = 0;
            __m256 a[9];
            a[0] = _mm256_loadu_ps(si);
            a[3] = _mm256_loadu_ps(si + width);
            a[6] = _mm256_loadu_ps(si + 2 * width);
            si += 6;
            a[1] = _mm256_permutevar8x32_ps(a[0], sr); a[4] = _mm256_permutevar8x32_ps(a[3], sr); a[7] = _mm256_permutevar8x32_ps(a[6], sr);
            a[2] = _mm256_permutevar8x32_ps(a[1], sr); a[5] = _mm256_permutevar8x32_ps(a[4], sr); a[8] = _mm256_permutevar8x32_ps(a[7], sr);
            SWAP(0, 1) SWAP(3, 4) SWAP(6, 7) SWAP(1, 2) SWAP(4, 5) SWAP(7, 8) SWAP(0, 1) SWAP(3, 4) SWAP(6, 7)
            MAXJ(0, 3) MAXJ(3, 6) SWAP(1, 4) MINI(4, 7) MAXJ(1, 4) MINI(5, 8) MINI(2, 5) SWAP(2, 4) MINI(4, 6) MAXJ(2, 4)
            __m256 r = _mm256_permutevar8x32_ps(a[4], sl);
            _mm256_maskstore_ps(di, store, r);
            di += 6;
        }
        ptrdiff_t ax = si - bp;
        if (ax >= 5) break;
        ax += 1;
        si -= ax;
        di -= ax;
        fix = 1;
    }
}
