/*
 * nl_oracle.h -- CPU restatement of mlnoga/nightlight's stacking hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C, line-by-line
 * restatement of the reference's Go arithmetic (file:line cited per function in
 * nl_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under nightlight_b200/ links, imports
 * or calls it; the product path is CUDA-only and fails loudly without its library.
 *
 * Parity pinning: the reference cannot be compiled here (no Go toolchain, un-vendored
 * modules), and its own tests pin only qsort (internal/qsort/qsort_test.go:25-53).
 * The oracle is pinned by that test, by the hand-derivable known-answer vectors of
 * SURVEY.md section 8c, and by an independent pure-Python float32 transliteration
 * (tests/pyref.py).  Everything the reference does not pin itself is "parity
 * unpinned by the reference" -- see DESIGN.md.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (SSE scalar fp32, never x87, no FMA).
 * Licence: the restated algorithms are GPL-3.0 (reference LICENSE); so is this file.
 */
#ifndef NL_ORACLE_H
#define NL_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* stack.go:33-42 */
enum { NLO_ST_MEDIAN = 0, NLO_ST_MEAN, NLO_ST_SIGMA, NLO_ST_WINSOR, NLO_ST_MAD, NLO_ST_LINFIT, NLO_ST_AUTO };
/* stack.go:57-63 */
enum { NLO_W_NONE = 0, NLO_W_EXPOSURE, NLO_W_INV_NOISE, NLO_W_INV_HFR };

/* star.Star, findstars.go:30-37 */
typedef struct {
    int32_t index;
    float   value;
    float   x, y;
    float   mass;
    float   hfr;
} nlo_star;

/* star.Transform2D, coord.go:52-59 */
typedef struct { float a, b, c, d, e, f; } nlo_transform;

/* ---- internal/qsort/qsort.go ---- */
void  nlo_qsort_f32(float *a, int n);
int   nlo_qpartition_f32(float *a, int n);
float nlo_qselect_f32(float *a, int n, int k);
float nlo_qselect_median_f32(float *a, int n);
float nlo_qselect_first_quartile_f32(float *a, int n);

/* ---- internal/stats/stats.go ---- */
void nlo_mean_stddev(const float *xs, int n, float *mean, float *stddev);
void nlo_linear_regression(const float *xs, const float *ys, int n, float *slope, float *intercept,
                           float *xmean, float *xstddev, float *ymean, float *ystddev);
float nlo_estimate_noise(const float *data, int32_t width, int32_t height);

/* ---- internal/ops/stack/stack.go: the reducers, one work package each ---- */
int  nlo_auto_select_mode(int n_frames);
void nlo_stack_median(const float *const *lights, int n, size_t len, float ref_loc, float *res);
void nlo_stack_mean(const float *const *lights, int n, size_t len, float ref_loc, float *res);
void nlo_stack_mean_weighted(const float *const *lights, const float *w, int n, size_t len, float ref_loc, float *res);
void nlo_stack_sigma(const float *const *lights, int n, size_t len, float ref_loc, float sig_lo, float sig_hi,
                     float *res, int32_t *clip_lo, int32_t *clip_hi);
void nlo_stack_sigma_weighted(const float *const *lights, const float *w, int n, size_t len, float ref_loc,
                              float sig_lo, float sig_hi, float *res, int32_t *clip_lo, int32_t *clip_hi);
void nlo_stack_mad_sigma(const float *const *lights, int n, size_t len, float ref_loc, float sig_lo, float sig_hi,
                         float *res, int32_t *clip_lo, int32_t *clip_hi);
void nlo_stack_winsor_sigma(const float *const *lights, int n, size_t len, float ref_loc, float sig_lo, float sig_hi,
                            float *res, int32_t *clip_lo, int32_t *clip_hi);
void nlo_stack_winsor_sigma_weighted(const float *const *lights, const float *w, int n, size_t len, float ref_loc,
                                     float sig_lo, float sig_hi, float *res, int32_t *clip_lo, int32_t *clip_hi);
void nlo_stack_linear_fit(const float *const *lights, int n, size_t len, float ref_loc, float sig_lo, float sig_hi,
                          float *res, int32_t *clip_lo, int32_t *clip_hi);

/* OpStack.Apply (stack.go:115-227): mode resolution, 8 MiB work packages, worker pool.
 * weights may be NULL.  threads<=0 -> all online CPUs.  Returns 0, or -1 invalid mode,
 * -2 MAD+weights (the reference panics there, stack.go:185). Clip totals are summed the
 * way the reference does (int32 per package) but returned widened to int64. */
int nlo_stack_apply(int mode, const float *const *lights, int n, size_t len, const float *weights,
                    float ref_loc, float sig_lo, float sig_hi, float *res,
                    int64_t *clip_lo, int64_t *clip_hi, int threads);

/* getWeights (stack.go:231-270), on already extracted per-frame scalars. Returns 0 or -1. */
int nlo_get_weights(int weighting, const float *exposure, const float *noise, const float *hfr, int n, float *w);

/* StackIncremental / StackIncrementalFinalize (stack.go:924-944) */
void nlo_stack_incremental(float *stack, const float *light, size_t len, float weight, int first);
void nlo_stack_incremental_finalize(float *stack, size_t len, float weight_sum);

/* OpStackBatches.partition arithmetic (stackbatches.go:121-183), permutation is an input elsewhere. */
int nlo_partition(int64_t num_frames, int64_t width, int64_t height, int64_t stack_memory_mb,
                  int64_t max_threads_in, int has_dark, int has_flat,
                  int64_t *num_batches, int64_t *batch_size, int64_t *max_threads);

/* ---- internal/star/coord.go, internal/fits/project.go ---- */
int  nlo_transform_invert(const nlo_transform *t, nlo_transform *inv);
void nlo_transform_apply(const nlo_transform *t, float x, float y, float *xo, float *yo);
int  nlo_new_transform2d(const float p[12], nlo_transform *t);
int  nlo_project(const float *src, int32_t sw, int32_t sh, float *dst, int32_t dw, int32_t dh,
                 const nlo_transform *trans, float oob);

/* ---- internal/median ---- */
float nlo_median9(float *a);
float nlo_median_f32(float *a, int n);
float nlo_gather_and_median(const float *data, int32_t len, int32_t index, const int32_t *mask, int nmask, float *buffer);
int   nlo_create_mask(int32_t width, float radius, int32_t *mask, int cap);

/* ---- amd64 numerics of the path's neighbours (nl_oracle_amd64.c): stats_amd64.s, noise_amd64.s,
 *      median3x3_amd64.s restated lane by lane, and the pure-Go definitions beside them ---- */
void   nlo_calc_min_mean_max_purego(const float *data, int64_t n, float *min, float *mean, float *max);
double nlo_calc_variance_purego(const float *data, int64_t n, float mean);
void   nlo_calc_min_mean_max_avx2(const float *data, int64_t n, float *min, float *mean, float *max);   /* n % 4 == 0 */
double nlo_calc_variance_avx2(const float *data, int64_t n, float mean);                                /* n % 4 == 0 */
void   nlo_stats(const float *data, int64_t n, int amd64, float out[4]);   /* {min, mean, max, stddev}, stats.go:102-153 */
float  nlo_estimate_noise_line_avx2(const float *rows3, int64_t width);
float  nlo_estimate_noise_amd64(const float *data, int32_t width, int32_t height);
void   nlo_median_filter3x3(float *out, const float *data, int32_t width, int32_t height, int amd64);
int64_t nlo_bad_pixel_map(const float *data, int64_t len, int32_t width, float sigma_low, float sigma_high, int amd64,
                          float *tmp, int32_t *bpm, int64_t cap, float stats[4]);

void   nlo_median_filter_sparse(float *data, int32_t len, int32_t width, const int32_t *indices, int64_t n);
int64_t nlo_op_bad_pixel(float *data, int64_t len, int32_t width, float sigma_low, float sigma_high, int amd64,
                         float *tmp, int32_t *bpm, float stats[4]);

/* ---- internal/star/findstars.go, internal/star/qsort.go ---- */
int   nlo_find_bright_pixels(const float *data, int32_t len, int32_t width, float threshold, int32_t radius,
                             nlo_star *stars, int cap);
int   nlo_reject_bad_pixels(nlo_star *stars, int n, const float *data, int32_t len, int32_t width,
                            float sigma, float median_diff_stddev);
void  nlo_qsort_stars_desc(nlo_star *a, int n);
int   nlo_filter_out_overlaps(nlo_star *stars, int n, int32_t width, int32_t height, int32_t radius);
float nlo_shift_to_center_of_mass(nlo_star *stars, int n, const float *data, int32_t len, int32_t width,
                                  float threshold, int32_t radius);
int   nlo_calc_and_filter_hfr(nlo_star *stars, int n, const float *data, int32_t len, int32_t width,
                              float radius, float location, float star_in_out, float *avg_hfr);
/* FindStars (findstars.go:59-100). median_diff_stddev replaces medianDiffStats.StdDev(); it must be
 * given when bp_sigma>0 (the reference's nil fallback is randomized, hence unpinned). */
int   nlo_find_stars(const float *data, int32_t len, int32_t width, float location, float scale, float star_sig,
                     float bp_sigma, float star_in_out, int32_t radius, float median_diff_stddev,
                     nlo_star *out, int cap, float *sum_of_shifts, float *avg_hfr);

/* ---- synthetic frames, SURVEY.md section 8d (not reference code) ---- */
uint32_t nlo_lowbias32(uint32_t x);
float    nlo_synth_sample(uint32_t p, uint32_t k, uint32_t seed);
void     nlo_synth_frame(float *dst, uint64_t p0, size_t len, uint32_t k, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
