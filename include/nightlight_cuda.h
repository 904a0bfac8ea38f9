/*
 * nightlight_cuda.h -- C ABI of libnightlight_cuda.so, the B200 (sm_100a) implementation of
 * mlnoga/nightlight's data-parallel stacking hot path.
 *
 * The reference has no FFI of its own: the seam is three Go call sites.
 * Every entry point below names the reference interface it replaces (file:line relative to the
 * reference root) and is what a cgo / ctypes binding for that call site binds
 * (INTEGRATION.md shows the cgo stubs).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns NL_OK (0) or a negative NL_E_* code; nl_last_error() returns the
 *     calling thread's message for the last failure;
 *   - "host" pointers are ordinary (pageable or pinned) CPU memory, "dev" pointers are CUDA device
 *     memory on the context's device;
 *   - a context owns one CUDA device and one stream; calls on one context are serialised on that
 *     stream, different contexts are independent (one context per goroutine / thread / rank);
 *   - there is no CPU fallback: without a CUDA device nl_ctx_create fails with NL_E_CUDA.
 */
#ifndef NIGHTLIGHT_CUDA_H
#define NIGHTLIGHT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NL_OK            0
#define NL_E_INVALID    (-1)   /* bad argument; for stacking: "invalid stacking mode" (stack.go:118-120) */
#define NL_E_CUDA       (-2)   /* CUDA runtime / driver error, or no device */
#define NL_E_UNSUPPORTED (-3)  /* MAD sigma with weights: the reference panics there (stack.go:185) */
#define NL_E_SINGULAR   (-4)   /* "Matrix has no inverse" (coord.go:160-163) */
#define NL_E_NOMEM      (-5)
#define NL_E_WEIGHTS    (-6)   /* getWeights error (stack.go:231-270), e.g. missing exposure */

/* Stacking modes, numbered like the reference's StackMode (internal/ops/stack/stack.go:33-42). */
enum {
    NL_ST_MEDIAN = 0, NL_ST_MEAN = 1, NL_ST_SIGMA = 2, NL_ST_WINSOR_SIGMA = 3,
    NL_ST_MAD_SIGMA = 4, NL_ST_LINEAR_FIT = 5, NL_ST_AUTO = 6
};
/* Weighting modes, like StackWeighting (stack.go:57-63). */
enum { NL_W_NONE = 0, NL_W_EXPOSURE = 1, NL_W_INVERSE_NOISE = 2, NL_W_INVERSE_HFR = 3 };

typedef struct nl_ctx nl_ctx;
typedef struct nl_stack_job nl_stack_job;

/* star.Star (internal/star/findstars.go:30-37), same field order and types. */
typedef struct {
    int32_t index;
    float   value;
    float   x, y;
    float   mass;
    float   hfr;
} nl_star;

/* ---- library / context ------------------------------------------------------------------- */
const char *nl_last_error(void);
int  nl_version(void);                            /* 10000*major + 100*minor + patch */
int  nl_device_count(int *count);
int  nl_ctx_create(int device, nl_ctx **ctx);     /* one device + one stream */
int  nl_ctx_destroy(nl_ctx *ctx);
int  nl_ctx_sync(nl_ctx *ctx);                    /* wait for the context's stream */
int  nl_ctx_stream(nl_ctx *ctx, void **stream);   /* the cudaStream_t, for event timing by the caller */
int  nl_ctx_device(nl_ctx *ctx, int *device);
/* free / total device memory: the budget OpStackBatches.partition (stackbatches.go:121-210) sizes its batches
 * from takes the place of the reference's StackMemoryMB (a share of host RAM) */
int  nl_ctx_mem_info(nl_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
int  nl_ctx_launch_count(nl_ctx *ctx, int64_t *launches);
/* Tuning knobs for A/B measurements and tests; the library reads NO environment variables.  Keys:
 *   "defer_passes"  "a,b,.."  after how many clipping passes (cumulative) each launch of the sigma / winsorized-sigma /
 *                   linear-fit kernels hands its unfinished columns to the next launch; "0" = one launch, "" = built-in
 *   "tile_width"    "32" | "16" | "8" | "1" | "0" (automatic): pixels per warp tile of the column kernel
 *   "stats_debug", "stats_force_replay"  "0" | "1": diagnostics of the frame statistics (nl_stats) */
int  nl_ctx_set_tuning(nl_ctx *ctx, const char *key, const char *value);

/* ---- stacking: replaces OpStack.Apply + Stack* (internal/ops/stack/stack.go:115-227, 274-918) --
 * A job holds the N frames (or one row stripe of them: `pixels` = stripe pixels) in device memory,
 * frame-major: frame i at dev_frames + i*pixels.  Go cannot pass [][]float32 through cgo, so frames
 * are handed over one at a time (&f[i].Data[lower], a pointer-free slice). */
int  nl_stack_begin(nl_ctx *ctx, int32_t n_frames, int64_t pixels, nl_stack_job **job);
int  nl_stack_put_frame(nl_stack_job *job, int32_t i, const float *host, int64_t count);   /* H2D, async on the stream */
int  nl_stack_frames_dev(nl_stack_job *job, float **dev_frames, int64_t *frame_stride);   /* device-resident producers */
int  nl_stack_job_shape(nl_stack_job *job, int32_t *n_frames, int64_t *pixels);
/* Runs one stacking pass.  mode/sigma/ref_frame_loc are OpStack's fields (stack.go:66-73); weights is
 * NULL (StWeightNone) or n_frames floats from nl_get_weights.  The result (pixels floats) and the two
 * clip counters (stack.go:140, widened to 64 bit) go to host memory; blocks until they are there.
 * Device memory: the sigma, winsorized-sigma and linear-fit modes keep a pool for columns whose late clipping
 * passes are finished by follow-up launches: up to 25 % of the job's frame bytes (125 % for the
 * linear fit), never more than half of the free device memory, allocated on the first such run and kept with the
 * job; without room for it the modes run in one launch. */
int  nl_stack_run(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                  float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high);
/* OpStack.Apply in ONE call for hosts that can pass all frame pointers at once: host_frames[i] points to frame
 * i (pixels floats; pinned memory uploads at full PCIe speed).  The image is cut into n_stripes row stripes
 * (row_pixels = image width; <= 0: 8 stripes) that alternate on two internal streams, so the upload of one
 * stripe overlaps the stacking of the previous one; blocks until host_out and the clip counters are filled. */
int  nl_stack_apply(nl_ctx *ctx, const float *const *host_frames, int32_t n_frames, int64_t pixels, int64_t row_pixels,
                    int32_t n_stripes, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                    float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high);
/* OpStack.Apply (stack.go:115) over SEVERAL devices in one call: host frame pointers in, ONE host image out.  The
 * image's rows are dealt to the n_ctx contexts (one per device) in contiguous blocks -- device g stacks rows
 * [g*H/G, (g+1)*H/G), the analogue of the reference's fan-out over pixel ranges (stack.go:134-147) -- and every device
 * pipelines its block in n_stripes row stripes like nl_stack_apply, on a host thread of its own.  No exchange between
 * devices.  Register the frames with nl_host_register (or allocate them pinned) for full PCIe speed. */
int  nl_stack_apply_multi(nl_ctx *const *ctxs, int32_t n_ctx, const float *const *host_frames, int32_t n_frames, int64_t pixels,
                          int64_t row_pixels, int32_t n_stripes, int32_t mode, const float *weights, float sigma_low,
                          float sigma_high, float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high);
/* nl_stack_apply keeps its two stripe lanes (device buffers of 2 x n_frames x stripe pixels) in the context
 * between calls; this frees them early (nl_ctx_destroy does it too). */
int  nl_stack_apply_release(nl_ctx *ctx);
/* Same, result left in device memory (dev_out: pixels floats), asynchronous on the context's stream;
 * the clip counters are readable after nl_ctx_sync via nl_stack_clip_counts.  dev_out == NULL: count-only run, the
 * kernels skip their result stores (trial stacks of the sigma goal-seek). */
int  nl_stack_run_dev(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                      float ref_frame_loc, float *dev_out);
/* Same, and the result is ALSO stored to n_peers (<= 8) further device buffers of `pixels` floats each:
 * multi-GPU row stripes pass, for every other rank, the address of this rank's stripe inside that
 * rank's gathered image (peer memory mapped with nl_ipc_open_handle), which fuses the reassembly of the
 * stacked image (the reference is single-process and has no counterpart) into the
 * kernel's epilogue instead of running an all-gather afterwards. */
int  nl_stack_run_dev_bcast(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                            float ref_frame_loc, float *dev_out, float *const *peer_outs, int32_t n_peers);
int  nl_stack_clip_counts(nl_stack_job *job, int64_t *clip_low, int64_t *clip_high);
/* one stacking pass that only counts what it clips; blocks until the two totals are there */
int  nl_stack_clip_counts_only(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                               int64_t *clip_low, int64_t *clip_high);
int  nl_stack_end(nl_stack_job *job);

/* ---- sigma goal-seek: FindSigmasAndStack / binarySearchAndStack / newtonMethodAndStack
 * (internal/ops/stack/stackfindsigma.go:27-170 -- DEAD code in the reference, inside a comment block: parity
 * unpinned).  Finds the clipping sigmas that reach target clip percentages: bisection on [1,11] per side for the
 * sigma / winsorized-sigma modes, Newton's method for the linear fit (which, like the reference, measures the high
 * side against the LOW target, :114,:155), at most 21 trial stacks.  The search is a state machine so that the trial
 * totals may come from anywhere (one job, the row stripes of several GPUs summed, a batch):
 *     nl_sigma_seek_begin(&s, ...);
 *     while (!s.done) { run a count-only stack at (s.trial_low, s.trial_high); nl_sigma_seek_step(&s, clipLow, clipHigh); }
 *     stack at (s.result_low, s.result_high)
 * Plain data, caller-allocated; only the fields named above are for the caller. */
typedef struct {
    int32_t mode;                       /* resolved mode (never NL_ST_AUTO) */
    int32_t done, converged, trials;
    float   trial_low, trial_high;      /* sigmas of the next trial stack */
    float   result_low, result_high;    /* valid when done */
    /* internal state */
    int32_t step, phase;
    float   perc_low, perc_high, total;
    float   low_l, low_r, low_m, high_l, high_r, high_m;
    float   sig_lo, sig_hi, d_l, d_h, new_lo;
} nl_sigma_seek;
int  nl_sigma_seek_begin(nl_sigma_seek *s, int32_t mode, int32_t n_frames, int64_t pixels, float clip_perc_low, float clip_perc_high);
int  nl_sigma_seek_step(nl_sigma_seek *s, int64_t clip_low, int64_t clip_high);   /* 1: done, 0: another trial, < 0: error */
/* FindSigmasAndStack over one resident job: count-only trials, then one stack at the sigmas found */
int  nl_find_sigmas_and_stack(nl_stack_job *job, int32_t mode, const float *weights, float ref_frame_loc, float clip_perc_low,
                              float clip_perc_high, float *host_out, int64_t *clip_low, int64_t *clip_high, float *sigma_low,
                              float *sigma_high, int32_t *trials);
/* autoSelectStackingMode (stack.go:45-55) */
int  nl_auto_select_mode(int32_t n_frames);
/* getWeights (stack.go:231-270) on the per-frame scalars it reads (Exposure, Stats.Noise(), HFR). */
int  nl_get_weights(int32_t weighting, const float *exposure, const float *noise, const float *hfr,
                    int32_t n_frames, float *weights);

/* ---- numerics of the frame statistics ------------------------------------------------------
 * The reference computes EstimateNoise, Stats.Min/Mean/Max/StdDev and MedianFilter3x3 with AVX2 assembly on
 * amd64 CPUs that have AVX2 (cpuid dispatch: noise_amd64.go:25-30, stats_amd64.go:24-45, median3x3_amd64.go:26-32)
 * and with pure-Go loops everywhere else; the two round differently (lane order, fused multiply-adds, min/max
 * operand roles).  A context reproduces one of them bit for bit; the default is the amd64 one.  Images too
 * narrow or arrays too ragged for the SIMD kernels (width < 8, length % 4 != 0: the assembly reads outside
 * its slice there) take the pure-Go definition in either setting. */
enum { NL_NUMERICS_AMD64 = 0, NL_NUMERICS_PUREGO = 1 };
int  nl_ctx_set_numerics(nl_ctx *ctx, int32_t numerics);
/* how many float64 summation chains had to be replayed in order because the interval around the parallel sum
 * straddled a float32 rounding boundary (diagnostics; see nl_prestats.cu) */
int  nl_ctx_exact_replays(nl_ctx *ctx, int64_t *replays);

/* ---- noise estimate: replaces stats.EstimateNoise (internal/stats/noise_amd64.go:25-43 + noise_amd64.s:75-192,
 * or noise.go:24-55 in pure-Go numerics),
 * the per-frame scalar behind StWeightInverseNoise (stack.go:247-259).  n_frames frames of width x height
 * pixels, frame i at dev_frames + i*frame_stride (e.g. the buffer of a stack job holding whole frames);
 * one launch for all frames, results to host memory. */
int  nl_estimate_noise_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, int32_t width,
                           int32_t height, float *host_noise);
int  nl_estimate_noise(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float *noise);

/* ---- frame statistics and bad-pixel map -----------------------------------
 * nl_median_filter3x3: median.MedianFilter3x3 (internal/median/median3x3_amd64.go:24-48, median3x3.go:26-110):
 *   interior pixels = median of their 3x3 neighbourhood, border rows and columns copied.
 * nl_stats: stats = {Min, Mean, Max, StdDev} of Stats (internal/stats/stats.go:102-153; calcMinMeanMax and
 *   calcVariance, stats_amd64.s:27-143 / stats.go:264-287): float64 sums in the reference's lane order,
 *   mean = float32(sum/len), stddev = float32(sqrt(sum of squared fp32 deviations / len)).
 * nl_bad_pixel_map: pre.BadPixelMap (internal/ops/pre/badpixels.go:32-51): tmp = data - median3x3(data),
 *   stats of tmp (the reference's medianDiffStats; stats[3] is the medianDiffStats.StdDev() that star
 *   detection takes, findstars.go:134-169), indices with tmp < -stddev*sigma_low or tmp > stddev*sigma_high
 *   in ascending order.  *count = number found; the first min(count, cap) are written to bpm.
 * Device pointers may have any float alignment, except dev_tmp of nl_bad_pixel_map_dev (16 bytes). */
int  nl_median_filter3x3(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float *host_out);
int  nl_median_filter3x3_dev(nl_ctx *ctx, const float *dev_data, int32_t width, int32_t height, float *dev_out);
int  nl_stats(nl_ctx *ctx, const float *host_data, int64_t len, float stats[4]);
int  nl_stats_dev(nl_ctx *ctx, const float *dev_data, int64_t len, float stats[4]);
int  nl_bad_pixel_map(nl_ctx *ctx, const float *host_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                      int32_t *host_bpm, int64_t cap, int64_t *count, float stats[4]);
int  nl_bad_pixel_map_dev(nl_ctx *ctx, const float *dev_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                          float *dev_tmp, int32_t *host_bpm, int64_t cap, int64_t *count, float stats[4]);
/* BadPixelMap of every frame of a resident stack (frame i at dev_frames + i*frame_stride, 16-byte aligned) in one call:
 * stats = n_frames x 4, counts = n_frames, frame i's list at host_bpm + i*cap. */
int  nl_bad_pixel_map_batch_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, int64_t len,
                                int32_t width, float sigma_low, float sigma_high, int32_t *host_bpm, int64_t cap, int64_t *counts,
                                float *stats);
/* OpBadPixel.Apply for monochrome frames (internal/ops/pre/preprocess.go:180-191): BadPixelMap, then
 * MedianFilterSparse (badpixels.go:79-85) repairs the listed pixels of host_data in place, one after the other.
 * Returns immediately when either sigma is 0, like the reference.  *removed = number of repaired pixels;
 * stats = the frame's MedianDiffStats (min, mean, max, stddev). */
int  nl_op_bad_pixel(nl_ctx *ctx, float *host_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                     int64_t *removed, float stats[4]);
/* The same for a frame that is already resident: dev_data holds the pixels of host_data; both copies are repaired,
 * so the operators that follow (nl_estimate_noise_dev, nl_find_stars_dev, nl_project_dev / _scatter_dev) work on
 * the device copy without another upload. */
int  nl_op_bad_pixel_dev(nl_ctx *ctx, float *dev_data, float *host_data, int64_t len, int32_t width, float sigma_low,
                         float sigma_high, int64_t *removed, float stats[4]);

/* ---- batches: replaces StackIncremental / StackIncrementalFinalize (stack.go:924-944) -------
 * acc = light*weight (first != 0) or acc += light*weight; then acc *= 1/weight_sum.  Device buffers. */
int  nl_stack_incremental_dev(nl_ctx *ctx, float *dev_acc, const float *dev_light, int64_t pixels, float weight, int first);
int  nl_stack_incremental_finalize_dev(nl_ctx *ctx, float *dev_acc, int64_t pixels, float weight_sum);

/* ---- resample: replaces (*Image).Project (internal/fits/project.go:26-76) ------------------
 * trans = Transform2D{A..F} (internal/star/coord.go:52-59); inverted on the host exactly like
 * Transform2D.Invert (coord.go:159-201); out_of_bounds = fill value (NaN for stacking). */
int  nl_transform_invert(const float trans[6], float inv[6]);
int  nl_project(nl_ctx *ctx, const float *host_src, int32_t src_w, int32_t src_h,
                float *host_dst, int32_t dst_w, int32_t dst_h, const float trans[6], float out_of_bounds);
int  nl_project_dev(nl_ctx *ctx, const float *dev_src, int32_t src_w, int32_t src_h,
                    float *dev_dst, int32_t dst_w, int32_t dst_h, const float trans[6], float out_of_bounds);

/* Resample fused with the histogram match that runs just before it (OpMatchHistogram -> Image.MatchHistogram,
 * internal/ops/post/postprocess.go:74-94, internal/fits/pixelops.go:601-612): every source sample is read as
 * d*multiplier + offset (mul, then add), which saves one full pass over the frame. */
int  nl_project_scaled(nl_ctx *ctx, const float *host_src, int32_t src_w, int32_t src_h, float *host_dst, int32_t dst_w,
                       int32_t dst_h, const float trans[6], float out_of_bounds, float multiplier, float offset);
int  nl_project_scaled_dev(nl_ctx *ctx, const float *dev_src, int32_t src_w, int32_t src_h, float *dev_dst, int32_t dst_w,
                           int32_t dst_h, const float trans[6], float out_of_bounds, float multiplier, float offset);
/* OpAlign over ALL frames of a resident stack in one launch (postprocess.go:142-191 runs Project once per frame from a
 * pool of goroutines): frame i is read at dev_src + i*src_stride and written at dev_dst + i*dst_stride -- e.g. straight
 * into slot i of a stack job (nl_stack_frames_dev).  trans = n_frames x 6 floats (Transform2D each); multipliers /
 * offsets = n_frames floats each (histogram match fused, as in nl_project_scaled) or both NULL. */
int  nl_project_batch_dev(nl_ctx *ctx, const float *dev_src, int64_t src_stride, int32_t src_w, int32_t src_h, float *dev_dst,
                          int64_t dst_stride, int32_t dst_w, int32_t dst_h, int32_t n_frames, const float *trans,
                          float out_of_bounds, const float *multipliers, const float *offsets);
/* Frame-sharded resample feeding row-sharded stacking.  Resamples one frame like
 * nl_project_scaled_dev (multiplier 1, offset 0 = plain Project) and stores destination rows
 * [stripe_row0[g], stripe_row0[g+1]) at stripe_frames[g] + frame_index * rows_g * dw, i.e. as frame `frame_index`
 * of the frame-major buffer of the stack job that owns stripe g (nl_stack_frames_dev).  Buffers of other GPUs are
 * passed as their peer mappings (nl_ipc_open_handle): the exchange between the two shardings rides on the resample's
 * stores.  The reference has no counterpart (one process: OpAlign fills f.Data, OpStack reads it, postprocess.go:142-191). */
int  nl_project_scatter_dev(nl_ctx *ctx, const float *dev_src, int32_t sw, int32_t sh, int32_t dw, int32_t dh, const float trans[6],
                            float oob, float multiplier, float offset, int32_t frame_index, void *const *stripe_frames,
                            const int32_t *stripe_row0, int32_t n_stripes);

/* ---- FITS pixel payload: replaces the conversion loops of internal/fits/read.go:176-443 (big-endian BITPIX
 * 8/16/32/64/-32/-64 -> fp32, v = float32(val)*Bscale + Bzero) and write.go:182-215 (fp32 -> big-endian,
 * NaN -> 0).  nl_stack_put_frame_raw uploads a frame as its raw payload (half the PCIe bytes for 16-bit
 * frames) and decodes it into the job on the device. */
int  nl_fits_decode(nl_ctx *ctx, const void *host_raw, int32_t bitpix, int64_t count, float bscale, float bzero, float *host_dst);
int  nl_fits_decode_dev(nl_ctx *ctx, const void *dev_raw, int32_t bitpix, int64_t count, float bscale, float bzero, float *dev_dst);
int  nl_fits_encode(nl_ctx *ctx, const float *host_src, int64_t count, void *host_raw);
int  nl_fits_encode_dev(nl_ctx *ctx, const float *dev_src, int64_t count, void *dev_raw);
int  nl_stack_put_frame_raw(nl_stack_job *job, int32_t i, const void *host_raw, int32_t bitpix, int64_t count, float bscale,
                            float bzero);

/* ---- star detection: replaces star.FindStars (internal/star/findstars.go:59-100) -----------
 * nl_find_bright = findBrightPixels (findstars.go:105-129): candidates in raster order.  *count is
 * the true number found; at most cap are stored. */
int  nl_find_bright(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float threshold,
                    int32_t radius, nl_star *out, int32_t cap, int32_t *count);
int  nl_find_bright_dev(nl_ctx *ctx, const float *dev_data, int32_t len, int32_t width, float threshold,
                        int32_t radius, nl_star *out, int32_t cap, int32_t *count);
/* The whole FindStars pipeline.  median_diff_stddev stands for medianDiffStats.StdDev(); it must be
 * given when bp_sigma > 0 (the reference's nil fallback draws a random sample, findstars.go:139-150). */
int  nl_find_stars(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float location, float scale,
                   float star_sig, float bp_sigma, float star_in_out, int32_t radius, float median_diff_stddev,
                   nl_star *out, int32_t cap, int32_t *count, float *sum_of_shifts, float *avg_hfr);
/* FindStars on a resident frame: the full-frame scan reads dev_data, the sparse per-star steps read host_data
 * (the same pixels). */
int  nl_find_stars_dev(nl_ctx *ctx, const float *dev_data, const float *host_data, int32_t len, int32_t width, float location,
                       float scale, float star_sig, float bp_sigma, float star_in_out, int32_t radius, float median_diff_stddev,
                       nl_star *out, int32_t cap, int32_t *count, float *sum_of_shifts, float *avg_hfr);

/* The same for ALL frames of a resident stack (frame i at dev_frames + i*frame_stride), as OpStarDetect runs over a
 * frame set (preprocess.go:440-465, one goroutine per frame):
 * nl_find_bright_batch_dev: one read of every frame (candidates go to per-row slots, then into raster order) and two
 *   host round trips for the whole set; thresholds = n_frames floats; frame i's candidates at host_out + i*cap.
 * nl_find_stars_batch_dev: that scan, then the sparse per-star steps of every frame on host threads reading
 *   host_frames[i] (the same pixels).  location, scale, median_diff_stddev, counts, sum_of_shifts, avg_hfr are arrays
 *   of n_frames entries; frame i's stars at out + i*cap.  seconds_device / seconds_host (may be NULL) report the time
 *   spent in the device scan and in the host steps. */
int  nl_find_bright_batch_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, int32_t len,
                              int32_t width, const float *thresholds, int32_t radius, nl_star *host_out, int32_t cap,
                              int32_t *counts);
int  nl_find_stars_batch_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride,
                             const float *const *host_frames, int32_t len, int32_t width, const float *location,
                             const float *scale, float star_sig, float bp_sigma, float star_in_out, int32_t radius,
                             const float *median_diff_stddev, nl_star *out, int32_t cap, int32_t *counts,
                             float *sum_of_shifts, float *avg_hfr, double *seconds_device, double *seconds_host);

/* The sparse, order-dependent steps of FindStars as the library runs them on host threads between its kernels, callable
 * on their own (no device, no context): for callers that keep part of FindStars in Go, and for the CPU parity tests.
 * nl_star_reject_bad_pixels_host: rejectBadPixels (findstars.go:134-169) the way the batched path runs it -- the test of
 *   every candidate whose 3x3 neighbourhood lies inside the frame computed independently (the device's part), the
 *   candidates of the first and last row replayed with the reference's carried-over gather buffer.
 * nl_star_sort_desc_host: QSortStarsDesc (star/qsort.go:25-55), the reference's unstable quicksort step for step.
 * nl_star_filter_overlaps_host: filterOutOverlaps (findstars.go:209-271), on the fine grid described in DESIGN.md 3.4.
 * Each works in place; *kept = stars remaining. */
int  nl_star_reject_bad_pixels_host(nl_star *stars, int32_t n, const float *data, int32_t len, int32_t width, float sigma,
                                    float median_diff_stddev, int32_t *kept);
int  nl_star_sort_desc_host(nl_star *stars, int32_t n);
int  nl_star_filter_overlaps_host(nl_star *stars, int32_t n, int32_t width, int32_t height, int32_t radius, int32_t *kept);

/* ---- synthetic frames (the workload generator, not reference code) --- */
int  nl_synth_fill_dev(nl_ctx *ctx, float *dev_dst, uint64_t p0, int64_t count, uint32_t frame, uint32_t seed);

/* ---- plain device memory helpers for bindings without a CUDA runtime of their own --------- */
int  nl_dev_alloc(nl_ctx *ctx, int64_t bytes, void **dev);
int  nl_dev_free(nl_ctx *ctx, void *dev);
/* cross-process peer mapping of a buffer from nl_dev_alloc (CUDA IPC, one process per GPU) */
int  nl_ipc_get_handle(nl_ctx *ctx, void *dev, unsigned char handle[64]);
int  nl_ipc_open_handle(nl_ctx *ctx, const unsigned char handle[64], void **dev);
int  nl_ipc_close_handle(nl_ctx *ctx, void *dev);
int  nl_host_alloc_pinned(int64_t bytes, void **host);
int  nl_host_free_pinned(void *host);
int  nl_host_register(void *host, int64_t bytes);     /* page-lock an existing buffer in place (cudaHostRegister) */
int  nl_host_unregister(void *host);
int  nl_memcpy_h2d(nl_ctx *ctx, void *dev, const void *host, int64_t bytes);   /* async on the stream */
int  nl_memcpy_d2h(nl_ctx *ctx, void *host, const void *dev, int64_t bytes);   /* async on the stream */
int  nl_memcpy_d2d(nl_ctx *ctx, void *dev_dst, const void *dev_src, int64_t bytes);   /* async on the stream */

#ifdef __cplusplus
}
#endif
#endif
