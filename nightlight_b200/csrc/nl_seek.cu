// nl_seek.cu -- goal-seek of the clipping sigmas for target clip percentages.
//
// Restated from the reference's FindSigmasAndStack / binarySearchAndStack / newtonMethodAndStack
// (internal/ops/stack/stackfindsigma.go:27-170).  That code sits inside a comment block in the reference -- nothing
// calls it, no test pins it: PARITY UNPINNED by construction.  The arithmetic is float32 like the Go code.
//
// The reference runs a full Stack() per trial and keeps the last one.  Here the frames stay resident in a stack job, a
// trial only needs the two clip totals (nl_stack_clip_counts_only: the column kernel with its result stores skipped),
// and the search itself is a small state machine the caller steps, so that several jobs (row stripes on several GPUs,
// batches) can feed one search with their summed totals.
#include "nl_internal.h"

using namespace nl;

namespace {

inline float perc_of(int64_t clipped, float total) { return (float)clipped * 100.0f / total; }    // :64-65, :110-111

}  // namespace

extern "C" {

int nl_sigma_seek_begin(nl_sigma_seek *s, int32_t mode, int32_t n_frames, int64_t pixels, float clip_perc_low, float clip_perc_high) {
    NL_REQUIRE(s && n_frames >= 1 && pixels >= 1, "bad argument");
    if (mode < NL_ST_MEDIAN || mode > NL_ST_AUTO) return set_error(NL_E_INVALID, "invalid stacking mode");
    if (mode == NL_ST_AUTO) mode = nl_auto_select_mode(n_frames);                  // :29-32
    *s = nl_sigma_seek{};
    s->mode = mode;
    s->perc_low = clip_perc_low;
    s->perc_high = clip_perc_high;
    s->total = (float)(pixels * (int64_t)n_frames);                                // float32(len(stack.Data)*len(lights))
    if (mode == NL_ST_SIGMA || mode == NL_ST_WINSOR_SIGMA) {                       // binarySearchAndStack, :50-55
        s->low_l = s->high_l = 1.0f;
        s->low_r = s->high_r = 11.0f;
        s->low_m = 0.5f * (s->low_l + s->low_r);
        s->high_m = 0.5f * (s->high_l + s->high_r);
        s->trial_low = s->low_m;
        s->trial_high = s->high_m;
    } else if (mode == NL_ST_LINEAR_FIT) {                                         // newtonMethodAndStack, :102
        s->sig_lo = s->sig_hi = 6.0f;
        s->trial_low = s->sig_lo;
        s->trial_high = s->sig_hi;
    } else {                                                                       // :41-44: the mode has no sigmas
        s->done = 1;
        s->result_low = s->result_high = 0.0f;
    }
    return NL_OK;
}

// Feeds the clip totals of the trial stack at (trial_low, trial_high).  Returns 1 when the search is over --
// (result_low, result_high) are the sigmas to stack with -- and 0 when another trial at the new (trial_low,
// trial_high) is needed; negative on error.
int nl_sigma_seek_step(nl_sigma_seek *s, int64_t clip_low, int64_t clip_high) {
    NL_REQUIRE(s, "NULL argument");
    if (s->done) return 1;
    s->trials++;
    const float eps = 0.005f;
    if (s->mode == NL_ST_SIGMA || s->mode == NL_ST_WINSOR_SIGMA) {
        const float pl = perc_of(clip_low, s->total), ph = perc_of(clip_high, s->total);
        const int dl = (int)(100.0f * pl + 0.5f) - (int)(100.0f * s->perc_low);       // :66-67
        const int dh = (int)(100.0f * ph + 0.5f) - (int)(100.0f * s->perc_high);
        if ((dl == 0 && dh == 0) || s->step >= 20) {                                  // :69-76
            s->converged = (dl == 0 && dh == 0);
            s->result_low = s->low_m; s->result_high = s->high_m;
            s->done = 1;
            return 1;
        }
        if (dl > 0) { s->low_l = s->low_m; s->low_m = 0.5f * (s->low_l + s->low_r); }              // :82-88
        else if (dl < 0) { s->low_r = s->low_m; s->low_m = 0.5f * (s->low_l + s->low_r); }
        if (dh > 0) { s->high_l = s->high_m; s->high_m = 0.5f * (s->high_l + s->high_r); }         // :91-97
        else if (dh < 0) { s->high_r = s->high_m; s->high_m = 0.5f * (s->high_l + s->high_r); }
        s->step++;
        s->trial_low = s->low_m; s->trial_high = s->high_m;
        return 0;
    }
    // Newton's method on (sigLow, sigHigh), three trials per iteration: at the point, at sigLow+eps, at sigHigh+eps
    auto finish = [&](bool converged) {
        s->converged = converged;
        s->result_low = s->sig_lo; s->result_high = s->sig_hi;
        s->done = 1;
        return 1;
    };
    if (s->phase == 0) {
        const float pl = perc_of(clip_low, s->total), ph = perc_of(clip_high, s->total);
        s->d_l = pl - s->perc_low;
        s->d_h = ph - s->perc_low;                         // the reference subtracts the LOW target here too (:114)
        const int dli = (int)(100.0f * s->d_l + 0.5f), dhi = (int)(100.0f * s->d_h + 0.5f);
        if (dli == 0 && dhi == 0) return finish(true);                                 // :120-123
        if (s->step >= 20) return finish(false);                                       // :124-127
        s->step++;                                                                     // :132
        s->phase = 1;
        s->trial_low = s->sig_lo + eps; s->trial_high = s->sig_hi;
        return 0;
    }
    if (s->phase == 1) {
        const float d2 = perc_of(clip_low, s->total) - s->perc_low;                    // :136-137
        const float diff = (d2 - s->d_l) / eps;
        if (diff == 0.0f) return finish(false);                                        // :139-142
        float nl = s->sig_lo - s->d_l / diff;
        if (nl < 0.1f) nl = 0.1f;
        if (nl > 20.0f) nl = 20.0f;
        s->new_lo = nl;
        s->step++;                                                                     // :150
        s->phase = 2;
        s->trial_low = s->sig_lo; s->trial_high = s->sig_hi + eps;
        return 0;
    }
    const float d3 = perc_of(clip_high, s->total) - s->perc_low;                       // :154-155 (again the low target)
    const float diff = (d3 - s->d_h) / eps;
    if (diff == 0.0f) return finish(false);                                            // :157-160
    float nh = s->sig_hi - s->d_h / diff;
    if (nh < 0.1f) nh = 0.1f;
    if (nh > 20.0f) nh = 20.0f;
    s->sig_lo = s->new_lo; s->sig_hi = nh;                                             // :168
    s->step++;                                                                         // the for loop's own i++
    s->phase = 0;
    s->trial_low = s->sig_lo; s->trial_high = s->sig_hi;
    return 0;
}

// FindSigmasAndStack over one resident job: count-only trials, then one stack at the sigmas found.
int nl_find_sigmas_and_stack(nl_stack_job *job, int32_t mode, const float *weights, float ref_frame_loc, float clip_perc_low,
                             float clip_perc_high, float *host_out, int64_t *clip_low, int64_t *clip_high, float *sigma_low,
                             float *sigma_high, int32_t *trials) {
    NL_REQUIRE(job, "NULL argument");
    int32_t n_frames = 0;
    int64_t pixels = 0;
    int rc = nl_stack_job_shape(job, &n_frames, &pixels);
    if (rc != NL_OK) return rc;
    NL_REQUIRE(pixels >= 1, "empty job");
    nl_sigma_seek s;
    rc = nl_sigma_seek_begin(&s, mode, n_frames, pixels, clip_perc_low, clip_perc_high);
    if (rc != NL_OK) return rc;
    while (!s.done) {
        int64_t lo = 0, hi = 0;
        rc = nl_stack_clip_counts_only(job, s.mode, weights, s.trial_low, s.trial_high, &lo, &hi);
        if (rc != NL_OK) return rc;
        rc = nl_sigma_seek_step(&s, lo, hi);
        if (rc < 0) return rc;
    }
    if (sigma_low) *sigma_low = s.result_low;
    if (sigma_high) *sigma_high = s.result_high;
    if (trials) *trials = s.trials;
    return nl_stack_run(job, s.mode, weights, s.result_low, s.result_high, ref_frame_loc, host_out, clip_low, clip_high);
}

}  // extern "C"
