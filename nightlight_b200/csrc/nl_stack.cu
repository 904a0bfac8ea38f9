// nl_stack.cu -- the stacking kernels and the job API (nl_stack_*).
//
// Replaces OpStack.Apply and the nine Stack* reducers of the reference
// (internal/ops/stack/stack.go:115-227, 274-918).  Data layout in HBM: one job buffer
// [n_frames][pixels] fp32, frame-major, so the column of pixel p is a stride-`pixels` gather and the
// 32 lanes of a warp read 128 contiguous bytes of each frame.
//
// Kernel families
//   stack_mean_kernel    StackMean / StackMeanWeighted: pure streaming, sequential sum in frame
//                        order per pixel, float4 (4 pixels per thread), no shared memory.
//   stack_column_kernel  every order-statistics mode: a warp owns a tile of S pixels (S = 32:
//                        lane == pixel == shared-memory bank), gathers the non-NaN samples of its
//                        pixels into shared memory [sample][lane], then every lane reduces its own
//                        column with the routines of nl_column.cuh, which reproduce the reference's
//                        quick-select permutation and fp32 evaluation order exactly.
// Persistent grid: warps pull tiles from a counter; warps of one SM are in different phases (staging /
// reducing), which overlaps the HBM stream with the latency-bound column work.
// The column kernel and its launcher live in nl_stack_kernel.cuh and are instantiated per mode family in
// nl_stack_sigma.cu / nl_stack_winsor.cu / nl_stack_linfit.cu.
#include "nl_stack_kernel.cuh"

#include <thread>

namespace nl {

// ---- StackMean / StackMeanWeighted (stack.go:307-366) ----------------------------------------
#ifndef NL_MEAN_UNROLL
#define NL_MEAN_UNROLL 8
#endif
#ifndef NL_MEAN_UNROLL_W
#define NL_MEAN_UNROLL_W 4
#endif
constexpr int MEAN_UNROLL = NL_MEAN_UNROLL;      // frames in flight per thread: unweighted 8 (0.998 of the copy bandwidth),
constexpr int MEAN_UNROLL_W = NL_MEAN_UNROLL_W;  // weighted 4 (more live registers per frame: 0.79 -> 0.84)
#ifndef NL_MEAN_MINB
#define NL_MEAN_MINB 1
#endif
template <bool W, int V>   // V pixels per thread (4: float4 path, 1: scalar path)
__global__ void __launch_bounds__(256, NL_MEAN_MINB) stack_mean_kernel(StackArgs a) {
    long long groups = (a.pixels + V - 1) / V;
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < groups;
         gidx += (long long)gridDim.x * blockDim.x) {
        long long p = gidx * V;
        float sum[V], wsum[V];
        int num[V];
#pragma unroll
        for (int c = 0; c < V; c++) { sum[c] = 0.0f; wsum[c] = 0.0f; num[c] = 0; }
        const float *src = a.frames + p;
#pragma unroll (W ? MEAN_UNROLL_W : MEAN_UNROLL)
        for (int k = 0; k < a.n; k++) {
            float v[V];
            if (V == 4) {
                float4 q = __ldcs(reinterpret_cast<const float4 *>(src + (long long)k * a.stride));
                v[0] = q.x; v[1 % V] = q.y; v[2 % V] = q.z; v[3 % V] = q.w;
            } else {
                v[0] = __ldcs(src + (long long)k * a.stride);
            }
            float w = W ? __ldg(a.weights + k) : 1.0f;
#pragma unroll
            for (int c = 0; c < V; c++) {
                if (v[c] == v[c]) {
                    if (W) { sum[c] = __fadd_rn(sum[c], __fmul_rn(v[c], w)); wsum[c] = __fadd_rn(wsum[c], w); }
                    else sum[c] = __fadd_rn(sum[c], v[c]);
                    num[c]++;
                }
            }
        }
        float r[V];
#pragma unroll
        for (int c = 0; c < V; c++)
            r[c] = num[c] == 0 ? a.ref_loc : __fdiv_rn(sum[c], W ? wsum[c] : (float)num[c]);
        if (!a.out) continue;
        if (V == 4) {
            const float4 q = make_float4(r[0], r[1 % V], r[2 % V], r[3 % V]);
            *reinterpret_cast<float4 *>(a.out + p) = q;
            for (int e = 0; e < a.n_peers; e++) *reinterpret_cast<float4 *>(a.peer_out[e] + p) = q;
        } else {
            store_result(a, p, r[0]);
        }
    }
}

__global__ void ramp_kernel(float *ramp, int n) {
    // ramp[2c], ramp[2c+1] = MeanStdDev(0..c-1), exactly as the reference computes it per pixel
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= n; c += gridDim.x * blockDim.x) {
        float m = 0.0f, s = 0.0f;
        if (c > 0) ramp_mean_stddev(c, m, s);
        ramp[2 * c] = m;
        ramp[2 * c + 1] = s;
    }
}

}  // namespace nl

using namespace nl;

namespace nl {

// The frame stack as a 2-D tensor [frame][pixel] for the TMA staging.  cuTensorMapEncodeTiled is a
// driver entry point; it is looked up through the runtime so that the library links cudart only.
static bool make_frame_tensor_map(nl_stack_job *job) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return false;
        }
        encode = (encode_fn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)job->pixels, (cuuint64_t)job->n};
    const cuuint64_t strides[1] = {(cuuint64_t)job->pixels * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)TMA_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return encode(&job->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, job->frames, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) == CUDA_SUCCESS;
}

template <bool W>
static int launch_mean(nl_stack_job *job, const StackArgs &args) {
    nl_ctx *ctx = job->ctx;
    // float4 path: frame rows (stride), the result and every peer copy 16-byte aligned; else one pixel per thread
    uintptr_t align = (uintptr_t)args.frames | (uintptr_t)args.out | (uintptr_t)(args.stride * 4);
    for (int e = 0; e < args.n_peers; e++) align |= (uintptr_t)args.peer_out[e];
    const bool vec = (args.pixels % 4) == 0 && (align & 15) == 0;
    const long long groups = vec ? args.pixels / 4 : args.pixels;
    long long grid = (groups + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (vec) stack_mean_kernel<W, 4><<<(unsigned)grid, 256, 0, ctx->stream>>>(args);
    else stack_mean_kernel<W, 1><<<(unsigned)grid, 256, 0, ctx->stream>>>(args);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

static int stack_launch(nl_stack_job *job, int mode, const float *host_weights, float sig_lo, float sig_hi,
                        float ref_loc, float *dev_out, float *const *peer_outs = nullptr, int n_peers = 0) {
    nl_ctx *ctx = job->ctx;
    if (mode < NL_ST_MEDIAN || mode > NL_ST_AUTO) return set_error(NL_E_INVALID, "invalid stacking mode");   // stack.go:118-120
    if (mode == NL_ST_AUTO) mode = auto_select_mode(job->n);
    const bool weighted = host_weights != nullptr;
    if (mode == NL_ST_MAD_SIGMA && weighted)
        return set_error(NL_E_UNSUPPORTED, "MADSigma stacking with weights is still unimplemented");         // stack.go:185
    NL_CUDA(cudaMemsetAsync(job->clip, 0, NL_JOB_COUNTERS * sizeof(unsigned long long), ctx->stream));
    if (job->active == 0) return NL_OK;
    if (weighted)
        NL_CUDA(cudaMemcpyAsync(job->weights, host_weights, sizeof(float) * job->n, cudaMemcpyHostToDevice, ctx->stream));
    if (mode == NL_ST_LINEAR_FIT && !job->ramp_ready) {
        ramp_kernel<<<(job->n + 128) / 128, 128, 0, ctx->stream>>>(job->ramp, job->n);
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        job->ramp_ready = true;
    }
    if (!job->tmap_ok && (job->pixels % 4) == 0 && job->pixels < (1ll << 31)) job->tmap_ok = make_frame_tensor_map(job);
    StackArgs a;
    a.use_tma = job->tmap_ok ? 1 : 0;
    a.frames = job->frames; a.stride = job->pixels; a.pixels = job->active; a.n = job->n;
    a.weights = weighted ? job->weights : nullptr; a.ramp = job->ramp;
    a.ref_loc = ref_loc; a.sig_lo = sig_lo; a.sig_hi = sig_hi; a.out = dev_out; a.clip = job->clip; a.tile_counter = job->clip + 2;
    a.defer_passes = 0; a.phase = 0; a.pool_in = StackArgs::Pool{nullptr, nullptr, nullptr, nullptr, job->clip + 3, 0};
    a.pool_out = a.pool_in; a.pool_tile_counter = job->clip + 4;
    a.stream_cache_blocks = 0;
    a.n_peers = n_peers;
    for (int e = 0; e < NL_MAX_PEERS; e++) a.peer_out[e] = e < n_peers ? peer_outs[e] : nullptr;
    switch (mode) {
    case NL_ST_MEDIAN: return launch_median(job, a);     // weights ignored, stack.go:160-161
    case NL_ST_MEAN: return weighted ? launch_mean<true>(job, a) : launch_mean<false>(job, a);
    case NL_ST_SIGMA: return launch_sigma(job, a, weighted);
    case NL_ST_WINSOR_SIGMA: return launch_winsor(job, a, weighted);
    case NL_ST_MAD_SIGMA: return launch_mad(job, a);
    case NL_ST_LINEAR_FIT: return launch_linfit(job, a);   // weights ignored, stack.go:188-189
    }
    return set_error(NL_E_INVALID, "invalid stacking mode");
}

}  // namespace nl

extern "C" {

int nl_auto_select_mode(int32_t n_frames) { return auto_select_mode(n_frames); }

// getWeights, stack.go:231-270
int nl_get_weights(int32_t weighting, const float *exposure, const float *noise, const float *hfr, int32_t n,
                   float *weights) {
    NL_REQUIRE(n >= 0, "n_frames < 0");
    if (weighting == NL_W_NONE) return NL_OK;
    NL_REQUIRE(weights, "weights is NULL");
    if (weighting == NL_W_EXPOSURE) {
        NL_REQUIRE(exposure, "exposure is NULL");
        for (int i = 0; i < n; i++) {
            if (exposure[i] == 0) return set_error(NL_E_WEIGHTS, "Missing exposure information for exposure-weighted stacking");
            weights[i] = exposure[i];
        }
        return NL_OK;
    }
    if (weighting == NL_W_INVERSE_NOISE || weighting == NL_W_INVERSE_HFR) {
        const float *x = weighting == NL_W_INVERSE_NOISE ? noise : hfr;
        NL_REQUIRE(x, "noise / hfr is NULL");
        float mn = 3.40282346638528859811704183484516925440e+38f, mx = -mn;
        for (int i = 0; i < n; i++) {
            if (x[i] < mn) mn = x[i];
            if (x[i] > mx) mx = x[i];
        }
        for (int i = 0; i < n; i++) weights[i] = 1.0f / (1.0f + 4.0f * (x[i] - mn) / (mx - mn));
        return NL_OK;
    }
    return set_error(NL_E_WEIGHTS, "invalid weighting mode %d", weighting);
}

int nl_stack_begin(nl_ctx *ctx, int32_t n_frames, int64_t pixels, nl_stack_job **out) {
    NL_REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    NL_REQUIRE(n_frames >= 1, "n_frames must be >= 1");
    NL_REQUIRE(pixels >= 0, "pixels must be >= 0");
    NL_GUARD(ctx);
    nl_stack_job *j = new nl_stack_job();
    j->ctx = ctx; j->n = n_frames; j->pixels = pixels; j->active = pixels;
    size_t frame_bytes = sizeof(float) * (size_t)n_frames * (size_t)(pixels > 0 ? pixels : 1);
    cudaError_t e = cudaMalloc(&j->frames, frame_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&j->out, sizeof(float) * (size_t)(pixels > 0 ? pixels : 1));
    if (e == cudaSuccess) e = cudaMalloc(&j->weights, sizeof(float) * (size_t)n_frames);
    if (e == cudaSuccess) e = cudaMalloc(&j->ramp, sizeof(float) * 2 * ((size_t)n_frames + 1));
    if (e == cudaSuccess) e = cudaMalloc(&j->clip, NL_JOB_COUNTERS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc(&j->clip_host, 2 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        nl_stack_end(j);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            return set_error(NL_E_NOMEM, "out of device memory for %d frames x %lld pixels", n_frames, (long long)pixels);
        }
        return cuda_fail(e, "nl_stack_begin allocation");
    }
    j->clip_host[0] = j->clip_host[1] = 0;
    *out = j;
    return NL_OK;
}

int nl_stack_put_frame(nl_stack_job *job, int32_t i, const float *host, int64_t count) {
    NL_REQUIRE(job && host, "NULL argument");
    NL_REQUIRE(i >= 0 && i < job->n, "frame index out of range");
    NL_REQUIRE(count == job->pixels, "frame size differs from the job's pixel count");
    NL_GUARD(job->ctx);
    NL_CUDA(cudaMemcpyAsync(job->frames + (size_t)i * job->pixels, host, sizeof(float) * (size_t)count,
                            cudaMemcpyHostToDevice, job->ctx->stream));
    return NL_OK;
}

// A frame handed over as its raw FITS payload (big-endian BITPIX samples): uploaded as is -- half the
// PCIe bytes for 16-bit camera frames -- and decoded into the job's frame slot on the device
// (read.go:176-443: v = float32(val)*Bscale + Bzero).
int nl_stack_put_frame_raw(nl_stack_job *job, int32_t i, const void *host_raw, int32_t bitpix, int64_t count, float bscale,
                           float bzero) {
    NL_REQUIRE(job && host_raw, "NULL argument");
    NL_REQUIRE(i >= 0 && i < job->n, "frame index out of range");
    NL_REQUIRE(count == job->pixels, "frame size differs from the job's pixel count");
    NL_REQUIRE(bitpix == 8 || bitpix == 16 || bitpix == 32 || bitpix == 64 || bitpix == -32 || bitpix == -64, "Unknown BITPIX value");
    if (count == 0) return NL_OK;
    nl_ctx *ctx = job->ctx;
    NL_GUARD(ctx);
    const size_t bytes = (size_t)count * (size_t)(bitpix < 0 ? -bitpix : bitpix) / 8;
    // two staging halves used alternately: the upload of frame i+1 may start while frame i is decoded
    const size_t half = (bytes + 255) & ~(size_t)255;
    int rc = ensure_scratch(ctx, 2 * half);
    if (rc != NL_OK) return rc;
    void *stage = (char *)ctx->scratch + (size_t)(i & 1) * half;
    NL_CUDA(cudaMemcpyAsync(stage, host_raw, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return fits_decode_launch(ctx, stage, bitpix, count, bscale, bzero, job->frames + (size_t)i * job->pixels);
}

int nl_stack_job_shape(nl_stack_job *job, int32_t *n_frames, int64_t *pixels) {
    NL_REQUIRE(job, "NULL argument");
    if (n_frames) *n_frames = job->n;
    if (pixels) *pixels = job->pixels;
    return NL_OK;
}

int nl_stack_frames_dev(nl_stack_job *job, float **dev_frames, int64_t *frame_stride) {
    NL_REQUIRE(job && dev_frames, "NULL argument");
    *dev_frames = job->frames;
    if (frame_stride) *frame_stride = job->pixels;
    return NL_OK;
}

int nl_stack_run_dev(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                     float ref_frame_loc, float *dev_out) {
    NL_REQUIRE(job, "NULL argument");              // dev_out == NULL: clip totals only, no image is written
    NL_GUARD(job->ctx);
    int rc = stack_launch(job, mode, weights, sigma_low, sigma_high, ref_frame_loc, dev_out);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(job->clip_host, job->clip, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            job->ctx->stream));
    return NL_OK;
}

int nl_stack_run_dev_bcast(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                           float ref_frame_loc, float *dev_out, float *const *peer_outs, int32_t n_peers) {
    NL_REQUIRE(job && (dev_out || job->pixels == 0), "NULL argument");
    NL_REQUIRE(n_peers >= 0 && n_peers <= NL_MAX_PEERS && (n_peers == 0 || peer_outs), "bad peer list");
    for (int e = 0; e < n_peers; e++) NL_REQUIRE(peer_outs[e], "NULL peer pointer");
    NL_GUARD(job->ctx);
    int rc = stack_launch(job, mode, weights, sigma_low, sigma_high, ref_frame_loc, dev_out, peer_outs, n_peers);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(job->clip_host, job->clip, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            job->ctx->stream));
    return NL_OK;
}

int nl_stack_clip_counts(nl_stack_job *job, int64_t *clip_low, int64_t *clip_high) {
    NL_REQUIRE(job, "NULL argument");
    if (clip_low) *clip_low = (int64_t)job->clip_host[0];
    if (clip_high) *clip_high = (int64_t)job->clip_host[1];
    return NL_OK;
}

// One stacking pass that only counts what it clips (the trial stacks of the sigma goal-seek): the column kernel
// runs as usual, the result stores are skipped.
int nl_stack_clip_counts_only(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                              int64_t *clip_low, int64_t *clip_high) {
    NL_REQUIRE(job, "NULL argument");
    int rc = nl_stack_run_dev(job, mode, weights, sigma_low, sigma_high, 0.0f, nullptr);
    if (rc != NL_OK) return rc;
    rc = nl_ctx_sync(job->ctx);
    if (rc != NL_OK) return rc;
    return nl_stack_clip_counts(job, clip_low, clip_high);
}

int nl_stack_run(nl_stack_job *job, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                 float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high) {
    NL_REQUIRE(job && (host_out || job->pixels == 0), "NULL argument");
    NL_GUARD(job->ctx);
    int rc = nl_stack_run_dev(job, mode, weights, sigma_low, sigma_high, ref_frame_loc, job->out);
    if (rc != NL_OK) return rc;
    if (job->pixels > 0)
        NL_CUDA(cudaMemcpyAsync(host_out, job->out, sizeof(float) * (size_t)job->pixels, cudaMemcpyDeviceToHost,
                                job->ctx->stream));
    NL_CUDA(cudaStreamSynchronize(job->ctx->stream));
    return nl_stack_clip_counts(job, clip_low, clip_high);
}

}  // extern "C"

namespace nl {

// ctx->lane_ctx[l]: a second / third stream with its own scratch on the context's device, created on first use and kept
// (nl_stack_apply's stripe lanes, the batched bad-pixel maps).  A lane inherits the settings of its context.
int lane_context(nl_ctx *ctx, int l, nl_ctx **out) {
    if (!ctx->lane_ctx[l]) {
        int rc = nl_ctx_create(ctx->device, &ctx->lane_ctx[l]);
        if (rc != NL_OK) return rc;
    }
    nl_ctx *lc = ctx->lane_ctx[l];
    lc->defer_override = ctx->defer_override; lc->defer_n = ctx->defer_n;
    for (int i = 0; i < 8; i++) lc->defer_at[i] = ctx->defer_at[i];
    lc->tile_width = ctx->tile_width;
    lc->linfit_stream = ctx->linfit_stream; lc->linfit_stream_ctas = ctx->linfit_stream_ctas; lc->linfit_stream_cache = ctx->linfit_stream_cache;
    lc->numerics = ctx->numerics;
    lc->stats_debug = ctx->stats_debug; lc->stats_force_replay = ctx->stats_force_replay;
    *out = lc;
    return NL_OK;
}

// The pixel range [p_begin, p_end) of an image stacked from host frames through one context: the range is cut into
// n_stripes row stripes that alternate on the context's two lanes (a lane = stream + job + result buffer), so the
// upload of stripe s+1 overlaps the stacking of stripe s and the download of stripe s-1 -- the device never holds more
// than two stripes.  The lanes are sized once for the largest stripe and kept in the context between calls; a ragged
// last stripe runs in the same lane on fewer active pixels.
static int stack_apply_range(nl_ctx *ctx, const float *const *host_frames, int32_t n_frames, int64_t p_begin, int64_t p_end,
                             int64_t row_pixels, int32_t n_stripes, int32_t mode, const float *weights, float sigma_low,
                             float sigma_high, float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high) {
    const int64_t pixels = p_end - p_begin;
    if (clip_low) *clip_low = 0;
    if (clip_high) *clip_high = 0;
    if (pixels <= 0) return NL_OK;
    if (row_pixels <= 0 || pixels % row_pixels != 0) row_pixels = pixels;       // no row structure known: one stripe
    const int64_t rows = pixels / row_pixels;
    if (n_stripes < 1) n_stripes = 8;
    if (n_stripes > rows) n_stripes = (int32_t)rows;
    const int64_t stripe_rows = (rows + n_stripes - 1) / n_stripes;             // all stripes but the last
    n_stripes = (int32_t)((rows + stripe_rows - 1) / stripe_rows);
    const int64_t lane_px = stripe_rows * row_pixels;
    const int lanes = n_stripes > 1 ? 2 : 1;
    int64_t tot_lo = 0, tot_hi = 0;
    int rc = NL_OK;
    auto collect = [&](int l) -> int {                      // wait for the lane's stripe and add its clip counts
        int r = nl_ctx_sync(ctx->lane_ctx[l]);
        if (r != NL_OK) return r;
        int64_t a = 0, b = 0;
        r = nl_stack_clip_counts(ctx->lane_job[l], &a, &b);
        tot_lo += a; tot_hi += b;
        return r;
    };
    for (int l = 0; l < lanes && rc == NL_OK; l++) {
        nl_ctx *lc = nullptr;
        rc = lane_context(ctx, l, &lc);
        if (rc == NL_OK && (lane_px > ctx->lane_px[l] || n_frames != ctx->lane_frames[l] || !ctx->lane_job[l])) {   // (re)size the lane
            if (ctx->lane_job[l]) { nl_stack_end(ctx->lane_job[l]); ctx->lane_job[l] = nullptr; }
            if (ctx->lane_out[l]) { nl_dev_free(ctx->lane_ctx[l], ctx->lane_out[l]); ctx->lane_out[l] = nullptr; }
            ctx->lane_px[l] = 0;
            rc = nl_stack_begin(ctx->lane_ctx[l], n_frames, lane_px, &ctx->lane_job[l]);
            if (rc == NL_OK) rc = nl_dev_alloc(ctx->lane_ctx[l], 4 * lane_px, (void **)&ctx->lane_out[l]);
            if (rc == NL_OK) { ctx->lane_px[l] = lane_px; ctx->lane_frames[l] = n_frames; }
        }
    }
    for (int32_t si = 0; si < n_stripes && rc == NL_OK; si++) {
        const int l = si % lanes;
        nl_stack_job *job = ctx->lane_job[l];
        const int64_t p0 = p_begin + si * lane_px;
        const int64_t px = (si + 1 < n_stripes ? lane_px : p_end - p0);
        if (si >= lanes) rc = collect(l);
        if (rc != NL_OK) break;
        NL_GUARD(job->ctx);
        job->active = px;                                    // frames stay job->pixels apart; the run covers px of them
        for (int32_t k = 0; k < n_frames && rc == NL_OK; k++) {
            if (!host_frames[k]) { rc = set_error(NL_E_INVALID, "frame %d is NULL", k); break; }
            cudaError_t e = cudaMemcpyAsync(job->frames + (size_t)k * job->pixels, host_frames[k] + p0, sizeof(float) * (size_t)px,
                                            cudaMemcpyHostToDevice, job->ctx->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "frame upload");
        }
        if (rc == NL_OK)
            rc = nl_stack_run_dev(job, mode, weights, sigma_low, sigma_high, ref_frame_loc, ctx->lane_out[l]);
        if (rc == NL_OK) rc = nl_memcpy_d2h(ctx->lane_ctx[l], host_out + p0, ctx->lane_out[l], 4 * px);
    }
    for (int l = 0; l < lanes && l < n_stripes && rc == NL_OK; l++) rc = collect(l);
    if (rc != NL_OK) {                                       // leave no copy in flight from or to the caller's buffers
        std::string err = nl_last_error();
        for (int l = 0; l < 2; l++)
            if (ctx->lane_ctx[l]) nl_ctx_sync(ctx->lane_ctx[l]);
        return set_error(rc, "%s", err.c_str());
    }
    if (clip_low) *clip_low = tot_lo;
    if (clip_high) *clip_high = tot_hi;
    return NL_OK;
}

}  // namespace nl

extern "C" {

// OpStack.Apply in one call for callers that can hand over all frames at once (C / C++ hosts; a Go host
// pins the slices and passes a C array of their addresses).
int nl_stack_apply(nl_ctx *ctx, const float *const *host_frames, int32_t n_frames, int64_t pixels, int64_t row_pixels,
                   int32_t n_stripes, int32_t mode, const float *weights, float sigma_low, float sigma_high,
                   float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high) {
    NL_REQUIRE(ctx && host_frames && n_frames >= 1 && pixels >= 0, "bad argument");
    NL_REQUIRE(host_out || pixels == 0, "NULL output");
    if (mode < NL_ST_MEDIAN || mode > NL_ST_AUTO) return set_error(NL_E_INVALID, "invalid stacking mode");
    return stack_apply_range(ctx, host_frames, n_frames, 0, pixels, row_pixels, n_stripes, mode, weights, sigma_low, sigma_high,
                             ref_frame_loc, host_out, clip_low, clip_high);
}

// OpStack.Apply over several devices in one call: the image's rows are dealt to the contexts in contiguous blocks
// (device g stacks rows [g*H/G, (g+1)*H/G), SURVEY.md 8e -- the analogue of the reference's fan-out over pixel ranges,
// stack.go:134-147), every device pipelines its block like nl_stack_apply on a host thread of its own, and all of them
// write their rows of the one host image.  No exchange between devices: a pixel's column never leaves its GPU.
int nl_stack_apply_multi(nl_ctx *const *ctxs, int32_t n_ctx, const float *const *host_frames, int32_t n_frames, int64_t pixels,
                         int64_t row_pixels, int32_t n_stripes, int32_t mode, const float *weights, float sigma_low,
                         float sigma_high, float ref_frame_loc, float *host_out, int64_t *clip_low, int64_t *clip_high) {
    NL_REQUIRE(ctxs && n_ctx >= 1 && n_ctx <= 64 && host_frames && n_frames >= 1 && pixels >= 0, "bad argument");
    NL_REQUIRE(host_out || pixels == 0, "NULL output");
    for (int g = 0; g < n_ctx; g++) NL_REQUIRE(ctxs[g], "NULL context");
    if (mode < NL_ST_MEDIAN || mode > NL_ST_AUTO) return set_error(NL_E_INVALID, "invalid stacking mode");
    if (clip_low) *clip_low = 0;
    if (clip_high) *clip_high = 0;
    if (pixels == 0) return NL_OK;
    if (row_pixels <= 0 || pixels % row_pixels != 0) row_pixels = pixels;
    const int64_t rows = pixels / row_pixels;
    std::vector<int> rc(n_ctx, NL_OK);
    std::vector<std::string> err(n_ctx);
    std::vector<int64_t> lo(n_ctx, 0), hi(n_ctx, 0);
    std::vector<std::thread> th;
    for (int g = 0; g < n_ctx; g++) {
        const int64_t r0 = rows * g / n_ctx, r1 = rows * (g + 1) / n_ctx;
        if (r1 == r0) continue;
        th.emplace_back([&, g, r0, r1]() {
            rc[g] = stack_apply_range(ctxs[g], host_frames, n_frames, r0 * row_pixels, r1 * row_pixels, row_pixels, n_stripes, mode,
                                      weights, sigma_low, sigma_high, ref_frame_loc, host_out, &lo[g], &hi[g]);
            if (rc[g] != NL_OK) err[g] = nl_last_error();                    // (thread-local: carry it to the caller's thread)
        });
    }
    for (auto &t : th) t.join();
    int64_t tl = 0, thi = 0;
    for (int g = 0; g < n_ctx; g++) {
        if (rc[g] != NL_OK) return set_error(rc[g], "device %d: %s", ctxs[g]->device, err[g].c_str());
        tl += lo[g]; thi += hi[g];
    }
    if (clip_low) *clip_low = tl;
    if (clip_high) *clip_high = thi;
    return NL_OK;
}

// releases the stripe lanes nl_stack_apply keeps between calls (also done by nl_ctx_destroy)
int nl_stack_apply_release(nl_ctx *ctx) {
    NL_REQUIRE(ctx, "ctx is NULL");
    for (int l = 0; l < 2; l++) {
        if (ctx->lane_job[l]) { nl_stack_end(ctx->lane_job[l]); ctx->lane_job[l] = nullptr; }
        if (ctx->lane_out[l]) { nl_dev_free(ctx->lane_ctx[l], ctx->lane_out[l]); ctx->lane_out[l] = nullptr; }
        if (ctx->lane_ctx[l]) { nl_ctx_destroy(ctx->lane_ctx[l]); ctx->lane_ctx[l] = nullptr; }
        ctx->lane_px[l] = 0; ctx->lane_frames[l] = 0;
    }
    for (int l = 2; l < NL_MAX_LANES; l++)
        if (ctx->lane_ctx[l]) { nl_ctx_destroy(ctx->lane_ctx[l]); ctx->lane_ctx[l] = nullptr; }
    return NL_OK;
}

int nl_stack_end(nl_stack_job *job) {
    if (!job) return NL_OK;
    NL_GUARD(job->ctx);
    cudaStreamSynchronize(job->ctx->stream);
    if (job->frames) cudaFree(job->frames);
    if (job->out) cudaFree(job->out);
    if (job->weights) cudaFree(job->weights);
    if (job->ramp) cudaFree(job->ramp);
    if (job->clip) cudaFree(job->clip);
    if (job->clip_host) cudaFreeHost(job->clip_host);
    if (job->pool) cudaFree(job->pool);
    delete job;
    return NL_OK;
}

// StackIncremental / StackIncrementalFinalize, stack.go:924-944
// four pixels per thread and iteration (float4) when both buffers are 16-byte aligned, two iterations in flight
__global__ void __launch_bounds__(256) incremental_vec_kernel(float *__restrict__ acc, const float *__restrict__ light, long long n,
                                                              float w, int first) {
    const long long n4 = n >> 2, step = (long long)gridDim.x * blockDim.x;
    float4 *a4 = reinterpret_cast<float4 *>(acc);
    const float4 *l4 = reinterpret_cast<const float4 *>(light);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += 2 * step) {
        const bool two = i + step < n4;
        const float4 x0 = __ldcs(l4 + i), x1 = two ? __ldcs(l4 + i + step) : make_float4(0, 0, 0, 0);
        float4 y0 = make_float4(0, 0, 0, 0), y1 = y0;
        if (!first) { y0 = a4[i]; if (two) y1 = a4[i + step]; }
        float4 r0, r1;
        r0.x = __fmul_rn(x0.x, w); r0.y = __fmul_rn(x0.y, w); r0.z = __fmul_rn(x0.z, w); r0.w = __fmul_rn(x0.w, w);
        r1.x = __fmul_rn(x1.x, w); r1.y = __fmul_rn(x1.y, w); r1.z = __fmul_rn(x1.z, w); r1.w = __fmul_rn(x1.w, w);
        if (!first) {
            r0.x = __fadd_rn(y0.x, r0.x); r0.y = __fadd_rn(y0.y, r0.y); r0.z = __fadd_rn(y0.z, r0.z); r0.w = __fadd_rn(y0.w, r0.w);
            r1.x = __fadd_rn(y1.x, r1.x); r1.y = __fadd_rn(y1.y, r1.y); r1.z = __fadd_rn(y1.z, r1.z); r1.w = __fadd_rn(y1.w, r1.w);
        }
        a4[i] = r0;
        if (two) a4[i + step] = r1;
    }
    // the up-to-three pixels behind the last whole float4
    const long long t = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) { const float v = __fmul_rn(light[t], w); acc[t] = first ? v : __fadd_rn(acc[t], v); }
}
__global__ void incremental_kernel(float *acc, const float *light, long long n, float w, int first) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float t = __fmul_rn(light[i], w);
        acc[i] = first ? t : __fadd_rn(acc[i], t);
    }
}
__global__ void __launch_bounds__(256) scale_vec_kernel(float *__restrict__ acc, long long n, float factor) {
    const long long n4 = n >> 2;
    float4 *a4 = reinterpret_cast<float4 *>(acc);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = a4[i];
        v.x = __fmul_rn(v.x, factor); v.y = __fmul_rn(v.y, factor); v.z = __fmul_rn(v.z, factor); v.w = __fmul_rn(v.w, factor);
        a4[i] = v;
    }
    const long long t = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) acc[t] = __fmul_rn(acc[t], factor);
}
__global__ void scale_kernel(float *acc, long long n, float factor) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc[i] = __fmul_rn(acc[i], factor);
}

int nl_stack_incremental_dev(nl_ctx *ctx, float *dev_acc, const float *dev_light, int64_t pixels, float weight, int first) {
    NL_REQUIRE(ctx && pixels >= 0, "bad argument");
    if (pixels == 0) return NL_OK;
    NL_REQUIRE(dev_acc && dev_light, "NULL argument");
    NL_GUARD(ctx);
    const bool vec = (((uintptr_t)dev_acc | (uintptr_t)dev_light) & 15) == 0;
    long long grid = ((vec ? pixels / 8 : pixels) + 255) / 256;
    if (grid > (long long)ctx->sm_count * 8) grid = (long long)ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    if (vec) incremental_vec_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(dev_acc, dev_light, pixels, weight, first);
    else incremental_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(dev_acc, dev_light, pixels, weight, first);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

int nl_stack_incremental_finalize_dev(nl_ctx *ctx, float *dev_acc, int64_t pixels, float weight_sum) {
    NL_REQUIRE(ctx && pixels >= 0, "bad argument");
    if (pixels == 0) return NL_OK;
    NL_REQUIRE(dev_acc, "NULL argument");
    NL_GUARD(ctx);
    float factor = 1.0f / weight_sum;    // stack.go:941
    const bool vec = ((uintptr_t)dev_acc & 15) == 0;
    long long grid = ((vec ? pixels / 4 : pixels) + 255) / 256;
    if (grid > (long long)ctx->sm_count * 8) grid = (long long)ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    if (vec) scale_vec_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(dev_acc, pixels, factor);
    else scale_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(dev_acc, pixels, factor);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

}  // extern "C"
