// nl_project.cu -- bilinear resample through an affine transform (the alignment gather).
// Replaces (*Image).Project (internal/fits/project.go:26-76) and Transform2D.Invert / Apply
// (internal/star/coord.go:141-201).  One thread per destination pixel; rows of the destination map
// to nearly-rows of the source, so warp reads are close to coalesced and the 2x2 footprints of
// neighbouring threads hit the same L1/L2 lines.  Algorithmic traffic: 4 B read + 4 B written per
// destination pixel.  fp32 in the reference's evaluation order, no FMA contraction.
#include "nl_internal.h"

#include <vector>

namespace nl {

struct Affine { float a, b, c, d, e, f; };

// One destination pixel, exactly the reference's expression shapes.
// SCALE: every gathered source sample first becomes d*mult + offset (mul, then add), i.e. the result is
// that of Image.MatchHistogram (internal/fits/pixelops.go:601-612) followed by Project, in one pass.
template <bool SCALE>
__device__ __forceinline__ float project_pixel(const float *__restrict__ src, int sw, int sh, const Affine &inv, float ax, float dx,
                                               int row, float oob, float mult, float offset) {
    // ax = A*x and dx = D*x are the same for every row of a thread's column (the products round once either way)
    const float y = (float)row;
    // coord.go:141-145: (A*x + B*y) + C
    const float px = __fadd_rn(__fadd_rn(ax, __fmul_rn(inv.b, y)), inv.c);
    const float py = __fadd_rn(__fadd_rn(dx, __fmul_rn(inv.e, y)), inv.f);
    const float fx = floorf(px), fy = floorf(py);
    float v = oob;
    // project.go:49-61; the float comparison form also rejects NaN and out-of-int32 coordinates
    if (fx >= 0.0f && fy >= 0.0f && fx < (float)(sw - 1) && fy < (float)(sh - 1)) {
        const int xl = (int)fx, yl = (int)fy;
        const float xr = __fsub_rn(px, fx), yr = __fsub_rn(py, fy);
        const float *s = src + ((unsigned)yl * (unsigned)sw + (unsigned)xl);     // (pixel counts are int32 in the reference)
        float d00 = __ldg(s), d10 = __ldg(s + 1), d01 = __ldg(s + sw), d11 = __ldg(s + sw + 1);
        if (SCALE) {
            d00 = __fadd_rn(__fmul_rn(d00, mult), offset); d10 = __fadd_rn(__fmul_rn(d10, mult), offset);
            d01 = __fadd_rn(__fmul_rn(d01, mult), offset); d11 = __fadd_rn(__fmul_rn(d11, mult), offset);
        }
        const float ox = __fsub_rn(1.0f, xr), oy = __fsub_rn(1.0f, yr);
        const float vyl = __fadd_rn(__fmul_rn(d00, ox), __fmul_rn(d10, xr));   // project.go:68-70
        const float vyh = __fadd_rn(__fmul_rn(d01, ox), __fmul_rn(d11, xr));
        v = __fadd_rn(__fmul_rn(vyl, oy), __fmul_rn(vyh, yr));
    }
    return v;
}

// One destination column x four rows per thread: the 32 lanes of a warp cover 32 consecutive
// destination pixels of a row, so every one of the 16 gathers of a thread is a (nearly) contiguous
// 128-byte warp access and every store a full 128-byte line; the four rows give each thread 16
// independent loads in flight and their 2x2 footprints share L1 lines with the rows above and below.
#ifndef NL_PROJ_ROWS
#define NL_PROJ_ROWS 4
#endif
#ifndef NL_PROJ_BX
#define NL_PROJ_BX 64
#endif
constexpr int PR = NL_PROJ_ROWS;          // destination rows per thread
constexpr int PBX = NL_PROJ_BX, PBY = 256 / NL_PROJ_BX;

template <bool SCALE>
__global__ void __launch_bounds__(256) project_kernel(const float *__restrict__ src, int sw, int sh, float *__restrict__ dst,
                                                      int dw, int dh, Affine inv, float oob, float mult, float offset) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row0 = (blockIdx.y * blockDim.y + threadIdx.y) * PR;
    if (col >= dw || row0 >= dh) return;
    const float x = (float)col, ax = __fmul_rn(inv.a, x), dx = __fmul_rn(inv.d, x);
    float v[PR];
#pragma unroll
    for (int r = 0; r < PR; r++) v[r] = (row0 + r < dh) ? project_pixel<SCALE>(src, sw, sh, inv, ax, dx, row0 + r, oob, mult, offset) : 0.0f;
#pragma unroll
    for (int r = 0; r < PR; r++)
        if (row0 + r < dh) __stcs(dst + (size_t)(row0 + r) * dw + col, v[r]);
}

// All frames of a resident stack in ONE launch (grid.z = frame): frame i is read at src + i*src_stride and written at
// dst + i*dst_stride (e.g. straight into slot i of a stack job) with its own inverse transform and histogram match.
struct BatchParam { Affine inv; float mult, offset; int scale; int pad; };

__global__ void __launch_bounds__(256) project_batch_kernel(const float *__restrict__ src, long long src_stride, int sw, int sh,
                                                            float *__restrict__ dst, long long dst_stride, int dw, int dh,
                                                            const BatchParam *__restrict__ params, float oob) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row0 = (blockIdx.y * blockDim.y + threadIdx.y) * PR;
    if (col >= dw || row0 >= dh) return;
    const BatchParam bp = params[blockIdx.z];
    const float *s = src + (long long)blockIdx.z * src_stride;
    float *d = dst + (long long)blockIdx.z * dst_stride;
    const float x = (float)col, ax = __fmul_rn(bp.inv.a, x), dx = __fmul_rn(bp.inv.d, x);
    float v[PR];
    if (bp.scale) {
#pragma unroll
        for (int r = 0; r < PR; r++) v[r] = (row0 + r < dh) ? project_pixel<true>(s, sw, sh, bp.inv, ax, dx, row0 + r, oob, bp.mult, bp.offset) : 0.0f;
    } else {
#pragma unroll
        for (int r = 0; r < PR; r++) v[r] = (row0 + r < dh) ? project_pixel<false>(s, sw, sh, bp.inv, ax, dx, row0 + r, oob, 1.0f, 0.0f) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < PR; r++)
        if (row0 + r < dh) __stcs(d + (size_t)(row0 + r) * dw + col, v[r]);
}

// Frame-sharded resample feeding row-sharded stacking: the destination rows of one frame are not
// stored into one image but straight into the stack jobs that own their row stripes -- the local job or, through
// peer mappings over NVLink, the jobs of the other GPUs.  The frame-major -> stripe-major exchange between the two
// shardings is thereby fused into the resample's stores: no staging image, no separate all-to-all pass.
struct Scatter {
    float *frame[NL_MAX_PEERS];       // where this frame's stripe g starts: job g's frame buffer + frame_index * stripe pixels
    int row0[NL_MAX_PEERS + 1];       // stripe g owns destination rows [row0[g], row0[g+1])
    int n;
};

template <bool SCALE>
__global__ void __launch_bounds__(256) project_scatter_kernel(const float *__restrict__ src, int sw, int sh, Scatter sc, int dw, int dh,
                                                              Affine inv, float oob, float mult, float offset) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row0 = (blockIdx.y * blockDim.y + threadIdx.y) * 4;
    if (col >= dw || row0 >= dh) return;
    const float x = (float)col, ax = __fmul_rn(inv.a, x), dx = __fmul_rn(inv.d, x);
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; r++) v[r] = (row0 + r < dh) ? project_pixel<SCALE>(src, sw, sh, inv, ax, dx, row0 + r, oob, mult, offset) : 0.0f;
    int g = 0;
    while (g + 1 < sc.n && row0 >= sc.row0[g + 1]) g++;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int row = row0 + r;
        if (row >= dh) break;
        while (g + 1 < sc.n && row >= sc.row0[g + 1]) g++;
        sc.frame[g][(size_t)(row - sc.row0[g]) * dw + col] = v[r];
    }
}

}  // namespace nl

using namespace nl;

extern "C" {

// Transform2D.Invert, coord.go:159-201 (fp32, same expression shapes)
int nl_transform_invert(const float t[6], float inv[6]) {
    NL_REQUIRE(t && inv, "NULL argument");
    const float A = t[0], B = t[1], C = t[2], D = t[3], E = t[4], F = t[5];
    volatile float eps = B * D - A * E;
    if (eps < 1e-8f && -eps < 1e-8f) return set_error(NL_E_SINGULAR, "Matrix has no inverse, epsilon=%g", (double)eps);
    volatile float bd = B * D, ae = A * E;
    volatile float det1 = bd - ae, det2 = ae - bd;
    volatile float ce = C * E, bf = B * F, cd = C * D, af = A * F;
    volatile float n1 = ce - bf, n2 = cd - af;
    inv[0] = -E / det1;
    inv[1] = B / det1;
    inv[2] = n1 / det1;
    inv[3] = -D / det2;
    inv[4] = A / det2;
    inv[5] = n2 / det2;
    return NL_OK;
}

static int project_launch(nl_ctx *ctx, const float *dev_src, int32_t sw, int32_t sh, float *dev_dst, int32_t dw, int32_t dh,
                          const float trans[6], float oob, bool scale, float mult, float offset) {
    NL_REQUIRE(ctx && trans, "NULL argument");
    NL_REQUIRE(sw >= 0 && sh >= 0 && dw >= 0 && dh >= 0, "negative image size");
    NL_REQUIRE((long long)sw * sh <= 0x7fffffffll && (long long)dw * dh <= 0x7fffffffll, "image larger than int32 pixels (fits.go:40)");
    float inv[6];
    int rc = nl_transform_invert(trans, inv);
    if (rc != NL_OK) return rc;
    if (dw == 0 || dh == 0) return NL_OK;
    NL_REQUIRE(dev_dst && (dev_src || sw == 0 || sh == 0), "NULL image pointer");
    NL_GUARD(ctx);
    Affine a{inv[0], inv[1], inv[2], inv[3], inv[4], inv[5]};
    dim3 block(PBX, PBY);
    dim3 grid((dw + block.x - 1) / block.x, (dh + PR * block.y - 1) / (PR * block.y));
    if (scale) project_kernel<true><<<grid, block, 0, ctx->stream>>>(dev_src, sw, sh, dev_dst, dw, dh, a, oob, mult, offset);
    else project_kernel<false><<<grid, block, 0, ctx->stream>>>(dev_src, sw, sh, dev_dst, dw, dh, a, oob, 1.0f, 0.0f);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

int nl_project_dev(nl_ctx *ctx, const float *dev_src, int32_t sw, int32_t sh, float *dev_dst, int32_t dw, int32_t dh,
                   const float trans[6], float oob) {
    return project_launch(ctx, dev_src, sw, sh, dev_dst, dw, dh, trans, oob, false, 1.0f, 0.0f);
}

int nl_project_scaled_dev(nl_ctx *ctx, const float *dev_src, int32_t sw, int32_t sh, float *dev_dst, int32_t dw, int32_t dh,
                          const float trans[6], float oob, float multiplier, float offset) {
    return project_launch(ctx, dev_src, sw, sh, dev_dst, dw, dh, trans, oob, true, multiplier, offset);
}

// OpAlign over all frames of a resident stack (postprocess.go:142-191 runs Project once per frame from a pool of
// goroutines): n_frames resamples in one launch.  trans = n_frames x 6 floats; multipliers / offsets = n_frames floats
// each or NULL (no histogram match); a multiplier of 1 with an offset of 0 also means "no match" for that frame.
int nl_project_batch_dev(nl_ctx *ctx, const float *dev_src, int64_t src_stride, int32_t sw, int32_t sh, float *dev_dst,
                         int64_t dst_stride, int32_t dw, int32_t dh, int32_t n_frames, const float *trans, float oob,
                         const float *multipliers, const float *offsets) {
    NL_REQUIRE(ctx && trans && n_frames >= 0, "bad argument");
    NL_REQUIRE(sw >= 0 && sh >= 0 && dw >= 0 && dh >= 0, "negative image size");
    NL_REQUIRE((long long)sw * sh <= 0x7fffffffll && (long long)dw * dh <= 0x7fffffffll, "image larger than int32 pixels (fits.go:40)");
    NL_REQUIRE(n_frames <= 65535, "more than 65535 frames in one batch");
    NL_REQUIRE((multipliers == nullptr) == (offsets == nullptr), "multipliers and offsets come together");
    std::vector<BatchParam> params((size_t)n_frames);
    for (int i = 0; i < n_frames; i++) {
        float inv[6];
        int rc = nl_transform_invert(trans + 6 * i, inv);
        if (rc != NL_OK) return rc;
        params[i].inv = Affine{inv[0], inv[1], inv[2], inv[3], inv[4], inv[5]};
        params[i].mult = multipliers ? multipliers[i] : 1.0f;
        params[i].offset = offsets ? offsets[i] : 0.0f;
        params[i].scale = (multipliers && !(multipliers[i] == 1.0f && offsets[i] == 0.0f)) ? 1 : 0;   // d*1 + 0 would turn -0 into +0
        params[i].pad = 0;
    }
    if (n_frames == 0 || dw == 0 || dh == 0) return NL_OK;
    NL_REQUIRE(dev_dst && (dev_src || sw == 0 || sh == 0), "NULL image pointer");
    NL_GUARD(ctx);
    int rc = ensure_scratch(ctx, sizeof(BatchParam) * (size_t)n_frames);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(ctx->scratch, params.data(), sizeof(BatchParam) * (size_t)n_frames, cudaMemcpyHostToDevice, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));                      // `params` is pageable host memory about to go away
    dim3 block(PBX, PBY);
    dim3 grid((dw + block.x - 1) / block.x, (dh + PR * block.y - 1) / (PR * block.y), n_frames);
    project_batch_kernel<<<grid, block, 0, ctx->stream>>>(dev_src, src_stride, sw, sh, dev_dst, dst_stride, dw, dh,
                                                         (const BatchParam *)ctx->scratch, oob);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

int nl_project_scatter_dev(nl_ctx *ctx, const float *dev_src, int32_t sw, int32_t sh, int32_t dw, int32_t dh, const float trans[6],
                           float oob, float multiplier, float offset, int32_t frame_index, void *const *stripe_frames,
                           const int32_t *stripe_row0, int32_t n_stripes) {
    NL_REQUIRE(ctx && trans && stripe_frames && stripe_row0, "NULL argument");
    NL_REQUIRE(sw >= 0 && sh >= 0 && dw >= 0 && dh >= 0 && frame_index >= 0, "negative size or index");
    NL_REQUIRE((long long)sw * sh <= 0x7fffffffll && (long long)dw * dh <= 0x7fffffffll, "image larger than int32 pixels (fits.go:40)");
    NL_REQUIRE(n_stripes >= 1 && n_stripes <= NL_MAX_PEERS, "stripe count out of range");
    NL_REQUIRE(stripe_row0[0] == 0 && stripe_row0[n_stripes] == dh, "stripes must cover rows [0, dh)");
    float inv[6];
    int rc = nl_transform_invert(trans, inv);
    if (rc != NL_OK) return rc;
    Scatter sc;
    sc.n = n_stripes;
    for (int g = 0; g < n_stripes; g++) {
        const int rows = stripe_row0[g + 1] - stripe_row0[g];
        NL_REQUIRE(rows >= 0, "stripe rows must ascend");
        NL_REQUIRE(stripe_frames[g] || rows == 0, "NULL stripe buffer");
        sc.row0[g] = stripe_row0[g];
        sc.frame[g] = (float *)stripe_frames[g] + (size_t)frame_index * (size_t)rows * (size_t)dw;
    }
    sc.row0[n_stripes] = dh;
    if (dw == 0 || dh == 0) return NL_OK;
    NL_REQUIRE(dev_src || sw == 0 || sh == 0, "NULL image pointer");
    NL_GUARD(ctx);
    Affine a{inv[0], inv[1], inv[2], inv[3], inv[4], inv[5]};
    dim3 block(64, 4);
    dim3 grid((dw + block.x - 1) / block.x, (dh + 4 * block.y - 1) / (4 * block.y));
    // multiplier 1 and offset 0 mean "no histogram match": d*1 + 0 would turn -0 into +0
    if (multiplier == 1.0f && offset == 0.0f)
        project_scatter_kernel<false><<<grid, block, 0, ctx->stream>>>(dev_src, sw, sh, sc, dw, dh, a, oob, 1.0f, 0.0f);
    else
        project_scatter_kernel<true><<<grid, block, 0, ctx->stream>>>(dev_src, sw, sh, sc, dw, dh, a, oob, multiplier, offset);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

static int project_host(nl_ctx *ctx, const float *host_src, int32_t sw, int32_t sh, float *host_dst, int32_t dw, int32_t dh,
                        const float trans[6], float oob, bool scale, float mult, float offset) {
    NL_REQUIRE(ctx && trans, "NULL argument");
    NL_REQUIRE(sw >= 0 && sh >= 0 && dw >= 0 && dh >= 0, "negative image size");
    NL_REQUIRE((long long)sw * sh <= 0x7fffffffll && (long long)dw * dh <= 0x7fffffffll, "image larger than int32 pixels (fits.go:40)");
    float inv[6];
    int rc = nl_transform_invert(trans, inv);
    if (rc != NL_OK) return rc;
    if (dw == 0 || dh == 0) return NL_OK;
    NL_REQUIRE(host_dst && (host_src || sw == 0 || sh == 0), "NULL image pointer");
    NL_GUARD(ctx);
    const size_t sbytes = sizeof(float) * (size_t)sw * sh, dbytes = sizeof(float) * (size_t)dw * dh;
    const size_t soff = (sbytes + 255) & ~(size_t)255;
    rc = ensure_scratch(ctx, soff + dbytes + 256);
    if (rc != NL_OK) return rc;
    float *ds = (float *)ctx->scratch, *dd = (float *)((char *)ctx->scratch + soff);
    if (sbytes) NL_CUDA(cudaMemcpyAsync(ds, host_src, sbytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = project_launch(ctx, ds, sw, sh, dd, dw, dh, trans, oob, scale, mult, offset);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(host_dst, dd, dbytes, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_project(nl_ctx *ctx, const float *host_src, int32_t sw, int32_t sh, float *host_dst, int32_t dw, int32_t dh,
               const float trans[6], float oob) {
    return project_host(ctx, host_src, sw, sh, host_dst, dw, dh, trans, oob, false, 1.0f, 0.0f);
}

int nl_project_scaled(nl_ctx *ctx, const float *host_src, int32_t sw, int32_t sh, float *host_dst, int32_t dw, int32_t dh,
                      const float trans[6], float oob, float multiplier, float offset) {
    return project_host(ctx, host_src, sw, sh, host_dst, dw, dh, trans, oob, true, multiplier, offset);
}

}  // extern "C"
