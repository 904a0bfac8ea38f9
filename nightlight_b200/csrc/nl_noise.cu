// nl_noise.cu -- Immerkaer noise estimate of whole frames on the device.
// Replaces stats.EstimateNoise (internal/stats/noise_amd64.go:25-43, noise.go:24-55): the 3x3 Laplacian
// [1 -2 1; -2 4 -2; 1 -2 1] of every interior pixel, absolute values summed per image row, the row sums
// summed sequentially over the rows, times sqrt(pi/2) / (6 (w-2)(h-2)), all in fp32.  It feeds the
// inverse-noise stacking weights (getWeights, stack.go:247-259; BASELINE configs[1]).  The row sum comes
// in the two orders the reference has (selected per context, nl_ctx_set_numerics):
//
//   amd64 (noise_amd64.s:75-192, what an amd64 build runs on any CPU with AVX2): the row is walked in
//     flights of six pixels, pixel x0+1+l in SIMD lane l; per lane four partial chains with fused
//     multiply-adds (y0 = d00, y0 = fma(d11,4,y0), y0 = fma(d22,1,y0), ...), added as (y3+y2) + (y1+y0);
//     every lane keeps its own running sum over the flights; a last flight flush with the row end covers
//     the remaining columns with the lanes already done masked off; the lanes are folded l^4, l^2, l^1.
//     `noise_rows_amd64_kernel`: one thread per (row, lane), eight threads per row, shuffles for the fold.
//   pure Go (noise.go:32-55): products accumulated one by one in row-major order, no FMA, one sequential
//     chain per row.  `noise_rows_kernel`: one thread per row with a sliding window in registers.
//
// Rows and frames run in parallel; every 32-byte sector a thread touches serves its next steps out of
// L1 and DRAM sees every row once.  Algorithmic bytes: 4 per pixel.
#include "nl_internal.h"

#include <math.h>

namespace nl {

__global__ void __launch_bounds__(128) noise_rows_kernel(const float *__restrict__ frames, long long stride, int w, int h,
                                                         float *__restrict__ row_sums) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int f = blockIdx.y;
    if (y >= h - 1) return;
    const float *r0 = frames + (long long)f * stride + (long long)(y - 1) * w, *r1 = r0 + w, *r2 = r1 + w;
    float a0 = __ldg(r0), a1 = __ldg(r0 + 1), b0 = __ldg(r1), b1 = __ldg(r1 + 1), c0 = __ldg(r2), c1 = __ldg(r2 + 1);
    float row_sum = 0.0f;
    for (int x = 1; x < w - 1; x++) {
        const float a2 = __ldg(r0 + x + 1), b2 = __ldg(r1 + x + 1), c2 = __ldg(r2 + x + 1);
        float conv = __fadd_rn(0.0f, a0);                       // noise.go:46-49: conv += data[i+o]*w, in offset order
        conv = __fadd_rn(conv, __fmul_rn(a1, -2.0f));
        conv = __fadd_rn(conv, a2);
        conv = __fadd_rn(conv, __fmul_rn(b0, -2.0f));
        conv = __fadd_rn(conv, __fmul_rn(b1, 4.0f));
        conv = __fadd_rn(conv, __fmul_rn(b2, -2.0f));
        conv = __fadd_rn(conv, c0);
        conv = __fadd_rn(conv, __fmul_rn(c1, -2.0f));
        conv = __fadd_rn(conv, c2);
        row_sum = __fadd_rn(row_sum, fabsf(conv));
        a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
    }
    row_sums[(long long)f * h + y] = row_sum;
}

// thread = (row y, SIMD lane l); lanes 6 and 7 of the reference's vectors only ever hold zero
__global__ void __launch_bounds__(128) noise_rows_amd64_kernel(const float *__restrict__ frames, long long stride, int w, int h,
                                                               float *__restrict__ row_sums) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = t & 7;
    int y = (t >> 3) + 1;
    const bool live = y < h - 1;
    if (!live) y = h - 2;                                   // keep the whole warp for the shuffles below
    const int f = blockIdx.y;
    const float *r0 = frames + (long long)f * stride + (long long)(y - 1) * w + l, *r1 = r0 + w, *r2 = r1 + w;
    float acc = 0.0f;
    const int bp = w - 7;
    const int flights = (bp + 5) / 6;                       // x0 = 0, 6, ... while x0 < bp
    const int ax = 6 * flights - bp;                        // noise_amd64.s:170-172
    const int extra = ax < 5;                               // one more flight at x0 = w-8, lanes > ax only
    if (l < 6) {
        for (int i = 0; i < flights + extra; i++) {
            const int x0 = i < flights ? 6 * i : w - 8;
            if (i == flights && l < ax + 1) break;
            const float *a = r0 + x0, *b = r1 + x0, *c = r2 + x0;
            float y0 = __ldg(a), y1 = __fmul_rn(__ldg(a + 1), -2.0f), y2 = __ldg(a + 2), y3 = __fmul_rn(__ldg(b), -2.0f);
            y0 = __fmaf_rn(__ldg(b + 1), 4.0f, y0);
            y1 = __fmaf_rn(__ldg(b + 2), -2.0f, y1);
            y2 = __fmaf_rn(__ldg(c), 1.0f, y2);
            y3 = __fmaf_rn(__ldg(c + 1), -2.0f, y3);
            y0 = __fmaf_rn(__ldg(c + 2), 1.0f, y0);
            y2 = __fadd_rn(y3, y2);
            y0 = __fadd_rn(y1, y0);
            y0 = __fadd_rn(y2, y0);
            acc = __fadd_rn(fabsf(y0), acc);
        }
    }
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));      // :183-190
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    if (live && l == 0) row_sums[(long long)f * h + y] = acc;
}

__global__ void noise_finalize_kernel(const float *__restrict__ row_sums, int n, int h, float factor, float *__restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    float sum = 0.0f;
    for (int y = 1; y < h - 1; y++) sum = __fadd_rn(sum, row_sums[(long long)f * h + y]);
    out[f] = __fmul_rn(sum, factor);
}

}  // namespace nl

using namespace nl;

extern "C" {

int nl_estimate_noise_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, int32_t width,
                          int32_t height, float *host_noise) {
    NL_REQUIRE(ctx && host_noise && n_frames >= 0, "bad argument");
    NL_REQUIRE(width >= 0 && height >= 0 && frame_stride >= (int64_t)width * height, "bad frame geometry");
    if (n_frames == 0) return NL_OK;
    NL_REQUIRE(dev_frames, "NULL frames");
    // noise.go:52: float32(math.Sqrt(0.5*math.Pi)) / (6 * float32(width-2) * float32(height-2))
    volatile float denom = 6.0f * (float)(width - 2);
    denom = denom * (float)(height - 2);
    const float factor = (float)sqrt(0.5 * M_PI) / denom;
    CtxGuard g(ctx);
    const size_t rows_bytes = sizeof(float) * (size_t)n_frames * (size_t)(height > 0 ? height : 1);
    int rc = ensure_scratch(ctx, ((rows_bytes + 255) & ~(size_t)255) + sizeof(float) * (size_t)n_frames);
    if (rc != NL_OK) return rc;
    float *row_sums = (float *)ctx->scratch;
    float *dev_out = (float *)((char *)ctx->scratch + ((rows_bytes + 255) & ~(size_t)255));
    if (height > 2 && width > 2) {
        // the AVX2 kernel needs eight columns (narrower rows make it read before the row): those take the Go loop
        if (ctx->numerics == NL_NUMERICS_AMD64 && width >= 8) {
            dim3 grid((unsigned)(((height - 2) * 8 + 127) / 128), (unsigned)n_frames);
            noise_rows_amd64_kernel<<<grid, 128, 0, ctx->stream>>>(dev_frames, frame_stride, width, height, row_sums);
        } else {
            dim3 grid((unsigned)((height - 2 + 127) / 128), (unsigned)n_frames);
            noise_rows_kernel<<<grid, 128, 0, ctx->stream>>>(dev_frames, frame_stride, width, height, row_sums);
        }
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    noise_finalize_kernel<<<(n_frames + 127) / 128, 128, 0, ctx->stream>>>(row_sums, n_frames, height, factor, dev_out);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    NL_CUDA(cudaMemcpyAsync(host_noise, dev_out, sizeof(float) * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_estimate_noise(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float *noise) {
    NL_REQUIRE(ctx && noise && len >= 0 && width > 0, "bad argument");
    NL_REQUIRE(host_data || len == 0, "NULL data");
    CtxGuard g(ctx);
    float *dev = nullptr;
    NL_CUDA(cudaMalloc(&dev, sizeof(float) * (size_t)(len > 0 ? len : 1)));
    cudaError_t e = cudaMemcpyAsync(dev, host_data, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? nl_estimate_noise_dev(ctx, dev, 1, len, width, len / width, noise) : cuda_fail(e, "noise upload");
    cudaFree(dev);
    return rc;
}

}  // extern "C"
