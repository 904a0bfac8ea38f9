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
//     `noise_rows_amd64_kernel`: one warp per row = four flights x eight SIMD lanes at a time.
//   pure Go (noise.go:32-55): products accumulated one by one in row-major order, no FMA, one sequential
//     chain per row.  `noise_rows_kernel`: one thread per row with a sliding window in registers (many
//     frames at once); `noise_rows_warp_kernel`: one warp per row (a single frame).
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

// A warp walks its row front to back with one dependent step per 24 or 32 pixels, so it cannot hide DRAM latency by
// itself: lanes 0..2 pull the line of rows y-1, y, y+1 that the walk reaches `ahead` floats later into L1.
__device__ __forceinline__ void prefetch_rows(const float *row0, int w, int x, int lane, const float *frame_end) {
    if (lane < 3) {
        const float *p = row0 + (long long)lane * w + x;
        if (p < frame_end) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }
}

// One warp per image row.  Lane = (q, l): flight q of a group of four, SIMD lane l of the reference's vector
// (lanes 6 and 7 of the vectors only ever hold zero).  The 32 lanes of the warp compute the |Laplacian| of 24
// consecutive pixels at once; the running sums of the reference are per SIMD lane over the flights in order,
// so the eight q = 0 lanes then add their column of four values one after the other.
__global__ void __launch_bounds__(256) noise_rows_amd64_kernel(const float *__restrict__ frames, long long stride, int w, int h,
                                                               float *__restrict__ row_sums) {
    const int lane = threadIdx.x & 31, l = lane & 7, q = lane >> 3;
    const int y = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) + 1;
    if (y >= h - 1) return;
    const int f = blockIdx.y;
    const float *r0 = frames + (long long)f * stride + (long long)(y - 1) * w + l, *r1 = r0 + w, *r2 = r1 + w;
    const int bp = w - 7;
    const int flights = (bp + 5) / 6;                       // x0 = 0, 6, ... while x0 < bp
    const int ax = 6 * flights - bp;                        // noise_amd64.s:170-172
    const int total = flights + (ax < 5);                   // one more flight at x0 = w-8 for lanes > ax
    float acc = 0.0f;
    const float *row0 = frames + (long long)f * stride + (long long)(y - 1) * w, *frame_end = frames + (long long)f * stride + (long long)w * h;
    for (int x = 0; x < 512; x += 32) prefetch_rows(row0, w, x, lane, frame_end);
#pragma unroll 2
    for (int i0 = 0; i0 < total; i0 += 4) {
        prefetch_rows(row0, w, 6 * i0 + 512, lane, frame_end);
        const int i = i0 + q;
        float v = 0.0f;
        if (l < 6 && i < total && !(i == flights && l < ax + 1)) {
            const int x0 = i < flights ? 6 * i : w - 8;
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
            v = fabsf(y0);
        }
        // masked and missing flights contribute +0 to a non-negative sum: no change
        const float v1 = __shfl_sync(0xffffffffu, v, l + 8), v2 = __shfl_sync(0xffffffffu, v, l + 16), v3 = __shfl_sync(0xffffffffu, v, l + 24);
        acc = __fadd_rn(v, acc);
        acc = __fadd_rn(v1, acc);
        acc = __fadd_rn(v2, acc);
        acc = __fadd_rn(v3, acc);
    }
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));      // :183-190
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    if (lane == 0) row_sums[(long long)f * h + y] = acc;
}

// Pure-Go order with one warp per row, for single frames (the thread-per-row kernel above needs tens of
// thousands of rows to fill the GPU): the lanes compute the Laplacians of 32 consecutive pixels, lane 0
// adds them to the row sum in pixel order.
__global__ void __launch_bounds__(256) noise_rows_warp_kernel(const float *__restrict__ frames, long long stride, int w, int h,
                                                              float *__restrict__ row_sums) {
    const int lane = threadIdx.x & 31;
    const int y = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) + 1;
    if (y >= h - 1) return;
    const int f = blockIdx.y;
    const float *r0 = frames + (long long)f * stride + (long long)(y - 1) * w, *r1 = r0 + w, *r2 = r1 + w;
    float row_sum = 0.0f;
    const float *frame_end = frames + (long long)f * stride + (long long)w * h;
    for (int x = 0; x < 512; x += 32) prefetch_rows(r0, w, x, lane, frame_end);
#pragma unroll 2
    for (int x0 = 1; x0 < w - 1; x0 += 32) {
        prefetch_rows(r0, w, x0 + 512, lane, frame_end);
        const int x = x0 + lane;
        float v = 0.0f;
        if (x < w - 1) {
            float conv = __fadd_rn(0.0f, __ldg(r0 + x - 1));
            conv = __fadd_rn(conv, __fmul_rn(__ldg(r0 + x), -2.0f));
            conv = __fadd_rn(conv, __ldg(r0 + x + 1));
            conv = __fadd_rn(conv, __fmul_rn(__ldg(r1 + x - 1), -2.0f));
            conv = __fadd_rn(conv, __fmul_rn(__ldg(r1 + x), 4.0f));
            conv = __fadd_rn(conv, __fmul_rn(__ldg(r1 + x + 1), -2.0f));
            conv = __fadd_rn(conv, __ldg(r2 + x - 1));
            conv = __fadd_rn(conv, __fmul_rn(__ldg(r2 + x), -2.0f));
            conv = __fadd_rn(conv, __ldg(r2 + x + 1));
            v = fabsf(conv);
        }
        const int cnt = min(32, w - 1 - x0);
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const float vk = __shfl_sync(0xffffffffu, v, k);
            if (k < cnt) row_sum = __fadd_rn(row_sum, vk);
        }
    }
    if (lane == 0) row_sums[(long long)f * h + y] = row_sum;
}

// noise_amd64.go:36-42 / noise.go:38-51: the row sums are added in row order: one CTA per frame stages them in
// shared memory (coalesced, all in flight at once), thread 0 then runs the chain of dependent additions from there
__global__ void __launch_bounds__(256) noise_finalize_kernel(const float *__restrict__ row_sums, int h, float factor,
                                                             float *__restrict__ out) {
    __shared__ float stage[4096];
    const int f = blockIdx.x;
    const float *rs = row_sums + (long long)f * h;
    float sum = 0.0f;
    for (int y0 = 1; y0 < h - 1; y0 += 4096) {
        const int cnt = min(4096, h - 1 - y0);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) stage[i] = rs[y0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int i = 0;
            for (; i + 4 <= cnt; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(stage + i);
                sum = __fadd_rn(sum, v.x); sum = __fadd_rn(sum, v.y); sum = __fadd_rn(sum, v.z); sum = __fadd_rn(sum, v.w);
            }
            for (; i < cnt; i++) sum = __fadd_rn(sum, stage[i]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[f] = __fmul_rn(sum, factor);
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
    NL_GUARD(ctx);
    const size_t rows_bytes = sizeof(float) * (size_t)n_frames * (size_t)(height > 0 ? height : 1);
    int rc = ensure_scratch(ctx, ((rows_bytes + 255) & ~(size_t)255) + sizeof(float) * (size_t)n_frames);
    if (rc != NL_OK) return rc;
    float *row_sums = (float *)ctx->scratch;
    float *dev_out = (float *)((char *)ctx->scratch + ((rows_bytes + 255) & ~(size_t)255));
    if (height > 2 && width > 2) {
        // the AVX2 kernel needs eight columns (narrower rows make it read before the row): those take the Go loop
        dim3 warp_grid((unsigned)((height - 2 + 7) / 8), (unsigned)n_frames);
        if (ctx->numerics == NL_NUMERICS_AMD64 && width >= 8) {
            noise_rows_amd64_kernel<<<warp_grid, 256, 0, ctx->stream>>>(dev_frames, frame_stride, width, height, row_sums);
        } else if ((long long)n_frames * (height - 2) < 64 * 1024) {
            noise_rows_warp_kernel<<<warp_grid, 256, 0, ctx->stream>>>(dev_frames, frame_stride, width, height, row_sums);
        } else {
            dim3 grid((unsigned)((height - 2 + 127) / 128), (unsigned)n_frames);
            noise_rows_kernel<<<grid, 128, 0, ctx->stream>>>(dev_frames, frame_stride, width, height, row_sums);
        }
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    noise_finalize_kernel<<<n_frames, 256, 0, ctx->stream>>>(row_sums, height, factor, dev_out);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    NL_CUDA(cudaMemcpyAsync(host_noise, dev_out, sizeof(float) * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_estimate_noise(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float *noise) {
    NL_REQUIRE(ctx && noise && len >= 0 && width > 0, "bad argument");
    NL_REQUIRE(host_data || len == 0, "NULL data");
    NL_GUARD(ctx);
    float *dev = nullptr;
    int rc = ensure_frame(ctx, 0, sizeof(float) * (size_t)(len > 0 ? len : 1), &dev);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(dev, host_data, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    return nl_estimate_noise_dev(ctx, dev, 1, len, width, len / width, noise);
}

}  // extern "C"
