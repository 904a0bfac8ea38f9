// nl_fits.cu -- FITS pixel payload decode / encode on the device (the two steps either side of the hot
// path).  Replaces the conversion loops of internal/fits/read.go:176-443
// (big-endian BITPIX 8/16/32/64/-32/-64 -> fp32 with v = float32(val)*Bscale + Bzero, mul then add) and
// write.go:182-215 (fp32 -> big-endian, NaN -> 0).  Uploading the RAW payload and decoding it on the
// device halves the PCIe traffic of 16-bit camera frames (2 instead of 4 bytes per sample).
// Header parsing, gzip and the running min / max / mean of the reader stay on the host.
// Streaming kernels: (|BITPIX|/8 + 4) bytes per sample.
#include "nl_internal.h"

namespace nl {

__device__ __forceinline__ unsigned bswap32(unsigned v) { return __byte_perm(v, 0, 0x0123); }

template <int BITPIX>
__device__ __forceinline__ float fits_value(const unsigned char *raw, long long i) {
    if (BITPIX == 8) return (float)raw[i];
    if (BITPIX == 16) {
        const unsigned short u = reinterpret_cast<const unsigned short *>(raw)[i];
        return (float)(short)(unsigned short)((u << 8) | (u >> 8));
    }
    if (BITPIX == 32) return (float)(int)bswap32(reinterpret_cast<const unsigned *>(raw)[i]);
    if (BITPIX == -32) return __uint_as_float(bswap32(reinterpret_cast<const unsigned *>(raw)[i]));
    const uint2 u = reinterpret_cast<const uint2 *>(raw)[i];
    const unsigned long long be = ((unsigned long long)bswap32(u.x) << 32) | bswap32(u.y);
    if (BITPIX == 64) return (float)(long long)be;
    return (float)__longlong_as_double((long long)be);                 // -64: float32(float64)
}

template <int BITPIX>
__global__ void __launch_bounds__(256) fits_decode_kernel(const unsigned char *__restrict__ raw, long long n, float bscale,
                                                          float bzero, float *__restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __fadd_rn(__fmul_rn(fits_value<BITPIX>(raw, i), bscale), bzero);     // read.go:196, 237, ...
}

// Four samples per thread for the two payload types that matter (16-bit camera frames, fp32 stacks): one 8- or 16-byte
// load and one 16-byte store; the up-to-three samples behind the last whole group take the scalar form.
template <int BITPIX>
__global__ void __launch_bounds__(256) fits_decode_vec_kernel(const unsigned char *__restrict__ raw, long long n, float bscale,
                                                              float bzero, float *__restrict__ dst) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float v[4];
        if (BITPIX == 16) {
            const uint2 u = __ldcs(reinterpret_cast<const uint2 *>(raw) + i);
            const unsigned w[2] = {u.x, u.y};
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const unsigned lo = w[k] & 0xffffu, hi = w[k] >> 16;
                v[2 * k] = (float)(short)(unsigned short)(((lo << 8) | (lo >> 8)) & 0xffffu);
                v[2 * k + 1] = (float)(short)(unsigned short)(((hi << 8) | (hi >> 8)) & 0xffffu);
            }
        } else {
            const uint4 u = __ldcs(reinterpret_cast<const uint4 *>(raw) + i);
            v[0] = __uint_as_float(bswap32(u.x)); v[1] = __uint_as_float(bswap32(u.y));
            v[2] = __uint_as_float(bswap32(u.z)); v[3] = __uint_as_float(bswap32(u.w));
        }
        float4 o;
        o.x = __fadd_rn(__fmul_rn(v[0], bscale), bzero); o.y = __fadd_rn(__fmul_rn(v[1], bscale), bzero);
        o.z = __fadd_rn(__fmul_rn(v[2], bscale), bzero); o.w = __fadd_rn(__fmul_rn(v[3], bscale), bzero);
        reinterpret_cast<float4 *>(dst)[i] = o;
    }
    const long long t = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) dst[t] = __fadd_rn(__fmul_rn(fits_value<BITPIX>(raw, t), bscale), bzero);
}

__global__ void __launch_bounds__(256) fits_encode_vec_kernel(const float *__restrict__ src, long long n, unsigned *__restrict__ raw) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 d = __ldcs(reinterpret_cast<const float4 *>(src) + i);
        if (d.x != d.x) d.x = 0.0f;                                    // write.go:192
        if (d.y != d.y) d.y = 0.0f;
        if (d.z != d.z) d.z = 0.0f;
        if (d.w != d.w) d.w = 0.0f;
        reinterpret_cast<uint4 *>(raw)[i] = make_uint4(bswap32(__float_as_uint(d.x)), bswap32(__float_as_uint(d.y)),
                                                       bswap32(__float_as_uint(d.z)), bswap32(__float_as_uint(d.w)));
    }
    const long long t = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) {
        float d = src[t];
        if (d != d) d = 0.0f;
        raw[t] = bswap32(__float_as_uint(d));
    }
}

__global__ void __launch_bounds__(256) fits_encode_kernel(const float *__restrict__ src, long long n, unsigned *__restrict__ raw) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d = src[i];
        if (d != d) d = 0.0f;                                          // write.go:192
        raw[i] = bswap32(__float_as_uint(d));
    }
}

int fits_decode_launch(nl_ctx *ctx, const void *dev_raw, int bitpix, long long n, float bscale, float bzero, float *dev_dst) {
    if (n == 0) return NL_OK;
    long long grid = (n + 255) / 256;
    if (grid > (long long)ctx->sm_count * 32) grid = (long long)ctx->sm_count * 32;
    const unsigned char *raw = (const unsigned char *)dev_raw;
    if ((bitpix == 16 || bitpix == -32) && (((uintptr_t)dev_raw | (uintptr_t)dev_dst) & 15) == 0) {
        long long vgrid = (n / 4 + 255) / 256;
        if (vgrid > (long long)ctx->sm_count * 16) vgrid = (long long)ctx->sm_count * 16;
        if (vgrid < 1) vgrid = 1;
        if (bitpix == 16) fits_decode_vec_kernel<16><<<(unsigned)vgrid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst);
        else fits_decode_vec_kernel<-32><<<(unsigned)vgrid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst);
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        return NL_OK;
    }
    switch (bitpix) {
    case 8: fits_decode_kernel<8><<<(unsigned)grid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst); break;
    case 16: fits_decode_kernel<16><<<(unsigned)grid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst); break;
    case 32: fits_decode_kernel<32><<<(unsigned)grid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst); break;
    case 64: fits_decode_kernel<64><<<(unsigned)grid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst); break;
    case -32: fits_decode_kernel<-32><<<(unsigned)grid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst); break;
    case -64: fits_decode_kernel<-64><<<(unsigned)grid, 256, 0, ctx->stream>>>(raw, n, bscale, bzero, dev_dst); break;
    default: return set_error(NL_E_INVALID, "Unknown BITPIX value %d", bitpix);   // read.go:168
    }
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

}  // namespace nl

using namespace nl;

extern "C" {

int nl_fits_decode_dev(nl_ctx *ctx, const void *dev_raw, int32_t bitpix, int64_t count, float bscale, float bzero, float *dev_dst) {
    NL_REQUIRE(ctx && count >= 0 && (count == 0 || (dev_raw && dev_dst)), "bad argument");
    NL_GUARD(ctx);
    return fits_decode_launch(ctx, dev_raw, bitpix, count, bscale, bzero, dev_dst);
}

int nl_fits_decode(nl_ctx *ctx, const void *host_raw, int32_t bitpix, int64_t count, float bscale, float bzero, float *host_dst) {
    NL_REQUIRE(ctx && count >= 0 && (count == 0 || (host_raw && host_dst)), "bad argument");
    NL_REQUIRE(bitpix == 8 || bitpix == 16 || bitpix == 32 || bitpix == 64 || bitpix == -32 || bitpix == -64, "Unknown BITPIX value");
    if (count == 0) return NL_OK;
    NL_GUARD(ctx);
    const size_t bytes_per = (size_t)(bitpix < 0 ? -bitpix : bitpix) / 8;
    const size_t raw_bytes = ((size_t)count * bytes_per + 255) & ~(size_t)255;
    int rc = ensure_scratch(ctx, raw_bytes + sizeof(float) * (size_t)count);
    if (rc != NL_OK) return rc;
    float *dev_dst = (float *)((char *)ctx->scratch + raw_bytes);
    NL_CUDA(cudaMemcpyAsync(ctx->scratch, host_raw, (size_t)count * bytes_per, cudaMemcpyHostToDevice, ctx->stream));
    rc = fits_decode_launch(ctx, ctx->scratch, bitpix, count, bscale, bzero, dev_dst);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(host_dst, dev_dst, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_fits_encode_dev(nl_ctx *ctx, const float *dev_src, int64_t count, void *dev_raw) {
    NL_REQUIRE(ctx && count >= 0 && (count == 0 || (dev_src && dev_raw)), "bad argument");
    if (count == 0) return NL_OK;
    NL_GUARD(ctx);
    if ((((uintptr_t)dev_src | (uintptr_t)dev_raw) & 15) == 0) {
        long long vgrid = (count / 4 + 255) / 256;
        if (vgrid > (long long)ctx->sm_count * 16) vgrid = (long long)ctx->sm_count * 16;
        if (vgrid < 1) vgrid = 1;
        fits_encode_vec_kernel<<<(unsigned)vgrid, 256, 0, ctx->stream>>>(dev_src, count, (unsigned *)dev_raw);
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        return NL_OK;
    }
    long long grid = (count + 255) / 256;
    if (grid > (long long)ctx->sm_count * 32) grid = (long long)ctx->sm_count * 32;
    fits_encode_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(dev_src, count, (unsigned *)dev_raw);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

int nl_fits_encode(nl_ctx *ctx, const float *host_src, int64_t count, void *host_raw) {
    NL_REQUIRE(ctx && count >= 0 && (count == 0 || (host_src && host_raw)), "bad argument");
    if (count == 0) return NL_OK;
    NL_GUARD(ctx);
    const size_t bytes = sizeof(float) * (size_t)count, half = (bytes + 255) & ~(size_t)255;
    int rc = ensure_scratch(ctx, 2 * half);
    if (rc != NL_OK) return rc;
    void *dev_raw = (char *)ctx->scratch + half;
    NL_CUDA(cudaMemcpyAsync(ctx->scratch, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = nl_fits_encode_dev(ctx, (const float *)ctx->scratch, count, dev_raw);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(host_raw, dev_raw, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

}  // extern "C"
