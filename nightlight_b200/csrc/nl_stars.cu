// nl_stars.cu -- star detection: the full-frame threshold scan on the GPU, the sparse per-star
// steps on the host.  Replaces star.FindStars (internal/star/findstars.go:59-100).
//
// The reference's detector is not a convolution: it is a raster scan for pixels above
// location + scale*starSig with a sequential same-row de-duplication against the last kept
// candidate (findstars.go:105-129), followed by sparse per-star work (3x3 median reject, unstable
// quicksort by mass, grid overlap filter, iterative centre of mass, half-flux radius).  Only the
// scan touches every pixel (4 B/pixel, HBM bound); it is the part that runs on the device:
//
//   bright_rows_kernel<false>  one warp per image row streams the row (coalesced 128 B per load,
//                              8 loads in flight per lane), ballots the hits and replays the
//                              reference's de-duplication state machine over the hit bits in raster
//                              order -- the dependency never crosses a row, because a candidate
//                              only merges with a predecessor of the same Y.  Emits the row's count.
//   row_offsets_kernel         exclusive prefix sum of the row counts (raster order of the output).
//   bright_rows_kernel<true>   the same walk again, now writing each row's candidates at its offset
//                              (the second read of the frame is served largely by the 126 MB L2).
#include "nl_internal.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
#include <vector>

namespace nl {

__host__ __device__ static inline void cswap(float &a, float &b) { if (a > b) { float t = a; a = b; b = t; } }

// MedianFloat32Slice9, median3x3.go:85-110 (the partially sorted buffer is an observable side effect)
__host__ __device__ static inline float median9(float *a) {
    cswap(a[0], a[1]); cswap(a[3], a[4]); cswap(a[6], a[7]);
    cswap(a[1], a[2]); cswap(a[4], a[5]); cswap(a[7], a[8]);
    cswap(a[0], a[1]); cswap(a[3], a[4]); cswap(a[6], a[7]);
    if (a[0] > a[3]) a[3] = a[0];
    if (a[3] > a[6]) a[6] = a[3];
    cswap(a[1], a[4]);
    if (a[4] > a[7]) a[4] = a[7];
    if (a[1] > a[4]) a[4] = a[1];
    if (a[5] > a[8]) a[5] = a[8];
    if (a[2] > a[5]) a[2] = a[5];
    cswap(a[2], a[4]);
    if (a[4] > a[6]) a[4] = a[6];
    if (a[2] > a[4]) a[4] = a[2];
    return a[4];
}


template <bool WRITE>
__global__ void __launch_bounds__(256) bright_rows_kernel(const float *__restrict__ data, int len, int width, int rows,
                                                          float threshold, int radius, int *__restrict__ row_count,
                                                          const int *__restrict__ row_offset, nl_star *__restrict__ out,
                                                          int cap) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= rows) return;
    const long long row0 = (long long)row * width;
    const int row_len = (int)min((long long)width, (long long)len - row0);
    const float *src = data + row0;
    const int base_out = WRITE ? row_offset[row] : 0;

    int count = 0;            // candidates kept in this row so far (warp-uniform)
    int last_x = 0;           // the last kept candidate of this row
    float last_v = 0.0f;
    constexpr int U = 8;
    for (int x0 = 0; x0 < row_len; x0 += 32 * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int x = x0 + u * 32 + lane;
            v[u] = x < row_len ? __ldcs(src + x) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int x = x0 + u * 32 + lane;
            unsigned hits = __ballot_sync(0xffffffffu, x < row_len && v[u] > threshold);
            while (hits) {                                   // raster order within the chunk
                const int b = __ffs(hits) - 1;
                hits &= hits - 1;
                const float hv = __shfl_sync(0xffffffffu, v[u], b);
                const int hx = x0 + u * 32 + b;
                // findstars.go:113-123: same row and within `radius` of the last kept candidate
                if (count > 0 && last_x >= hx - radius) {
                    if (last_v >= hv) continue;              // keep the older, brighter one
                } else {
                    count++;
                }
                last_x = hx; last_v = hv;
                if (WRITE && lane == 0) {
                    const long long slot = (long long)base_out + count - 1;
                    if (slot < cap) {
                        nl_star s;
                        s.index = (int)(row0 + hx); s.value = hv; s.x = (float)hx; s.y = (float)row;
                        s.mass = hv; s.hfr = 1.0f;
                        out[slot] = s;
                    }
                }
            }
        }
    }
    if (!WRITE && lane == 0) row_count[row] = count;
}

// exclusive scan of the row counts; rows <= a few 10^4, one CTA is plenty
__global__ void __launch_bounds__(1024) row_offsets_kernel(const int *__restrict__ row_count, int *__restrict__ row_offset,
                                                           int rows, int *__restrict__ total) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < rows; base += 1024) {
        const int i = base + threadIdx.x;
        const int c = i < rows ? row_count[i] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;     // exclusive
        }
        __syncthreads();
        const int excl = carry + warp_sums[warp] + incl - c;
        if (i < rows) row_offset[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// ---- all frames of a resident stack at once, ONE read of every frame ------------------------------------------------
// bright_rows_slots_kernel: the walk of bright_rows_kernel over frame blockIdx.y, writing the row's candidates into the
// row's own K slots while counting (rows with more than K candidates keep counting; their frame is redone by the
// two-pass scan).  row_offsets_batch_kernel: exclusive scan of the row counts per frame.  bright_compact_kernel: the
// slots of every row move to the frame's raster-ordered list.
constexpr int BRIGHT_SLOTS = 32;

__global__ void __launch_bounds__(256) bright_rows_slots_kernel(const float *__restrict__ frames, long long frame_stride, int len, int width,
                                                                int rows, const float *__restrict__ thresholds, int radius,
                                                                int *__restrict__ row_count, nl_star *__restrict__ slots) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= rows) return;
    const int frame = blockIdx.y;
    const float threshold = thresholds[frame];
    const long long row0 = (long long)row * width;
    const int row_len = (int)min((long long)width, (long long)len - row0);
    const float *src = frames + (long long)frame * frame_stride + row0;
    nl_star *my = slots + ((long long)frame * rows + row) * BRIGHT_SLOTS;
    int count = 0, last_x = 0;
    float last_v = 0.0f;
    constexpr int U = 8;
    for (int x0 = 0; x0 < row_len; x0 += 32 * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int x = x0 + u * 32 + lane;
            v[u] = x < row_len ? __ldcs(src + x) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int x = x0 + u * 32 + lane;
            unsigned hits = __ballot_sync(0xffffffffu, x < row_len && v[u] > threshold);
            while (hits) {
                const int b = __ffs(hits) - 1;
                hits &= hits - 1;
                const float hv = __shfl_sync(0xffffffffu, v[u], b);
                const int hx = x0 + u * 32 + b;
                if (count > 0 && last_x >= hx - radius) {            // findstars.go:113-123
                    if (last_v >= hv) continue;
                } else {
                    count++;
                }
                last_x = hx; last_v = hv;
                if (lane == 0 && count <= BRIGHT_SLOTS) {
                    nl_star st;
                    st.index = (int)(row0 + hx); st.value = hv; st.x = (float)hx; st.y = (float)row;
                    st.mass = hv; st.hfr = 1.0f;
                    my[count - 1] = st;
                }
            }
        }
    }
    if (lane == 0) row_count[(long long)frame * rows + row] = count;
}

__global__ void __launch_bounds__(1024) row_offsets_batch_kernel(const int *__restrict__ row_count, int *__restrict__ row_offset, int rows,
                                                                 int *__restrict__ totals, int *__restrict__ overflow) {
    // one CTA per frame: exclusive scan of its row counts; totals[frame] = candidates, overflow[frame] = rows beyond their slots
    __shared__ int warp_sums[32];
    __shared__ int carry, over;
    const int *rc = row_count + (long long)blockIdx.x * rows;
    int *ro = row_offset + (long long)blockIdx.x * rows;
    if (threadIdx.x == 0) { carry = 0; over = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < rows; base += 1024) {
        const int i = base + threadIdx.x;
        const int c = i < rows ? rc[i] : 0;
        if (c > BRIGHT_SLOTS) atomicAdd(&over, 1);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry + warp_sums[warp] + incl - c;
        if (i < rows) ro[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[blockIdx.x] = carry; overflow[blockIdx.x] = over; }
}

// With bp_thr (one threshold per frame): rejectBadPixels' test (findstars.go:134-169) for every candidate whose 3x3
// neighbourhood lies inside the frame -- flag 1: keep, 0: reject -- while the candidate is at hand; flag 2 marks the
// candidates of the first and last row, whose gather buffer keeps entries of the candidate before them (the host
// replays those).
__global__ void __launch_bounds__(256) bright_compact_kernel(const nl_star *__restrict__ slots, const int *__restrict__ row_count,
                                                             const int *__restrict__ row_offset, int rows, nl_star *__restrict__ list,
                                                             long long list_stride, const float *__restrict__ frames,
                                                             long long frame_stride, int len, int width,
                                                             const float *__restrict__ bp_thr, unsigned char *__restrict__ flags) {
    // one warp per row: lane i moves candidate i of the row to the frame's list (frames with overflowing rows are redone)
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= rows) return;
    const long long fr = (long long)blockIdx.y * rows + row;
    const int c = row_count[fr];
    if (lane < c && lane < BRIGHT_SLOTS) {
        const long long dst = (long long)row_offset[fr] + lane;
        if (dst < list_stride) {
            const nl_star s = slots[fr * BRIGHT_SLOTS + lane];
            list[(long long)blockIdx.y * list_stride + dst] = s;
            if (bp_thr) {
                unsigned char flag = 2;
                const long long idx = s.index;
                if (idx - width - 1 >= 0 && idx + width + 1 < len) {
                    const float *d = frames + (long long)blockIdx.y * frame_stride;
                    float b[9];
#pragma unroll
                    for (int y = -1; y <= 1; y++)
#pragma unroll
                        for (int x = -1; x <= 1; x++) b[(y + 1) * 3 + (x + 1)] = d[idx + (long long)y * width + x];
                    const float med = median9(b);
                    const float diff = d[idx] - med;
                    const float thr = bp_thr[blockIdx.y];
                    flag = (diff < thr && -diff < thr) ? 1 : 0;
                }
                flags[(long long)blockIdx.y * list_stride + dst] = flag;
            }
        }
    }
}

// ---- the two window passes of FindStars on the device, one thread per star, all frames of a batch in one launch -------
// shiftToCenterOfMass (findstars.go:274-322): up to ten rounds of first moments over the (2r+1)^2 window around the
// star, summed in raster order in fp32 exactly like the reference (a thread walks its window sequentially; the
// windows of neighbouring stars overlap in L1/L2).  calcAndFilterHalfFluxRadius (findstars.go:327-396): the half-flux
// radius over the disc and the inner / outer mass test.  Both update the star records in place (mapped pinned host
// memory: a few megabytes for a whole batch) and leave the order-dependent parts -- the sums of the shifts and of the
// radii, the compaction of the survivors -- to the host.
// int32(float) as Go on amd64 (and the C oracle) computes it: CVTTSS2SI yields 0x80000000 for NaN and out-of-range
// values, where the GPU's conversion saturates
__device__ __forceinline__ int go_int32(float v) {
    return (v >= -2147483648.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000;
}

__global__ void __launch_bounds__(128) star_com_kernel(const float *__restrict__ frames, long long frame_stride, int len, int width,
                                                       const float *__restrict__ thresholds, int radius, nl_star *__restrict__ lists,
                                                       long long list_stride, const int *__restrict__ counts, float *__restrict__ shifts) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[frame]) return;
    const float *data = frames + (long long)frame * frame_stride;
    const float threshold = thresholds[frame];
    nl_star s = lists[(long long)frame * list_stride + i];
    float shift_sq = 3.40282346638528859811704183484516925440e+38f;
    for (int round = 0; shift_sq > 0.0001f && round < 10; round++) {
        float xm = 0.0f, ym = 0.0f, mass = 0.0f;
        for (int y = -radius; y <= radius; y++) {
            const float fy = (float)y;
            for (int x = -radius; x <= radius; x++) {
                const int index = s.index + y * width + x;
                float value = 0.0f;
                if (index >= 0 && index < len) {
                    value = __ldg(data + index) - threshold;
                    if (value < 0) value = 0;
                }
                xm += (float)x * value;
                ym += fy * value;
                mass += value;
            }
        }
        const int x = s.index % width, y = s.index / width;
        if (mass == 0.0f) mass = 1e-8f;
        const float dx = xm / mass, dy = ym / mass;
        const float nx = (float)x + dx, ny = (float)y + dy;
        const float pdx = nx - s.x, pdy = ny - s.y;
        shift_sq = pdx * pdx + pdy * pdy;
        const int index = (int)((unsigned)s.index + (unsigned)width * (unsigned)go_int32(dy + 0.5f) + (unsigned)go_int32(dx + 0.5f));   // (wraps like Go)
        float value = 0.0f;
        if (index >= 0 && index < len) value = __ldg(data + index);
        s.index = index; s.value = value; s.x = nx; s.y = ny; s.mass = mass; s.hfr = 0.0f;
    }
    lists[(long long)frame * list_stride + i] = s;
    shifts[(long long)frame * list_stride + i] = sqrtf(shift_sq);
}

__global__ void __launch_bounds__(128) star_hfr_kernel(const float *__restrict__ frames, long long frame_stride, int len, int width,
                                                       const float *__restrict__ locations, float radius, float star_in_out,
                                                       nl_star *__restrict__ lists, long long list_stride, const int *__restrict__ counts,
                                                       unsigned char *__restrict__ keep) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[frame]) return;
    const float *data = frames + (long long)frame * frame_stride;
    const float location = locations[frame];
    nl_star s = lists[(long long)frame * list_stride + i];
    float moment = 0.0f, mass = 0.0f;
    int pixels = 0;
    const int rad = (int)ceil((double)radius);
    int lim = (int)ceil((double)(radius + 1e-8f) * (double)(radius + 1e-8f));
    for (int y = -rad; y <= rad; y++)
        for (int x = -rad; x <= rad; x++) {
            const int dsq = x * x + y * y;
            if (dsq > lim) continue;
            const float distance = (float)sqrt((double)dsq);
            const int index = s.index + y * width + x;
            float value = 0.0f;
            if (index >= 0 && index < len) {
                const float v = __ldg(data + index) - location;
                if (v > 0) value = v;
            }
            moment += distance * value;
            mass += value;
            pixels++;
        }
    if (mass == 0.0f) mass = 1e-8f;
    const float hfr = moment / mass;
    bool ok = !(hfr > radius);
    if (ok) {
        float inner_mass = 0.0f;
        int inner_pixels = 0;
        const int irad = (int)ceil((double)hfr);
        lim = (int)ceil((double)(hfr * hfr));
        for (int y = -irad; y <= irad; y++)
            for (int x = -irad; x <= irad; x++) {
                const int dsq = x * x + y * y;
                if (dsq > lim) continue;
                const int index = s.index + y * width + x;
                float value = 0.0f;
                if (index >= 0 && index < len) {
                    const float v = __ldg(data + index) - location;
                    if (v > 0) value = v;
                }
                inner_mass += value;
                inner_pixels++;
            }
        const float outer_mass = mass - inner_mass;
        const int outer_pixels = pixels - inner_pixels;
        if (inner_mass * (float)outer_pixels <= star_in_out * outer_mass * (float)inner_pixels) ok = false;
    }
    if (ok) {
        s.hfr = hfr;
        s.mass = mass;
        lists[(long long)frame * list_stride + i] = s;
    }
    keep[(long long)frame * list_stride + i] = ok ? 1 : 0;
}

int exclusive_scan_launch(nl_ctx *ctx, const int *dev_counts, int *dev_offsets, int n, int *dev_total) {
    row_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(dev_counts, dev_offsets, n, dev_total);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

// Both passes queued back to back with ONE host round trip: the count and the candidates land in mapped pinned host
// memory (the write pass stores them there directly, a few thousand 24-byte records), sized from the previous frames'
// yield; a frame that overflows it repeats the write pass into a larger buffer.
// *count = candidates in the frame; `stars` receives min(count, keep_max) of them in raster order.
static int bright_scan(nl_ctx *ctx, const float *dev_data, int len, int width, float threshold, int radius, int keep_max,
                       std::vector<nl_star> *stars_vec, nl_star *stars_ptr, int *count) {
    *count = 0;
    if (len == 0) return NL_OK;
    const int rows = (len + width - 1) / width;
    const size_t ints = (size_t)2 * rows + 1;
    int rc = ensure_scratch(ctx, (ints * sizeof(int) + 255) & ~(size_t)255);
    if (rc != NL_OK) return rc;
    int *row_count = (int *)ctx->scratch, *row_offset = row_count + rows, *total = row_offset + rows;
    const int threads = 256, warps_per_cta = threads / 32;
    const unsigned grid = (unsigned)((rows + warps_per_cta - 1) / warps_per_cta);
    rc = ensure_pinned(ctx, 64 + sizeof(nl_star) * 16384);
    if (rc != NL_OK) return rc;
    bright_rows_kernel<false><<<grid, threads, 0, ctx->stream>>>(dev_data, len, width, rows, threshold, radius, row_count,
                                                                nullptr, nullptr, 0);
    NL_CUDA(cudaGetLastError());
    row_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(row_count, row_offset, rows, total);
    NL_CUDA(cudaGetLastError());
    ctx->launches += 2;
    for (int attempt = 0; attempt < 2; attempt++) {
        const int slots = (int)std::min<size_t>((ctx->pinned_bytes - 64) / sizeof(nl_star), (size_t)0x7fffffff);
        volatile int *host_total = (volatile int *)ctx->pinned;
        nl_star *host_list = (nl_star *)((char *)ctx->pinned + 64);
        nl_star *dev_list = (nl_star *)((char *)ctx->pinned_dev + 64);
        bright_rows_kernel<true><<<grid, threads, 0, ctx->stream>>>(dev_data, len, width, rows, threshold, radius, row_count,
                                                                   row_offset, dev_list, slots);
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        NL_CUDA(cudaMemcpyAsync((void *)host_total, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
        const int n = *host_total;
        *count = n;
        const int keep = n < keep_max ? n : keep_max;
        if (keep > slots) {                                   // more candidates than the pinned list holds: grow, write again
            rc = ensure_pinned(ctx, 64 + sizeof(nl_star) * ((size_t)keep + (size_t)keep / 2));
            if (rc != NL_OK) return rc;
            continue;
        }
        if (stars_vec) { stars_vec->resize((size_t)(keep > 0 ? keep : 1)); stars_ptr = stars_vec->data(); }
        if (keep > 0) memcpy(stars_ptr, host_list, sizeof(nl_star) * (size_t)keep);
        return NL_OK;
    }
    return set_error(NL_E_CUDA, "star scan: candidate list did not fit after growing");
}

static int find_bright_dev(nl_ctx *ctx, const float *dev_data, int len, int width, float threshold, int radius,
                           nl_star *host_out, int cap, int *count) {
    return bright_scan(ctx, dev_data, len, width, threshold, radius, cap, nullptr, host_out, count);
}

// ---- sparse per-star steps on the host -------------------------------------------------------

// CreateMask, findstars.go:187-200
static std::vector<int32_t> create_mask(int32_t width, float radius) {
    std::vector<int32_t> mask;
    const int32_t rad = (int32_t)radius;
    for (int32_t y = -rad; y <= rad; y++)
        for (int32_t x = -rad; x <= rad; x++) {
            const float dist = (float)sqrt((double)(y * y + x * x));
            if (dist <= radius + 1e-8f) mask.push_back(y * width + x);
        }
    return mask;
}

// rejectBadPixels with medianDiffStats given, findstars.go:134-169.  GatherAndMedian (gather.go:26-38)
// takes the median of the WHOLE 9-entry buffer even when fewer neighbours were in range, so entries
// left over from the previous candidate take part at the image borders; the buffer persists here too.
static int reject_bad_pixels(nl_star *stars, int n, const float *data, int32_t len, int32_t width, float sigma,
                             float median_diff_stddev) {
    const std::vector<int32_t> mask = create_mask(width, 1.5f);
    float buffer[16] = {0};
    const float threshold = median_diff_stddev * sigma;
    int remaining = 0;
    for (int i = 0; i < n; i++) {
        const nl_star s = stars[i];
        int num = 0;
        for (int32_t o : mask) {
            const int32_t io = s.index + o;
            if (io >= 0 && io < len) buffer[num++] = data[io];
        }
        const float med = median9(buffer);            // the mask of radius 1.5 always has 9 entries
        const float diff = data[s.index] - med;
        if (diff < threshold && -diff < threshold) stars[remaining++] = s;
    }
    return remaining;
}

// rejectBadPixels with the test of the interior candidates done on the device (bright_compact_kernel): flag 1 keeps,
// 0 rejects, 2 = a candidate of the first or last row.  Its gather buffer is only partly overwritten, so the verdict
// depends on what the candidate before it left there: replayed from the last candidate that filled the whole buffer.
static int reject_bad_pixels_flagged(nl_star *stars, int n, const unsigned char *flags, const float *data, int32_t len, int32_t width,
                                     float sigma, float median_diff_stddev) {
    const std::vector<int32_t> mask = create_mask(width, 1.5f);
    const float threshold = median_diff_stddev * sigma;
    int remaining = 0;
    std::vector<int32_t> chain;          // the last candidate that filled the whole buffer and the first/last-row candidates since
    for (int i = 0; i < n; i++) {
        const int32_t idx = stars[i].index;
        bool keep = flags[i] == 1;
        if (flags[i] == 2) {
            chain.push_back(idx);
            float buffer[16] = {0};
            for (int32_t c : chain) {
                int num = 0;
                for (int32_t o : mask) {
                    const int32_t io = c + o;
                    if (io >= 0 && io < len) buffer[num++] = data[io];
                }
                const float med = median9(buffer);
                const float diff = data[c] - med;
                keep = diff < threshold && -diff < threshold;
            }
        } else {
            chain.clear();
            chain.push_back(idx);
        }
        if (keep) stars[remaining++] = stars[i];
    }
    return remaining;
}

// pre.MedianFilterSparse, badpixels.go:79-85: the listed pixels are replaced one after the other, in place, by the
// median of their radius-1.5 neighbourhood (a repaired pixel is seen by the repairs after it); sparse, host side
void median_filter_sparse_host(float *data, int32_t len, int32_t width, const int32_t *indices, int64_t n) {
    const std::vector<int32_t> mask = create_mask(width, 1.5f);
    float buffer[16] = {0};
    for (int64_t k = 0; k < n; k++) {
        const int32_t i = indices[k];
        int num = 0;
        for (int32_t o : mask) {
            const int32_t io = i + o;
            if (io >= 0 && io < len) buffer[num++] = data[io];
        }
        data[i] = median9(buffer);
    }
}

// QSortStarsDesc, star/qsort.go:25-55: unstable Hoare quicksort by mass, descending.  The order it leaves equal masses
// in is part of the result (the overlap filter keeps greedily in that order), so it is this algorithm, step for step --
// run on (mass, position) keys of 8 bytes instead of the 32-byte records, which then move once.
struct StarKey { float mass; int pos; };
static void qsort_keys_desc(StarKey *a, int n) {
    while (n > 1) {
        const float pivot = a[(n - 1) >> 1].mass;
        int l = -1, r = n;
        for (;;) {
            do l++; while (a[l].mass > pivot);
            do r--; while (a[r].mass < pivot);
            if (l >= r) break;
            const StarKey t = a[l]; a[l] = a[r]; a[r] = t;
        }
        qsort_keys_desc(a, r + 1);      // left part recursively, right part iteratively
        a += r + 1;
        n -= r + 1;
    }
}
static void qsort_stars_desc(nl_star *a, int n) {
    if (n < 2) return;
    std::vector<StarKey> keys((size_t)n);
    for (int i = 0; i < n; i++) keys[i] = StarKey{a[i].mass, i};
    qsort_keys_desc(keys.data(), n);
    std::vector<nl_star> tmp(a, a + n);
    for (int i = 0; i < n; i++) a[i] = tmp[keys[i].pos];
}

// filterOutOverlaps, findstars.go:209-271: greedy keep in the given order; 256 px bins, each a list
// in insertion order
static int filter_out_overlaps_bins256(nl_star *stars, int n, int32_t width, int32_t height, int32_t radius) {
    const int32_t bin = 256;
    const int32_t xbins = (width + bin - 1) / bin, ybins = (height + bin - 1) / bin;
    std::vector<std::vector<int>> bins((size_t)(xbins > 0 && ybins > 0 ? xbins * ybins : 0));
    const int32_t r2 = radius * radius;
    int kept = 0;
    for (int i = 0; i < n; i++) {
        const nl_star s = stars[i];
        const int32_t xc = (int32_t)(s.x + 0.5f) / bin, yc = (int32_t)(s.y + 0.5f) / bin;
        bool skip = false;
        for (int32_t dy = -1; dy <= 1 && !skip; dy++) {
            if (yc + dy < 0 || yc + dy >= ybins) continue;
            for (int32_t dx = -1; dx <= 1 && !skip; dx++) {
                if (xc + dx < 0 || xc + dx >= xbins) continue;
                for (int j : bins[(size_t)((xc + dx) + (yc + dy) * xbins)]) {
                    const float xd = s.x - stars[j].x, yd = s.y - stars[j].y;
                    const int32_t sq = (int32_t)(xd * xd + yd * yd + 0.5f);
                    if (sq <= r2) { skip = true; break; }
                }
            }
        }
        if (skip) continue;
        stars[kept] = s;
        // the reference indexes its bin table unguarded here (findstars.go:255) and would panic for a
        // star whose centre left the image; such a star is kept but not binned
        if (xc >= 0 && xc < xbins && yc >= 0 && yc < ybins) bins[(size_t)(xc + yc * xbins)].push_back(kept);
        kept++;
    }
    return kept;
}

// The same filter on a finer grid.  A candidate is dropped when a star kept (and binned) before it has
// int32(dx^2 + dy^2 + 0.5) <= radius^2, i.e. lies closer than radius + 1 in x and in y; the reference finds such a star
// because it is in the same or an adjacent 256-pixel bin.  With cells of at least radius + 1 pixels it is just as
// surely in the same or an adjacent cell, so scanning 3 x 3 cells gives the same verdict for every candidate -- over a
// handful of stars instead of the hundreds a 256-pixel bin of a dense candidate list holds (the scan of a frame's
// ~20 000 candidates was the largest host step of the batched star detection).  Which star matches does not matter,
// only whether one does; a star the reference does not bin (centre outside the image) is not binned here either.
static int filter_out_overlaps(nl_star *stars, int n, int32_t width, int32_t height, int32_t radius) {
    const int32_t bin = 256;
    if (radius < 0 || radius + 1 > bin / 2 || width <= 0 || height <= 0) return filter_out_overlaps_bins256(stars, n, width, height, radius);
    const int32_t xbins = (width + bin - 1) / bin, ybins = (height + bin - 1) / bin;
    const int32_t cs = radius + 1 > 16 ? radius + 1 : 16;                         // cell size in pixels
    // a binned star has int32(x + 0.5) in (-256, xbins*256): shifted by 256 its cell index is >= 0
    const int32_t xcells = (xbins * bin + bin) / cs + 1, ycells = (ybins * bin + bin) / cs + 1;
    std::vector<int> head((size_t)xcells * ycells, -1), next((size_t)n, -1);
    auto floordiv = [](long long a, long long b) { return (a >= 0 ? a / b : -((-a + b - 1) / b)); };
    const int32_t r2 = radius * radius;
    int kept = 0;
    for (int i = 0; i < n; i++) {
        const nl_star s = stars[i];
        const int32_t xi = (int32_t)(s.x + 0.5f), yi = (int32_t)(s.y + 0.5f);
        const long long cx = floordiv((long long)xi + bin, cs), cy = floordiv((long long)yi + bin, cs);
        bool skip = false;
        for (long long yy = cy - 1; yy <= cy + 1 && !skip; yy++) {
            if (yy < 0 || yy >= ycells) continue;
            for (long long xx = cx - 1; xx <= cx + 1 && !skip; xx++) {
                if (xx < 0 || xx >= xcells) continue;
                for (int j = head[(size_t)(xx + yy * xcells)]; j >= 0; j = next[j]) {
                    const float xd = s.x - stars[j].x, yd = s.y - stars[j].y;
                    const int32_t sq = (int32_t)(xd * xd + yd * yd + 0.5f);
                    if (sq <= r2) { skip = true; break; }
                }
            }
        }
        if (skip) continue;
        stars[kept] = s;
        const int32_t xc = xi / bin, yc = yi / bin;                                 // the reference's bin (findstars.go:255)
        if (xc >= 0 && xc < xbins && yc >= 0 && yc < ybins) {
            const size_t cell = (size_t)(cx + cy * xcells);
            next[kept] = head[cell];
            head[cell] = kept;
        }
        kept++;
    }
    return kept;
}

// shiftToCenterOfMass, findstars.go:274-322
static float shift_to_center_of_mass(nl_star *stars, int n, const float *data, int32_t len, int32_t width,
                                     float threshold, int32_t radius) {
    float sum_of_shifts = 0.0f;
    for (int i = 0; i < n; i++) {
        nl_star s = stars[i];
        float shift_sq = 3.40282346638528859811704183484516925440e+38f;
        for (int32_t round = 0; shift_sq > 0.0001f && round < 10; round++) {
            float xm = 0.0f, ym = 0.0f, mass = 0.0f;
            for (int32_t y = -radius; y <= radius; y++)
                for (int32_t x = -radius; x <= radius; x++) {
                    const int32_t index = s.index + y * width + x;
                    float value = 0.0f;
                    if (index >= 0 && index < len) {
                        value = data[index] - threshold;
                        if (value < 0) value = 0;
                    }
                    xm += (float)x * value;
                    ym += (float)y * value;
                    mass += value;
                }
            const int32_t x = s.index % width, y = s.index / width;
            if (mass == 0.0f) mass = 1e-8f;
            const float dx = xm / mass, dy = ym / mass;
            const float nx = (float)x + dx, ny = (float)y + dy;
            const float pdx = nx - s.x, pdy = ny - s.y;
            shift_sq = pdx * pdx + pdy * pdy;
            const int32_t index = s.index + width * (int32_t)(dy + 0.5f) + (int32_t)(dx + 0.5f);
            float value = 0.0f;
            if (index >= 0 && index < len) value = data[index];
            s.index = index; s.value = value; s.x = nx; s.y = ny; s.mass = mass; s.hfr = 0.0f;
            stars[i] = s;
        }
        sum_of_shifts += sqrtf(shift_sq);
    }
    return sum_of_shifts;
}

// calcAndFilterHalfFluxRadius, findstars.go:327-396
static int calc_and_filter_hfr(nl_star *stars, int n, const float *data, int32_t len, int32_t width, float radius,
                               float location, float star_in_out, float *avg_hfr) {
    int remaining = 0;
    float avg = 0.0f;
    for (int i = 0; i < n; i++) {
        nl_star s = stars[i];
        float moment = 0.0f, mass = 0.0f;
        int32_t pixels = 0;
        const int32_t rad = (int32_t)ceil((double)radius);
        int32_t lim = (int32_t)ceil((double)(radius + 1e-8f) * (double)(radius + 1e-8f));
        for (int32_t y = -rad; y <= rad; y++)
            for (int32_t x = -rad; x <= rad; x++) {
                const int32_t dsq = x * x + y * y;
                if (dsq > lim) continue;
                const float distance = (float)sqrt((double)dsq);
                const int32_t index = s.index + y * width + x;
                float value = 0.0f;
                if (index >= 0 && index < len) {
                    const float v = data[index] - location;
                    if (v > 0) value = v;
                }
                moment += distance * value;
                mass += value;
                pixels++;
            }
        if (mass == 0.0f) mass = 1e-8f;
        const float hfr = moment / mass;
        if (hfr > radius) continue;
        float inner_mass = 0.0f;
        int32_t inner_pixels = 0;
        const int32_t irad = (int32_t)ceil((double)hfr);
        lim = (int32_t)ceil((double)(hfr * hfr));
        for (int32_t y = -irad; y <= irad; y++)
            for (int32_t x = -irad; x <= irad; x++) {
                const int32_t dsq = x * x + y * y;
                if (dsq > lim) continue;
                const int32_t index = s.index + y * width + x;
                float value = 0.0f;
                if (index >= 0 && index < len) {
                    const float v = data[index] - location;
                    if (v > 0) value = v;
                }
                inner_mass += value;
                inner_pixels++;
            }
        const float outer_mass = mass - inner_mass;
        const int32_t outer_pixels = pixels - inner_pixels;
        if (inner_mass * (float)outer_pixels <= star_in_out * outer_mass * (float)inner_pixels) continue;
        s.hfr = hfr;
        s.mass = mass;
        stars[remaining++] = s;
        avg += hfr;
    }
    avg /= (float)remaining;
    *avg_hfr = avg;
    return remaining;
}

// the sparse per-star steps of FindStars (findstars.go:66-99) on `n` raster-ordered candidates; returns the star count
static int find_stars_sparse(nl_star *stars, int n, const float *host_data, int32_t len, int32_t width, float location, float scale,
                             float star_sig, float bp_sigma, float star_in_out, int32_t radius, float median_diff_stddev, nl_star *out,
                             int32_t cap, float *sum_of_shifts, float *avg_hfr) {
    int m = n;
    *sum_of_shifts = 0.0f; *avg_hfr = 0.0f;
    if (bp_sigma > 0) m = reject_bad_pixels(stars, m, host_data, len, width, bp_sigma, median_diff_stddev);
    qsort_stars_desc(stars, m);
    m = filter_out_overlaps(stars, m, width, len / width, radius);
    *sum_of_shifts = shift_to_center_of_mass(stars, m, host_data, len, width, location + scale * star_sig * 0.5f, radius);
    qsort_stars_desc(stars, m);
    m = filter_out_overlaps(stars, m, width, len / width, radius);
    m = calc_and_filter_hfr(stars, m, host_data, len, width, (float)radius, location, star_in_out, avg_hfr);
    for (int i = 0; i < m && i < cap; i++) out[i] = stars[i];
    return m;
}

}  // namespace nl

using namespace nl;

extern "C" {

int nl_find_bright_dev(nl_ctx *ctx, const float *dev_data, int32_t len, int32_t width, float threshold, int32_t radius,
                       nl_star *out, int32_t cap, int32_t *count) {
    NL_REQUIRE(ctx && count, "NULL argument");
    NL_REQUIRE(len >= 0 && width > 0 && cap >= 0, "bad size");
    NL_REQUIRE(out || cap == 0, "out is NULL");
    NL_REQUIRE(dev_data || len == 0, "data is NULL");
    NL_GUARD(ctx);
    return find_bright_dev(ctx, dev_data, len, width, threshold, radius, out, cap, count);
}

int nl_find_bright(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float threshold, int32_t radius,
                   nl_star *out, int32_t cap, int32_t *count) {
    NL_REQUIRE(ctx && count, "NULL argument");
    NL_REQUIRE(len >= 0 && width > 0 && cap >= 0, "bad size");
    NL_REQUIRE(out || cap == 0, "out is NULL");
    NL_REQUIRE(host_data || len == 0, "data is NULL");
    *count = 0;
    if (len == 0) return NL_OK;
    NL_GUARD(ctx);
    float *dev = nullptr;
    int rc = ensure_frame(ctx, 0, sizeof(float) * (size_t)len, &dev);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(dev, host_data, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    return find_bright_dev(ctx, dev, len, width, threshold, radius, out, cap, count);
}

// FindStars with the frame already on the device (dev_data) and in host memory (host_data, the same pixels): the
// full-frame scan reads the device copy, the sparse per-star steps the host copy
int nl_find_stars_dev(nl_ctx *ctx, const float *dev_data, const float *host_data, int32_t len, int32_t width, float location,
                      float scale, float star_sig, float bp_sigma, float star_in_out, int32_t radius, float median_diff_stddev,
                      nl_star *out, int32_t cap, int32_t *count, float *sum_of_shifts, float *avg_hfr) {
    NL_REQUIRE(ctx && count && sum_of_shifts && avg_hfr, "NULL argument");
    NL_REQUIRE(len >= 0 && width > 0 && cap >= 0, "bad size");
    NL_REQUIRE(out || cap == 0, "out is NULL");
    // findstars.go:61: threshold = location + scale*starSig
    const float threshold = location + scale * star_sig;
    int32_t n = 0;
    *count = 0; *sum_of_shifts = 0.0f; *avg_hfr = 0.0f;
    NL_REQUIRE((host_data && dev_data) || len == 0, "data is NULL");
    NL_GUARD(ctx);
    std::vector<nl_star> stars(1);
    if (len > 0) {
        int rc = bright_scan(ctx, dev_data, len, width, threshold, radius, 0x7fffffff, &stars, nullptr, &n);
        if (rc != NL_OK) return rc;
    }
    *count = find_stars_sparse(stars.data(), n, host_data, len, width, location, scale, star_sig, bp_sigma, star_in_out, radius,
                               median_diff_stddev, out, cap, sum_of_shifts, avg_hfr);
    return NL_OK;
}

}  // extern "C"

namespace nl {

// findBrightPixels of ALL frames of a resident stack (frame i at dev_frames + i*frame_stride) with one read of every
// frame and two host round trips in total: per-row slots, scan of the row counts, compaction into raster order straight
// into mapped pinned host memory.  On return frame i's first min(counts[i], keep_cap) candidates are at *list + i * *stride
// (host memory owned by the context, valid until the next batched scan).
// bp_thresholds (optional, per frame): the compaction also runs rejectBadPixels' test; *flags = its result per candidate
// ([frame][stride], see bright_compact_kernel), flags_valid[i] = 0 for a frame that was redone by the two-pass scan.
static int bright_scan_batch(nl_ctx *ctx, const float *dev_frames, int n_frames, long long frame_stride, int len, int width,
                             const float *thresholds, int radius, int keep_cap, nl_star **list, long long *stride, int32_t *counts,
                             const float *bp_thresholds = nullptr, unsigned char **flags = nullptr, std::vector<char> *flags_valid = nullptr) {
    const int rows = (len + width - 1) / width;
    const size_t fr = (size_t)n_frames * rows;
    // scratch: row counts | row offsets | thresholds | totals | overflow ; ctx->list: the row slots
    const size_t ints = 2 * fr + 4 * (size_t)n_frames;
    int rc = ensure_scratch(ctx, (ints * sizeof(int) + 255) & ~(size_t)255);
    if (rc != NL_OK) return rc;
    int *row_count = (int *)ctx->scratch, *row_offset = row_count + fr;
    float *dthr = (float *)(row_offset + fr);
    int *totals = (int *)(dthr + n_frames), *overflow = totals + n_frames;
    float *dbp = (float *)(overflow + n_frames);
    const size_t slot_bytes = fr * BRIGHT_SLOTS * sizeof(nl_star);
    if (ctx->list_bytes < slot_bytes) {
        if (ctx->list) { NL_CUDA(cudaStreamSynchronize(ctx->stream)); NL_CUDA(cudaFree(ctx->list)); ctx->list = nullptr; ctx->list_bytes = 0; }
        NL_CUDA(cudaMalloc(&ctx->list, slot_bytes));
        ctx->list_bytes = slot_bytes;
    }
    nl_star *slots = (nl_star *)ctx->list;
    auto ensure_batch_pinned = [&](size_t bytes) -> int {
        if (ctx->batch_pinned_bytes >= bytes) return NL_OK;
        if (ctx->batch_pinned) { NL_CUDA(cudaStreamSynchronize(ctx->stream)); NL_CUDA(cudaFreeHost(ctx->batch_pinned)); ctx->batch_pinned = nullptr; ctx->batch_pinned_bytes = 0; }
        NL_CUDA(cudaHostAlloc(&ctx->batch_pinned, bytes, cudaHostAllocMapped));
        NL_CUDA(cudaHostGetDevicePointer(&ctx->batch_pinned_dev, ctx->batch_pinned, 0));
        ctx->batch_pinned_bytes = bytes;
        return NL_OK;
    };
    const size_t head = (2 * sizeof(int) * (size_t)n_frames + 2 * sizeof(float) * (size_t)n_frames + 255) & ~(size_t)255;   // totals | overflow | thresholds | bad-pixel thresholds
    const size_t per_cand = sizeof(nl_star) + 1;                      // a record and its flag
    rc = ensure_batch_pinned(head + per_cand * (size_t)n_frames * 4096 + 256);
    if (rc != NL_OK) return rc;
    float *h_thr = (float *)((char *)ctx->batch_pinned + 2 * sizeof(int) * (size_t)n_frames);
    memcpy(h_thr, thresholds, sizeof(float) * (size_t)n_frames);
    if (bp_thresholds) memcpy(h_thr + n_frames, bp_thresholds, sizeof(float) * (size_t)n_frames);
    NL_CUDA(cudaMemcpyAsync(dthr, h_thr, sizeof(float) * (size_t)n_frames, cudaMemcpyHostToDevice, ctx->stream));
    if (bp_thresholds) NL_CUDA(cudaMemcpyAsync(dbp, h_thr + n_frames, sizeof(float) * (size_t)n_frames, cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256, wpc = threads / 32;
    dim3 grid((unsigned)((rows + wpc - 1) / wpc), (unsigned)n_frames);
    bright_rows_slots_kernel<<<grid, threads, 0, ctx->stream>>>(dev_frames, frame_stride, len, width, rows, dthr, radius, row_count, slots);
    NL_CUDA(cudaGetLastError());
    row_offsets_batch_kernel<<<(unsigned)n_frames, 1024, 0, ctx->stream>>>(row_count, row_offset, rows, totals, overflow);
    NL_CUDA(cudaGetLastError());
    ctx->launches += 2;
    NL_CUDA(cudaMemcpyAsync(ctx->batch_pinned, totals, 2 * sizeof(int) * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));                       // round trip 1: how many candidates per frame
    std::vector<int> host_tot(2 * (size_t)n_frames);
    memcpy(host_tot.data(), ctx->batch_pinned, 2 * sizeof(int) * (size_t)n_frames);
    int max_keep = 1;
    for (int i = 0; i < n_frames; i++) {
        counts[i] = host_tot[i];
        const int keep = host_tot[i] < keep_cap ? host_tot[i] : keep_cap;
        if (keep > max_keep) max_keep = keep;
    }
    rc = ensure_batch_pinned(head + per_cand * (size_t)n_frames * max_keep + 256);
    if (rc != NL_OK) return rc;
    nl_star *host_list = (nl_star *)((char *)ctx->batch_pinned + head);
    nl_star *dev_list = (nl_star *)((char *)ctx->batch_pinned_dev + head);
    const size_t flags_off = head + sizeof(nl_star) * (size_t)n_frames * max_keep;
    bright_compact_kernel<<<grid, threads, 0, ctx->stream>>>(slots, row_count, row_offset, rows, dev_list, max_keep, dev_frames, frame_stride,
                                                             len, width, bp_thresholds ? dbp : nullptr,
                                                             (unsigned char *)ctx->batch_pinned_dev + flags_off);
    if (flags) *flags = (unsigned char *)ctx->batch_pinned + flags_off;
    if (flags_valid) flags_valid->assign((size_t)n_frames, bp_thresholds ? 1 : 0);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    NL_CUDA(cudaStreamSynchronize(ctx->stream));                       // round trip 2: the candidates have landed in host memory
    // frames with a row beyond its slots (dense star fields, hot columns): the two-pass scan of that frame
    for (int i = 0; i < n_frames; i++) {
        if (host_tot[n_frames + i] == 0) continue;
        int n = 0;
        rc = bright_scan(ctx, dev_frames + (size_t)i * frame_stride, len, width, thresholds[i], radius, max_keep, nullptr,
                         host_list + (size_t)i * max_keep, &n);
        if (rc != NL_OK) return rc;
        if (n > max_keep && max_keep < keep_cap) {
            // (the recount found more than the slots let the first pass see: grow the lists and redo the whole batch)
            return bright_scan_batch(ctx, dev_frames, n_frames, frame_stride, len, width, thresholds, radius, keep_cap, list, stride, counts,
                                     bp_thresholds, flags, flags_valid);
        }
        counts[i] = n;
        if (flags_valid) (*flags_valid)[i] = 0;
    }
    *list = host_list;
    *stride = max_keep;
    return NL_OK;
}

}  // namespace nl

extern "C" {

// ---- the sparse host steps on their own (no device) ------------------------------------------------------------------
int nl_star_reject_bad_pixels_host(nl_star *stars, int32_t n, const float *data, int32_t len, int32_t width, float sigma,
                                   float median_diff_stddev, int32_t *kept) {
    NL_REQUIRE(kept && n >= 0 && len >= 0 && width > 0 && (stars || n == 0) && (data || len == 0), "bad argument");
    // the flags bright_compact_kernel writes: every candidate with its whole neighbourhood inside the frame on its own
    std::vector<unsigned char> flags((size_t)n);
    const float thr = median_diff_stddev * sigma;
    for (int i = 0; i < n; i++) {
        const long long idx = stars[i].index;
        unsigned char flag = 2;
        if (idx - width - 1 >= 0 && idx + width + 1 < len) {
            float b[9];
            for (int y = -1; y <= 1; y++)
                for (int x = -1; x <= 1; x++) b[(y + 1) * 3 + (x + 1)] = data[idx + (long long)y * width + x];
            const float med = median9(b);
            const float diff = data[idx] - med;
            flag = (diff < thr && -diff < thr) ? 1 : 0;
        }
        flags[i] = flag;
    }
    *kept = reject_bad_pixels_flagged(stars, n, flags.data(), data, len, width, sigma, median_diff_stddev);
    return NL_OK;
}

int nl_star_sort_desc_host(nl_star *stars, int32_t n) {
    NL_REQUIRE(n >= 0 && (stars || n == 0), "bad argument");
    qsort_stars_desc(stars, n);
    return NL_OK;
}

int nl_star_filter_overlaps_host(nl_star *stars, int32_t n, int32_t width, int32_t height, int32_t radius, int32_t *kept) {
    NL_REQUIRE(kept && n >= 0 && (stars || n == 0), "bad argument");
    *kept = filter_out_overlaps(stars, n, width, height, radius);
    return NL_OK;
}

// findBrightPixels of all frames of a resident stack; host_out receives frame i's candidates at host_out + i*cap (the
// first min(count, cap)); counts[i] = candidates found in frame i.
int nl_find_bright_batch_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, int32_t len, int32_t width,
                             const float *thresholds, int32_t radius, nl_star *host_out, int32_t cap, int32_t *counts) {
    NL_REQUIRE(ctx && counts && thresholds && n_frames >= 0, "bad argument");
    NL_REQUIRE(len >= 0 && width > 0 && cap >= 0, "bad size");
    NL_REQUIRE(host_out || cap == 0, "out is NULL");
    NL_REQUIRE(dev_frames || len == 0, "data is NULL");
    for (int i = 0; i < n_frames; i++) counts[i] = 0;
    if (len == 0 || n_frames == 0) return NL_OK;
    NL_REQUIRE(n_frames <= 65535, "more than 65535 frames in one batch");
    NL_GUARD(ctx);
    nl_star *list = nullptr;
    long long stride = 0;
    int rc = bright_scan_batch(ctx, dev_frames, n_frames, frame_stride, len, width, thresholds, radius, cap, &list, &stride, counts);
    if (rc != NL_OK) return rc;
    for (int i = 0; i < n_frames; i++) {
        const int keep = counts[i] < cap ? counts[i] : cap;
        if (keep > 0) memcpy(host_out + (size_t)i * cap, list + (size_t)i * stride, sizeof(nl_star) * (size_t)keep);
    }
    return NL_OK;
}

// FindStars over all frames of a resident stack.  Device: the batched scan, the centre-of-mass iterations and the
// half-flux radius (one thread per star, all frames per launch).  Host threads (one frame each): the order-dependent
// sparse steps between them -- bad-pixel rejection, the unstable quicksort by mass, the greedy overlap filter, the
// sequential sums -- working in place on the candidate lists in mapped pinned memory; host_frames[i] are the host
// copies of the frames (the bad-pixel rejection gathers from them).  Per-frame inputs and outputs are arrays of n_frames
// entries; out receives frame i's stars at out + i*cap.
int nl_find_stars_batch_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, const float *const *host_frames,
                            int32_t len, int32_t width, const float *location, const float *scale, float star_sig, float bp_sigma,
                            float star_in_out, int32_t radius, const float *median_diff_stddev, nl_star *out, int32_t cap,
                            int32_t *counts, float *sum_of_shifts, float *avg_hfr, double *seconds_device, double *seconds_host) {
    NL_REQUIRE(ctx && counts && sum_of_shifts && avg_hfr && location && scale && host_frames, "NULL argument");
    NL_REQUIRE(n_frames >= 0 && len >= 0 && width > 0 && cap >= 0 && radius >= 0, "bad size");
    NL_REQUIRE(n_frames <= 65535, "more than 65535 frames in one batch");
    NL_REQUIRE(bp_sigma <= 0 || median_diff_stddev, "median_diff_stddev is needed when bp_sigma > 0");
    for (int i = 0; i < n_frames; i++) { counts[i] = 0; sum_of_shifts[i] = 0.0f; avg_hfr[i] = 0.0f; }
    if (seconds_device) *seconds_device = 0.0;
    if (seconds_host) *seconds_host = 0.0;
    if (n_frames == 0 || len == 0) return NL_OK;
    NL_REQUIRE(dev_frames, "data is NULL");
    NL_GUARD(ctx);
    typedef std::chrono::steady_clock clk;
    double t_dev = 0.0, t_host = 0.0;
    auto since = [](clk::time_point t) { return std::chrono::duration<double>(clk::now() - t).count(); };
    unsigned hw = std::thread::hardware_concurrency();
    const int n_threads = (int)std::max(1u, std::min(hw ? hw : 1u, (unsigned)n_frames));
    auto each_frame = [&](const std::function<void(int)> &f) {
        std::atomic<int> next{0};
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&]() { for (int i = next++; i < n_frames; i = next++) f(i); });
        for (auto &t : th) t.join();
    };

    // ---- device: the scan.  Every candidate is needed: the filters follow.
    auto t0 = clk::now();
    std::vector<float> thr((size_t)n_frames);
    for (int i = 0; i < n_frames; i++) thr[i] = location[i] + scale[i] * star_sig;            // findstars.go:61
    std::vector<int32_t> m((size_t)n_frames);
    nl_star *list = nullptr;
    long long stride = 0;
    std::vector<float> bp_thr;
    if (bp_sigma > 0) {
        bp_thr.resize((size_t)n_frames);
        for (int i = 0; i < n_frames; i++) bp_thr[i] = median_diff_stddev[i] * bp_sigma;         // findstars.go:137
    }
    unsigned char *flags = nullptr;
    std::vector<char> flags_valid;
    int rc = bright_scan_batch(ctx, dev_frames, n_frames, frame_stride, len, width, thr.data(), radius, 0x7fffffff, &list, &stride, m.data(),
                               bp_sigma > 0 ? bp_thr.data() : nullptr, &flags, &flags_valid);
    if (rc != NL_OK) return rc;
    nl_star *dev_list = (nl_star *)((char *)ctx->batch_pinned_dev + ((char *)list - (char *)ctx->batch_pinned));
    t_dev += since(t0);

    // ---- host: rejectBadPixels, QSortStarsDesc, filterOutOverlaps (findstars.go:66-74)
    t0 = clk::now();
    each_frame([&](int i) {
        nl_star *st = list + (size_t)i * stride;
        int k = m[i];
        if (bp_sigma > 0) {
            if (flags_valid[i]) k = reject_bad_pixels_flagged(st, k, flags + (size_t)i * stride, host_frames[i], len, width, bp_sigma, median_diff_stddev[i]);
            else k = reject_bad_pixels(st, k, host_frames[i], len, width, bp_sigma, median_diff_stddev[i]);
        }
        qsort_stars_desc(st, k);
        m[i] = filter_out_overlaps(st, k, width, len / width, radius);
    });
    t_host += since(t0);
    if (ctx->stats_debug) fprintf(stderr, "find_stars_batch: %d host threads, reject+sort+overlaps %.3f ms\n", n_threads, since(t0) * 1e3);

    // ---- device: shiftToCenterOfMass.  Per-frame scalars and per-star outputs in a second pinned block.
    t0 = clk::now();
    const size_t per_frame = ((sizeof(int) + 2 * sizeof(float)) * (size_t)n_frames + 255) & ~(size_t)255;
    const size_t aux_bytes = per_frame + (sizeof(float) + 1) * (size_t)n_frames * (size_t)stride + 256;
    rc = ensure_pinned(ctx, aux_bytes);
    if (rc != NL_OK) return rc;
    int *h_counts = (int *)ctx->pinned;
    float *h_thr2 = (float *)(h_counts + n_frames), *h_loc = h_thr2 + n_frames;
    float *h_shift = (float *)((char *)ctx->pinned + per_frame);
    unsigned char *h_keep = (unsigned char *)(h_shift + (size_t)n_frames * stride);
    char *d_aux = (char *)ctx->pinned_dev;
    int max_m = 1;
    for (int i = 0; i < n_frames; i++) {
        h_counts[i] = m[i];
        h_thr2[i] = location[i] + scale[i] * star_sig * 0.5f;                                   // findstars.go:77
        h_loc[i] = location[i];
        if (m[i] > max_m) max_m = m[i];
    }
    {
        dim3 grid((unsigned)((max_m + 127) / 128), (unsigned)n_frames);
        star_com_kernel<<<grid, 128, 0, ctx->stream>>>(dev_frames, frame_stride, len, width, (const float *)(d_aux + ((char *)h_thr2 - (char *)ctx->pinned)),
                                                       radius, dev_list, stride, (const int *)d_aux,
                                                       (float *)(d_aux + ((char *)h_shift - (char *)ctx->pinned)));
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    t_dev += since(t0);

    // ---- host: the sum of the shifts in star order, then sort and overlap filter again (findstars.go:78-84)
    t0 = clk::now();
    each_frame([&](int i) {
        nl_star *st = list + (size_t)i * stride;
        float sos = 0.0f;
        for (int k = 0; k < m[i]; k++) sos += h_shift[(size_t)i * stride + k];
        sum_of_shifts[i] = sos;
        qsort_stars_desc(st, m[i]);
        m[i] = filter_out_overlaps(st, m[i], width, len / width, radius);
    });
    t_host += since(t0);
    if (ctx->stats_debug) fprintf(stderr, "find_stars_batch: shifts+sort+overlaps %.3f ms\n", since(t0) * 1e3);

    // ---- device: calcAndFilterHalfFluxRadius
    t0 = clk::now();
    max_m = 1;
    for (int i = 0; i < n_frames; i++) { h_counts[i] = m[i]; if (m[i] > max_m) max_m = m[i]; }
    {
        dim3 grid((unsigned)((max_m + 127) / 128), (unsigned)n_frames);
        star_hfr_kernel<<<grid, 128, 0, ctx->stream>>>(dev_frames, frame_stride, len, width, (const float *)(d_aux + ((char *)h_loc - (char *)ctx->pinned)),
                                                       (float)radius, star_in_out, dev_list, stride, (const int *)d_aux,
                                                       (unsigned char *)(d_aux + ((char *)h_keep - (char *)ctx->pinned)));
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    t_dev += since(t0);

    // ---- host: the survivors in order, their average radius (findstars.go:384-395), the caller's array
    t0 = clk::now();
    each_frame([&](int i) {
        nl_star *st = list + (size_t)i * stride;
        int remaining = 0;
        float avg = 0.0f;
        for (int k = 0; k < m[i]; k++) {
            if (!h_keep[(size_t)i * stride + k]) continue;
            st[remaining++] = st[k];
            avg += st[remaining - 1].hfr;
        }
        avg /= (float)remaining;
        avg_hfr[i] = avg;
        counts[i] = remaining;
        for (int k = 0; k < remaining && k < cap; k++) out[(size_t)i * cap + k] = st[k];
    });
    t_host += since(t0);
    if (ctx->stats_debug) fprintf(stderr, "find_stars_batch: survivors %.3f ms\n", since(t0) * 1e3);
    if (seconds_device) *seconds_device = t_dev;
    if (seconds_host) *seconds_host = t_host;
    return NL_OK;
}

int nl_find_stars(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float location, float scale,
                  float star_sig, float bp_sigma, float star_in_out, int32_t radius, float median_diff_stddev,
                  nl_star *out, int32_t cap, int32_t *count, float *sum_of_shifts, float *avg_hfr) {
    NL_REQUIRE(ctx && len >= 0 && (host_data || len == 0), "bad argument");
    NL_GUARD(ctx);
    float *dev = nullptr;
    if (len > 0) {
        int rc = ensure_frame(ctx, 0, sizeof(float) * (size_t)len, &dev);
        if (rc != NL_OK) return rc;
        NL_CUDA(cudaMemcpyAsync(dev, host_data, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    }
    return nl_find_stars_dev(ctx, dev, host_data, len, width, location, scale, star_sig, bp_sigma, star_in_out, radius,
                             median_diff_stddev, out, cap, count, sum_of_shifts, avg_hfr);
}

}  // extern "C"
