// nl_prestats.cu -- the frame statistics that feed the path:
//   median.MedianFilter3x3        internal/median/median3x3.go:26-110, median3x3_amd64.go:24-48, median3x3_amd64.s
//   Stats.Min/Mean/Max/StdDev     internal/stats/stats.go:102-153 -> calcMinMeanMax / calcVariance
//                                 (stats.go:264-287 pure Go; stats_amd64.s:27-143 AVX2)
//   pre.BadPixelMap               internal/ops/pre/badpixels.go:32-51 (its medianDiffStats.StdDev() is the
//                                 bad-pixel threshold of star detection, findstars.go:134-169)
//
// Two numerics, selected per context (nl_ctx_set_numerics): the reference as built for amd64 runs the
// AVX2 assembly whenever the CPU has AVX2; every other build runs the pure-Go loops.  They differ in
//  * the operand roles of min/max (VMINPS/VMAXPS return their second source on NaN and on +-0 ties;
//    the Go code compares and keeps), and
//  * the float64 summation: four interleaved lanes (element i -> lane i%4) folded (0+1)+(2+3), or one
//    chain.  Each lane is a sequential chain of up to millions of rounded float64 additions.
//
// A sequentially rounded chain cannot be reproduced by a parallel sum in general, but its float32
// consumer can: the kernels compute every lane sum in parallel together with a rigorous bound on how
// far the sequentially rounded chain can be from it (u * sum_k |partial sum_k|, bounded per 1024-element
// record from the record's prefix and its sum of magnitudes).  mean = float32(sum/n) and
// stddev = float32(sqrt(sum/n)) are monotone in the sum, so when both ends of the interval round to
// the same float32 the result is proven bit-identical to the reference's.  Otherwise (about 1 frame in
// 10^3) the lanes are replayed exactly, in order, by `stats_exact_kernel` (one thread per lane).
//
// Algorithmic bytes: median filter 8 B/pixel (4 read, 4 written); stats 4 B/pixel per pass (two passes:
// the variance needs the rounded mean); bad-pixel scan 4 B/pixel per pass (count, write).
#include "nl_internal.h"

#include <float.h>
#include <math.h>
#include <string.h>

namespace nl {

// ---- median of nine ------------------------------------------------------------------------

// Go-assembler `VMINPS src2, src1, dst`: dst = src1 < src2 ? src1 : src2
__device__ __forceinline__ float minps(float src1, float src2) { return src1 < src2 ? src1 : src2; }
__device__ __forceinline__ float maxps(float src1, float src2) { return src1 > src2 ? src1 : src2; }

// exchange / max-into-j / min-into-i steps of the network, in either numerics
template <bool AMD64> __device__ __forceinline__ void net_s(float &ai, float &aj) {
    if (AMD64) { const float lo = minps(aj, ai), hi = maxps(aj, ai); ai = lo; aj = hi; }
    else if (ai > aj) { const float t = ai; ai = aj; aj = t; }
}
template <bool AMD64> __device__ __forceinline__ void net_x(float ai, float &aj) {
    if (AMD64) aj = maxps(aj, ai); else if (ai > aj) aj = ai;
}
template <bool AMD64> __device__ __forceinline__ void net_n(float &ai, float aj) {
    if (AMD64) ai = minps(aj, ai); else if (ai > aj) ai = aj;
}

// median3x3.go:85-110 / median3x3_amd64.s:124-213, a[] in row-major gather order
template <bool AMD64> __device__ __forceinline__ float median9(float a0, float a1, float a2, float a3, float a4, float a5,
                                                               float a6, float a7, float a8) {
    net_s<AMD64>(a0, a1); net_s<AMD64>(a3, a4); net_s<AMD64>(a6, a7);
    net_s<AMD64>(a1, a2); net_s<AMD64>(a4, a5); net_s<AMD64>(a7, a8);
    net_s<AMD64>(a0, a1); net_s<AMD64>(a3, a4); net_s<AMD64>(a6, a7);
    net_x<AMD64>(a0, a3); net_x<AMD64>(a3, a6);
    net_s<AMD64>(a1, a4);
    net_n<AMD64>(a4, a7); net_x<AMD64>(a1, a4);
    net_n<AMD64>(a5, a8); net_n<AMD64>(a2, a5);
    net_s<AMD64>(a2, a4);
    net_n<AMD64>(a4, a6); net_x<AMD64>(a2, a4);
    return a4;
}

// One column x four rows per thread with the three-column window of six rows in registers; lanes
// cover consecutive columns, so loads and stores are full 128-byte lines and the neighbouring
// columns come from L1.  DIFF: writes data - median (badpixels.go:34-35) instead of the median.
template <bool AMD64, bool DIFF>
__global__ void __launch_bounds__(256) median3x3_kernel(const float *__restrict__ data, int w, int h, float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * 4;
    if (x >= w || y0 >= h) return;
    const int xl = x > 0 ? x - 1 : x, xr = x < w - 1 ? x + 1 : x;
    float l[6], c[6], r[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int y = y0 - 1 + k;
        y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
        const float *row = data + (size_t)y * w;
        l[k] = __ldg(row + xl); c[k] = __ldg(row + x); r[k] = __ldg(row + xr);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = y0 + k;
        if (y >= h) break;
        float m = c[k + 1];                                       // border rows and columns are copied
        if (x > 0 && x < w - 1 && y > 0 && y < h - 1)
            m = median9<AMD64>(l[k], c[k], r[k], l[k + 1], c[k + 1], r[k + 1], l[k + 2], c[k + 2], r[k + 2]);
        out[(size_t)y * w + x] = DIFF ? __fsub_rn(c[k + 1], m) : m;
    }
}

// ---- min / mean / max / variance -------------------------------------------------------------

// Ordered running minimum (or maximum) as a function of the incoming value, closed under composition:
//   AMD64: kind 0 -> f(m) = minps(m, c); kind 1 (a NaN was seen: the kernel restarts from the element
//          after it) -> f(m) = c; kind 2 -> identity.
//   pure Go: `if v < min { min = v }`: NaN never enters; c = NaN stands for the identity.
struct Ext { float c; int kind; };

template <bool AMD64, bool MAX> __device__ __forceinline__ Ext ext_push(Ext s, float x) {
    if (AMD64) {
        if (s.kind == 2) { s.c = x; s.kind = x != x ? 1 : 0; return s; }
        s.c = MAX ? maxps(s.c, x) : minps(s.c, x);
        if (x != x) s.kind = 1;
        return s;
    }
    if (x == x && (s.c != s.c || (MAX ? x > s.c : x < s.c))) s.c = x;
    return s;
}
// a = the earlier part of the sequence, b = the later one
template <bool AMD64, bool MAX> __device__ __forceinline__ Ext ext_join(Ext a, Ext b) {
    if (AMD64) {
        if (b.kind == 2) return a;
        if (a.kind == 2 || b.kind == 1) return b;
        a.c = MAX ? maxps(a.c, b.c) : minps(a.c, b.c);
        return a;
    }
    if (a.c != a.c || (MAX ? b.c > a.c : b.c < a.c)) a.c = b.c;
    return a;
}

struct StatRec {            // one record = what one warp saw: 1024 vectors of four consecutive elements
    double sum[4], mag[4];  // per lane: sum of the terms, sum of their magnitudes
    Ext mn[4], mx[4];
};

constexpr int REC_VECS = 1024;      // float4 vectors per record (32 per thread)

// term of the chain: the element itself (MODE 0, calcMinMeanMax) or its squared distance from the
// rounded mean, subtracted in fp32 and squared in float64 (MODE 1, calcVariance)
template <int MODE> __device__ __forceinline__ double stat_term(float x, float mean) {
    if (MODE == 0) return (double)x;
    const double d = (double)__fsub_rn(x, mean);
    return __dmul_rn(d, d);
}

// AMD64: lane j of the vectors is its own chain (LANES = 4).  Pure Go: one chain over all elements;
// a thread's 32 vectors are 128 consecutive elements, kept in slot 0.
template <bool AMD64, int MODE>
__global__ void __launch_bounds__(256) stats_records_kernel(const float4 *__restrict__ data, long long n_vecs, float mean,
                                                            StatRec *__restrict__ recs) {
    const int lane = threadIdx.x & 31;
    const long long rec = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long v0 = rec * REC_VECS + lane * 32;
    if (rec * REC_VECS >= n_vecs) return;
    constexpr int L = AMD64 ? 4 : 1;
    double sum[L], mag[L];
    Ext mn[L], mx[L];
#pragma unroll
    for (int j = 0; j < L; j++) { sum[j] = 0.0; mag[j] = 0.0; mn[j] = Ext{AMD64 ? 0.0f : NAN, 2}; mx[j] = mn[j]; }
    for (int i0 = 0; i0 < 32; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = v0 + i0 + i < n_vecs ? __ldcs(data + v0 + i0 + i) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (v0 + i0 + i >= n_vecs) break;
            const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int s = AMD64 ? j : 0;
                const double t = stat_term<MODE>(e[j], mean);
                sum[s] = __dadd_rn(sum[s], t);
                mag[s] = __dadd_rn(mag[s], fabs(t));
                if (MODE == 0) { mn[s] = ext_push<AMD64, false>(mn[s], e[j]); mx[s] = ext_push<AMD64, true>(mx[s], e[j]); }
            }
        }
    }
    // ordered tree over the 32 threads of the record (thread t holds the part before thread t+1's)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < L; j++) {
            const double s2 = __shfl_down_sync(0xffffffffu, sum[j], o), m2 = __shfl_down_sync(0xffffffffu, mag[j], o);
            sum[j] = __dadd_rn(sum[j], s2);
            mag[j] = __dadd_rn(mag[j], m2);
            if (MODE == 0) {
                Ext b;
                b.c = __shfl_down_sync(0xffffffffu, mn[j].c, o); b.kind = __shfl_down_sync(0xffffffffu, mn[j].kind, o);
                mn[j] = ext_join<AMD64, false>(mn[j], b);
                b.c = __shfl_down_sync(0xffffffffu, mx[j].c, o); b.kind = __shfl_down_sync(0xffffffffu, mx[j].kind, o);
                mx[j] = ext_join<AMD64, true>(mx[j], b);
            }
        }
    }
    if (lane == 0) {
        StatRec r;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int s = j < L ? j : 0;
            r.sum[j] = j < L ? sum[s] : 0.0; r.mag[j] = j < L ? mag[s] : 0.0;
            r.mn[j] = j < L ? mn[s] : Ext{NAN, 2}; r.mx[j] = j < L ? mx[s] : Ext{NAN, 2};
        }
        recs[rec] = r;
    }
}

struct StatFold {           // per lane: the parallel sum, the bound on the sequential chain's distance from it, extremes
    double sum[4], bound[4];
    float mn[4], mx[4];
};

// Folds the records in order: 4 lanes x 32 workers, each worker a contiguous run of records.
// chain_len = elements one record adds to a chain (1024 per lane, or 4096 for the single chain).
template <bool AMD64>
__global__ void __launch_bounds__(128) stats_fold_kernel(const StatRec *__restrict__ recs, int n_recs, double chain_len,
                                                         StatFold *__restrict__ out) {
    __shared__ double w_sum[4][32], w_wsum[4][32], w_len[4][32];
    __shared__ Ext w_mn[4][32], w_mx[4][32];
    const int j = threadIdx.x >> 5, k = threadIdx.x & 31;
    const int per = (n_recs + 31) / 32;
    const int lo = k * per, hi = min(n_recs, lo + per);
    double prefix = 0.0, weighted = 0.0;      // weighted = sum over records of chain_len * (|local prefix| + magnitudes)
    Ext mn{AMD64 ? 0.0f : NAN, 2}, mx = mn;
    for (int c = lo; c < hi; c++) {
        weighted += chain_len * (fabs(prefix) + recs[c].mag[j]);
        prefix += recs[c].sum[j];
        mn = ext_join<AMD64, false>(mn, recs[c].mn[j]);
        mx = ext_join<AMD64, true>(mx, recs[c].mx[j]);
    }
    w_sum[j][k] = prefix; w_wsum[j][k] = weighted; w_len[j][k] = chain_len * (double)max(hi - lo, 0);
    w_mn[j][k] = mn; w_mx[j][k] = mx;
    __syncthreads();
    if (k == 0) {
        double total = 0.0, bound = 0.0;
        Ext tmn{AMD64 ? 0.0f : NAN, 2}, tmx = tmn;
        for (int q = 0; q < 32; q++) {
            bound += w_wsum[j][q] + w_len[j][q] * fabs(total);
            total += w_sum[j][q];
            tmn = ext_join<AMD64, false>(tmn, w_mn[j][q]);
            tmx = ext_join<AMD64, true>(tmx, w_mx[j][q]);
        }
        out->sum[j] = total;
        // u = 2^-53 per rounded addition of the reference's chain; the same order of error again for this
        // kernel's own float64 sums (each shorter than one record), and head-room for second-order terms
        out->bound[j] = bound * (2.0 * 1.001 * 1.1102230246251565e-16);
        out->mn[j] = tmn.c; out->mx[j] = tmx.c;
    }
}

// The chains replayed exactly: thread j of the first warp adds its lane's terms one by one in the
// reference's order; the warp stages 32 vectors at a time through shared memory.
template <bool AMD64, int MODE>
__global__ void __launch_bounds__(32) stats_exact_kernel(const float4 *__restrict__ data, long long n_vecs, float mean,
                                                         double *__restrict__ out) {
    constexpr int B = 128;                       // vectors per batch: four loads in flight per thread
    __shared__ float4 stage[2][B];
    const int lane = threadIdx.x;
    double sum = 0.0;
    float4 nxt[4];
#pragma unroll
    for (int q = 0; q < 4; q++) nxt[q] = q * 32 + lane < n_vecs ? data[q * 32 + lane] : make_float4(0, 0, 0, 0);
    int buf = 0;
    for (long long g0 = 0; g0 < n_vecs; g0 += B, buf ^= 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) stage[buf][q * 32 + lane] = nxt[q];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (g0 + B + q * 32 + lane < n_vecs) nxt[q] = data[g0 + B + q * 32 + lane];
        const int cnt = (int)min((long long)B, n_vecs - g0);
        if (AMD64) {
            if (lane < 4) {
                const float *col = reinterpret_cast<const float *>(stage[buf]) + lane;
#pragma unroll 8
                for (int i = 0; i < cnt; i++) sum = __dadd_rn(sum, stat_term<MODE>(col[4 * i], mean));
            }
        } else if (lane == 0) {
            const float *el = reinterpret_cast<const float *>(stage[buf]);
#pragma unroll 8
            for (int i = 0; i < 4 * cnt; i++) sum = __dadd_rn(sum, stat_term<MODE>(el[i], mean));
        }
        __syncwarp();
    }
    if (lane < 4) out[lane] = sum;
}

// ---- bad-pixel scan: indices of t < lo || t > hi in ascending order ------------------------------

constexpr int SEG = 4096;   // elements per warp segment

template <bool WRITE>
__global__ void __launch_bounds__(256) outlier_scan_kernel(const float *__restrict__ t, long long n, float lo, float hi,
                                                           int *__restrict__ seg_count, const int *__restrict__ seg_offset,
                                                           int32_t *__restrict__ out, long long cap) {
    const int lane = threadIdx.x & 31;
    const long long seg = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long base = seg * SEG;
    if (base >= n) return;
    long long pos = WRITE ? seg_offset[seg] : 0;
    int count = 0;
    for (int i0 = 0; i0 < SEG; i0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const long long i = base + i0 + u * 32 + lane;
            v[u] = i < n ? __ldcs(t + i) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const long long i = base + i0 + u * 32 + lane;
            const bool bad = i < n && (v[u] < lo || v[u] > hi);       // badpixels.go:45
            const unsigned m = __ballot_sync(0xffffffffu, bad);
            if (WRITE) {
                const long long slot = pos + __popc(m & ((1u << lane) - 1));
                if (bad && slot < cap) out[slot] = (int32_t)i;
                pos += __popc(m);
            } else {
                count += __popc(m);
            }
        }
    }
    if (!WRITE && lane == 0) seg_count[seg] = count;
}

// ---- host side ---------------------------------------------------------------------------------

static inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// sum of the four lane sums in the kernels' fold order, and the width of the interval around it
static void fold_lanes(const StatFold &f, double *total, double *bound) {
    const double t = (f.sum[2] + f.sum[3]) + (f.sum[0] + f.sum[1]);
    double b = 0.0, m = 0.0;
    for (int j = 0; j < 4; j++) { b += f.bound[j]; m += fabs(f.sum[j]); }
    b += 8.0 * DBL_EPSILON * (m + b);          // the three folding additions, on both sides
    *total = t;
    *bound = b * 1.000001 + DBL_MIN;
}

// Launches one pass (MODE 0: sums + extremes; MODE 1: squared deviations) over n elements and returns
// the reference's chain total: proven from the interval when `resolve` accepts both ends, else replayed.
template <int MODE, typename Resolve>
static int stats_pass(nl_ctx *ctx, const float *dev, long long n, bool amd64, float mean, StatFold *fold_host, double *total,
                      Resolve same_result) {
    const long long n_vecs = n / 4;
    const int n_recs = (int)((n_vecs + REC_VECS - 1) / REC_VECS);
    const size_t rec_bytes = ((size_t)n_recs * sizeof(StatRec) + 255) & ~(size_t)255;
    int rc = ensure_scratch(ctx, rec_bytes + 512);
    if (rc != NL_OK) return rc;
    StatRec *recs = (StatRec *)ctx->scratch;
    StatFold *fold = (StatFold *)((char *)ctx->scratch + rec_bytes);
    double *exact = (double *)((char *)fold + 256);
    const unsigned grid = (unsigned)((n_recs + 7) / 8);
    const float4 *d4 = (const float4 *)dev;
    if (amd64) {
        stats_records_kernel<true, MODE><<<grid, 256, 0, ctx->stream>>>(d4, n_vecs, mean, recs);
        stats_fold_kernel<true><<<1, 128, 0, ctx->stream>>>(recs, n_recs, (double)REC_VECS, fold);
    } else {
        stats_records_kernel<false, MODE><<<grid, 256, 0, ctx->stream>>>(d4, n_vecs, mean, recs);
        stats_fold_kernel<false><<<1, 128, 0, ctx->stream>>>(recs, n_recs, 4.0 * REC_VECS, fold);
    }
    NL_CUDA(cudaGetLastError());
    ctx->launches += 2;
    NL_CUDA(cudaMemcpyAsync(fold_host, fold, sizeof(StatFold), cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    double t, b;
    fold_lanes(*fold_host, &t, &b);
    if (b == b && same_result(t - b, t + b)) { *total = t; return NL_OK; }
    if (amd64) stats_exact_kernel<true, MODE><<<1, 32, 0, ctx->stream>>>(d4, n_vecs, mean, exact);
    else stats_exact_kernel<false, MODE><<<1, 32, 0, ctx->stream>>>(d4, n_vecs, mean, exact);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    ctx->exact_replays++;
    double lanes[4];
    NL_CUDA(cudaMemcpyAsync(lanes, exact, sizeof(lanes), cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    *total = amd64 ? (lanes[2] + lanes[3]) + (lanes[0] + lanes[1]) : lanes[0];
    return NL_OK;
}

// Stats.Min / Mean / Max / StdDev of n device floats -> out = {min, mean, max, stddev}
static int stats_dev(nl_ctx *ctx, const float *dev, long long n, float out[4]) {
    NL_REQUIRE(n >= 1, "statistics of an empty array");
    NL_REQUIRE(n < ((long long)1 << 40), "array too long");
    // the AVX2 loops read whole vectors; a length that is not a multiple of four makes the reference read past
    // its slice, so such arrays take the pure-Go definition (as do all arrays in pure-Go numerics)
    const bool amd64 = ctx->numerics == NL_NUMERICS_AMD64 && n % 4 == 0;
    const double dn = (double)n;
    StatFold f;
    double total = 0.0;
    // a tail of n % 4 elements only exists in pure-Go numerics; it is appended to the single chain on the host
    float tail[3] = {0, 0, 0};
    const int n_tail = (int)(n % 4);
    if (n_tail) {
        NL_CUDA(cudaMemcpyAsync(tail, dev + (n - n_tail), sizeof(float) * n_tail, cudaMemcpyDeviceToHost, ctx->stream));
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    float first = 0.0f;
    NL_CUDA(cudaMemcpyAsync(&first, dev, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));

    // a tail changes the chain's last few additions: the interval test would need them too, so arrays with a
    // tail fold it in exactly here, after an exact or proven body
    auto finish_sum = [&](double body, int mode, float mean) {
        for (int i = 0; i < n_tail; i++) {
            if (mode == 0) body += (double)tail[i];
            else { volatile float d = tail[i] - mean; const double dd = (double)d; volatile double sq = dd * dd; body += sq; }
        }
        return body;
    };
    auto mean_of = [&](double s) { return (float)(s / dn); };
    int rc;
    if (n >= 4) {
        rc = stats_pass<0>(ctx, dev, n - n_tail, amd64, 0.0f, &f, &total, [&](double lo, double hi) {
            return f32_bits(mean_of(finish_sum(lo, 0, 0.0f))) == f32_bits(mean_of(finish_sum(hi, 0, 0.0f)));
        });
        if (rc != NL_OK) return rc;
    } else {
        for (int j = 0; j < 4; j++) { f.mn[j] = NAN; f.mx[j] = NAN; }
    }
    const float mean = mean_of(finish_sum(total, 0, 0.0f));
    float mn, mx;
    if (amd64) {
        // stats_amd64.s:66-77: lanes (0,1) and (2,3), then across; src1 is the lower lane
        auto mnps = [](float a, float b) { return a < b ? a : b; };
        auto mxps = [](float a, float b) { return a > b ? a : b; };
        mn = mnps(mnps(f.mn[0], f.mn[1]), mnps(f.mn[2], f.mn[3]));
        mx = mxps(mxps(f.mx[0], f.mx[1]), mxps(f.mx[2], f.mx[3]));
    } else {
        // stats.go:265-273: starts from data[0]; a NaN there stays; the tail continues the same loop
        mn = f.mn[0]; mx = f.mx[0];
        for (int i = 0; i < n_tail; i++) {
            if (tail[i] == tail[i] && (mn != mn || tail[i] < mn)) mn = tail[i];
            if (tail[i] == tail[i] && (mx != mx || tail[i] > mx)) mx = tail[i];
        }
        if (first != first) mn = mx = first;
    }
    double var_total = 0.0;
    auto std_of = [&](double s) { return (float)sqrt(s / dn); };
    if (n >= 4) {
        StatFold fv;
        rc = stats_pass<1>(ctx, dev, n - n_tail, amd64, mean, &fv, &var_total, [&](double lo, double hi) {
            return f32_bits(std_of(finish_sum(lo < 0.0 ? 0.0 : lo, 1, mean))) == f32_bits(std_of(finish_sum(hi, 1, mean)));
        });
        if (rc != NL_OK) return rc;
    }
    out[0] = mn; out[1] = mean; out[2] = mx;
    out[3] = std_of(finish_sum(var_total, 1, mean));       // stats.go:147-149
    return NL_OK;
}

static int median_launch(nl_ctx *ctx, const float *dev_data, int w, int h, float *dev_out, bool diff) {
    if (w <= 0 || h <= 0) return NL_OK;
    // the AVX2 line kernel needs eight columns; narrower images read before the row there, take the Go loop
    const bool amd64 = ctx->numerics == NL_NUMERICS_AMD64 && w >= 8;
    dim3 block(64, 4), grid((w + 63) / 64, (h + 15) / 16);
    if (amd64 && diff) median3x3_kernel<true, true><<<grid, block, 0, ctx->stream>>>(dev_data, w, h, dev_out);
    else if (amd64) median3x3_kernel<true, false><<<grid, block, 0, ctx->stream>>>(dev_data, w, h, dev_out);
    else if (diff) median3x3_kernel<false, true><<<grid, block, 0, ctx->stream>>>(dev_data, w, h, dev_out);
    else median3x3_kernel<false, false><<<grid, block, 0, ctx->stream>>>(dev_data, w, h, dev_out);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

}  // namespace nl

using namespace nl;

extern "C" {

int nl_ctx_set_numerics(nl_ctx *ctx, int32_t numerics) {
    NL_REQUIRE(ctx && (numerics == NL_NUMERICS_AMD64 || numerics == NL_NUMERICS_PUREGO), "bad numerics");
    ctx->numerics = numerics;
    return NL_OK;
}

int nl_ctx_exact_replays(nl_ctx *ctx, int64_t *replays) {
    NL_REQUIRE(ctx && replays, "NULL argument");
    *replays = ctx->exact_replays;
    return NL_OK;
}

int nl_median_filter3x3_dev(nl_ctx *ctx, const float *dev_data, int32_t width, int32_t height, float *dev_out) {
    NL_REQUIRE(ctx && width >= 0 && height >= 0, "bad argument");
    NL_REQUIRE((dev_data && dev_out) || width == 0 || height == 0, "NULL image pointer");
    CtxGuard g(ctx);
    return median_launch(ctx, dev_data, width, height, dev_out, false);
}

int nl_median_filter3x3(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float *host_out) {
    NL_REQUIRE(ctx && len >= 0 && width > 0 && len % width == 0, "bad image geometry");
    if (len == 0) return NL_OK;
    NL_REQUIRE(host_data && host_out, "NULL image pointer");
    CtxGuard g(ctx);
    const size_t bytes = sizeof(float) * (size_t)len, off = (bytes + 255) & ~(size_t)255;
    int rc = ensure_scratch(ctx, 2 * off);
    if (rc != NL_OK) return rc;
    float *din = (float *)ctx->scratch, *dout = (float *)((char *)ctx->scratch + off);
    NL_CUDA(cudaMemcpyAsync(din, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = median_launch(ctx, din, width, len / width, dout, false);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(host_out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_stats_dev(nl_ctx *ctx, const float *dev_data, int64_t len, float stats[4]) {
    NL_REQUIRE(ctx && dev_data && stats, "NULL argument");
    CtxGuard g(ctx);
    return stats_dev(ctx, dev_data, len, stats);
}

int nl_stats(nl_ctx *ctx, const float *host_data, int64_t len, float stats[4]) {
    NL_REQUIRE(ctx && host_data && stats && len >= 1, "bad argument");
    CtxGuard g(ctx);
    float *dev = nullptr;
    NL_CUDA(cudaMalloc(&dev, sizeof(float) * (size_t)len));
    cudaError_t e = cudaMemcpyAsync(dev, host_data, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? stats_dev(ctx, dev, len, stats) : cuda_fail(e, "stats upload");
    cudaFree(dev);
    return rc;
}

// BadPixelMap on a device frame.  dev_tmp (len floats) receives data - median3x3(data).
int nl_bad_pixel_map_dev(nl_ctx *ctx, const float *dev_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                         float *dev_tmp, int32_t *host_bpm, int64_t cap, int64_t *count, float stats[4]) {
    NL_REQUIRE(ctx && count && stats && width > 0 && len >= 1 && len % width == 0, "bad argument");
    NL_REQUIRE(len <= INT32_MAX, "frame too large for int32 indices");
    NL_REQUIRE(dev_data && dev_tmp && (host_bpm || cap == 0) && cap >= 0, "NULL pointer");
    CtxGuard g(ctx);
    *count = 0;
    int rc = median_launch(ctx, dev_data, width, (int)(len / width), dev_tmp, true);
    if (rc != NL_OK) return rc;
    rc = stats_dev(ctx, dev_tmp, len, stats);
    if (rc != NL_OK) return rc;
    volatile float lo = -stats[3] * sigma_low, hi = stats[3] * sigma_high;    // badpixels.go:38-39
    const long long n_seg = (len + SEG - 1) / SEG;
    rc = ensure_scratch(ctx, ((size_t)(2 * n_seg + 1) * sizeof(int) + 255) & ~(size_t)255);
    if (rc != NL_OK) return rc;
    int *seg_count = (int *)ctx->scratch, *seg_offset = seg_count + n_seg, *total = seg_offset + n_seg;
    const unsigned grid = (unsigned)((n_seg + 7) / 8);
    outlier_scan_kernel<false><<<grid, 256, 0, ctx->stream>>>(dev_tmp, len, lo, hi, seg_count, nullptr, nullptr, 0);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    rc = exclusive_scan_launch(ctx, seg_count, seg_offset, (int)n_seg, total);
    if (rc != NL_OK) return rc;
    int n_bad = 0;
    NL_CUDA(cudaMemcpyAsync(&n_bad, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = n_bad;
    const long long keep = n_bad < cap ? n_bad : cap;
    if (keep <= 0) return NL_OK;
    const size_t need = sizeof(int32_t) * (size_t)keep;
    if (ctx->list_bytes < need) {
        if (ctx->list) { NL_CUDA(cudaFree(ctx->list)); ctx->list = nullptr; ctx->list_bytes = 0; }
        const size_t grow = need + need / 2 + 4096;
        NL_CUDA(cudaMalloc(&ctx->list, grow));
        ctx->list_bytes = grow;
    }
    outlier_scan_kernel<true><<<grid, 256, 0, ctx->stream>>>(dev_tmp, len, lo, hi, seg_count, seg_offset, (int32_t *)ctx->list, keep);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    NL_CUDA(cudaMemcpyAsync(host_bpm, ctx->list, need, cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_bad_pixel_map(nl_ctx *ctx, const float *host_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                     int32_t *host_bpm, int64_t cap, int64_t *count, float stats[4]) {
    NL_REQUIRE(ctx && host_data && len >= 1, "bad argument");
    CtxGuard g(ctx);
    float *dev = nullptr;
    const size_t bytes = sizeof(float) * (size_t)len, off = (bytes + 255) & ~(size_t)255;
    NL_CUDA(cudaMalloc(&dev, 2 * off));
    cudaError_t e = cudaMemcpyAsync(dev, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? nl_bad_pixel_map_dev(ctx, dev, len, width, sigma_low, sigma_high, (float *)((char *)dev + off),
                                                     host_bpm, cap, count, stats)
                              : cuda_fail(e, "bad-pixel map upload");
    cudaFree(dev);
    return rc;
}

}  // extern "C"
