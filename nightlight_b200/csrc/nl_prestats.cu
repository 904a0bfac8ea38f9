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
#include <thread>
#include <atomic>
#include <string>

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

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

// Windows without NaN and without zeros have one well-defined median value (equal non-zero values are the same bits
// whichever comparator wins), so any exact median will do there.  Sorting each window ROW once (three comparators,
// shared by the three output rows that contain it) leaves median9 = med3(max of the row minima, med3 of the row
// medians, min of the row maxima): 21 min/max per pixel instead of the network's 31.
__device__ __forceinline__ void sort3(float &a, float &b, float &c) {
    float t = fminf(a, b); b = fmaxf(a, b); a = t;
    t = fminf(b, c); c = fmaxf(b, c); b = t;
    t = fminf(a, b); b = fmaxf(a, b); a = t;
}
__device__ __forceinline__ float med3(float a, float b, float c) { return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c)); }

// NaN or a zero of either sign: (bits << 1) is 0 for zeros and above 0xff000000 for NaN
__device__ __forceinline__ bool nan_or_zero(float v) { const uint32_t b = __float_as_uint(v) << 1; return b == 0u || b > 0xff000000u; }

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
    bool special[6];                                              // a NaN or a zero in this row of the window
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int y = y0 - 1 + k;
        y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
        const float *row = data + (size_t)y * w;
        l[k] = __ldg(row + xl); c[k] = __ldg(row + x); r[k] = __ldg(row + xr);
        special[k] = nan_or_zero(l[k]) | nan_or_zero(c[k]) | nan_or_zero(r[k]);
    }
    float lo[6], md[6], hi[6];                                    // every window row sorted once
#pragma unroll
    for (int k = 0; k < 6; k++) { lo[k] = l[k]; md[k] = c[k]; hi[k] = r[k]; sort3(lo[k], md[k], hi[k]); }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = y0 + k;
        if (y >= h) break;
        float m = c[k + 1];                                       // border rows and columns are copied
        if (x > 0 && x < w - 1 && y > 0 && y < h - 1) {
            if (special[k] | special[k + 1] | special[k + 2])
                m = median9<AMD64>(l[k], c[k], r[k], l[k + 1], c[k + 1], r[k + 1], l[k + 2], c[k + 2], r[k + 2]);
            else
                m = med3(fmaxf(fmaxf(lo[k], lo[k + 1]), lo[k + 2]), med3(md[k], md[k + 1], md[k + 2]),
                         fminf(fminf(hi[k], hi[k + 1]), hi[k + 2]));
        }
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

// A node of the ordered fold: what a contiguous run of chain elements contributes, per lane.
//   S   parallel sum of the run's terms
//   A   upper bound on the sum over the run's elements of |partial sum up to that element|, partial sums counted
//       from the start of the run: a leaf of at most 1024 elements has A = len * (sum of magnitudes)
//   P   upper bound on |partial sum| anywhere in the run (leaf: the sum of magnitudes)
//   len chain elements in the run
// join(a, b) for a before b: S = a.S + b.S, A = a.A + b.A + b.len * |a.S|, P = max(a.P, |a.S| + b.P), len = a.len + b.len.  At the root, u * A
// bounds the distance of the sequentially rounded chain from the exact sum (u = 2^-53 per addition).
struct Node {
    double S[4], A[4];
    double P[4];            // upper bound on |partial sum| anywhere in the run, partial sums counted from its start
    Ext mn[4], mx[4];
    double len;
};

struct StatOut {            // what the host reads back after both passes
    float mn, mean, mx, stddev;
    int mean_decided, std_decided;
    double total[2], bound[2];
    double lane_sum[2][4], lane_pmax[2][4];     // per pass: the lanes' parallel sums and bounds on their partial sums
};

constexpr int REC_VECS = 1024;      // float4 vectors per warp (32 per thread)
constexpr int FOLD_WORKERS = 64;

// term of the chain: the element itself (MODE 0, calcMinMeanMax) or its squared distance from the
// rounded mean, subtracted in fp32 and squared in float64 (MODE 1, calcVariance)
template <int MODE> __device__ __forceinline__ double stat_term(float x, float mean) {
    if (MODE == 0) return (double)x;
    const double d = (double)__fsub_rn(x, mean);
    return __dmul_rn(d, d);
}

template <bool AMD64> __device__ __forceinline__ Ext ext_empty() { return Ext{AMD64 ? 0.0f : NAN, 2}; }

__device__ __forceinline__ uint32_t dev_f32_bits(float f) { return __float_as_uint(f); }

// One pass over the array (MODE 0: sums and extremes; MODE 1: squared deviations from out->mean).
// AMD64: lane j of the vectors is its own chain.  Pure Go: one chain over all elements (slot 0); a thread's 32
// vectors are 128 consecutive elements.  Every CTA folds its eight warps into one node; the CTA that finishes last
// folds all nodes in order, appends the up-to-three tail elements of a ragged pure-Go array, and decides the
// float32 result from the interval [total - bound, total + bound] (see the file header).
template <bool AMD64, int MODE>
__global__ void __launch_bounds__(256) stats_pass_kernel(const float4 *__restrict__ data, long long n_vecs, long long n,
                                                         Node *__restrict__ nodes, unsigned *__restrict__ counter,
                                                         StatOut *__restrict__ out) {
    constexpr int L = AMD64 ? 4 : 1;
    __shared__ double sh_S[4][FOLD_WORKERS], sh_A[4][FOLD_WORKERS], sh_P[4][FOLD_WORKERS], sh_len[FOLD_WORKERS];
    __shared__ Ext sh_mn[4][FOLD_WORKERS], sh_mx[4][FOLD_WORKERS];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float mean = MODE == 1 ? out->mean : 0.0f;
    const long long rec = (long long)blockIdx.x * 8 + warp;
    const long long v0 = rec * REC_VECS + lane * 32;
    double sum[L], mag[L];
    Ext mn[L], mx[L];
#pragma unroll
    for (int j = 0; j < L; j++) { sum[j] = 0.0; mag[j] = 0.0; mn[j] = ext_empty<AMD64>(); mx[j] = mn[j]; }
    if (rec * REC_VECS < n_vecs) {
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
    }
    // ordered tree over the 32 threads of the warp (thread t holds the part before thread t+1's)
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
        long long vecs = n_vecs - rec * REC_VECS;
        vecs = vecs < 0 ? 0 : (vecs > REC_VECS ? REC_VECS : vecs);
        const double len = (double)(AMD64 ? vecs : 4 * vecs);
        sh_len[warp] = len;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int s = j < L ? j : 0;
            sh_S[j][warp] = j < L ? sum[s] : 0.0;
            sh_A[j][warp] = j < L ? (double)(AMD64 ? REC_VECS : 4 * REC_VECS) * mag[s] : 0.0;    // (a full leaf's length also for the last, shorter one)
            sh_P[j][warp] = j < L ? mag[s] : 0.0;
            sh_mn[j][warp] = j < L ? mn[s] : ext_empty<AMD64>();
            sh_mx[j][warp] = j < L ? mx[s] : ext_empty<AMD64>();
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {                       // the eight warps of this CTA, in order, into one node
        const int j = threadIdx.x;
        double S = 0.0, A = 0.0, P = 0.0, len = 0.0;
        Ext tmn = ext_empty<AMD64>(), tmx = tmn;
        for (int w = 0; w < 8; w++) {
            A += sh_A[j][w] + sh_len[w] * fabs(S);
            P = fmax(P, fabs(S) + sh_P[j][w]);
            S += sh_S[j][w];
            len += sh_len[w];
            tmn = ext_join<AMD64, false>(tmn, sh_mn[j][w]);
            tmx = ext_join<AMD64, true>(tmx, sh_mx[j][w]);
        }
        Node *nd = nodes + blockIdx.x;
        nd->S[j] = S; nd->A[j] = A; nd->P[j] = P; nd->mn[j] = tmn; nd->mx[j] = tmx;
        if (j == 0) nd->len = len;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // the last CTA: nodes in order, 64 workers x 4 lanes, then an ordered tree over the workers
    {
        const int j = threadIdx.x & 3, k = threadIdx.x >> 2;
        const int n_nodes = (int)gridDim.x;
        const int per = (n_nodes + FOLD_WORKERS - 1) / FOLD_WORKERS;
        const int lo = min(n_nodes, k * per), hi = min(n_nodes, lo + per);
        double S = 0.0, A = 0.0, P = 0.0, len = 0.0;
        Ext tmn = ext_empty<AMD64>(), tmx = tmn;
        const volatile Node *vn = nodes;
        for (int c = lo; c < hi; c++) {
            const double cS = vn[c].S[j], cA = vn[c].A[j], cP = vn[c].P[j], cl = vn[c].len;
            P = fmax(P, fabs(S) + cP);
            Ext cmn, cmx;
            cmn.c = vn[c].mn[j].c; cmn.kind = vn[c].mn[j].kind; cmx.c = vn[c].mx[j].c; cmx.kind = vn[c].mx[j].kind;
            A += cA + cl * fabs(S);
            S += cS;
            len += cl;
            tmn = ext_join<AMD64, false>(tmn, cmn);
            tmx = ext_join<AMD64, true>(tmx, cmx);
        }
        sh_S[j][k] = S; sh_A[j][k] = A; sh_P[j][k] = P; sh_mn[j][k] = tmn; sh_mx[j][k] = tmx;
        if (j == 0) sh_len[k] = len;
        __syncthreads();
        for (int d = 1; d < FOLD_WORKERS; d <<= 1) {
            double nS = 0, nA = 0, nP = 0, nl = 0;
            Ext nmn = tmn, nmx = tmx;
            const bool act = (k % (2 * d)) == 0;
            if (act) {
                nA = sh_A[j][k] + sh_A[j][k + d] + sh_len[k + d] * fabs(sh_S[j][k]);
                nP = fmax(sh_P[j][k], fabs(sh_S[j][k]) + sh_P[j][k + d]);
                nS = sh_S[j][k] + sh_S[j][k + d];
                nl = sh_len[k] + sh_len[k + d];
                nmn = ext_join<AMD64, false>(sh_mn[j][k], sh_mn[j][k + d]);
                nmx = ext_join<AMD64, true>(sh_mx[j][k], sh_mx[j][k + d]);
            }
            __syncthreads();
            if (act) {
                sh_S[j][k] = nS; sh_A[j][k] = nA; sh_P[j][k] = nP; sh_mn[j][k] = nmn; sh_mx[j][k] = nmx;
                if (j == 0) sh_len[k] = nl;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x != 0) return;
    *counter = 0;                                              // ready for the next pass
    // the lanes, folded like stats_amd64.s:80-84; the interval around the total
    double total = (sh_S[2][0] + sh_S[3][0]) + (sh_S[0][0] + sh_S[1][0]);
    double bound = 0.0, m = 0.0, pm = 0.0;
    for (int j = 0; j < 4; j++) { bound += sh_A[j][0]; m += fabs(sh_S[j][0]); pm += sh_P[j][0]; }
    // u = 2^-53 per rounded addition of the reference's chain: u * A.  This kernel's own float64 sums: an element passes
    // through at most 45 additions whose results stay below its leaf's sum of magnitudes (<= A / 1024 in total: 5 % of A),
    // and one addition per node plus the worker tree at prefix level (results <= P).  Head-room for second-order terms;
    // then the three folding additions.
    bound = (1.05 * bound + ((double)gridDim.x + 400.0) * pm) * (1.001 * 1.1102230246251565e-16);
    bound += 8.0 * 2.220446049250313e-16 * (m + bound);
    bound = bound * 1.000001 + 2.2250738585072014e-308;
    out->total[MODE] = total;
    out->bound[MODE] = bound;
    for (int j = 0; j < 4; j++) { out->lane_sum[MODE][j] = sh_S[j][0]; out->lane_pmax[MODE][j] = sh_P[j][0]; }
    const double dn = (double)n;
    const int n_tail = AMD64 ? 0 : (int)(n - 4 * n_vecs);       // a ragged pure-Go array: its last elements continue the chain
    const float *tail = reinterpret_cast<const float *>(data) + 4 * n_vecs;
    double lo = total - bound, hi = total + bound;
    if (MODE == 1 && lo < 0.0) lo = 0.0;
    for (int i = 0; i < n_tail; i++) {
        const double t = stat_term<MODE>(tail[i], mean);
        lo = __dadd_rn(lo, t); hi = __dadd_rn(hi, t);
    }
    if (MODE == 0) {
        const float mlo = __double2float_rn(__ddiv_rn(lo, dn)), mhi = __double2float_rn(__ddiv_rn(hi, dn));
        out->mean = mlo;
        out->mean_decided = bound == bound && dev_f32_bits(mlo) == dev_f32_bits(mhi);
        float mn_, mx_;
        if (AMD64) {
            // stats_amd64.s:66-77: lanes (0,1) and (2,3), then across; src1 is the lower lane
            mn_ = minps(minps(sh_mn[0][0].c, sh_mn[1][0].c), minps(sh_mn[2][0].c, sh_mn[3][0].c));
            mx_ = maxps(maxps(sh_mx[0][0].c, sh_mx[1][0].c), maxps(sh_mx[2][0].c, sh_mx[3][0].c));
        } else {
            // stats.go:265-273: the loop starts from data[0] (a NaN there stays) and the tail continues it
            Ext a = sh_mn[0][0], b = sh_mx[0][0];
            for (int i = 0; i < n_tail; i++) { a = ext_push<false, false>(a, tail[i]); b = ext_push<false, true>(b, tail[i]); }
            const float first = reinterpret_cast<const float *>(data)[0];
            mn_ = first != first ? first : a.c;
            mx_ = first != first ? first : b.c;
        }
        out->mn = mn_; out->mx = mx_;
    } else {
        const float slo = __double2float_rn(__dsqrt_rn(__ddiv_rn(lo, dn))), shi = __double2float_rn(__dsqrt_rn(__ddiv_rn(hi, dn)));
        out->stddev = slo;                                      // stats.go:147-149
        out->std_decided = bound == bound && dev_f32_bits(slo) == dev_f32_bits(shi);
    }
}

// Second-level proof for sums with heavy cancellation (a mean many orders below the spread of the data: the
// interval above is relative to the magnitudes summed, far wider than the float32 grid of such a mean).  Let
// G = 2^g be the ulp of the largest binade any partial sum of the chain can reach (from Node::P).  A term that is a
// multiple of G is added without rounding to a partial sum that is a multiple of G, in any order; only "fine" terms
// (lowest set bit below G) start rounding, and each of them accounts for at most one geometric series of roundings
// bounded by G in the reference's chain and by 2G in the parallel sum.  So |chain - parallel sum| <= 3 * G * (number of
// fine terms) per lane -- zero for data on a grid (camera ADUs, differences of them), a few dozen ulps otherwise.
// This kernel counts the fine terms per lane.
template <bool AMD64, int MODE>
__global__ void __launch_bounds__(256) stats_fine_count_kernel(const float4 *__restrict__ data, long long n_vecs, float mean, int g0,
                                                               int g1, int g2, int g3, unsigned long long *__restrict__ counts) {
    const int g[4] = {g0, AMD64 ? g1 : g0, AMD64 ? g2 : g0, AMD64 ? g3 : g0};
    unsigned cnt[4] = {0, 0, 0, 0};
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n_vecs; v += (long long)gridDim.x * blockDim.x) {
        const float4 q = __ldcs(data + v);
        const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double t = stat_term<MODE>(e[j], mean);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
            const int ex = (int)((bits >> 52) & 0x7ff);
            const unsigned long long mant = bits & 0xfffffffffffffull;
            int lsb;                                                  // exponent of the lowest set bit of t
            if (ex == 0x7ff) lsb = -100000;                           // Inf / NaN: never provable
            else if (ex == 0) lsb = mant ? -1074 + (__ffsll((long long)mant) - 1) : 100000;     // zero adds nothing
            else lsb = ex - 1075 + (__ffsll((long long)(mant | (1ull << 52))) - 1);
            cnt[AMD64 ? j : 0] += lsb < g[j] ? 1u : 0u;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        unsigned c = cnt[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + j, (unsigned long long)c);
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


// Stats.Min / Mean / Max / StdDev of n device floats -> out = {min, mean, max, stddev}.
// Both passes are queued back to back (the second reads the mean the first one left on the device); one read-back.
// An undecided interval replays that chain in order and, for the mean, repeats the second pass.
static int stats_dev(nl_ctx *ctx, const float *dev, long long n, float out[4]) {
    NL_REQUIRE(n >= 1, "statistics of an empty array");
    NL_REQUIRE(n < ((long long)1 << 40), "array too long");
    // the AVX2 loops read whole vectors; a length that is not a multiple of four makes the reference read past
    // its slice, so such arrays take the pure-Go definition (as do all arrays in pure-Go numerics)
    const bool amd64 = ctx->numerics == NL_NUMERICS_AMD64 && n % 4 == 0;
    const long long n_vecs = n / 4;
    const int n_tail = (int)(n - 4 * n_vecs);
    const unsigned grid = (unsigned)std::max<long long>(1, (n_vecs + 8 * REC_VECS - 1) / (8 * REC_VECS));
    const size_t node_bytes = ((size_t)grid * sizeof(Node) + 255) & ~(size_t)255;
    int rc = ensure_scratch(ctx, node_bytes + 1280);
    if (rc != NL_OK) return rc;
    Node *nodes = (Node *)ctx->scratch;
    StatOut *dout = (StatOut *)((char *)ctx->scratch + node_bytes);
    unsigned *counter = (unsigned *)((char *)dout + 256);
    double *exact = (double *)((char *)dout + 512);
    const float4 *d4 = (const float4 *)dev;
    NL_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    auto pass0 = [&]() {
        if (amd64) stats_pass_kernel<true, 0><<<grid, 256, 0, ctx->stream>>>(d4, n_vecs, n, nodes, counter, dout);
        else stats_pass_kernel<false, 0><<<grid, 256, 0, ctx->stream>>>(d4, n_vecs, n, nodes, counter, dout);
        ctx->launches++;
    };
    auto pass1 = [&]() {
        if (amd64) stats_pass_kernel<true, 1><<<grid, 256, 0, ctx->stream>>>(d4, n_vecs, n, nodes, counter, dout);
        else stats_pass_kernel<false, 1><<<grid, 256, 0, ctx->stream>>>(d4, n_vecs, n, nodes, counter, dout);
        ctx->launches++;
    };
    // the chain of pass `mode` replayed in order (plus the tail of a ragged pure-Go array, on the host)
    float tail[3] = {0, 0, 0};
    auto replay = [&](int mode, float mean, double *total) -> int {
        if (amd64) {
            if (mode == 0) stats_exact_kernel<true, 0><<<1, 32, 0, ctx->stream>>>(d4, n_vecs, mean, exact);
            else stats_exact_kernel<true, 1><<<1, 32, 0, ctx->stream>>>(d4, n_vecs, mean, exact);
        } else {
            if (mode == 0) stats_exact_kernel<false, 0><<<1, 32, 0, ctx->stream>>>(d4, n_vecs, mean, exact);
            else stats_exact_kernel<false, 1><<<1, 32, 0, ctx->stream>>>(d4, n_vecs, mean, exact);
        }
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        ctx->exact_replays++;
        double lanes[4];
        NL_CUDA(cudaMemcpyAsync(lanes, exact, sizeof(lanes), cudaMemcpyDeviceToHost, ctx->stream));
        if (n_tail) NL_CUDA(cudaMemcpyAsync(tail, dev + 4 * n_vecs, sizeof(float) * n_tail, cudaMemcpyDeviceToHost, ctx->stream));
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
        double t = amd64 ? (lanes[2] + lanes[3]) + (lanes[0] + lanes[1]) : lanes[0];
        for (int i = 0; i < n_tail; i++) {
            if (mode == 0) t += (double)tail[i];
            else { volatile float d = tail[i] - mean; const double dd = (double)d; volatile double sq = dd * dd; t += sq; }
        }
        *total = t;
        return NL_OK;
    };
    pass0();
    pass1();
    NL_CUDA(cudaGetLastError());
    StatOut h;
    NL_CUDA(cudaMemcpyAsync(&h, dout, sizeof(StatOut), cudaMemcpyDeviceToHost, ctx->stream));
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    const double dn = (double)n;
    if (ctx->stats_debug)
        fprintf(stderr, "nl_stats: n=%lld amd64=%d mean_decided=%d total0=%.17g bound0=%.3g (rel %.3g) std_decided=%d total1=%.17g bound1=%.3g (rel %.3g)\n",
                n, (int)amd64, h.mean_decided, h.total[0], h.bound[0], h.bound[0] / fabs(h.total[0]), h.std_decided, h.total[1], h.bound[1],
                h.bound[1] / fabs(h.total[1]));
    // second-level proof (see stats_fine_count_kernel): the chain total of pass `mode` from the parallel lane sums when
    // few terms can round at all.  true: *total is within a proven distance that does not change the float32 result.
    unsigned long long *dcounts = (unsigned long long *)((char *)dout + 768);
    auto prove = [&](int mode, float mean, const StatOut &so, double *total) -> int {       // 1 proven, 0 not, < 0 error
        int g[4];
        for (int j = 0; j < 4; j++) {
            const double p = so.lane_pmax[mode][j] * 1.0000001;
            if (!(p == p) || p > 1e300) return 0;
            g[j] = (p > 0.0 ? ilogb(p) + 1 : -1074) + 1 - 53 + 1;       // ulp of the binade above 2 * max |partial sum|
        }
        if (cudaMemsetAsync(dcounts, 0, 4 * sizeof(unsigned long long), ctx->stream) != cudaSuccess) return -1;
        const unsigned cgrid = (unsigned)std::min<long long>(std::max<long long>(1, (n_vecs + 255) / 256), (long long)ctx->sm_count * 8);
        if (amd64) {
            if (mode == 0) stats_fine_count_kernel<true, 0><<<cgrid, 256, 0, ctx->stream>>>(d4, n_vecs, mean, g[0], g[1], g[2], g[3], dcounts);
            else stats_fine_count_kernel<true, 1><<<cgrid, 256, 0, ctx->stream>>>(d4, n_vecs, mean, g[0], g[1], g[2], g[3], dcounts);
        } else {
            if (mode == 0) stats_fine_count_kernel<false, 0><<<cgrid, 256, 0, ctx->stream>>>(d4, n_vecs, mean, g[0], g[1], g[2], g[3], dcounts);
            else stats_fine_count_kernel<false, 1><<<cgrid, 256, 0, ctx->stream>>>(d4, n_vecs, mean, g[0], g[1], g[2], g[3], dcounts);
        }
        if (cudaGetLastError() != cudaSuccess) return -1;
        ctx->launches++;
        unsigned long long counts[4];
        if (cudaMemcpyAsync(counts, dcounts, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
        if (n_tail && cudaMemcpyAsync(tail, dev + 4 * n_vecs, sizeof(float) * n_tail, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
        double b = 0.0;
        for (int j = 0; j < (amd64 ? 4 : 1); j++) b += 4.0 * (double)counts[j] * ldexp(1.0, g[j]);
        const double *ls = so.lane_sum[mode];
        double t = amd64 ? (ls[2] + ls[3]) + (ls[0] + ls[1]) : ls[0];
        if (b > 0.0) b += 4.0 * DBL_EPSILON * (fabs(ls[0]) + fabs(ls[1]) + fabs(ls[2]) + fabs(ls[3]) + b);      // the lane fold
        auto finish = [&](double v) {
            for (int i = 0; i < n_tail; i++) {
                if (mode == 0) v += (double)tail[i];
                else { volatile float d = tail[i] - mean; const double dd = (double)d; volatile double sq = dd * dd; v += sq; }
            }
            return mode == 0 ? (float)(v / dn) : (float)sqrt(v / dn);
        };
        double lo = t - b, hi = t + b;
        if (mode == 1 && lo < 0.0) lo = 0.0;
        const float flo = finish(lo), fhi = finish(hi);
        uint32_t ulo, uhi;
        memcpy(&ulo, &flo, 4); memcpy(&uhi, &fhi, 4);
        if (ctx->stats_debug)
            fprintf(stderr, "nl_stats: second-level proof pass %d: fine terms %llu %llu %llu %llu, grid 2^%d, bound %.3g -> %s\n", mode, counts[0],
                    counts[1], counts[2], counts[3], g[0], b, ulo == uhi ? "proven" : "not proven");
        if (ulo != uhi) return 0;
        *total = t;                               // (without the tail: the caller appends it like the replay path does)
        return 1;
    };
    const bool force_replay = ctx->stats_force_replay;      // nl_ctx_set_tuning "stats_force_replay": time the in-order replay
    if (force_replay) h.mean_decided = h.std_decided = 0;
    if (!h.mean_decided) {
        double t;
        const int proven = force_replay ? 0 : prove(0, 0.0f, h, &t);
        if (proven < 0) return cuda_fail(cudaGetLastError(), "second-level proof");
        if (proven) {
            for (int i = 0; i < n_tail; i++) t += (double)tail[i];
        } else {
            rc = replay(0, 0.0f, &t);
            if (rc != NL_OK) return rc;
        }
        h.mean = (float)(t / dn);
        NL_CUDA(cudaMemcpyAsync(&dout->mean, &h.mean, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        pass1();
        NL_CUDA(cudaGetLastError());
        const float mean = h.mean;
        NL_CUDA(cudaMemcpyAsync(&h, dout, sizeof(StatOut), cudaMemcpyDeviceToHost, ctx->stream));
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
        h.mean = mean;
        if (force_replay) h.std_decided = 0;
    }
    if (!h.std_decided) {
        double t;
        const int proven = force_replay ? 0 : prove(1, h.mean, h, &t);
        if (proven < 0) return cuda_fail(cudaGetLastError(), "second-level proof");
        if (proven) {
            for (int i = 0; i < n_tail; i++) { volatile float d = tail[i] - h.mean; const double dd = (double)d; volatile double sq = dd * dd; t += sq; }
        } else {
            rc = replay(1, h.mean, &t);
            if (rc != NL_OK) return rc;
        }
        h.stddev = (float)sqrt(t / dn);
    }
    out[0] = h.mn; out[1] = h.mean; out[2] = h.mx; out[3] = h.stddev;
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
    NL_GUARD(ctx);
    return median_launch(ctx, dev_data, width, height, dev_out, false);
}

int nl_median_filter3x3(nl_ctx *ctx, const float *host_data, int32_t len, int32_t width, float *host_out) {
    NL_REQUIRE(ctx && len >= 0 && width > 0 && len % width == 0, "bad image geometry");
    if (len == 0) return NL_OK;
    NL_REQUIRE(host_data && host_out, "NULL image pointer");
    NL_GUARD(ctx);
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
    NL_GUARD(ctx);
    if ((uintptr_t)dev_data & 15) {
        // the passes read float4 vectors: a frame that does not start on 16 bytes (e.g. frame k of a stack job whose
        // pixel count is not a multiple of four) is staged into the context's aligned frame buffer first
        NL_REQUIRE(len >= 1, "statistics of an empty array");
        float *aligned = nullptr;
        int rc = ensure_frame(ctx, 0, sizeof(float) * (size_t)len, &aligned);
        if (rc != NL_OK) return rc;
        NL_CUDA(cudaMemcpyAsync(aligned, dev_data, sizeof(float) * (size_t)len, cudaMemcpyDeviceToDevice, ctx->stream));
        dev_data = aligned;
    }
    return stats_dev(ctx, dev_data, len, stats);
}

int nl_stats(nl_ctx *ctx, const float *host_data, int64_t len, float stats[4]) {
    NL_REQUIRE(ctx && host_data && stats && len >= 1, "bad argument");
    NL_GUARD(ctx);
    float *dev = nullptr;
    int rc = ensure_frame(ctx, 0, sizeof(float) * (size_t)len, &dev);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(dev, host_data, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    return stats_dev(ctx, dev, len, stats);
}

// BadPixelMap on a device frame.  dev_tmp (len floats) receives data - median3x3(data).
int nl_bad_pixel_map_dev(nl_ctx *ctx, const float *dev_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                         float *dev_tmp, int32_t *host_bpm, int64_t cap, int64_t *count, float stats[4]) {
    NL_REQUIRE(ctx && count && stats && width > 0 && len >= 1 && len % width == 0, "bad argument");
    NL_REQUIRE(len <= INT32_MAX, "frame too large for int32 indices");
    NL_REQUIRE(dev_data && dev_tmp && (host_bpm || cap == 0) && cap >= 0, "NULL pointer");
    NL_REQUIRE(((uintptr_t)dev_tmp & 15) == 0, "dev_tmp must be 16-byte aligned (its statistics are read as float4 vectors)");
    NL_GUARD(ctx);
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
    // the write pass follows at once, into mapped pinned host memory sized from earlier frames: one host round trip
    rc = ensure_pinned(ctx, 64 + sizeof(int32_t) * 65536);
    if (rc != NL_OK) return rc;
    for (int attempt = 0; attempt < 2; attempt++) {
        const long long slots = (long long)((ctx->pinned_bytes - 64) / sizeof(int32_t));
        volatile int *host_total = (volatile int *)ctx->pinned;
        outlier_scan_kernel<true><<<grid, 256, 0, ctx->stream>>>(dev_tmp, len, lo, hi, seg_count, seg_offset,
                                                                (int32_t *)((char *)ctx->pinned_dev + 64), slots);
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
        NL_CUDA(cudaMemcpyAsync((void *)host_total, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
        const int n_bad = *host_total;
        *count = n_bad;
        const long long keep = n_bad < cap ? n_bad : cap;
        if (keep > slots) {                                   // more than the pinned list holds: grow, write again
            rc = ensure_pinned(ctx, 64 + sizeof(int32_t) * (size_t)(keep + keep / 2));
            if (rc != NL_OK) return rc;
            continue;
        }
        if (keep > 0) memcpy(host_bpm, (char *)ctx->pinned + 64, sizeof(int32_t) * (size_t)keep);
        return NL_OK;
    }
    return set_error(NL_E_CUDA, "bad-pixel map: list did not fit after growing");
}

// BadPixelMap of every frame of a resident stack in one call (frame i at dev_frames + i*frame_stride).  A frame's chain
// stays what nl_bad_pixel_map_dev runs -- its float64 statistics decide host-side whether a chain must be replayed --
// but up to six frames are in flight on the context's helper lanes, and the caller crosses the language boundary once.
// stats = n_frames x 4, counts = n_frames, host_bpm receives frame i's list at host_bpm + i*cap.
int nl_bad_pixel_map_batch_dev(nl_ctx *ctx, const float *dev_frames, int32_t n_frames, int64_t frame_stride, int64_t len, int32_t width,
                               float sigma_low, float sigma_high, int32_t *host_bpm, int64_t cap, int64_t *counts, float *stats) {
    NL_REQUIRE(ctx && counts && stats && n_frames >= 0, "bad argument");
    NL_REQUIRE(dev_frames || n_frames == 0, "NULL frames");
    NL_REQUIRE(frame_stride % 4 == 0 && ((uintptr_t)dev_frames & 15) == 0, "frames must start on 16 bytes");
    NL_GUARD(ctx);
    // several frames in flight: the context and its lanes (own stream, scratch and pinned read-back buffers each), one
    // host thread per lane -- a frame's chain is a handful of short kernels and two host round trips, which the other
    // lanes fill
    constexpr int WORKERS = 1 + NL_MAX_LANES;
    nl_ctx *wc[WORKERS] = {ctx};
    const int workers = n_frames < WORKERS ? (n_frames < 1 ? 1 : n_frames) : WORKERS;
    for (int w = 1; w < workers; w++) {
        int rc = lane_context(ctx, w - 1, &wc[w]);
        if (rc != NL_OK) return rc;
    }
    NL_CUDA(cudaStreamSynchronize(ctx->stream));          // the frames were produced on the context's stream
    std::atomic<int> next{0};
    int64_t replays_of_lanes[WORKERS] = {0};
    int rcs[WORKERS] = {NL_OK};
    std::string msgs[WORKERS];
    auto work = [&](int w) {
        nl_ctx *c = wc[w];
        CtxGuard g(c);
        float *tmp = nullptr;
        int rc = g.ok ? ensure_frame(c, 1, sizeof(float) * (size_t)len, &tmp) : set_error(NL_E_CUDA, "cudaSetDevice(%d) failed", c->device);
        const int64_t launches0 = c->launches.load(), replays0 = c->exact_replays;
        for (int i = next++; rc == NL_OK && i < n_frames; i = next++)
            rc = nl_bad_pixel_map_dev(c, dev_frames + (size_t)i * frame_stride, len, width, sigma_low, sigma_high, tmp,
                                      host_bpm ? host_bpm + (size_t)i * cap : nullptr, host_bpm ? cap : 0, counts + i, stats + 4 * i);
        if (rc != NL_OK) msgs[w] = nl_last_error();       // (the message is per thread)
        if (w > 0) {                                       // the lanes' work counts as the context's
            ctx->launches += c->launches.load() - launches0;
            replays_of_lanes[w] = c->exact_replays - replays0;
        }
        rcs[w] = rc;
    };
    std::thread th[WORKERS];
    for (int w = 1; w < workers; w++) th[w] = std::thread(work, w);
    work(0);
    for (int w = 1; w < workers; w++) th[w].join();
    for (int w = 1; w < workers; w++) ctx->exact_replays += replays_of_lanes[w];
    for (int w = 0; w < workers; w++)
        if (rcs[w] != NL_OK) return w == 0 ? rcs[w] : set_error(rcs[w], "%s", msgs[w].c_str());
    return NL_OK;
}

int nl_bad_pixel_map(nl_ctx *ctx, const float *host_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                     int32_t *host_bpm, int64_t cap, int64_t *count, float stats[4]) {
    NL_REQUIRE(ctx && host_data && len >= 1, "bad argument");
    NL_GUARD(ctx);
    float *dev = nullptr, *tmp = nullptr;
    const size_t bytes = sizeof(float) * (size_t)len;
    int rc = ensure_frame(ctx, 0, bytes, &dev);
    if (rc == NL_OK) rc = ensure_frame(ctx, 1, bytes, &tmp);
    if (rc != NL_OK) return rc;
    NL_CUDA(cudaMemcpyAsync(dev, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return nl_bad_pixel_map_dev(ctx, dev, len, width, sigma_low, sigma_high, tmp, host_bpm, cap, count, stats);
}

// scatter of repaired pixels into the device copy of a frame
__global__ void patch_kernel(float *__restrict__ data, const int32_t *__restrict__ idx, const float *__restrict__ val, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) data[idx[i]] = val[i];
}

// OpBadPixel.Apply for monochrome frames, preprocess.go:180-191: BadPixelMap on the device, then the listed pixels
// repaired in place (sparse and sequential: on the host).  host_data is updated; *removed = len(bpm).
static int op_bad_pixel(nl_ctx *ctx, float *dev_data, bool upload, float *host_data, int64_t len, int32_t width, float sigma_low,
                        float sigma_high, int64_t *removed, float stats[4]) {
    NL_REQUIRE(ctx && host_data && removed && stats && len >= 1 && width > 0, "bad argument");
    *removed = 0;
    if (sigma_low == 0.0f || sigma_high == 0.0f) return NL_OK;          // :181-183
    NL_GUARD(ctx);
    const size_t bytes = sizeof(float) * (size_t)len;
    float *tmp = nullptr;
    int rc = ensure_frame(ctx, 1, bytes, &tmp);
    if (rc != NL_OK) return rc;
    if (upload) {
        rc = ensure_frame(ctx, 0, bytes, &dev_data);
        if (rc != NL_OK) return rc;
        NL_CUDA(cudaMemcpyAsync(dev_data, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    int64_t cap = len / 100 + 1024, count = 0;                          // badpixels.go:42
    std::vector<int32_t> bpm((size_t)cap);
    rc = nl_bad_pixel_map_dev(ctx, dev_data, len, width, sigma_low, sigma_high, tmp, bpm.data(), cap, &count, stats);
    if (rc == NL_OK && count > cap) {
        cap = count;
        bpm.resize((size_t)cap);
        rc = nl_bad_pixel_map_dev(ctx, dev_data, len, width, sigma_low, sigma_high, tmp, bpm.data(), cap, &count, stats);
    }
    if (rc != NL_OK) return rc;
    median_filter_sparse_host(host_data, (int32_t)len, width, bpm.data(), count);
    *removed = count;
    if (!upload && count > 0) {
        // keep the device copy identical to the repaired host frame: scatter the final values of the repaired pixels
        // (a pixel listed once has one final value; the list is ascending and duplicate-free)
        std::vector<float> val((size_t)count);
        for (int64_t k = 0; k < count; k++) val[(size_t)k] = host_data[bpm[(size_t)k]];
        const long long slots = (count + 63) & ~63ll;
        if ((size_t)(2 * slots) * 4 > bytes) {                          // more than half the frame is bad: send it whole
            NL_CUDA(cudaMemcpyAsync(dev_data, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            int32_t *didx = (int32_t *)tmp;                             // the difference image is not needed any more
            float *dval = tmp + slots;
            NL_CUDA(cudaMemcpyAsync(didx, bpm.data(), sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
            NL_CUDA(cudaMemcpyAsync(dval, val.data(), sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
            patch_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(dev_data, didx, dval, count);
            NL_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        NL_CUDA(cudaStreamSynchronize(ctx->stream));                    // bpm and val go out of scope
    }
    return NL_OK;
}

// OpBadPixel.Apply for monochrome frames, preprocess.go:180-191: BadPixelMap on the device, then the listed pixels
// repaired in place (sparse and sequential: on the host).  host_data is updated; *removed = len(bpm).
int nl_op_bad_pixel(nl_ctx *ctx, float *host_data, int64_t len, int32_t width, float sigma_low, float sigma_high,
                    int64_t *removed, float stats[4]) {
    return op_bad_pixel(ctx, nullptr, true, host_data, len, width, sigma_low, sigma_high, removed, stats);
}

// The same with the frame already resident (dev_data holds the pixels of host_data): both copies are repaired, so
// the operators that follow (noise estimate, star detection, resample) run on the device copy without a new upload.
int nl_op_bad_pixel_dev(nl_ctx *ctx, float *dev_data, float *host_data, int64_t len, int32_t width, float sigma_low,
                        float sigma_high, int64_t *removed, float stats[4]) {
    NL_REQUIRE(dev_data, "NULL device frame");
    return op_bad_pixel(ctx, dev_data, false, host_data, len, width, sigma_low, sigma_high, removed, stats);
}

}  // extern "C"
