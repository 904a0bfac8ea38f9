// nl_column.cuh -- per-pixel column reducers of the stacking hot path.
//
// One pixel's samples across the N frames form a "column".  The reference reduces each column on
// the CPU with quick-select + sequential fp32 sums (internal/ops/stack/stack.go:274-918,
// internal/qsort/qsort.go:68-126, internal/stats/stats.go:246-261,569-586).  The result of the
// mean-type modes depends on the element ORDER left behind by that quick-select and by the
// swap-with-last clip loop, because mean and sigma are sequential fp32 sums over the permuted buffer.
// These routines therefore reproduce the reference's permutation and evaluation order exactly, but
// are organised for SIMT: a column lives in shared memory with a compile-time element stride S
// (S = pixels per warp tile, so lane == bank), and the quick-select is a flattened state machine
// that keeps the 32 lanes of a warp (32 different pixels) in one loop instead of nested
// data-dependent loops.
//
// Everything here is __host__ __device__ so tests can compile the very same code for the CPU
// (tests/host_emul.cpp, S = 1) and compare it against the oracle without a GPU.
// Compile with -fmad=false: Go/amd64 never contracts a*b+c.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NL_HD __host__ __device__ __forceinline__
#else
#define NL_HD inline
#endif

// Warp-wide "any lane" vote.  The flattened loops below keep the 32 lanes of a warp (32 different
// pixels) in ONE loop that runs until the slowest lane is done; on the host (stride 1, one column
// at a time) the vote is the lane's own flag.  Every routine that votes must be called by all 32
// lanes of the warp (lanes without a pixel pass n = 0).
#if defined(__CUDA_ARCH__)
#define NL_ANY(x) (__any_sync(0xffffffffu, (x)))
#define NL_WARP_MIN(x) (__reduce_min_sync(0xffffffffu, (x)))
#define NL_WARP_MAX(x) (__reduce_max_sync(0xffffffffu, (x)))
#define NL_SYNCWARP() __syncwarp()
#else
#define NL_ANY(x) (x)
#define NL_WARP_MIN(x) (x)
#define NL_WARP_MAX(x) (x)
#define NL_SYNCWARP() ((void)0)
#endif

namespace nl {

NL_HD float nl_sqrtf(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);      // float32(math.Sqrt(float64(x))) == correctly rounded sqrtf
#else
    return sqrtf(x);
#endif
}
NL_HD float nl_divf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
NL_HD float nl_mulf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);    // never contracted into an FMA
#else
    return a * b;
#endif
}
NL_HD float nl_addf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
NL_HD float nl_subf(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}

// Stacking modes and weighting modes, numbered like the reference (stack.go:33-42, 57-63).
enum StackMode { ST_MEDIAN = 0, ST_MEAN = 1, ST_SIGMA = 2, ST_WINSOR = 3, ST_MAD = 4, ST_LINFIT = 5, ST_AUTO = 6 };

// stack.go:45-55
NL_HD int auto_select_mode(int n_frames) {
    if (n_frames >= 25) return ST_LINFIT;
    if (n_frames >= 15) return ST_WINSOR;
    if (n_frames >= 6) return ST_SIGMA;
    return ST_MEAN;
}

NL_HD int lowest_bit(unsigned x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// Column memory.  On the device a column lives in shared memory and is addressed by 32-bit
// shared-window byte addresses with immediate element offsets (one LDS per access, no address
// arithmetic); on the host (tests) it is a plain float array.  The accesses of the quick-select are
// `volatile` asm so that loads and stores keep their program order.
// ---------------------------------------------------------------------------------------------
template <int S>
struct ColMem {
#if defined(__CUDA_ARCH__)
    typedef unsigned pos_t;
    static __device__ __forceinline__ pos_t at(const float *a) { return (unsigned)__cvta_generic_to_shared(a); }
    template <int J>
    static __device__ __forceinline__ float ld(pos_t p) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(p), "n"(J * S * 4));
        return v;
    }
    static __device__ __forceinline__ void st(pos_t p, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(p), "f"(v)); }
    static __device__ __forceinline__ pos_t add(pos_t p, int elems) { return p + (unsigned)(elems * (S * 4)); }
    // raw position arithmetic: one element = UNIT raw units (bytes here), so that the bookkeeping of the
    // quick-select needs no scaling of pointer differences
    static constexpr int UNIT = S * 4;
    static __device__ __forceinline__ pos_t add_raw(pos_t p, int raw) { return p + (unsigned)raw; }
    static __device__ __forceinline__ int diff_raw(pos_t a, pos_t b) { return (int)(a - b); }
#else
    typedef float *pos_t;
    static pos_t at(const float *a) { return const_cast<float *>(a); }
    template <int J>
    static float ld(pos_t p) { return p[J * S]; }
    static void st(pos_t p, float v) { *p = v; }
    static pos_t add(pos_t p, int elems) { return p + elems * S; }
    static constexpr int UNIT = S;                         // raw units = floats on the host
    static pos_t add_raw(pos_t p, int raw) { return p + raw; }
    static int diff_raw(pos_t a, pos_t b) { return (int)(a - b); }
#endif
};

// ---------------------------------------------------------------------------------------------
// Quick-select, exact permutation of qsort.go:94-126 (Hoare partition, pivot a[(l+r)>>1]).
// k is 1-based.
//
// SIMT form.  The 32 lanes of a warp run 32 different columns, so the nested data-dependent loops
// of the reference are flattened into ONE loop of branch-free steps that all lanes execute until
// the slowest is done.  The two scans of a Hoare round are independent (nothing is stored between
// them), so they advance in lock step.  Every step looks at a window of QW samples under each
// scan pointer: a pointer that is not at a stop jumps to the first stop inside its window (or
// over the whole window); when both pointers are at their stops and have not crossed, the step
// swaps the two samples and keeps scanning the rest of the two windows.  The windows were loaded
// before the swap was stored, which is only a problem when the pointers are closer than a window
// (`near`): then the step advances by the one slot the swap itself consumes and the next step
// re-reads memory.  Window slots beyond a scan's guaranteed stop (the pivot slot, or a slot
// swapped earlier) may lie outside the partition or even the column (callers pad QW-1 slots on
// both sides of the buffers); their flags are never used.  A lane whose scans have crossed idles
// until the next check for finished partitions; a lane whose selection is finished, or that has
// no samples, is parked on slot 0 (l == r, pivot = a[0], which must not be a NaN: callers store a
// 0 there for an empty column).  GATE adds an explicit `active` term for lanes that alias another
// lane's column (narrow tiles).
// ---------------------------------------------------------------------------------------------
#ifndef NL_QW
#define NL_QW 4
#endif
constexpr int QW = NL_QW;           // samples per scan window

template <typename M, int N, int DIR, int J = 0>
struct WindowLoader {
    static NL_HD void run(typename M::pos_t p, float *w) {
        w[J] = M::template ld<DIR * J>(p);
        WindowLoader<M, N, DIR, J + 1>::run(p, w);
    }
};
template <typename M, int N, int DIR>
struct WindowLoader<M, N, DIR, N> {
    static NL_HD void run(typename M::pos_t, float *) {}
};
template <typename M, int N, int DIR>
NL_HD void load_window(typename M::pos_t p, float *w) { WindowLoader<M, N, DIR>::run(p, w); }

#ifndef NL_QSTEPS
#define NL_QSTEPS 6
#endif
constexpr int QSTEPS = NL_QSTEPS;   // steps between two checks for crossed scans

template <int S, bool GATE = false>
NL_HD float qselect(float *a, int n, int k) {
    typedef ColMem<S> M;
    typedef typename M::pos_t P;
    constexpr int U = M::UNIT;                               // raw position units per element (power of two)
    P left = M::at(a), right = M::add(left, n > 0 ? n - 1 : 0);
    bool active = n > 1;
    P l = left, r = right;
    float pivot = M::template ld<0>(M::add(left, (n > 0 ? n - 1 : 0) >> 1));
    int kraw = k * U;                                        // k, the partition sizes and the scan distances in raw units
    bool any_active = NL_ANY(active);
    while (any_active) {
#pragma unroll
        for (int u = 0; u < QSTEPS; u++) {
            float lw[QW], rw[QW];                            // the two windows
            load_window<M, QW, 1>(l, lw);
            load_window<M, QW, -1>(r, rw);
            const bool sl = lw[0] >= pivot;                  // left scan stops here  (qsort.go:104-108)
            const bool sr = rw[0] <= pivot;                  // right scan stops here (qsort.go:109-113)
            int tl = QW * U, tr = QW * U;                    // first stop among the window slots 1..QW-1
#pragma unroll
            for (int j = QW - 1; j >= 1; j--) {
                tl = (lw[j] >= pivot) ? j * U : tl;
                tr = (rw[j] <= pivot) ? j * U : tr;
            }
            const int d = M::diff_raw(r, l);
            const bool sw = sl & sr & (d > 0);               // both stopped, not crossed: swap (qsort.go:114-115)
            if (sw) { M::st(l, rw[0]); M::st(r, lw[0]); }
            const bool near = d < QW * U;                    // the windows saw slots the swap has just changed
            const bool step = sw & near;
            int dl = (!sl | sw) ? (step ? U : tl) : 0;
            int dr = (!sr | sw) ? (step ? U : tr) : 0;
            if (GATE) { dl = active ? dl : 0; dr = active ? dr : 0; }
            l = M::add_raw(l, dl);
            r = M::add_raw(r, -dr);
        }
        const bool cross = active & (M::template ld<0>(l) >= pivot) & (M::template ld<0>(r) <= pivot) & !(l < r);
        if (NL_ANY(cross)) {                                 // qsort.go:114: partition index = r
            if (cross) {
                const int offset = M::diff_raw(r, left) + U;
                if (kraw <= offset) right = r;
                else { left = M::add_raw(r, U); kraw -= offset; }
                active = left < right;
                // a[(left+right)>>1]: half the distance, rounded down to a whole element
                pivot = M::template ld<0>(M::add_raw(left, (M::diff_raw(right, left) >> 1) & ~(U - 1)));
                l = left;
                r = active ? right : left;
            }
            any_active = NL_ANY(active);                     // lanes only ever finish in here
        }
    }
    return M::template ld<0>(left);
}

// qsort.go:68-82 QSelectMedianFloat32
template <int S, bool GATE = false>
NL_HD float qselect_median(float *a, int n) {
    int k = (n >> 1) + 1;
    float upper = qselect<S, GATE>(a, n, k);
    if (n & 1) return upper;
    float lower = a[0];
    for (int i = 1; i < k - 1; i++) {          // `if a[i]>lower { lower=a[i] }`: a comparison, not fmaxf -- it keeps
        const float v = a[i * S];              // the FIRST of equal values, which decides between -0.0 and +0.0
        lower = v > lower ? v : lower;
    }
    return nl_mulf(0.5f, nl_addf(lower, upper));
}

// ---------------------------------------------------------------------------------------------
// Median by value.  Where only the VALUE of the median matters -- StackMedian (stack.go:274-303) and
// the median of the absolute deviations in StackMADSigma (stack.go:566-572) -- the permutation the
// reference's quick-select leaves behind is irrelevant and any exact selection returns the same
// bits.  This one is built for SIMT: a narrowing level of REGULAR passes that all 32 lanes run in lock
// step without divergence (a second level can be enabled; it measured no faster), then the flattened
// quick-select on the few samples that are left.
// A level sorts 16 evenly spaced samples of the lane's list in registers, picks NP of them around
// the position where the wanted rank should fall as pivots, counts the list against the pivots
// (one pass), and compacts the one interval that contains the wanted rank to the front of the
// buffer (one pass), remembering the largest sample dropped below it (the lower median of an even
// column may be exactly that one).  The list is destroyed, the result is exact for any data: ties,
// infinities, lists that do not shrink (all samples equal) just leave more work to the quick-select.
// ---------------------------------------------------------------------------------------------
NL_HD void cswap_minmax(float &a, float &b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}

// bitonic network on 16 registers (ascending); all indices are compile-time constants
NL_HD void sort16(float (&s)[16]) {
#pragma unroll
    for (int k = 2; k <= 16; k <<= 1) {
#pragma unroll
        for (int st = k >> 1; st >= 1; st >>= 1) {
            const bool mirror = st == (k >> 1);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int lo = ((i & ~(st - 1)) << 1) | (i & (st - 1));
                const int hi = mirror ? (lo ^ (k - 1)) : (lo | st);
                cswap_minmax(s[lo], s[hi]);
            }
        }
    }
}

// s[idx] for a run-time idx, -inf / +inf outside 0..15
NL_HD float pick16(const float (&s)[16], int idx) {
    float v = idx < 0 ? -INFINITY : INFINITY;
#pragma unroll
    for (int j = 0; j < 16; j++) v = (idx == j) ? s[j] : v;
    return v;
}

// one narrowing level; `go` lanes take part (w > 16), the others pass through untouched
template <int S, int NP>
NL_HD void narrow_level(float *g, bool go, int &w, int &r, float &below) {
    float s[16];
    const int wl = go ? w : 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s[j] = g[(go ? (int)(((2 * j + 1) * wl) >> 5) : 0) * S];   // centre of the j-th sixteenth
    sort16(s);
    const int t = go ? ((r - 1) * 16) / wl : 0;             // sample position of the wanted rank
    float piv[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) piv[q] = pick16(s, t + 2 * q - (NP - 1));   // NP=2: t-1,t+1;  NP=4: t-3,t-1,t+1,t+3
    int cnt[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) cnt[q] = 0;
#pragma unroll 4
    for (int i = 0; i < wl; i++) {
        const float v = g[i * S];
#pragma unroll
        for (int q = 0; q < NP; q++) cnt[q] += (v < piv[q]) ? 1 : 0;
    }
    // the interval [lo, hi) that holds rank r; cb = samples below lo
    float lo = -INFINITY, hi = INFINITY;
    int cb = 0;
    bool found = false;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        if (!found) {
            if (r <= cnt[q]) { hi = piv[q]; found = true; }
            else { lo = piv[q]; cb = cnt[q]; }
        }
    }
    // compaction, branch free: four samples are loaded ahead of the stores (the write slot never
    // overtakes the read slot), every sample is stored at the write slot and the slot only advances
    // for samples inside the interval
    int w2 = 0;
    for (int i = 0; i < wl; i += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = g[(i + u) * S];      // may read up to 3 slots past wl (inside the padded buffer)
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool ok = i + u < wl;
            const bool lt = v[u] < lo;
            const bool in = ok & !lt & (!found | (v[u] < hi));   // open-ended last interval: keeps +inf samples
            below = (ok & lt) ? fmaxf(below, v[u]) : below;
            if (ok) g[w2 * S] = v[u];                            // (w2 <= i + u: an already consumed slot, or the sample's own)
            w2 += in ? 1 : 0;
        }
    }
    if (go) { w = w2; r -= cb; }
}

template <int S, bool GATE = false>
NL_HD float median_by_value(float *g, int n) {
    const int k = (n >> 1) + 1;
    int w = n, r = k;
    float below = -INFINITY;
#ifndef NL_MED_NP1
#define NL_MED_NP1 6
#endif
#ifndef NL_MED_NP2
#define NL_MED_NP2 4
#endif
#ifndef NL_MED_T2
#define NL_MED_T2 0
#endif
    if (NL_ANY(w > 48)) narrow_level<S, NL_MED_NP1>(g, w > 48, w, r, below);
    if (NL_MED_T2 > 0 && NL_ANY(w > NL_MED_T2)) narrow_level<S, NL_MED_NP2>(g, w > NL_MED_T2, w, r, below);
    const float upper = qselect<S, GATE>(g, w, r);           // r-th smallest of what is left
    if (n & 1) return upper;
    float lower = below;                                      // rank r-1: inside the list, or the largest sample dropped below it
    for (int i = 0; i < r - 1; i++) lower = fmaxf(lower, g[i * S]);
    return nl_mulf(0.5f, nl_addf(lower, upper));
}

// stats.go:246-261 MeanStdDev: two sequential fp32 sums in buffer order, population sigma.
template <int S>
NL_HD void mean_stddev(const float *a, int n, float &mean, float &sd) {
    float s = 0.0f;
#pragma unroll 8
    for (int i = 0; i < n; i++) s = nl_addf(s, a[i * S]);
    float fn = (float)n;
    float m = nl_divf(s, fn);
    float v = 0.0f;
#pragma unroll 8
    for (int i = 0; i < n; i++) {
        float d = nl_subf(a[i * S], m);
        v = nl_addf(v, nl_mulf(d, d));
    }
    v = nl_divf(v, fn);
    mean = m;
    sd = nl_sqrtf(v);
}

// The clip loop shared by the sigma and winsor variants (stack.go:411-424, 495-514, 674-689,
// 779-798): an out-of-bounds sample is overwritten by the last one, the slice shrinks and slot j
// is tested again.  W: weights travel with the values -- as the FRAME INDEX of every sample (IDX =
// uint8_t up to 256 frames, else uint16_t), a quarter / half of the shared memory an fp32 copy of
// the weights would take; the weight itself is looked up when the weighted mean is formed.
// SIMT form: samples that are in bounds are only ever stepped over by the reference loop, so the
// buffer is scanned 32 slots at a time into a bit mask of out-of-bounds slots (regular, branch
// free), and only the set bits are then resolved one by one in ascending order exactly like the
// reference does (replace by the last sample, re-test the slot, stop at the shrinking end).
// Reads up to 31 slots past `cur`: buffers are padded to a multiple of 32 slots.
// One-sided scan: right after the quick-select the buffer is partitioned around slot km1 = n>>1 (the
// upper median): slots below hold values <= median, slots from km1 on values >= median.  With
// non-negative sigmas (lo <= median <= hi) a slot below km1 can only violate the lower bound and a slot
// from km1 on only the upper bound, so whole 32-slot blocks on either side need one comparison per
// sample instead of two (`onesided` is warp-uniform; blocks that straddle some lane's km1 test both).
template <int SIDE, int S>
NL_HD unsigned clip_mask32(const float *g, int b, float lo, float hi) {
    unsigned bad = 0;
#pragma unroll
    for (int u = 0; u < 32; u++) {
        const float v = g[(b + u) * S];
        const bool o = SIDE < 0 ? (v < lo) : (SIDE > 0 ? (v > hi) : ((v < lo) | (v > hi)));
        bad |= o ? (1u << u) : 0u;
    }
    return bad;
}

template <int S, bool W, typename IDX>
NL_HD int clip_pass(float *g, IDX *gw, int cur, float lo, float hi, int &ncl, int &nch, int km1 = 0, bool onesided = false) {
    const int kmin = onesided ? NL_WARP_MIN(cur > 0 ? km1 : 0x7fffffff) : 0;
    const int kmax = onesided ? NL_WARP_MAX(cur > 0 ? km1 : 0) : 0x7fffffff;
    for (int b = 0; NL_ANY(b < cur); b += 32) {
        unsigned bad;
        if (b + 32 <= kmin) bad = clip_mask32<-1, S>(g, b, lo, hi);
        else if (b >= kmax) bad = clip_mask32<1, S>(g, b, lo, hi);
        else bad = clip_mask32<0, S>(g, b, lo, hi);
        const int rem = cur - b;
        bad &= rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
        while (NL_ANY(bad != 0)) {
            if (bad != 0) {
                const int j = b + lowest_bit(bad);
                if (j >= cur) {
                    bad = 0;                            // the slice has shrunk below the remaining slots
                } else {
                    if (g[j * S] < lo) ncl++; else nch++;   // a set bit means slot j is out of bounds
                    cur--;
                    const float t = g[cur * S];         // g[j] = g[last]; shrink; test slot j again
                    g[j * S] = t;
                    if (W) gw[j * S] = gw[cur * S];
                    if (!((t < lo) | (t > hi)) | (j >= cur)) bad &= bad - 1;
                }
            }
        }
    }
    return cur;
}

// weighted mean of the survivors in buffer order (stack.go:518-524, 802-808); gw = frame index of
// every surviving sample, wtab = the per-frame weights
template <int S, typename IDX>
NL_HD float weighted_mean(const float *g, const IDX *gw, const float *wtab, int cur) {
    float ws = 0.0f, wsum = 0.0f;
    for (int i = 0; i < cur; i++) {
#if defined(__CUDA_ARCH__)
        float w = __ldg(wtab + gw[i * S]);
#else
        float w = wtab[gw[i * S]];
#endif
        ws = nl_addf(ws, nl_mulf(g[i * S], w));
        wsum = nl_addf(wsum, w);
    }
    return nl_divf(ws, wsum);
}

// inner winsorisation loop (stack.go:649-672, 754-777).  The reference clamps a COPY of the column
// again and again; a composition of clamps onto intervals is itself a clamp, onto
// [clamp(L,lo,hi), clamp(H,lo,hi)], and clamping never rounds, so the copy after any number of rounds
// is clamp(g[i], L, H) with the cumulative bounds (L, H) -- recomputed on the fly here, which saves
// the second shared-memory buffer (and its traffic) the copy would need.  `changed` counts, like the
// reference, the samples the CURRENT round moved.
// clamp as two min/max instructions.  Against the reference's `if w<lo {w=lo} else if w>hi {w=hi}` this
// can only differ in the sign of a zero (fmax(-0,+0)), which neither the strict comparisons that count
// `changed` nor the sums of squares below can see.
NL_HD float clampmm(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// One-sided clamps: winsor_sigma runs right after the quick-select, when the buffer is partitioned around slot
// km1 = cur >> 1 (the upper median): slots below hold values <= median, slots from km1 on values >= median.  All bounds
// of all rounds satisfy L, lo <= median <= H, hi, so below km1 only the lower bound can act and from km1 on only the
// upper one -- one FMNMX and one compare per sample instead of four and one.  `changed` (the samples the round moved:
// previous copy < lo or > hi) becomes: sample < lo counted when the lower bound rose (L < lo), sample > hi when the
// upper one fell.  kmin / kmax: warp-uniform ends of the two one-sided ranges; the slots between them do both sides.
template <int S>
NL_HD float winsor_sigma(const float *g, int cur, float median, float sd, int km1) {
    float L = -INFINITY, H = INFINITY;
    const int kmin = NL_WARP_MIN(cur > 0 ? km1 : 0x7fffffff), kmax = NL_WARP_MAX(cur > 0 ? km1 : 0);
    int e0 = cur < kmin ? cur : kmin;                       // [0, e0): lower bound only
    int e1 = cur < kmax ? cur : kmax;                       // [e0, e1): both;  [e1, cur): upper bound only
    if (!(fabsf(median) < INFINITY)) { e0 = 0; e1 = cur; }  // a median that overflowed (or NaN) orders nothing: both sides everywhere
    for (;;) {
        const float lo = nl_subf(median, nl_mulf(1.5f, sd));
        const float hi = nl_addf(median, nl_mulf(1.5f, sd));
        const bool rise = L < lo, fall = H > hi;            // which bound moves this round
        const float Ln = clampmm(L, lo, hi), Hn = clampmm(H, lo, hi);
        int changed = 0;
        // clamp and first sum of MeanStdDev fused: the sum runs over the clamped values in index order
        float s = 0.0f;
        int i = 0;
#pragma unroll 8
        for (; i < e0; i++) {
            const float x = g[i * S];
            changed += (rise & (x < lo)) ? 1 : 0;
            s = nl_addf(s, fmaxf(x, Ln));
        }
        for (; i < e1; i++) {
            const float x = g[i * S];
            changed += ((rise & (x < lo)) | (fall & (x > hi))) ? 1 : 0;
            s = nl_addf(s, clampmm(x, Ln, Hn));
        }
#pragma unroll 8
        for (; i < cur; i++) {
            const float x = g[i * S];
            changed += (fall & (x > hi)) ? 1 : 0;
            s = nl_addf(s, fminf(x, Hn));
        }
        L = Ln;
        H = Hn;
        const float fn = (float)cur;
        const float m = nl_divf(s, fn);
        float var = 0.0f;
        i = 0;
#pragma unroll 8
        for (; i < e0; i++) {
            const float d = nl_subf(fmaxf(g[i * S], L), m);
            var = nl_addf(var, nl_mulf(d, d));
        }
        for (; i < e1; i++) {
            const float d = nl_subf(clampmm(g[i * S], L, H), m);
            var = nl_addf(var, nl_mulf(d, d));
        }
#pragma unroll 8
        for (; i < cur; i++) {
            const float d = nl_subf(fminf(g[i * S], H), m);
            var = nl_addf(var, nl_mulf(d, d));
        }
        var = nl_divf(var, fn);
        const float old = sd;
        sd = nl_mulf(1.134f, nl_sqrtf(var));
        const float factor = nl_divf(fabsf(nl_subf(sd, old)), old);
        if (changed == 0 || factor <= 0.0005f) break;
    }
    return sd;
}

// In-place ascending sort of a column.  The reference sorts with its Hoare quicksort
// (qsort.go:26-32); the sorted array is unique, so any correct sort is bit-exact.  SIMT form: a
// bitonic sorting network in its ascending-only formulation (every compare-exchange puts the
// smaller sample at the lower index: the first stage of each merge mirrors the upper half instead
// of sorting it downwards), which makes slots at or beyond n behave like +infinity without being
// stored: an exchange whose upper slot is >= n is skipped.  The network is data independent, so the
// 32 lanes (32 columns of different length n <= nmax) run it in lock step without divergence.
// nmax: the longest column of the warp (host: n).
// one stage of the network: compare-exchange partners `lo` (bit s clear) and lo|s, or the mirrored
// partner lo ^ (k-1) in the first stage of a merge; K, SS compile-time -> the index arithmetic folds
template <int S, int K, int SS, int P>
NL_HD void sort_stage(float *a, int n) {
    constexpr bool mirror = SS == (K >> 1);
#pragma unroll 8
    for (int i = 0; i < (P >> 1); i++) {
        const int lo = ((i & ~(SS - 1)) << 1) | (i & (SS - 1));
        const int hi = mirror ? (lo ^ (K - 1)) : (lo | SS);
        if (hi < n) {
            const float x = a[lo * S], y = a[hi * S];
            a[lo * S] = fminf(x, y);
            a[hi * S] = fmaxf(x, y);
        }
    }
}
// Register blocking: the comparators of every stage with distance <= 8 stay inside an aligned block of 16
// consecutive samples, so a lane loads a block into registers (positions >= n as +inf), runs those stages there
// (two FMNMX per comparator, no shared-memory traffic, no guards) and stores the block back.
template <int SS> NL_HD void reg_clean(float (&v)[16]) {          // half-cleaner of distance SS
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int lo = ((i & ~(SS - 1)) << 1) | (i & (SS - 1)), hi = lo | SS;
        const float x = v[lo], y = v[hi];
        v[lo] = fminf(x, y); v[hi] = fmaxf(x, y);
    }
}
template <int K> NL_HD void reg_mirror(float (&v)[16]) {          // first stage of the merge of sorted K/2-runs
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int lo = ((i & ~((K >> 1) - 1)) << 1) | (i & ((K >> 1) - 1)), hi = lo ^ (K - 1);
        const float x = v[lo], y = v[hi];
        v[lo] = fminf(x, y); v[hi] = fmaxf(x, y);
    }
}
// FULL: sort every block of 16 (merges 2, 4, 8, 16); else: the last four stages (distance 8, 4, 2, 1) of a wider merge
template <int S, bool FULL>
NL_HD void sort_blocks16(float *a, int n, int nmax) {
    for (int b = 0; b < nmax; b += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = b + j < n ? a[(b + j) * S] : INFINITY;
        if (FULL) {
            reg_mirror<2>(v);
            reg_mirror<4>(v); reg_clean<1>(v);
            reg_mirror<8>(v); reg_clean<2>(v); reg_clean<1>(v);
            reg_mirror<16>(v); reg_clean<4>(v); reg_clean<2>(v); reg_clean<1>(v);
        } else {
            reg_clean<8>(v); reg_clean<4>(v); reg_clean<2>(v); reg_clean<1>(v);
        }
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (b + j < n) a[(b + j) * S] = v[j];
    }
}
// stages of merge K with distance SS .. 16 in shared memory
template <int S, int K, int SS, int P>
struct WideStages {
    static NL_HD void run(float *a, int n) {
        sort_stage<S, K, SS, P>(a, n);
        WideStages<S, K, (SS >> 1), P>::run(a, n);
    }
};
template <int S, int K, int P>
struct WideStages<S, K, 8, P> {
    static NL_HD void run(float *, int) {}
};
template <int S, int K, int P, bool LIVE = (K <= P)>
struct MergeUp {
    static NL_HD void run(float *a, int n, int nmax) {
        WideStages<S, K, (K >> 1), P>::run(a, n);
        sort_blocks16<S, false>(a, n, nmax);
        MergeUp<S, (K << 1), P>::run(a, n, nmax);
    }
};
template <int S, int K, int P>
struct MergeUp<S, K, P, false> {
    static NL_HD void run(float *, int, int) {}
};
template <int S, int P>
NL_HD void sort_static(float *a, int n, int nmax) {
    sort_blocks16<S, true>(a, n, nmax);
    MergeUp<S, 32, P>::run(a, n, nmax);
}

// q, Q: a tile narrower than the warp leaves 32/S lanes per column; for the long columns that need such tiles the
// Q = 32/S lanes of a column share the sort -- helper q takes the 16-blocks and the comparators congruent to q
// modulo Q (comparators of one stage are independent; with the [sample][S] layout the Q helpers of all S columns
// hit 32 different banks), one __syncwarp between stages.  Every lane passes its column's n (not its own sample
// count).  Host and 32-pixel tiles: q = 0, Q = 1.
template <int S, bool FULL>
NL_HD void sort_blocks16_coop(float *a, int n, int nmax, int q, int Q) {
    for (int b = 16 * q; b < nmax; b += 16 * Q) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = b + j < n ? a[(b + j) * S] : INFINITY;
        if (FULL) {
            reg_mirror<2>(v);
            reg_mirror<4>(v); reg_clean<1>(v);
            reg_mirror<8>(v); reg_clean<2>(v); reg_clean<1>(v);
            reg_mirror<16>(v); reg_clean<4>(v); reg_clean<2>(v); reg_clean<1>(v);
        } else {
            reg_clean<8>(v); reg_clean<4>(v); reg_clean<2>(v); reg_clean<1>(v);
        }
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (b + j < n) a[(b + j) * S] = v[j];
    }
    NL_SYNCWARP();
}

template <int S>
NL_HD void sort_column(float *a, int n, int nmax, int q = 0, int Q = 1) {
    // (no NaNs and min/max instead of a swap: an exchange of equal values or of -0/+0 cannot be seen
    // in the sorted sequence of values)
    if (Q == 1) {
        if (nmax <= 32) { sort_static<S, 32>(a, n, nmax); return; }
        if (nmax <= 64) { sort_static<S, 64>(a, n, nmax); return; }
        if (nmax <= 128) { sort_static<S, 128>(a, n, nmax); return; }
        if (nmax <= 256) { sort_static<S, 256>(a, n, nmax); return; }
    }
    int P = 32;
    while (P < nmax) P <<= 1;
    sort_blocks16_coop<S, true>(a, n, nmax, q, Q);
    for (int k = 32; k <= P; k <<= 1) {
        for (int s = k >> 1; s >= 16; s >>= 1) {
            const bool mirror = s == (k >> 1);
#pragma unroll 4
            for (int i = q; i < (P >> 1); i += Q) {
                const int lo = ((i & ~(s - 1)) << 1) | (i & (s - 1));
                const int hi = mirror ? (lo ^ (k - 1)) : (lo | s);
                if (hi < n) {
                    const float x = a[lo * S], y = a[hi * S];
                    a[lo * S] = fminf(x, y);
                    a[hi * S] = fmaxf(x, y);
                }
            }
            NL_SYNCWARP();
        }
        sort_blocks16_coop<S, false>(a, n, nmax, q, Q);
    }
}

// MeanStdDev of xs = 0,1,..,n-1 (stats.go:246-261 applied to StackLinearFit's xs, stack.go:836-839)
NL_HD void ramp_mean_stddev(int n, float &mean, float &sd) {
    float s = 0.0f;
    for (int i = 0; i < n; i++) s = nl_addf(s, (float)i);
    float fn = (float)n;
    float m = nl_divf(s, fn);
    float v = 0.0f;
    for (int i = 0; i < n; i++) {
        float d = nl_subf((float)i, m);
        v = nl_addf(v, nl_mulf(d, d));
    }
    v = nl_divf(v, fn);
    mean = m;
    sd = nl_sqrtf(v);
}

// ---------------------------------------------------------------------------------------------
// The reducers.  g: gathered non-NaN samples in frame order, cur > 0 of them.  Scratch buffers as
// noted.  Clip counters accumulate into ncl / nch.
// ---------------------------------------------------------------------------------------------

// stack.go:372-436 StackSigma / stack.go:442-531 StackSigmaWeighted.  In the weighted variant the
// quick-select permutes the values but not the weights (stack.go:487 hands it gatheredCur only);
// the weights move in the clip loop alone.  Reproduced as is.
// All 32 lanes of a warp call these together (the quick-select and the clip pass vote); `cur` may
// be 0 for a lane without samples, whose result is then meaningless.  A lane that has finished
// keeps walking through the remaining passes of its neighbours with an empty column.
template <int S, bool W, typename IDX>
NL_HD float reduce_sigma(float *g, IDX *gw, const float *wtab, int &cur, float sig_lo, float sig_hi, int &ncl, int &nch,
                 int max_passes = 0, bool *pending = nullptr) {
    // max_passes > 0: stop after that many clipping passes; *pending tells which columns are not finished.
    // A column's state between passes is exactly (g[0..cur), gw[0..cur), cur): calling again resumes it.
    bool done = cur == 0;
    float result = 0.0f;
    int pass = 0;
    while (NL_ANY(!done)) {
        if (max_passes > 0 && pass == max_passes) break;
        pass++;
        const int m = done ? 0 : cur;
        const float median = qselect_median<S, (S < 32)>(g, m);
        float mean, sd;
        mean_stddev<S>(g, m, mean, sd);
        const float lo = nl_subf(median, nl_mulf(sig_lo, sd));
        const float hi = nl_addf(median, nl_mulf(sig_hi, sd));
        const int left = clip_pass<S, W, IDX>(g, gw, m, lo, hi, ncl, nch, m >> 1, sig_lo >= 0.0f && sig_hi >= 0.0f);
        if (!done) {
            if (left == cur || left <= 1) {
                result = W ? weighted_mean<S, IDX>(g, gw, wtab, left) : mean;
                done = true;
            }
            cur = left;
        }
    }
    if (pending) *pending = !done;
    return result;
}

// stack.go:611-705 StackWinsorSigma / stack.go:710-829 StackWinsorSigmaWeighted
template <int S, bool W, typename IDX>
NL_HD float reduce_winsor(float *g, IDX *gw, const float *wtab, int &cur, float sig_lo, float sig_hi, int &ncl, int &nch,
                 int max_passes = 0, bool *pending = nullptr) {
    // max_passes > 0: stop after that many clipping passes; *pending tells which columns are not finished.
    // A column's state between passes is exactly (g[0..cur), gw[0..cur), cur): calling again resumes it.
    bool done = cur == 0;
    float result = 0.0f;
    int pass = 0;
    while (NL_ANY(!done)) {
        if (max_passes > 0 && pass == max_passes) break;
        pass++;
        const int m = done ? 0 : cur;
        const float median = qselect_median<S, (S < 32)>(g, m);
        float mean, sd;
        mean_stddev<S>(g, m, mean, sd);
        if (NL_ANY(m > 0)) { const float wsd = winsor_sigma<S>(g, m, median, sd, m >> 1); if (m > 0) sd = wsd; }
        const float lo = nl_subf(median, nl_mulf(sig_lo, sd));
        const float hi = nl_addf(median, nl_mulf(sig_hi, sd));
        const int left = clip_pass<S, W, IDX>(g, gw, m, lo, hi, ncl, nch, m >> 1, sig_lo >= 0.0f && sig_hi >= 0.0f);
        if (!done) {
            if (left == cur || left <= 1) {
                result = W ? weighted_mean<S, IDX>(g, gw, wtab, left) : mean;
                done = true;
            }
            cur = left;
        }
    }
    if (pending) *pending = !done;
    return result;
}

// stack.go:536-605 StackMADSigma (single pass; 0/0 -> NaN when everything is clipped, as in Go)
template <int S>
NL_HD float reduce_mad(float *g, float *ad, int cur, float sig_lo, float sig_hi, int &ncl, int &nch) {
    float median = qselect_median<S, (S < 32)>(g, cur);
    for (int i = 0; i < cur; i++) ad[i * S] = fabsf(nl_subf(g[i * S], median));
    if (cur == 0 && S == 32) ad[0] = 0.0f;         // an ungated parked lane compares slot 0 with itself: never a NaN
    float mad = median_by_value<S, (S < 32)>(ad, cur);      // only the value matters: ad is scratch
    float sd = nl_mulf(mad, 1.4826f);
    float lo = nl_subf(median, nl_mulf(sig_lo, sd));
    float hi = nl_addf(median, nl_mulf(sig_hi, sd));
    cur = clip_pass<S, false, unsigned char>(g, nullptr, cur, lo, hi, ncl, nch, cur >> 1, sig_lo >= 0.0f && sig_hi >= 0.0f);
    float s = 0.0f;
    for (int i = 0; i < cur; i++) s = nl_addf(s, g[i * S]);
    return nl_divf(s, (float)cur);
}

// stack.go:834-918 StackLinearFit.  ramp[2*c], ramp[2*c+1] = MeanStdDev of 0..c-1 (precomputed per
// length by ramp_mean_stddev; the reference recomputes it per pixel, stats.go:570).
// The reference rejects a sample by overwriting its slot with the current front sample and then
// drops the front slots (stack.go:889-909), and sorts again.  Every surviving value ends up in the
// slice exactly once, so the next round's SORTED array is the current sorted array with the rejected
// slots removed -- a stable compaction, no second sort.  All 32 lanes call this together (nmax = the
// longest column of the warp); cur may be 0.
template <int S>
NL_HD float reduce_linfit(float *g, int &cur, int nmax, const float *ramp, float sig_lo, float sig_hi, int &ncl, int &nch,
                          bool sorted = false, int max_iters = 0, bool *pending = nullptr) {
    // max_iters > 0: stop after that many rejection rounds (< 0: right after the sort); *pending tells which columns are
    // not finished.  A column's state between rounds is its sorted survivors g[0..cur): calling again with sorted = true
    // resumes it.
    if (!sorted) {
#if defined(__CUDA_ARCH__)
        if (S < 32) {
            // narrow tile: the 32/S lanes that alias a column sort it together (they hold no samples of their own)
            const int lane = threadIdx.x & 31;
            const int n_col = __shfl_sync(0xffffffffu, cur, lane % S);
            sort_column<S>(g, n_col, nmax, lane / S, 32 / S);
        } else
#endif
            sort_column<S>(g, cur, nmax);
    }
    float mean = 0.0f;
    bool done = cur == 0;
    if (max_iters < 0) {                                  // sort only: every column with samples is handed on
        if (pending) *pending = !done;
        return mean;
    }
    // sum of the samples in index order: the first chain of MeanStdDev (stats.go:247-250).  After the first
    // round it rides on the compaction of the survivors, which visits them in exactly that order.
    float ysum = 0.0f;
#pragma unroll 8
    for (int i = 0; i < cur; i++) ysum = nl_addf(ysum, g[i * S]);
    int round = 0;
    while (NL_ANY(!done)) {
        if (max_iters > 0 && round == max_iters) {
            // a column the last round emptied is not handed on (an empty column means "no samples" to the caller):
            // its next round is mean = 0/0 and nothing left to reject
            if (!done && cur == 0) { mean = nl_divf(ysum, 0.0f); done = true; }
            break;
        }
        round++;
        const int m = done ? 0 : cur;
        // LinearRegression(xs, ys), stats.go:569-586
        const float xm = ramp[2 * m], xsd = ramp[2 * m + 1];
        // the second pass of MeanStdDev(ys) (stats.go:251-259) with the covariance sum of LinearRegression
        // (stats.go:575-579) riding on it: independent sequential chains, each in the reference's order
        const float fm = (float)m;
        const float ym = nl_divf(ysum, fm);
        float yvar = 0.0f, corr = 0.0f, fi = 0.0f;       // fi = float32(i), exact below 2^24
#pragma unroll 8
        for (int i = 0; i < m; i++) {
            const float d = nl_subf(g[i * S], ym);
            yvar = nl_addf(yvar, nl_mulf(d, d));
            corr = nl_addf(corr, nl_mulf(nl_subf(fi, xm), d));
            fi += 1.0f;
        }
        const float ysd = nl_sqrtf(nl_divf(yvar, fm));
        corr = nl_divf(corr, nl_mulf(nl_mulf(xsd, ysd), nl_addf(fm, 1.0f)));
        const float slope = nl_divf(nl_mulf(corr, ysd), xsd);
        const float icpt = nl_subf(ym, nl_mulf(slope, xm));
        // mean absolute residual, stack.go:878-886
        float sigma = 0.0f;
        fi = 0.0f;
#pragma unroll 4
        for (int i = 0; i < m; i++) {
            const float lin = nl_addf(nl_mulf(fi, slope), icpt);
            sigma = nl_addf(sigma, fabsf(nl_subf(g[i * S], lin)));
            fi += 1.0f;
        }
        sigma = nl_divf(sigma, fm);
        // rejection (stack.go:889-909) fused with the compaction of the survivors and their sum
        int w = 0;
        const float lob = nl_mulf(sig_lo, sigma), hib = nl_mulf(sig_hi, sigma);
        float nsum = 0.0f;
        fi = 0.0f;
#pragma unroll 4
        for (int i = 0; i < m; i++) {
            const float v = g[i * S];
            const float lin = nl_addf(nl_mulf(fi, slope), icpt);
            fi += 1.0f;
            const bool low = nl_subf(lin, v) > lob;
            const bool high = !low && nl_subf(v, lin) > hib;
            ncl += low ? 1 : 0;
            nch += high ? 1 : 0;
            if (!(low | high)) { g[w * S] = v; w++; nsum = nl_addf(nsum, v); }
        }
        if (!done) {
            mean = ym;
            if (w == cur || cur < 3) done = true;          // left == 0 || len < 3
            cur = w;
            ysum = nsum;
        }
    }
    if (pending) *pending = !done;
    return mean;
}

}  // namespace nl
