// nl_stack_kernel.cuh -- the column kernel of the order-statistics stacking modes and its launcher, shared by
// the per-mode translation units (nl_stack_sigma.cu, nl_stack_winsor.cu, nl_stack_linfit.cu: one object file per
// mode family keeps the build parallel) and by nl_stack.cu (job API, mean kernels, dispatch).
#pragma once

#include "nl_internal.h"
#include "nl_column.cuh"

#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched through the runtime, libcuda is not linked)
#include <stdlib.h>
#include <string>
#include <vector>

namespace nl {

struct StackArgs {
    const float *frames;     // [n][stride]
    long long stride;        // elements between frames
    long long pixels;
    int n;
    const float *weights;    // [n] or nullptr
    const float *ramp;       // [2*(n+1)] MeanStdDev of 0..c-1, linear fit only
    float ref_loc, sig_lo, sig_hi;
    float *out;              // [pixels]
    int use_tma;             // tiles are staged by TMA tensor copies (16-byte aligned frame rows)
    float *peer_out[NL_MAX_PEERS];   // further copies of the result (peer-mapped stripes of the gathered image)
    int n_peers;
    unsigned long long *clip;   // [2] low, high
    unsigned long long *tile_counter;   // next tile of the dynamic scheduler (zeroed per launch)
    // Deferral of late passes (sigma / winsorized sigma / linear fit, 32-pixel tiles): pixels need different
    // numbers of clipping passes (rejection rounds), and a warp would walk all its 32 lanes through the passes of
    // its slowest pixel.  A launch therefore stops after `defer_passes` passes and moves the columns that are not
    // finished -- their state is exactly the (permuted or sorted) survivors and their count -- into a pool in
    // global memory; the next launch runs the same kernel over that pool, 32 unfinished columns to a warp, and may
    // defer again into a second pool.  The arithmetic of a column is unchanged.
    int stream_cache_blocks; // streaming linear fit: 32-sample blocks of every column kept in shared memory
    int defer_passes;        // 0: run every column to the end
    int phase;               // 0: tiles of the frame stack; >= 1: tiles of the input pool
    struct Pool {
        float *samples;      // [cap/S][npad][S] sample columns, S = the tile width of the launch
        void *idx;           // weighted modes: the frame-index columns, same layout
        long long *pixel;    // [cap] pixel of a slot
        int *cur;            // [cap] survivors of a slot
        unsigned long long *count;     // slots handed out (may exceed cap: the excess finished in place)
        long long cap;       // slots, a multiple of 32
    } pool_in, pool_out;
    unsigned long long *pool_tile_counter;   // scheduler of the input pool's tiles
};


__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }

// ---- TMA staging of a tile: the frame stack is described by a 2-D tensor map [frame][pixel]; one
// tensor copy moves a [32 frames x 32 pixels] box (32 rows of 128 bytes) straight from HBM/L2 into
// the warp's [sample][lane] slab in shared memory, N/32 boxes per tile, all in flight at once and
// completing on the warp's mbarrier.  Out-of-range pixels and frames arrive as NaN (the map's fill
// mode), which is exactly "no sample" for the reducers.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned mb, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
    return ok != 0;
}
// one [TMA_ROWS frames x 32 pixels] box of the frame stack -> shared memory (SASS: UTMALDG)
constexpr int TMA_ROWS = 32;
__device__ __forceinline__ void tma_box_g2s(unsigned dst, const CUtensorMap *tmap, int pixel0, int frame0, unsigned mb) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(pixel0), "r"(frame0), "r"(mb) : "memory");
}

// Result store.  Multi-GPU: the reassembly of the stacked image is fused into this epilogue -- besides
// the local copy, every warp stores its 128-byte result segment straight into the gathered image of
// each peer GPU (peer-mapped memory over NVLink), so no separate all-gather pass runs afterwards.
__device__ __forceinline__ void store_result(const StackArgs &a, long long p, float v) {
    if (!a.out) return;                                  // count-only run (sigma goal-seek trials): clip totals, no image
    a.out[p] = v;
    for (int e = 0; e < a.n_peers; e++) a.peer_out[e][p] = v;
}

// ---- order-statistics modes ------------------------------------------------------------------
// shared-memory bytes per column slot: the fp32 samples, the MAD scratch column, and (weighted
// modes) the frame index of every sample.  (Measured: moving the index column to an L2-resident scratch in global
// memory gives the weighted modes the seven warps per SM of the unweighted ones, but the clip loop's dependent
// load/store pairs then wait on L2: sigma_w 10.4 -> 14.9 ms, winsor_w 14.9 -> 15.4 ms per 512-row stripe.  Rejected.
// Also measured: no index column at all, the moved slots kept as a (slot, position) table in the slots the clip loop
// frees and the frames rebuilt in one walk per launch -- seven warps, bit-exact, but +23 % instructions in divergent
// code: 9.8 / 13.1 ms at 256 frames, and 1.2-1.5 x SLOWER below 200 frames.  Rejected: profiles/r02_select_levers.md.)
template <int MODE, bool W, typename IDX> struct SlotBytes {
    static constexpr int value = 4 * (MODE == ST_MAD ? 2 : 1) + (W ? (int)sizeof(IDX) : 0);
};

template <int MODE, bool W, int S, typename IDX>
__global__ void __launch_bounds__(256) stack_column_kernel(StackArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n = a.n;
    const int npad = (n + 31) & ~31;                              // clip_pass scans whole 32-slot blocks
    constexpr int SB = SlotBytes<MODE, W, IDX>::value;
    // column element i of this lane's pixel at g[i*S]; lanes beyond a narrow tile alias a valid
    // column but never get samples (cur = 0) and never store
    // (a gap of QW-1 rows in front of, between and behind the warp slabs: the quick-select windows may
    // read, never use, up to QW-1 slots outside a column -- they land in a gap nobody writes, never in
    // another warp's live data)
    constexpr size_t GAP = (size_t)(QW - 1) * S * 4;
    const size_t slab_bytes = (size_t)SB * S * npad;
    char *region = reinterpret_cast<char *>(smem) + GAP + (size_t)warp * (slab_bytes + GAP);
    float *g = reinterpret_cast<float *>(region) + (lane % S);
    float *sc = g + (size_t)S * npad;                             // MAD scratch column
    IDX *gw = reinterpret_cast<IDX *>(region + (size_t)4 * (MODE == ST_MAD ? 2 : 1) * S * npad) + (lane % S);
    (void)sc; (void)gw;

    constexpr bool DEFER = MODE == ST_SIGMA || MODE == ST_WINSOR || MODE == ST_LINFIT;
    const bool pool_phase = DEFER && a.phase >= 1;
    long long pool_slots = 0;
    if (pool_phase) {
        const unsigned long long c = *a.pool_in.count;
        pool_slots = c < (unsigned long long)a.pool_in.cap ? (long long)c : a.pool_in.cap;
    }
    const long long tiles = pool_phase ? (pool_slots + S - 1) / S : (a.pixels + S - 1) / S;
    int ncl = 0, nch = 0;

    // one mbarrier per warp for the TMA staging, in the last 64 bytes of the last gap (32-pixel tiles only:
    // at N = 256 the seven slabs and eight gaps fill the 232 448 bytes a CTA may use to the byte; a window
    // load of the last warp that strays there reads the barrier words as meaningless sample data)
    const unsigned mb = smem_u32(reinterpret_cast<char *>(smem) + GAP + (size_t)(blockDim.x >> 5) * (slab_bytes + GAP) - 64) + 8u * warp;
    const bool tma_tiles = S == 32 && a.use_tma;
    unsigned tma_phase = 0;
    if (tma_tiles) {
        if (lane == 0) mbar_init(mb, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // make the init visible to the async proxy
        __syncwarp();
    }

    // dynamic tile scheduler: column work varies from pixel to pixel, so warps pull tiles from a counter
    unsigned long long *const counter = pool_phase ? a.pool_tile_counter : a.tile_counter;
    auto next_tile = [&]() {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(counter, 1ull);
        return (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    for (long long t = next_tile(); t < tiles; t = next_tile()) {
        long long p = t * S + lane;
        bool valid = lane < S && p < a.pixels;
        int cur = 0;
        bool negzero = false;            // median mode: a -0.0 sample makes the SIGN of a zero median depend on the permutation
        if (pool_phase) {
            // a tile of the pool: S deferred columns, laid out [sample][slot] like a slab
            const long long slot = t * S + lane;
            valid = lane < S && slot < pool_slots;
            p = valid ? a.pool_in.pixel[slot] : 0;
            cur = valid ? a.pool_in.cur[slot] : 0;
            const int rows = __reduce_max_sync(0xffffffffu, cur);
            const float *src = a.pool_in.samples + t * ((long long)npad * S) + lane;
            for (int i0 = 0; i0 < rows; i0 += 32) {               // 32 row segments in flight per warp
                float v[32];
#pragma unroll
                for (int u = 0; u < 32; u++) v[u] = lane < S ? ld_stream(src + (long long)(i0 + u) * S) : 0.0f;   // (npad is a multiple of 32)
                if (lane < S) {
#pragma unroll
                    for (int u = 0; u < 32; u++) g[(i0 + u) * S] = v[u];
                }
            }
            if (W && lane < S) {
                const IDX *srcw = reinterpret_cast<const IDX *>(a.pool_in.idx) + t * ((long long)npad * S) + lane;
#pragma unroll 8
                for (int i = 0; i < rows; i++) gw[i * S] = srcw[(long long)i * S];
            }
        } else if (tma_tiles) {
            // (the slab was last touched by this warp's generic-proxy loads and stores: order them
            // before the async-proxy writes of the tensor copies)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_expect_tx(mb, (unsigned)npad * (S * 4));
            __syncwarp();
            const unsigned slab = smem_u32(region);
            for (int j = lane; j * TMA_ROWS < npad; j += 32)
                tma_box_g2s(slab + (unsigned)j * (TMA_ROWS * S * 4), &tmap, (int)(t * S), j * TMA_ROWS, mb);
            while (!mbar_try_wait(mb, tma_phase)) {}
            tma_phase ^= 1;
            // drop the NaNs in frame order, in place (stack.go:380-387); nothing moves until the first NaN
#pragma unroll 8
            for (int k = 0; k < n; k++) {
                const float v = g[k * S];
                if (cur != k) g[cur * S] = v;
                if (W) gw[cur * S] = (IDX)k;
                if (MODE == ST_MEDIAN) negzero |= __float_as_uint(v) == 0x80000000u;
                cur += (v == v) ? 1 : 0;
            }
            if (!valid) cur = 0;      // a job run on fewer pixels than it was sized for: the columns beyond are not NaN-filled
        } else if (valid) {
            // unaligned frame rows or narrow tiles: gather the non-NaN samples through registers, 32 loads
            // in flight per lane, each a 128-byte row segment per warp
            const float *src = a.frames + p;
            int k = 0;
            for (; k + 32 <= n; k += 32) {
                float v[32];
#pragma unroll
                for (int u = 0; u < 32; u++) v[u] = ld_stream(src + (long long)(k + u) * a.stride);
#pragma unroll
                for (int u = 0; u < 32; u++) {
                    g[cur * S] = v[u];
                    if (W) gw[cur * S] = (IDX)(k + u);
                    if (MODE == ST_MEDIAN) negzero |= __float_as_uint(v[u]) == 0x80000000u;
                    cur += (v[u] == v[u]) ? 1 : 0;
                }
            }
            for (; k < n; k++) {
                float v = ld_stream(src + (long long)k * a.stride);
                if (MODE == ST_MEDIAN) negzero |= __float_as_uint(v) == 0x80000000u;
                if (v == v) {
                    g[cur * S] = v;
                    if (W) gw[cur * S] = (IDX)k;
                    cur++;
                }
            }
        }
        if (cur == 0 && lane < S) g[0] = 0.0f;                    // parked lanes compare slot 0 with itself
        __syncwarp();
        const int cur0 = cur;
        bool spilled = false;
        float res;
        if (MODE == ST_MEDIAN) {
            // stack.go:274-303.  Only the value of the median matters -- except for the sign of a zero: with
            // both -0.0 and +0.0 among the samples the reference's result carries the sign its permutation
            // happens to leave at the median slot, so such (rare) tiles take the emulated quick-select.
            if (__any_sync(0xffffffffu, negzero)) res = qselect_median<S, (S < 32)>(g, cur);
            else res = median_by_value<S, (S < 32)>(g, cur);
        } else if (MODE == ST_MAD) {
            res = reduce_mad<S>(g, sc, cur, a.sig_lo, a.sig_hi, ncl, nch);
        } else {
            // sigma, winsorized sigma, linear fit.  (One call site in a loop: a second inlined copy of the reducer
            // costs more in instruction fetch than the rare second round does.)
            int limit = DEFER ? a.defer_passes : 0;
            int c = cur;
            bool mine = true;                       // this lane's column is still to be reduced here
            bool sorted = pool_phase;               // linear fit: pooled columns are sorted already
            res = 0.0f;
            for (;;) {
                bool pending = false;
                float r;
                if (MODE == ST_SIGMA) r = reduce_sigma<S, W, IDX>(g, gw, a.weights, c, a.sig_lo, a.sig_hi, ncl, nch, limit, &pending);
                else if (MODE == ST_WINSOR) r = reduce_winsor<S, W, IDX>(g, gw, a.weights, c, a.sig_lo, a.sig_hi, ncl, nch, limit, &pending);
                else r = reduce_linfit<S>(g, c, __reduce_max_sync(0xffffffffu, c), a.ramp, a.sig_lo, a.sig_hi, ncl, nch, sorted, limit, &pending);
                if (mine) res = r;
                const unsigned pm = limit != 0 ? __ballot_sync(0xffffffffu, pending) : 0u;
                if (pm == 0u) break;
                // hand the unfinished columns to the pool: consecutive slots for this warp's columns
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.pool_out.count, (unsigned long long)__popc(pm));
                base = __shfl_sync(0xffffffffu, base, 0);
                const long long slot = (long long)base + __popc(pm & ((1u << lane) - 1u));
                spilled = pending && slot < a.pool_out.cap;
                if (spilled) {
                    float *dst = a.pool_out.samples + (slot / S) * ((long long)npad * S) + (slot % S);
                    for (int i = 0; i < c; i++) dst[(long long)i * S] = g[i * S];
                    if (W) {
                        IDX *dstw = reinterpret_cast<IDX *>(a.pool_out.idx) + (slot / S) * ((long long)npad * S) + (slot % S);
                        for (int i = 0; i < c; i++) dstw[(long long)i * S] = gw[i * S];
                    }
                    a.pool_out.pixel[slot] = p;
                    a.pool_out.cur[slot] = c;
                }
                // a full pool: the columns that found no slot finish here in a second round, the other lanes parked
                mine = pending && !spilled;
                if (!__any_sync(0xffffffffu, mine)) break;
                if (!mine) c = 0;
                limit = 0;
                sorted = true;
            }
        }
        if (valid && !spilled) store_result(a, p, cur0 == 0 ? a.ref_loc : res);   // stack.go:388-397
        __syncwarp();
    }
    if (MODE >= ST_SIGMA) {
        // clip totals (stack.go:193-198): warp reduce, one atomic pair per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ncl += __shfl_xor_sync(0xffffffffu, ncl, o);
            nch += __shfl_xor_sync(0xffffffffu, nch, o);
        }
        if (lane == 0 && (ncl | nch)) {
            atomicAdd(a.clip + 0, (unsigned long long)ncl);
            atomicAdd(a.clip + 1, (unsigned long long)nch);
        }
    }
}


// ---- linear fit, streaming rejection rounds (narrow tiles: more than 256 frames) -----------------------------------
// The rejection rounds of StackLinearFit (stack.go:849-915) only ever walk a column front to back -- three sequential
// fp32 chains per round (variance + covariance, mean absolute residual, the survivors' sum) -- so, once a column is
// SORTED, they need no per-lane dynamic indexing and no shared-memory residency: this kernel streams the sorted columns
// of a pool from global memory (L2), 32 columns to a warp, lane = column, every load a coalesced row of the pool tile.
// Rejected samples are not compacted away (that would need a per-lane store index): a bit mask of the surviving
// samples lives in shared memory ([word][lane], 4 bytes per 32 samples) and every operation of a sample is predicated on
// its bit; the x coordinate of a sample -- its index among the survivors -- is a running per-lane count.  The
// arithmetic, operation for operation, is reduce_linfit's (nl_column.cuh).  With no column slab in shared memory the
// SM holds 16+ warps instead of 7 x 8 lanes: the long columns that need 8-pixel tiles in stack_column_kernel (a quarter
// of the lanes busy) run with all lanes busy here.
// Launch r: up to `defer_passes` rounds (0: to the end); unfinished columns are written, compacted, to pool_out and
// regrouped by the next launch.
#ifndef NL_LINFIT_STREAM_WARPS
#define NL_LINFIT_STREAM_WARPS 6
#endif
constexpr int LINFIT_STREAM_WARPS = NL_LINFIT_STREAM_WARPS;
constexpr int LINFIT_POOL_S = 8;           // pool tile width of the streaming linear fit: 8 pixels = one 32-byte sector per sample row

// Sort of the long columns for the streaming linear fit: ONE COLUMN PER WARP, the column in registers.  Lane l holds the
// R consecutive samples l*R .. l*R+R-1 (32 R >= n_frames; NaNs and the padding enter as +inf and sort to the end, where
// they are dropped), and the warp runs a bitonic sorting network over the 32 R values: the stages whose partners are
// less than R apart are compare-exchanges between a lane's own registers (two FMNMX each), the wider ones one shuffle
// and one FMNMX per sample.  Any correct sort gives the reference's sorted column (qsort.go:26-32 sorts in place; the
// sorted array is unique up to the order of equal values and signed zeros, which no later sum can see).
// A warp sorts the 8 adjacent pixels of a tile one after the other -- the 32-byte sectors its scattered loads touch hold
// exactly those 8 pixels, so every frame sector is fetched from HBM once -- and writes every sorted column, with its
// pixel and sample count, to slot = pixel of the pool ([tile][sample][8]).
// compare-exchange across lanes: the lower lane of a pair keeps the smaller value (one compare, one select)
__device__ __forceinline__ float keep_min_or_max(float v, float o, bool lower) { return ((v > o) == lower) ? o : v; }

// Ascending-only formulation of the bitonic network: the first stage of every merge pairs sample e with its mirror
// e ^ (k-1), the remaining stages are half-cleaners (e, e ^ s); in all of them the lower index gets the minimum, so
// the comparators inside a lane are two FMNMX with nothing to decide, and across lanes a shuffle, a compare and a select.
// The merges wider than a lane's R samples all run the same code -- mirror across lanes, half-cleaners across lanes,
// then the log2(R) half-cleaners inside the lane -- with only the lane masks changing, so they are ONE loop body: the
// whole network is ~0.9 k instructions of code instead of 3.8 k fully unrolled (which stalled 1.9 cycles per issue on
// instruction fetch with twelve warps per SM at different places in it).
template <int R>
__device__ __forceinline__ void half_clean_in_lane(float (&v)[R]) {       // stages s = R/2 ... 1
#pragma unroll
    for (int s = R >> 1; s >= 1; s >>= 1) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((r & s) == 0) {
                const float x = v[r], y = v[r | s];
                v[r] = fminf(x, y);
                v[r | s] = fmaxf(x, y);
            }
        }
    }
}

template <int R>
__device__ __forceinline__ void bitonic_sort_lane_major(float (&v)[R], int lane) {
    // merges of 2 ... R samples: inside the lane, fully static
#pragma unroll
    for (int k = 2; k <= R; k <<= 1) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((r & (k >> 1)) == 0) {
                const float x = v[r], y = v[r ^ (k - 1)];
                v[r] = fminf(x, y);
                v[r ^ (k - 1)] = fmaxf(x, y);
            }
        }
#pragma unroll
        for (int s = k >> 2; s >= 1; s >>= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if ((r & s) == 0) {
                    const float x = v[r], y = v[r | s];
                    v[r] = fminf(x, y);
                    v[r | s] = fmaxf(x, y);
                }
            }
        }
    }
    // merges of 2R ... 32R samples: kl = lanes per merge
#pragma unroll 1
    for (int kl = 2; kl <= 32; kl <<= 1) {
        {   // mirror across lanes: sample (lane, r) meets (lane ^ (kl - 1), R-1-r)
            const int lm = kl - 1;
            const bool lower = (lane & (kl >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < R / 2; r++) {
                const float o0 = __shfl_xor_sync(0xffffffffu, v[R - 1 - r], lm);
                const float o1 = __shfl_xor_sync(0xffffffffu, v[r], lm);
                v[r] = keep_min_or_max(v[r], o0, lower);
                v[R - 1 - r] = keep_min_or_max(v[R - 1 - r], o1, lower);
            }
        }
#pragma unroll 1
        for (int ls = kl >> 2; ls >= 1; ls >>= 1) {                      // half-cleaners across lanes
            const bool lower = (lane & ls) == 0;
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = keep_min_or_max(v[r], __shfl_xor_sync(0xffffffffu, v[r], ls), lower);
        }
        half_clean_in_lane<R>(v);
    }
}

// ---- asynchronous staging of pool rows (cp.async, 16 bytes = half a sample row of one 8-pixel pool tile) -----------
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// mask |= BIT when x > y: a compare and one predicated OR with an immediate
template <unsigned BIT>
__device__ __forceinline__ void or_bit_if_gt(unsigned &mask, float x, float y) {
    asm("{\n .reg .pred p;\n setp.gt.f32 p, %1, %2;\n @p or.b32 %0, %0, %3;\n}" : "+r"(mask) : "f"(x), "f"(y), "n"(BIT));
}
template <int J>
struct RejectTests {       // samples J .. 31 of a block: the two rejection tests of StackLinearFit, bits into the masks
    static __device__ __forceinline__ void run(const float (&v)[32], unsigned w, float slope, float icpt, float lob, float hib,
                                               float &fi, unsigned &lowm, unsigned &highm) {
        const float lin = nl_addf(nl_mulf(fi, slope), icpt);
        or_bit_if_gt<(1u << J)>(lowm, nl_subf(lin, v[J]), lob);        // lin - y > sigmaLow * sigma
        or_bit_if_gt<(1u << J)>(highm, nl_subf(v[J], lin), hib);       // y - lin > sigmaHigh * sigma
        if ((w >> J) & 1u) fi += 1.0f;
        RejectTests<J + 1>::run(v, w, slope, icpt, lob, hib, fi, lowm, highm);
    }
};
template <>
struct RejectTests<32> {
    static __device__ __forceinline__ void run(const float (&)[32], unsigned, float, float, float, float, float &, unsigned &, unsigned &) {}
};
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int LINFIT_SORT_CTAS = 6;         // resident CTAs per SM (shared memory: one tile slab each)
constexpr int LINFIT_SORT_WARPS = 2;        // warps per CTA: they stage one tile together and sort four of its columns each
                                            // (all control flow is CTA-uniform, so the compiler treats the shuffles as converged)
// shared-memory slab of a CTA: the tile's 32 R sample rows of 8 pixels, 4 words of padding after every 32 rows (the
// lanes of a warp then read / write a column's samples l*R + r four to a bank instead of all 32 in one)
template <int R> struct SortSlab { static constexpr int WORDS = 32 * R * 8 + 4 * R; };
__device__ __forceinline__ int sort_slab_word(int e, int c) { return e * 8 + c + 4 * (e >> 5); }

template <int R>
__global__ void __launch_bounds__(LINFIT_SORT_WARPS * 32) linfit_sort_kernel(StackArgs a) {
    constexpr int S = LINFIT_POOL_S;
    constexpr int T = LINFIT_SORT_WARPS * 32;
    extern __shared__ __align__(128) float sort_smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5, tid = threadIdx.x;
    float *slab = sort_smem;
    const unsigned slab_addr = smem_u32(slab);
    const int npad = (a.n + 31) & ~31;
    const long long tiles = (a.pixels + S - 1) / S;
    if (blockIdx.x == 0 && tid == 0) *a.pool_out.count = (unsigned long long)a.pixels;      // slot = pixel: every pixel has a column
    // 16-byte copies need frame rows that start on 16 bytes and whole tiles
    const bool aligned = (a.stride % 4) == 0 && (reinterpret_cast<uintptr_t>(a.frames) & 15) == 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long p0 = t * S;
        const int cols = (int)((a.pixels - p0) < S ? (a.pixels - p0) : S);
        // ---- stage the tile: sample row k = the 8 pixels of frame k, one 32-byte sector
        if (aligned && cols == S) {
            const float *src = a.frames + p0;
            for (int q = tid; q < 2 * a.n; q += T) {
                const int k = q >> 1, h = q & 1;
                cp_async16(slab_addr + 4u * (unsigned)sort_slab_word(k, h * 4), src + (long long)k * a.stride + h * 4);
            }
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            for (int q = tid; q < 8 * a.n; q += T) {
                const int k = q >> 3, c = q & 7;
                slab[sort_slab_word(k, c)] = c < cols ? __ldg(a.frames + (long long)k * a.stride + p0 + c) : 0.0f;
            }
        }
        __syncthreads();
        int my_valid = 0;                                            // lane i < 4: the sample count of this warp's column i
#pragma unroll 1
        for (int i = 0; i < S / LINFIT_SORT_WARPS; i++) {            // (a ragged last tile sorts its zero-filled columns too)
            const int c = wic * (S / LINFIT_SORT_WARPS) + i;
            float v[R];
            int valid = 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int e = lane * R + r;
                const float x = e < a.n ? slab[sort_slab_word(e, c)] : __int_as_float(0x7fc00000);
                const bool ok = x == x;
                valid += ok ? 1 : 0;
                v[r] = ok ? x : INFINITY;                          // (a real +inf sample equals the padding: either may stay)
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
            bitonic_sort_lane_major<R>(v, lane);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int e = lane * R + r;
                if (e < npad) slab[sort_slab_word(e, c)] = v[r];
            }
            if (lane == i) my_valid = valid;
        }
        __syncthreads();
        // ---- the sorted tile goes out as it came in: whole sectors
        float *tile_out = a.pool_out.samples + t * ((long long)npad * S);
        for (int q = tid; q < 2 * npad; q += T) {
            const int k = q >> 1, h = q & 1;
            const float4 x = *reinterpret_cast<const float4 *>(slab + sort_slab_word(k, h * 4));
            __stcg(reinterpret_cast<float4 *>(tile_out + (long long)k * S + h * 4), x);
        }
        const int c_mine = wic * (S / LINFIT_SORT_WARPS) + lane;
        if (lane < S / LINFIT_SORT_WARPS && c_mine < cols) {
            a.pool_out.pixel[p0 + c_mine] = p0 + c_mine;
            a.pool_out.cur[p0 + c_mine] = my_valid;
        }
        __syncthreads();
    }
}

// A warp's view of its 32 columns: four consecutive pool tiles of 8 pixels ([sample][8] each).  Block b = samples
// 32b .. 32b+31 of all 32 columns = 32 rows x 4 tiles x 32 bytes; it is copied into a [row][32 lanes] buffer in shared
// memory by 8 cp.async of 16 bytes per lane (coalesced: whole 32-byte sectors), two buffers deep, so the copy of the
// next block runs while the lanes work on the current one.  Lane l copies, for i = 0..7, the 16-byte half h = l & 1 of
// row 4i + (l >> 3) of tile (l & 7) >> 1: source and destination advance by constants from one chunk to the next.
struct BlockStream {
    const float *lane_src;          // this lane's chunk 0 of block 0
    unsigned lane_dst;              // shared-window address of this lane's chunk 0 in buffer 0
    __device__ __forceinline__ void init(const float *first_tile, int npad, unsigned buf) {
        const int lane = threadIdx.x & 31;
        const int lrow = lane >> 3, t = (lane & 7) >> 1, h = lane & 1;
        lane_src = first_tile + (long long)t * ((long long)npad * 8) + (lrow * 8 + h * 4);
        lane_dst = buf + (unsigned)(lrow * 128 + t * 32 + h * 16);
    }
    __device__ __forceinline__ void issue(int b, int which) const {
        const float *src = lane_src + (long long)b * 256;
        const unsigned dst = lane_dst + (unsigned)which * 4096u;
#pragma unroll
        for (int i = 0; i < 8; i++) cp_async16(dst + (unsigned)i * 512u, src + i * 32);
        cp_async_commit();
    }
};

// walks the blocks of the warp's columns: f(b, v) gets the 32 samples of block b of THIS lane's column.  The first
// `cb` blocks of every column stay in shared memory for the whole launch (FILL: this walk puts them there) -- the kernel
// is bound by the bytes it moves, every cached block saves three reads per round; the others come through a ring of
// STREAM_DEPTH buffers: the copies of the next STREAM_DEPTH-1 blocks are in flight while the lanes work on one, and the
// ring is started before the cached blocks are worked on, so their arithmetic covers its first latency.
#ifndef NL_STREAM_DEPTH
#define NL_STREAM_DEPTH 3
#endif
constexpr int STREAM_DEPTH = NL_STREAM_DEPTH;
template <bool FILL, typename F>
__device__ __forceinline__ void stream_blocks(const BlockStream &bs, float *stage, float *cache, int cb, int b0, int nblk, F &&f) {
    // blocks [b0, nblk) of the columns (warp-uniform: blocks whose samples are rejected in every lane are not read)
    const int lane = threadIdx.x & 31;
    const int first = FILL ? b0 : (cb < nblk ? (cb > b0 ? cb : b0) : nblk);          // blocks [first, nblk) are streamed
#pragma unroll
    for (int d = 0; d < STREAM_DEPTH - 1; d++) {
        if (first + d < nblk) bs.issue(first + d, d); else cp_async_commit();
    }
#pragma unroll 1
    for (int b = b0; b < first; b++) {
        const float *src = cache + b * 1024 + lane;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = src[j * 32];
        f(b, v);
    }
#pragma unroll 1
    for (int b = first; b < nblk; b++) {
        const int rel = b - first, ahead = b + STREAM_DEPTH - 1;
        if (ahead < nblk) bs.issue(ahead, (rel + STREAM_DEPTH - 1) % STREAM_DEPTH); else cp_async_commit();     // (an empty group keeps the count)
        cp_async_wait<STREAM_DEPTH - 1>();
        __syncwarp();
        const float *src = stage + (rel % STREAM_DEPTH) * 1024 + lane;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = src[j * 32];
        if (FILL && b < cb) {
            float *dst = cache + b * 1024 + lane;
#pragma unroll
            for (int j = 0; j < 32; j++) dst[j * 32] = v[j];
        }
        f(b, v);
        __syncwarp();                                               // the buffer is refilled STREAM_DEPTH-1 blocks later
    }
    cp_async_wait<0>();
}

template <int S>
__global__ void __launch_bounds__(LINFIT_STREAM_WARPS * 32) linfit_rounds_kernel(StackArgs a) {
    static_assert(S == 8, "the streaming rounds read pools of 8-pixel tiles");
    extern __shared__ __align__(128) unsigned char stream_smem[];      // per warp: STREAM_DEPTH 4 KiB ring buffers, the cached blocks, the survivor masks [word][lane]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npad = (a.n + 31) & ~31;
    const int nwords = npad >> 5;
    const int cb = a.stream_cache_blocks;
    const size_t per_warp = 4096 * (size_t)(STREAM_DEPTH + cb) + (size_t)nwords * 128;
    float *stage = reinterpret_cast<float *>(stream_smem + (size_t)warp * per_warp);
    float *cache = stage + 1024 * STREAM_DEPTH;
    unsigned *alive = reinterpret_cast<unsigned *>(stream_smem + (size_t)warp * per_warp + 4096 * (size_t)(STREAM_DEPTH + cb)) + lane;   // word w at alive[w * 32]
    const unsigned long long cnt = *a.pool_in.count;
    const long long pool_slots = cnt < (unsigned long long)a.pool_in.cap ? (long long)cnt : a.pool_in.cap;
    const long long groups = (pool_slots + 31) / 32;
    int ncl = 0, nch = 0;
    auto next_group = [&]() {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(a.pool_tile_counter, 1ull);
        return (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    for (long long grp = next_group(); grp < groups; grp = next_group()) {
        const long long slot = grp * 32 + lane;
        const bool valid = slot < pool_slots;
        const long long p = valid ? a.pool_in.pixel[slot] : 0;
        int cur = valid ? a.pool_in.cur[slot] : 0;
        const int cur0 = cur;
        // the warp's four tiles: slots grp*32 .. grp*32+31 (the pool's capacity is a multiple of 32 slots, so all four exist)
        BlockStream bs;
        bs.init(a.pool_in.samples + grp * 4 * ((long long)npad * S), npad, smem_u32(stage));
        const int nblk = (__reduce_max_sync(0xffffffffu, cur) + 31) >> 5;
        // the survivors' masks, and the sum of the samples in index order (first chain of MeanStdDev, stats.go:247-250)
        float ysum = 0.0f;
        stream_blocks<true>(bs, stage, cache, cb, 0, nblk, [&](int b, const float (&v)[32]) {
            const int rem = cur - b * 32;
            const unsigned w = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
            alive[b * 32] = w;
#pragma unroll
            for (int j = 0; j < 32; j++)
                if ((w >> j) & 1u) ysum = nl_addf(ysum, v[j]);
        });
        float mean = 0.0f;
        bool done = cur == 0;
        bool mine = true, spilled = false;
        float res = 0.0f;
        int limit = a.defer_passes;
        int blo = 0, bhi = nblk;                                       // blocks that still hold a survivor in some lane
        for (;;) {
            int round = 0;
            while (__any_sync(0xffffffffu, !done)) {
                if (limit > 0 && round == limit) {
                    // a column the last round emptied is not handed on: its next round is mean = 0/0, nothing left to reject
                    if (!done && cur == 0) { mean = nl_divf(ysum, 0.0f); done = true; }
                    break;
                }
                round++;
                const int m = done ? 0 : cur;
                // LinearRegression(xs, ys), stats.go:569-586, xs = 0..m-1: the second pass of MeanStdDev(ys) with the
                // covariance sum riding on it, each chain in the reference's order
                const float xm = __ldg(a.ramp + 2 * m), xsd = __ldg(a.ramp + 2 * m + 1);
                const float fm = (float)m;
                const float ym = nl_divf(ysum, fm);
                float yvar = 0.0f, corr = 0.0f, fi = 0.0f;
                stream_blocks<false>(bs, stage, cache, cb, blo, bhi, [&](int b, const float (&v)[32]) {
                    const unsigned w = done ? 0u : alive[b * 32];
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if ((w >> j) & 1u) {
                            const float d = nl_subf(v[j], ym);
                            yvar = nl_addf(yvar, nl_mulf(d, d));
                            corr = nl_addf(corr, nl_mulf(nl_subf(fi, xm), d));
                            fi += 1.0f;
                        }
                    }
                });
                const float ysd = nl_sqrtf(nl_divf(yvar, fm));
                corr = nl_divf(corr, nl_mulf(nl_mulf(xsd, ysd), nl_addf(fm, 1.0f)));
                const float slope = nl_divf(nl_mulf(corr, ysd), xsd);
                const float icpt = nl_subf(ym, nl_mulf(slope, xm));
                // mean absolute residual, stack.go:878-886
                float sigma = 0.0f;
                fi = 0.0f;
                stream_blocks<false>(bs, stage, cache, cb, blo, bhi, [&](int b, const float (&v)[32]) {
                    const unsigned w = done ? 0u : alive[b * 32];
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if ((w >> j) & 1u) {
                            const float lin = nl_addf(nl_mulf(fi, slope), icpt);
                            sigma = nl_addf(sigma, fabsf(nl_subf(v[j], lin)));
                            fi += 1.0f;
                        }
                    }
                });
                sigma = nl_divf(sigma, fm);
                // rejection (stack.go:889-909): the rejected samples leave the mask, the survivors' sum rides along.
                // Branch free: the two tests of a sample set bits of a low and a high mask, counted per block.
                const float lob = nl_mulf(a.sig_lo, sigma), hib = nl_mulf(a.sig_hi, sigma);
                float nsum = 0.0f;
                int left = 0;
                int first_b = 0x7fffffff, last_b = -1;                 // this lane's first and last block with a survivor
                fi = 0.0f;
                stream_blocks<false>(bs, stage, cache, cb, blo, bhi, [&](int b, const float (&v)[32]) {
                    const unsigned w = done ? 0u : alive[b * 32];
                    // (the tests of dead samples set garbage bits, masked below; x only advances over survivors)
                    unsigned lowm = 0u, highm = 0u;
                    RejectTests<0>::run(v, w, slope, icpt, lob, hib, fi, lowm, highm);
                    lowm &= w;
                    highm &= w & ~lowm;                                              // `else if`: low wins
                    {
                        const unsigned ok = w & ~(lowm | highm);
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if ((ok >> j) & 1u) nsum = nl_addf(nsum, v[j]);
                    }
                    ncl += __popc(lowm);
                    nch += __popc(highm);
                    const unsigned keep = w & ~(lowm | highm);
                    if (keep != w) alive[b * 32] = keep;
                    left += __popc(keep);
                    if (keep != 0u) { first_b = first_b < b ? first_b : b; last_b = b; }
                });
                {
                    const int nlo = __reduce_min_sync(0xffffffffu, first_b), nhi = __reduce_max_sync(0xffffffffu, last_b) + 1;
                    if (nlo < nhi) { blo = nlo; bhi = nhi; }           // (no survivor anywhere: every lane is done after this round)
                }
                if (!done) {
                    mean = ym;
                    if (left == cur || cur < 3) done = true;          // nothing rejected || len < 3
                    cur = left;
                    ysum = nsum;
                }
            }
            const bool pending = !done;
            if (mine) res = mean;
            const unsigned pm = limit > 0 ? __ballot_sync(0xffffffffu, pending) : 0u;
            if (pm == 0u) break;
            // hand the unfinished columns to the next launch: their survivors, compacted, into consecutive slots
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.pool_out.count, (unsigned long long)__popc(pm));
            base = __shfl_sync(0xffffffffu, base, 0);
            const long long oslot = (long long)base + __popc(pm & ((1u << lane) - 1u));
            spilled = pending && oslot < a.pool_out.cap;
            const long long wslot = spilled ? oslot : 0;
            float *dst = a.pool_out.samples + (wslot / S) * ((long long)npad * S) + (wslot % S);
            int wpos = 0;
            stream_blocks<false>(bs, stage, cache, cb, blo, bhi, [&](int b, const float (&v)[32]) {
                const unsigned w = spilled ? alive[b * 32] : 0u;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if ((w >> j) & 1u) { dst[(long long)wpos * S] = v[j]; wpos++; }
                }
            });
            if (spilled) {
                a.pool_out.pixel[oslot] = p;
                a.pool_out.cur[oslot] = cur;
            }
            // a full pool: the columns that found no slot finish here, the other lanes parked
            mine = pending && !spilled;
            if (!__any_sync(0xffffffffu, mine)) break;
            if (!mine) done = true;
            limit = 0;
        }
        if (valid && !spilled) store_result(a, p, cur0 == 0 ? a.ref_loc : res);     // stack.go:388-397: no samples at all
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ncl += __shfl_xor_sync(0xffffffffu, ncl, o);
        nch += __shfl_xor_sync(0xffffffffu, nch, o);
    }
    if (lane == 0 && (ncl | nch)) {
        atomicAdd(a.clip + 0, (unsigned long long)ncl);
        atomicAdd(a.clip + 1, (unsigned long long)nch);
    }
}

}  // namespace nl

struct nl_stack_job {
    nl_ctx *ctx = nullptr;
    int n = 0;
    long long pixels = 0;             // capacity = element stride between frames
    long long active = 0;             // pixels a run covers (<= pixels; nl_stack_apply runs ragged last stripes in a lane sized once)
    float *frames = nullptr;          // [n][pixels]
    float *out = nullptr;             // [pixels], used by nl_stack_run
    float *weights = nullptr;         // [n]
    float *ramp = nullptr;            // [2*(n+1)]
    bool ramp_ready = false;
    unsigned long long *clip = nullptr;   // [NL_JOB_COUNTERS] device: clip low, clip high, tile counter, then per deferral round
                                          // the slots it handed out and the tile counter of the round that consumes them
    unsigned long long *clip_host = nullptr;   // [2] pinned
    alignas(64) CUtensorMap tmap;     // [n][pixels] fp32, box 32 frames x 32 pixels, NaN fill
    bool tmap_ok = false;
    // the two pools of deferred columns (StackArgs): one allocation, each carved into samples | pixel | cur | indices
    void *pool = nullptr;
    long long pool_cap[2] = {0, 0};   // slots
    int pool_idx_bytes = 0;           // bytes per index element the pools were sized for (0: none)
    bool pool_failed = false;         // allocation failed once: run without deferral
};

namespace nl {

template <int MODE, bool W, int S, typename IDX>
inline int launch_column(nl_stack_job *job, const StackArgs &args) {
    nl_ctx *ctx = job->ctx;
    constexpr int SB = SlotBytes<MODE, W, IDX>::value;
    const size_t gap = (size_t)(QW - 1) * S * sizeof(float);      // in front of, between and behind the warp slabs
    const size_t per_warp = (size_t)SB * S * ((job->n + 31) & ~31) + gap;
    const size_t cap = (size_t)ctx->max_smem_optin;
    if (per_warp + gap > cap) return set_error(NL_E_INVALID, "n_frames %d too large for shared memory at tile %d", job->n, S);
    // warps per CTA (at most 8): the split that puts the most warps on an SM -- the kernel lives on latency hiding
    // across warps, and e.g. 128 frames fit 13 slabs per SM: two CTAs of 6 warps beat one of 8
    const size_t sm_bytes = (size_t)ctx->smem_per_sm;
    int warps = 1, best_total = 0;
    for (int w = 1; w <= 8; w++) {
        const size_t need = per_warp * w + gap;
        if (need > cap) break;
        int ctas = (int)(sm_bytes / (need + 1024));               // 1 KiB per CTA is reserved by the system
        if (ctas > 32) ctas = 32;
        if (ctas * w >= best_total) { best_total = ctas * w; warps = w; }
    }
    const size_t smem = per_warp * warps + gap;                   // (the tile mbarriers live in the tail of the last gap)
    auto kern = stack_column_kernel<MODE, W, S, IDX>;
    NL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 0;
    NL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, warps * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const long long tiles = (args.pixels + S - 1) / S;
    long long grid = (long long)ctx->sm_count * ctas_per_sm;
    const long long need = (tiles + warps - 1) / warps;
    if (grid > need && args.phase == 0) grid = need;       // (the pool's tile count is only known on the device)
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, warps * 32, smem, ctx->stream>>>(args, job->tmap);
    NL_CUDA(cudaGetLastError());
    ctx->launches++;
    return NL_OK;
}

// Deferral schedule: after how many passes (cumulative) each launch hands its unfinished columns on; the launch
// after the last entry runs to the end.  Measured on the synthetic workload (512-row stripe of 256 frames):
//   sigma clipping settles after three passes for four pixels in five and after four for nearly all others:
//     7.5 -> 6.9 ms with {3};  winsorized clipping settles one pass earlier: 11.8 -> 10.7 ms with {2};
//   the linear fit needs 4 .. 40 rejection rounds per pixel (mean 12, slowest of 32 pixels: 26.6), so its columns
//     are regrouped several times.
// nl_ctx_set_tuning(ctx, "defer_passes", "a,b,..") replaces the schedule (A/B measurements, tests); "0" switches the deferral off.
struct DeferSchedule { int n; int at[8]; double frac[2]; };
inline DeferSchedule defer_schedule(int mode, const nl_ctx *ctx, int n_frames = 256) {
    DeferSchedule d{0, {0}, {0.25, 0.0}};
    if (mode == ST_SIGMA) {
        // 256 frames: one regrouping after three passes.  Shorter columns settle earlier and profit from regrouping after
        // every pass (same sample count, synthetic workload): 16 frames 7.48 -> 4.95 ms, 32 frames 5.96 -> 4.54 ms with
        // {1,2,3}; 64 frames 4.97 -> 4.48 ms, 96 frames 4.80 -> 4.65 ms with {2,3,4}; from 128 frames on {3} wins
        // (5.23 vs 5.43 ms).  The first pool of the short schedules takes every column (nearly all are still open).
        if (n_frames <= 48) { d.n = 3; d.at[0] = 1; d.at[1] = 2; d.at[2] = 3; d.frac[0] = 1.0; d.frac[1] = 0.75; }
        else if (n_frames <= 96) { d.n = 3; d.at[0] = 2; d.at[1] = 3; d.at[2] = 4; d.frac[0] = 1.0; d.frac[1] = 0.75; }
        else { d.n = 1; d.at[0] = 3; }
    }
    else if (mode == ST_WINSOR) { d.n = 1; d.at[0] = 2; }     // (measured best or within 3 % of the best from 16 to 256 frames)
    else if (mode == ST_LINFIT) {
        // short columns need fewer rounds (a round rejects at least one sample, and there are fewer to reject), so
        // their columns are regrouped earlier and more often.  Measured on the synthetic workload, same sample count:
        //   32 frames: 9.15 -> 6.80 ms;  64 frames x 6000x4000: 19.9 -> 18.3 ms;  128 and 256 frames: the long schedule wins
        //   (20.2 vs 22.3 ms, 34.2 vs 38.7 ms with the 64-frame schedule)
        static const int at_long[6] = {8, 12, 16, 20, 24, 30}, at_mid[8] = {4, 6, 8, 10, 12, 16, 20, 26}, at_short[7] = {3, 5, 7, 9, 12, 15, 20};
        const int *at = n_frames <= 40 ? at_short : (n_frames <= 96 ? at_mid : at_long);
        d.n = n_frames <= 40 ? 7 : (n_frames <= 96 ? 8 : 6);
        for (int i = 0; i < d.n; i++) d.at[i] = at[i];
        d.frac[0] = n_frames <= 96 ? 1.0 : 0.75;          // (at the first early regrouping nearly every column is still open)
        d.frac[1] = n_frames <= 96 ? 0.85 : 0.5;
    }
    if (ctx->defer_override) {
        d.n = ctx->defer_n;
        for (int i = 0; i < d.n; i++) d.at[i] = ctx->defer_at[i];
        if (d.n > 1) { d.frac[0] = 1.0; d.frac[1] = 1.0; } else if (d.n == 1) { d.frac[0] = d.at[0] < 3 ? 1.0 : 0.25; }
    }
    return d;
}

// Two pools (a launch reads one and fills the other); pool k holds up to frac[k] of the job's pixels, columns that
// find no slot finish in place.  false: no memory for them, run without deferral.
inline bool ensure_pools(nl_stack_job *job, int idx_bytes, const double frac[2], StackArgs::Pool pools[2]) {
    if (job->pool_failed) return false;
    const long long npad = (job->n + 31) & ~31;
    const size_t per_slot = (size_t)npad * (4 + (size_t)idx_bytes) + sizeof(long long) + sizeof(int);
    long long want[2];
    for (int k = 0; k < 2; k++) {
        want[k] = frac[k] > 0.0 ? (((long long)((double)job->pixels * frac[k]) + 31) & ~31ll) : 0;
        if (frac[k] > 0.0 && want[k] < 32) want[k] = 32;
    }
    if (!job->pool || job->pool_idx_bytes < idx_bytes || job->pool_cap[0] < want[0] || job->pool_cap[1] < want[1]) {
        if (job->pool) { cudaStreamSynchronize(job->ctx->stream); cudaFree(job->pool); job->pool = nullptr; }
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); job->pool_failed = true; return false; }
        // never take more than half of what is free; shrink both pools alike
        const double need = (double)(want[0] + want[1]) * (double)per_slot;
        if (need > 0.5 * (double)free_b) {
            const double f = 0.5 * (double)free_b / need;
            for (int k = 0; k < 2; k++) want[k] = (long long)((double)want[k] * f) & ~31ll;
        }
        if (want[0] < 32 || cudaMalloc(&job->pool, (size_t)(want[0] + want[1]) * per_slot + 512) != cudaSuccess) {
            cudaGetLastError();
            job->pool = nullptr;
            job->pool_failed = true;
            return false;
        }
        job->pool_cap[0] = want[0];
        job->pool_cap[1] = want[1];
        job->pool_idx_bytes = idx_bytes;
    }
    char *base = (char *)job->pool;
    for (int k = 0; k < 2; k++) {
        const size_t cap = (size_t)job->pool_cap[k];
        pools[k].samples = (float *)base;
        pools[k].pixel = (long long *)(base + cap * (size_t)npad * 4);
        pools[k].cur = (int *)(base + cap * (size_t)npad * 4 + cap * sizeof(long long));
        pools[k].idx = base + cap * (size_t)npad * 4 + cap * (sizeof(long long) + sizeof(int));
        pools[k].cap = job->pool_cap[k];
        pools[k].count = nullptr;
        base += (cap * per_slot + 255) & ~(size_t)255;
    }
    return true;
}

// One launch, or -- deferral of late passes (see StackArgs) -- launch 0 over the frame stack and one launch per pool
// generation.
// Linear fit of long columns (more than 256 frames, up to 1024): launch 0 = linfit_sort_kernel (one column per warp, in
// registers) hands every sorted column to pool 0; then linfit_rounds_kernel streams the pools, regrouping the unfinished
// columns between launches (schedule of the linear fit, cumulative rounds).  Pool 0 holds every pixel's column once
// (+100 % of the frame bytes), pool 1 those that are unfinished after the first rounds.  *done stays false when there is
// no memory for the pools or the columns are longer than the sort holds: the caller falls back to the in-place kernel.
inline int launch_linfit_stream(nl_stack_job *job, const StackArgs &args, bool *done) {
    constexpr int S = LINFIT_POOL_S;
    *done = false;
    nl_ctx *ctx = job->ctx;
    if (ctx->linfit_stream == 0 || job->n > 1024) return NL_OK;       // (the register sort holds 32 x 32 samples)
    DeferSchedule d = defer_schedule(ST_LINFIT, ctx);
    if (ctx->defer_override && d.n == 0) return NL_OK;                 // "0": the single-launch kernel was asked for
    const int nwords = ((job->n + 31) & ~31) >> 5;
    // shared memory per warp: the ring, the survivor masks, and as many cached blocks of the columns as one CTA per SM holds
    const size_t fixed = 4096 * (size_t)STREAM_DEPTH + (size_t)nwords * 32 * sizeof(unsigned);
    if ((size_t)LINFIT_STREAM_WARPS * fixed > (size_t)ctx->max_smem_optin) return NL_OK;
    int cache_blocks = (int)(((size_t)ctx->max_smem_optin / LINFIT_STREAM_WARPS - fixed) / 4096);
    if (cache_blocks > nwords) cache_blocks = nwords;
    if (ctx->linfit_stream_cache >= 0 && cache_blocks > ctx->linfit_stream_cache) cache_blocks = ctx->linfit_stream_cache;
    const size_t smem = (size_t)LINFIT_STREAM_WARPS * (fixed + 4096 * (size_t)cache_blocks);
    const double frac[2] = {1.0, 0.85};
    StackArgs::Pool pools[2];
    if (!ensure_pools(job, 0, frac, pools) || pools[0].cap < ((job->pixels + 31) & ~31ll)) return NL_OK;
    if (pools[1].cap < 32) d.n = 0;                                    // no second pool: sort, then one launch to the end
    if (!ctx->defer_override) {
        // every regrouping launch ends in a tail of half-idle SMs: small jobs regroup once or not at all
        const long long groups = (job->pixels + 31) / 32, lanes = (long long)ctx->sm_count * LINFIT_STREAM_WARPS;
        if (groups < 2 * lanes) d.n = 0;
        else if (groups < 8 * lanes) { d.n = 1; d.at[0] = 12; }
    }
    auto kern = linfit_rounds_kernel<S>;
    NL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 0;
    NL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, LINFIT_STREAM_WARPS * 32, smem));
    if (ctas_per_sm < 1) return NL_OK;
    if (ctx->linfit_stream_ctas > 0 && ctas_per_sm > ctx->linfit_stream_ctas) ctas_per_sm = ctx->linfit_stream_ctas;
    const unsigned grid = (unsigned)(ctx->sm_count * ctas_per_sm);
    StackArgs a2 = args;
    a2.phase = 0;
    a2.stream_cache_blocks = cache_blocks;
    a2.pool_out = pools[0];
    a2.pool_out.count = job->clip + 3;
    {
        const long long tiles = (args.pixels + S - 1) / S;
        const bool small = job->n <= 512;
        const size_t ssmem = (size_t)(small ? SortSlab<16>::WORDS : SortSlab<32>::WORDS) * sizeof(float);
        long long sgrid = tiles;
        const long long cap = (long long)ctx->sm_count * LINFIT_SORT_CTAS * (small ? 2 : 1);
        if (sgrid > cap) sgrid = cap;
        if (sgrid < 1) sgrid = 1;
        if (small) {
            NL_CUDA(cudaFuncSetAttribute(linfit_sort_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
            linfit_sort_kernel<16><<<(unsigned)sgrid, LINFIT_SORT_WARPS * 32, ssmem, ctx->stream>>>(a2);
        } else {
            NL_CUDA(cudaFuncSetAttribute(linfit_sort_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
            linfit_sort_kernel<32><<<(unsigned)sgrid, LINFIT_SORT_WARPS * 32, ssmem, ctx->stream>>>(a2);
        }
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    for (int r = 1; r <= d.n + 1; r++) {
        a2.phase = r;
        a2.defer_passes = r <= d.n ? d.at[r - 1] - (r > 1 ? d.at[r - 2] : 0) : 0;
        a2.pool_in = pools[(r + 1) & 1];
        a2.pool_in.count = job->clip + 3 + 2 * (r - 1);
        a2.pool_tile_counter = job->clip + 4 + 2 * (r - 1);
        a2.pool_out = pools[r & 1];
        a2.pool_out.count = job->clip + 3 + 2 * r;
        kern<<<grid, LINFIT_STREAM_WARPS * 32, smem, ctx->stream>>>(a2);
        NL_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    *done = true;
    return NL_OK;
}

template <int MODE, bool W, int S, typename IDX>
inline int launch_deferred(nl_stack_job *job, const StackArgs &args) {
    if constexpr (MODE == ST_LINFIT && S < 32) {
        bool done = false;
        int rc = launch_linfit_stream(job, args, &done);
        if (rc != NL_OK || done) return rc;
    }
    if (MODE == ST_SIGMA || MODE == ST_WINSOR || MODE == ST_LINFIT) {
        DeferSchedule d = defer_schedule(MODE, job->ctx, job->n);
        // several regrouping launches only pay with many tiles per warp (each launch ends in a tail of half-idle SMs):
        // measured, 1024 frames x 65 536 pixels: 15.5 ms in one launch, 16.6 ms with six regroupings; x 1 M pixels: 223 -> 213 ms
        // (the clipping modes keep their first regrouping: one extra launch)
        if (d.n > 1 && !job->ctx->defer_override && (job->pixels + S - 1) / S < 32ll * job->ctx->sm_count * 8) d.n = MODE == ST_LINFIT ? 0 : 1;
        StackArgs::Pool pools[2];
        if (d.n > 0 && ensure_pools(job, W ? (int)sizeof(IDX) : 0, d.frac, pools)) {
            StackArgs a2 = args;
            const int rounds = pools[1].cap >= 32 ? d.n : 1;          // without a second pool: defer once
            for (int r = 0; r <= rounds; r++) {
                a2.phase = r;
                a2.defer_passes = r < rounds ? d.at[r] - (r ? d.at[r - 1] : 0) : 0;
                a2.pool_in = pools[(r + 1) & 1];                      // what launch r-1 filled
                a2.pool_in.count = job->clip + 3 + 2 * (r > 0 ? r - 1 : 0);
                a2.pool_tile_counter = job->clip + 4 + 2 * (r > 0 ? r - 1 : 0);
                a2.pool_out = pools[r & 1];
                a2.pool_out.count = job->clip + 3 + 2 * r;
                int rc = launch_column<MODE, W, S, IDX>(job, a2);
                if (rc != NL_OK) return rc;
            }
            return NL_OK;
        }
    }
    return launch_column<MODE, W, S, IDX>(job, args);
}

template <int MODE, bool W, typename IDX>
inline int launch_column_i(nl_stack_job *job, const StackArgs &args) {
    constexpr int SB = SlotBytes<MODE, W, IDX>::value;
    const size_t per_pixel = (size_t)SB * ((job->n + 31) & ~31);
    const size_t gap = (size_t)(QW - 1) * sizeof(float);           // per pixel of tile width
    const size_t cap = (size_t)job->ctx->max_smem_optin;
    // Tile width: a wide tile uses every lane of a warp but needs SB*npad*S bytes per warp, and the kernel
    // lives on latency hiding across warps (each column is a serial dependency chain).  Score = columns
    // that make progress per cycle ~ min(warps, 8) * S; e.g. N=256 -> 32 pixels x 7 warps, N=1024 ->
    // 8 pixels x 7 warps instead of 32 pixels x 1 warp.
    const int widths[4] = {32, 16, 8, 1};
    int best = 0;
    double best_score = -1;
    for (int wdt : widths) {
        const size_t per_warp = (per_pixel + gap) * wdt;
        if (per_warp + gap * wdt > cap) continue;
        size_t warps = (cap - gap * wdt) / per_warp;
        if (warps > 64) warps = 64;
        const double score = (double)(warps > 8 ? 8 : warps) * wdt;
        if (score > best_score) { best_score = score; best = wdt; }
    }
    if (job->ctx->tile_width) {                                    // nl_ctx_set_tuning "tile_width" (A/B measurements)
        const int wdt = job->ctx->tile_width;
        if ((wdt == 32 || wdt == 16 || wdt == 8 || wdt == 1) && (per_pixel + 2 * gap) * wdt <= cap) best = wdt;
    }
    switch (best) {
    case 32: return launch_deferred<MODE, W, 32, IDX>(job, args);
    case 16: return launch_deferred<MODE, W, 16, IDX>(job, args);
    case 8: return launch_deferred<MODE, W, 8, IDX>(job, args);
    case 1: return launch_deferred<MODE, W, 1, IDX>(job, args);
    }
    return set_error(NL_E_INVALID, "n_frames %d too large for shared memory", job->n);
}

template <int MODE, bool W>
inline int launch_column_s(nl_stack_job *job, const StackArgs &args) {
    if (!W || job->n <= 256) return launch_column_i<MODE, W, unsigned char>(job, args);
    if (job->n <= 65536) return launch_column_i<MODE, W, unsigned short>(job, args);
    return set_error(NL_E_INVALID, "weighted stacking of more than 65536 frames is not supported");
}

// per-mode launchers, one translation unit each
int launch_median(nl_stack_job *job, const StackArgs &args);
int launch_sigma(nl_stack_job *job, const StackArgs &args, bool weighted);
int launch_winsor(nl_stack_job *job, const StackArgs &args, bool weighted);
int launch_mad(nl_stack_job *job, const StackArgs &args);
int launch_linfit(nl_stack_job *job, const StackArgs &args);

}  // namespace nl
