// nl_internal.h -- shared declarations of libnightlight_cuda.so (not installed; the public
// interface is include/nightlight_cuda.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>

#include "../../include/nightlight_cuda.h"

#define NL_MAX_PEERS 8
#define NL_MAX_LANES 5       // helper contexts (stream + scratch) a context keeps on its device
#define NL_JOB_COUNTERS 32     // device counters of a stack job: clip low/high, tile counter, two per deferral round

struct nl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int max_smem_optin = 0;          // bytes of dynamic shared memory one CTA may opt in to
    int smem_per_sm = 0;
    std::atomic<int64_t> launches{0};
    int numerics = NL_NUMERICS_AMD64;   // which build of the reference the frame statistics reproduce (nl_ctx_set_numerics)
    int64_t exact_replays = 0;          // float64 chains that had to be replayed in order (nl_prestats.cu)
    // scratch owned by the context, grown on demand (star scan)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *frame[2] = {nullptr, nullptr};   // whole-frame staging of the host-pointer entry points (grown on demand, kept)
    size_t frame_bytes[2] = {0, 0};
    void *pinned = nullptr;          // mapped pinned host memory: the star scan writes its count and candidates straight into it
    void *pinned_dev = nullptr;      // (its device address)
    size_t pinned_bytes = 0;
    void *batch_pinned = nullptr;    // mapped pinned host memory of the batched star scan: [frame][stride] candidate records
    void *batch_pinned_dev = nullptr;
    size_t batch_pinned_bytes = 0;
    void *list = nullptr;            // candidate list of the star scan (kept apart from `scratch`, which holds the row offsets)
    size_t list_bytes = 0;
    // nl_stack_apply keeps its two stripe lanes (context + job + result buffer each) between calls:
    // allocating and freeing multi-GiB device buffers per call would cost more than the stack itself
    nl_ctx *lane_ctx[NL_MAX_LANES] = {nullptr};      // (the stripe lanes are the first two; the batched bad-pixel maps use all)
    struct nl_stack_job *lane_job[2] = {nullptr, nullptr};
    float *lane_out[2] = {nullptr, nullptr};
    int64_t lane_px[2] = {0, 0};
    int lane_frames[2] = {0, 0};
    // tuning knobs for A/B measurements and tests (nl_ctx_set_tuning); the library reads no environment variables
    bool defer_override = false;     // "defer_passes": the deferral schedule below replaces the built-in one
    int defer_n = 0;
    int defer_at[8] = {0};
    int tile_width = 0;              // "tile_width": 32 / 16 / 8 / 1 forces the tile width of the column kernel, 0 = automatic
    int linfit_stream = 1;           // "linfit_stream": 0 = long columns take the in-place linear-fit kernel instead of the streaming rounds
    int linfit_stream_cache = -1;    // "linfit_stream_cache": at most this many cached blocks per column (-1 = what fits)
    int linfit_stream_ctas = 0;      // "linfit_stream_ctas": CTAs per SM of the streaming rounds kernel (0 = what fits)
    bool stats_debug = false;        // "stats_debug": the frame statistics print their interval proofs to stderr
    bool stats_force_replay = false; // "stats_force_replay": always replay the float64 chains in order
};

namespace nl {

int set_error(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
int ensure_scratch(nl_ctx *ctx, size_t bytes);
int lane_context(nl_ctx *ctx, int l, nl_ctx **out);                    // nl_stack.cu: ctx->lane_ctx[l], created on first use
int ensure_frame(nl_ctx *ctx, int slot, size_t bytes, float **out);     // nl_api.cu
int ensure_pinned(nl_ctx *ctx, size_t bytes);                           // nl_api.cu: ctx->pinned (mapped), at least `bytes`
int exclusive_scan_launch(nl_ctx *ctx, const int *dev_counts, int *dev_offsets, int n, int *dev_total);   // nl_stars.cu
void median_filter_sparse_host(float *data, int32_t len, int32_t width, const int32_t *indices, int64_t n);     // nl_stars.cu
int fits_decode_launch(nl_ctx *ctx, const void *dev_raw, int bitpix, long long n, float bscale, float bzero, float *dev_dst);

#define NL_CUDA(call)                                        \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return nl::cuda_fail(e__, #call); \
    } while (0)

#define NL_REQUIRE(cond, msg)                                \
    do {                                                     \
        if (!(cond)) return nl::set_error(NL_E_INVALID, "%s", msg); \
    } while (0)

struct CtxGuard {   // make the context's device current for the calling thread
    int prev = -1;
    bool ok = true;
    explicit CtxGuard(const nl_ctx *c) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != c->device) ok = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~CtxGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// every entry point that touches the device: make the context's device current, fail when that is impossible
#define NL_GUARD(c)                                          \
    nl::CtxGuard g(c);                                       \
    if (!g.ok) return nl::set_error(NL_E_CUDA, "cudaSetDevice(%d) failed", (c)->device)

}  // namespace nl
