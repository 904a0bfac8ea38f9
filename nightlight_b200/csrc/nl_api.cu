// nl_api.cu -- context, error reporting and memory helpers of libnightlight_cuda.so.
// The reference is one Go process with goroutines; the one boundary this
// library adds is Go <-> C ABI <-> CUDA.  A context = one device + one stream.
#include "nl_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace nl {

static thread_local std::string g_last_error;

int set_error(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    return set_error(NL_E_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int ensure_scratch(nl_ctx *ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return NL_OK;
    if (ctx->scratch) {
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
        NL_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    NL_CUDA(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return NL_OK;
}

int ensure_frame(nl_ctx *ctx, int slot, size_t bytes, float **out) {
    if (ctx->frame_bytes[slot] < bytes) {
        if (ctx->frame[slot]) {
            NL_CUDA(cudaStreamSynchronize(ctx->stream));
            NL_CUDA(cudaFree(ctx->frame[slot]));
            ctx->frame[slot] = nullptr;
            ctx->frame_bytes[slot] = 0;
        }
        NL_CUDA(cudaMalloc(&ctx->frame[slot], bytes + 256));
        ctx->frame_bytes[slot] = bytes;
    }
    *out = (float *)ctx->frame[slot];
    return NL_OK;
}

int ensure_pinned(nl_ctx *ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return NL_OK;
    if (ctx->pinned) {
        NL_CUDA(cudaStreamSynchronize(ctx->stream));
        NL_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
    }
    NL_CUDA(cudaHostAlloc(&ctx->pinned, bytes, cudaHostAllocMapped));
    NL_CUDA(cudaHostGetDevicePointer(&ctx->pinned_dev, ctx->pinned, 0));
    ctx->pinned_bytes = bytes;
    return NL_OK;
}

}  // namespace nl

using namespace nl;

extern "C" {

const char *nl_last_error(void) { return g_last_error.c_str(); }

int nl_version(void) { return 100; }   // 0.1.0

int nl_device_count(int *count) {
    NL_REQUIRE(count, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
    return NL_OK;
}

int nl_ctx_create(int device, nl_ctx **out) {
    NL_REQUIRE(out, "ctx out pointer is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(NL_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return set_error(NL_E_INVALID, "device %d out of range [0,%d)", device, count);
    // the caller's current device is left as it was (a host that uses CUDA itself must not end up on another GPU)
    struct Restore {
        int prev = -1;
        Restore() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
        ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    NL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NL_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_error(NL_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                         prop.major, prop.minor);
    nl_ctx *c = new nl_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    c->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    *out = c;
    return NL_OK;
}

int nl_ctx_destroy(nl_ctx *ctx) {
    if (!ctx) return NL_OK;
    nl_stack_apply_release(ctx);
    NL_GUARD(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->list) cudaFree(ctx->list);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->batch_pinned) cudaFreeHost(ctx->batch_pinned);
    for (int k = 0; k < 2; k++)
        if (ctx->frame[k]) cudaFree(ctx->frame[k]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NL_OK;
}

int nl_ctx_sync(nl_ctx *ctx) {
    NL_REQUIRE(ctx, "ctx is NULL");
    NL_GUARD(ctx);
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    return NL_OK;
}

int nl_ctx_stream(nl_ctx *ctx, void **stream) {
    NL_REQUIRE(ctx && stream, "NULL argument");
    *stream = (void *)ctx->stream;
    return NL_OK;
}

int nl_ctx_device(nl_ctx *ctx, int *device) {
    NL_REQUIRE(ctx && device, "NULL argument");
    *device = ctx->device;
    return NL_OK;
}

int nl_ctx_mem_info(nl_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes) {
    NL_REQUIRE(ctx, "ctx is NULL");
    NL_GUARD(ctx);
    size_t f = 0, t = 0;
    NL_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return NL_OK;
}

// Tuning knobs for A/B measurements and tests.  The library reads no environment variables: whatever changes its
// behaviour is set explicitly, per context.
int nl_ctx_set_tuning(nl_ctx *ctx, const char *key, const char *value) {
    NL_REQUIRE(ctx && key, "NULL argument");
    const std::string k = key, v = value ? value : "";
    if (k == "defer_passes") {
        if (v.empty()) { ctx->defer_override = false; return NL_OK; }          // back to the built-in schedule
        ctx->defer_override = true;
        ctx->defer_n = 0;
        for (const char *q = v.c_str(); *q && ctx->defer_n < 8;) {
            const int x = atoi(q);
            if (x > (ctx->defer_n ? ctx->defer_at[ctx->defer_n - 1] : 0)) ctx->defer_at[ctx->defer_n++] = x;
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        return NL_OK;
    }
    if (k == "tile_width") {
        const int w = atoi(v.c_str());
        NL_REQUIRE(w == 0 || w == 32 || w == 16 || w == 8 || w == 1, "tile_width must be 0 (automatic), 32, 16, 8 or 1");
        ctx->tile_width = w;
        return NL_OK;
    }
    if (k == "linfit_stream") { ctx->linfit_stream = atoi(v.c_str()); return NL_OK; }
    if (k == "linfit_stream_cache") { ctx->linfit_stream_cache = atoi(v.c_str()); return NL_OK; }
    if (k == "linfit_stream_ctas") { ctx->linfit_stream_ctas = atoi(v.c_str()); return NL_OK; }
    if (k == "stats_debug") { ctx->stats_debug = atoi(v.c_str()) != 0; return NL_OK; }
    if (k == "stats_force_replay") { ctx->stats_force_replay = atoi(v.c_str()) != 0; return NL_OK; }
    return set_error(NL_E_INVALID, "unknown tuning key '%s'", key);
}

int nl_ctx_launch_count(nl_ctx *ctx, int64_t *launches) {
    NL_REQUIRE(ctx && launches, "NULL argument");
    *launches = ctx->launches.load();
    return NL_OK;
}

int nl_dev_alloc(nl_ctx *ctx, int64_t bytes, void **dev) {
    NL_REQUIRE(ctx && dev && bytes >= 0, "bad argument");
    NL_GUARD(ctx);
    cudaError_t e = cudaMalloc(dev, (size_t)(bytes > 0 ? bytes : 1));
    if (e == cudaErrorMemoryAllocation) return set_error(NL_E_NOMEM, "cudaMalloc of %lld bytes failed", (long long)bytes);
    NL_CUDA(e);
    return NL_OK;
}

int nl_dev_free(nl_ctx *ctx, void *dev) {
    NL_REQUIRE(ctx, "ctx is NULL");
    NL_GUARD(ctx);
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    NL_CUDA(cudaFree(dev));
    return NL_OK;
}

// Cross-process peer mapping (one process per GPU): the owner exports a cudaMalloc'ed buffer, every
// other rank opens it and gets a pointer it can hand to nl_stack_run_dev_bcast as a peer stripe.
int nl_ipc_get_handle(nl_ctx *ctx, void *dev, unsigned char handle[64]) {
    NL_REQUIRE(ctx && dev && handle, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    NL_GUARD(ctx);
    cudaIpcMemHandle_t h;
    NL_CUDA(cudaIpcGetMemHandle(&h, dev));
    memcpy(handle, &h, 64);
    return NL_OK;
}

int nl_ipc_open_handle(nl_ctx *ctx, const unsigned char handle[64], void **dev) {
    NL_REQUIRE(ctx && dev && handle, "NULL argument");
    NL_GUARD(ctx);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    NL_CUDA(cudaIpcOpenMemHandle(dev, h, cudaIpcMemLazyEnablePeerAccess));
    return NL_OK;
}

int nl_ipc_close_handle(nl_ctx *ctx, void *dev) {
    NL_REQUIRE(ctx, "ctx is NULL");
    NL_GUARD(ctx);
    NL_CUDA(cudaStreamSynchronize(ctx->stream));
    NL_CUDA(cudaIpcCloseMemHandle(dev));
    return NL_OK;
}

int nl_host_alloc_pinned(int64_t bytes, void **host) {
    NL_REQUIRE(host && bytes >= 0, "bad argument");
    cudaError_t e = cudaHostAlloc(host, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocPortable);
    if (e == cudaErrorMemoryAllocation) return set_error(NL_E_NOMEM, "cudaHostAlloc of %lld bytes failed", (long long)bytes);
    NL_CUDA(e);
    return NL_OK;
}

// Page-lock an existing host buffer in place (e.g. a Go []float32 that stays alive and is not moved:
// Go's heap is non-moving) so that nl_stack_put_frame / nl_project copy from it at full PCIe speed.
int nl_host_register(void *host, int64_t bytes) {
    NL_REQUIRE(host && bytes > 0, "bad argument");
    NL_CUDA(cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable));
    return NL_OK;
}

int nl_host_unregister(void *host) {
    NL_CUDA(cudaHostUnregister(host));
    return NL_OK;
}

int nl_host_free_pinned(void *host) {
    NL_CUDA(cudaFreeHost(host));
    return NL_OK;
}

int nl_memcpy_h2d(nl_ctx *ctx, void *dev, const void *host, int64_t bytes) {
    NL_REQUIRE(ctx && bytes >= 0, "bad argument");
    NL_GUARD(ctx);
    NL_CUDA(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    return NL_OK;
}

int nl_memcpy_d2h(nl_ctx *ctx, void *host, const void *dev, int64_t bytes) {
    NL_REQUIRE(ctx && bytes >= 0, "bad argument");
    NL_GUARD(ctx);
    NL_CUDA(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return NL_OK;
}

int nl_memcpy_d2d(nl_ctx *ctx, void *dev_dst, const void *dev_src, int64_t bytes) {
    NL_REQUIRE(ctx && bytes >= 0, "bad argument");
    NL_GUARD(ctx);
    NL_CUDA(cudaMemcpyAsync(dev_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return NL_OK;
}

}  // extern "C"
