// ctx.cu — context lifetime, error text, workspaces.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <new>

int uco_fail(uco_b200_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

void* uco_ws(uco_b200_ctx* ctx, int slot, size_t bytes) {
    if ((size_t)slot >= ctx->dev.size()) ctx->dev.resize(slot + 1);
    uco_dev_buf& b = ctx->dev[slot];
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return b.p;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        uco_fail(ctx, UCO_E_NOMEM, "cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        b.p = nullptr;
        return nullptr;
    }
    b.cap = want;
    return b.p;
}

void* uco_pinned(uco_b200_ctx* ctx, int slot, size_t bytes) {
    if ((size_t)slot >= ctx->pin.size()) ctx->pin.resize(slot + 1);
    uco_dev_buf& b = ctx->pin[slot];
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return b.p;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&b.p, want);
    if (e != cudaSuccess) {
        uco_fail(ctx, UCO_E_NOMEM, "cudaMallocHost(%zu) -> %s", want, cudaGetErrorString(e));
        b.p = nullptr;
        return nullptr;
    }
    b.cap = want;
    return b.p;
}

cudaError_t uco_sleep_sync(uco_b200_ctx* ctx) {
    static const bool spin = getenv("UCO_SPIN_SYNC") != nullptr;   // A/B switch for measurements
    if (!ctx->sleep_event || spin) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->sleep_event, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->sleep_event);
}

extern "C" {

uco_b200_ctx* uco_b200_create(int device, int flags) {
    (void)flags;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    uco_b200_ctx* ctx = new (std::nothrow) uco_b200_ctx();
    if (!ctx) return nullptr;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return nullptr;
    }
    cudaEventCreateWithFlags(&ctx->sleep_event, cudaEventBlockingSync | cudaEventDisableTiming);
    ctx->dev.resize(WS_COUNT);
    ctx->pin.resize(WS_COUNT);
    return ctx;
}

void uco_b200_destroy(uco_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    uco_orb_state_free(ctx);
    uco_ba_state_free(ctx);
    for (auto& b : ctx->dev)
        if (b.p) cudaFree(b.p);
    for (auto& b : ctx->pin)
        if (b.p) cudaFreeHost(b.p);
    if (ctx->sleep_event) cudaEventDestroy(ctx->sleep_event);
    if (ctx->stage_event) cudaEventDestroy(ctx->stage_event);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* uco_b200_last_error(const uco_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no usable CUDA device?)"; }
void* uco_b200_stream(uco_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int uco_b200_sync(uco_b200_ctx* ctx) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UCO_OK;
}
uint64_t uco_b200_launch_count(const uco_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int uco_b200_version(void) { return 100; }
void uco_b200_set_profiling(uco_b200_ctx* ctx, int on) { if (ctx) ctx->profiling = on; }

}  // extern "C"
