// context.cu -- ctx lifecycle, error text, pinned host memory, phase timings.
#include "dge_internal.cuh"
#include <cstring>

static thread_local std::string g_tls_error;

void dge_set_error(dge_ctx *ctx, const std::string &msg) {
    g_tls_error = msg;
    if (ctx) ctx->err = msg;
}
int dge_fail(dge_ctx *ctx, int code, const std::string &msg) {
    dge_set_error(ctx, msg);
    return code;
}

static void ctx_teardown(dge_ctx *ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->tev0) cudaEventDestroy(ctx->tev0);
    if (ctx->tev1) cudaEventDestroy(ctx->tev1);
    delete ctx;
}
void dge_ctx_retain(dge_ctx *ctx) { if (ctx) ctx->refs++; }
void dge_ctx_release(dge_ctx *ctx) {
    if (ctx && --ctx->refs == 0) ctx_teardown(ctx);
}

extern "C" {

int dge_version(void) { return 100; }

int dge_create(int device, dge_ctx **out) {
    if (!out) return dge_fail(nullptr, DGE_E_INVALID, "dge_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return dge_fail(nullptr, DGE_E_NO_DEVICE,
                        std::string("dge_create: no CUDA device (") + cudaGetErrorString(e) +
                            "); libdge has no CPU fallback");
    if (device < 0 || device >= n)
        return dge_fail(nullptr, DGE_E_INVALID, "dge_create: device ordinal out of range");
    cudaDeviceProp prop;
    DGE_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return dge_fail(nullptr, DGE_E_NO_DEVICE,
                        std::string("dge_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                            std::to_string(prop.minor) + "; libdge is built for sm_100a (B200) only");
    DGE_CUDA(nullptr, cudaSetDevice(device));
    dge_ctx *ctx = new dge_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->tev0) != cudaSuccess || cudaEventCreate(&ctx->tev1) != cudaSuccess) {
        delete ctx;
        return dge_fail(nullptr, DGE_E_CUDA, "dge_create: stream/event creation failed");
    }
    // stream-ordered allocator: a pool of this ctx's own that keeps freed blocks (no trim at synchronisation points);
    // the device's default pool -- shared with torch / NCCL living in the same process -- is not touched
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
        ctx->pool = nullptr;   // fall back to the default pool with its default (trimming) threshold
        cudaGetLastError();
    }
    *out = ctx;
    return DGE_OK;
}

void dge_destroy(dge_ctx *ctx) {
    if (!ctx || ctx->closed) return;
    ctx->closed = true;
    dge_comm_destroy(ctx);
    // handles that are still alive keep the stream and the pool alive; their *_free releases the last reference
    dge_ctx_release(ctx);
}

const char *dge_last_error(const dge_ctx *ctx) { return ctx ? ctx->err.c_str() : g_tls_error.c_str(); }

void *dge_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void dge_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int dge_phase_ms(const dge_ctx *ctx, const char *phase, float *ms) {
    if (!ctx || !phase || !ms) return DGE_E_INVALID;
    auto it = ctx->phase_ms.find(phase);
    if (it == ctx->phase_ms.end()) return DGE_E_INVALID;
    *ms = it->second;
    return DGE_OK;
}

int64_t dge_kernel_launches(const dge_ctx *ctx) { return ctx ? ctx->launches : 0; }

int dge_timer_start(dge_ctx *ctx) {
    if (!ctx) return DGE_E_INVALID;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DGE_CUDA(ctx, cudaEventRecord(ctx->tev0, ctx->stream));
    return DGE_OK;
}

int dge_timer_stop(dge_ctx *ctx, float *ms) {
    if (!ctx || !ms) return DGE_E_INVALID;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    DGE_CUDA(ctx, cudaEventRecord(ctx->tev1, ctx->stream));
    DGE_CUDA(ctx, cudaEventSynchronize(ctx->tev1));
    DGE_CUDA(ctx, cudaEventElapsedTime(ms, ctx->tev0, ctx->tev1));
    return DGE_OK;
}

} // extern "C"
