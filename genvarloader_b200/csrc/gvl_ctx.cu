// gvl_ctx.cu -- context, workspace and error plumbing of the C ABI (include/gvl_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "gvl_internal.cuh"

namespace gvl {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

template <typename T>
static int grow(T *&p, int64_t &cap_elems_unused, int64_t need, int64_t old_cap) {
    (void)cap_elems_unused;
    if (need <= old_cap && p) return GVL_OK;
    if (p) GVL_CUDA(cudaFree(p));
    p = nullptr;
    GVL_CUDA(cudaMalloc(&p, sizeof(T) * (size_t)need));
    return GVL_OK;
}

int ensure_rows(gvl_ctx *ctx, int64_t n_work) {
    if (n_work <= ctx->rows_cap) return GVL_OK;
    int64_t cap = n_work + n_work / 2 + 64, dummy = 0;
    int rc;
    if ((rc = grow(ctx->rows, dummy, cap, 0))) return rc;
    if ((rc = grow(ctx->tile_off, dummy, cap + 1, 0))) return rc;
    if ((rc = grow(ctx->row_len, dummy, cap, 0))) return rc;
    ctx->rows_cap = cap;
    return GVL_OK;
}

int ensure_records(gvl_ctx *ctx, int64_t n_rec) {
    if (n_rec <= ctx->rec_cap) return GVL_OK;
    int64_t cap = n_rec + n_rec / 2 + 1024, dummy = 0;
    int rc;
    if ((rc = grow(ctx->rec.a, dummy, cap, 0))) return rc;
    if ((rc = grow(ctx->rec.n, dummy, cap, 0))) return rc;
    if ((rc = grow(ctx->rec.src, dummy, cap, 0))) return rc;
    if ((rc = grow(ctx->rec.resume, dummy, cap, 0))) return rc;
    if ((rc = grow(ctx->rec.vidx, dummy, cap, 0))) return rc;
    if ((rc = grow(ctx->rec.vpos, dummy, cap, 0))) return rc;
    ctx->rec_cap = cap;
    return GVL_OK;
}

}  // namespace gvl

using namespace gvl;

extern "C" {

const char *gvl_last_error(void) { return g_err; }

int64_t gvl_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int gvl_ctx_create(int device, gvl_ctx **out) {
    if (!out) return fail(GVL_ERR_ARG, "gvl_ctx_create: out is NULL");
    int n = 0;
    GVL_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(GVL_ERR_ARG, "gvl_ctx_create: device %d of %d", device, n);
    GVL_CUDA(cudaSetDevice(device));
    gvl_ctx *ctx = new gvl_ctx();
    ctx->device = device;
    ctx->own_stream = nullptr;
    ctx->rows = nullptr;
    ctx->rows_cap = 0;
    ctx->rec = RecArrays{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ctx->rec_cap = 0;
    ctx->dev_words = nullptr;
    ctx->host_words = nullptr;
    ctx->tile_off = nullptr;
    ctx->row_len = nullptr;
    ctx->plan_valid = false;
    ctx->n_work = 0;
    ctx->fixed_len = -1;
    ctx->total = -1;
    ctx->plan_out_offsets = nullptr;
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    ctx->host_out_offsets_dev = nullptr;
    memset(&ctx->host_tab, 0, sizeof(ctx->host_tab));
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dev_words, sizeof(int64_t) * W_COUNT);
    if (e == cudaSuccess) e = cudaMemset(ctx->dev_words, 0, sizeof(int64_t) * W_COUNT);
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->host_words, sizeof(int64_t) * W_COUNT, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        gvl_ctx_destroy(ctx);
        return fail(GVL_ERR_CUDA, "gvl_ctx_create: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return GVL_OK;
}

void gvl_ctx_destroy(gvl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->rows);
    cudaFree(ctx->tile_off);
    cudaFree(ctx->row_len);
    cudaFree(ctx->rec.a);
    cudaFree(ctx->rec.n);
    cudaFree(ctx->rec.src);
    cudaFree(ctx->rec.resume);
    cudaFree(ctx->rec.vidx);
    cudaFree(ctx->rec.vpos);
    cudaFree(ctx->dev_words);
    if (ctx->host_words) cudaFreeHost(ctx->host_words);
    for (auto &kv : ctx->statics) cudaFree(kv.second.dev);
    for (auto &s : ctx->scratch) cudaFree(s.first);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int gvl_ctx_check(gvl_ctx *ctx, gvl_stream stream) {
    if (!ctx) return fail(GVL_ERR_ARG, "gvl_ctx_check: ctx is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    GVL_CUDA(cudaSetDevice(ctx->device));
    GVL_CUDA(cudaMemcpyAsync(ctx->host_words, ctx->dev_words, sizeof(int64_t) * W_COUNT, cudaMemcpyDeviceToHost, st));
    GVL_CUDA(cudaStreamSynchronize(st));
    if (ctx->host_words[W_STATUS] != 0) {
        int64_t s = ctx->host_words[W_STATUS];
        cudaMemsetAsync(ctx->dev_words + W_STATUS, 0, sizeof(int64_t), st);
        return fail(GVL_ERR_CAPACITY, "device workspace overflow (status=%lld): raise max_records", (long long)s);
    }
    return GVL_OK;
}

}  // extern "C"
