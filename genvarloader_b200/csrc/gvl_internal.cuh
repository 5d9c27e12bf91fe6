// gvl_internal.cuh -- helpers private to the library (not part of the C ABI).
#pragma once
#include "gvl_common.cuh"

namespace gvl {

int fail(int code, const char *fmt, ...);
void count_launch(int n = 1);
int ensure_rows(gvl_ctx *ctx, gvl_workspace &ws, int64_t n_work);
int ensure_records(gvl_ctx *ctx, gvl_workspace &ws, int64_t n_rec);

#define GVL_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return gvl::fail(GVL_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define GVL_LAUNCH_CHECK()                                                                              \
    do {                                                                                                \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess)                                                                          \
            return gvl::fail(GVL_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
        gvl::count_launch();                                                                            \
    } while (0)

// ---- small device helpers ----------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// last index i in [lo, hi) with arr[i] <= key, or lo-1 if none.  Warp-cooperative 32-ary search:
// every lane must call it with identical arguments; every lane gets the result.
__device__ __forceinline__ int64_t warp_upper_le(const int32_t *__restrict__ arr, int64_t lo, int64_t hi, int32_t key) {
    // invariant: answer in [lo-1, hi-1]
    const int lane = lane_id();
    while (hi - lo > 32) {
        int64_t step = (hi - lo + 31) / 32;  // 32 probes at lo + (lane+1)*step - 1 (clamped)
        int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
        bool ok = (idx < hi) ? (arr[idx] <= key) : false;
        unsigned m = __ballot_sync(0xffffffffu, ok);
        int cnt = __popc(m);  // probes are monotone: first `cnt` lanes are <= key
        int64_t new_lo = lo + (int64_t)cnt * step;
        int64_t new_hi = imin64(hi, lo + (int64_t)(cnt + 1) * step - 1);
        // elements [lo, new_lo) are all <= key; element at new_hi (if < hi) is > key
        lo = new_lo;
        hi = new_hi;
        if (lo > hi) hi = lo;
    }
    int64_t idx = lo + lane;
    bool ok = (idx < hi) ? (arr[idx] <= key) : false;
    unsigned m = __ballot_sync(0xffffffffu, ok);
    return lo + __popc(m) - 1;
}

}  // namespace gvl
