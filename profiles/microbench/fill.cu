// Write-bandwidth microbenchmarks (not part of the library): how fast can B200 absorb a pure store stream?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o profiles/microbench/libfill.so profiles/microbench/fill.cu
#include <cstdint>
#include <cuda_runtime.h>

template <int W>  // bytes per store: 16 or 32
__global__ void fill_kernel(uint8_t *out, int64_t n_bytes, uint32_t v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * W;
    for (int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W; o + W <= n_bytes; o += stride) {
        if (W == 16) {
            *reinterpret_cast<uint4 *>(out + o) = make_uint4(v, v, v, v);
        } else {
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(out + o), "r"(v) : "memory");
        }
    }
}

// one CTA per contiguous tile (like the execute kernel): each warp writes 1 KiB per step
template <int W>
__global__ void fill_tiled_kernel(uint8_t *out, int64_t tile_bytes, uint32_t v) {
    uint8_t *base = out + (int64_t)blockIdx.x * tile_bytes;
    for (int64_t o = (int64_t)threadIdx.x * W; o + W <= tile_bytes; o += (int64_t)blockDim.x * W) {
        if (W == 16) {
            *reinterpret_cast<uint4 *>(base + o) = make_uint4(v, v, v, v);
        } else {
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(base + o), "r"(v) : "memory");
        }
    }
}

// lane-blocked stores: every lane owns `LB` contiguous bytes (LB / 32 256-bit stores), lanes LB bytes apart -- what a
// kernel does whose threads walk consecutive output values with a carried state (track execute, round 2)
template <int LB>
__global__ void fill_blocked_kernel(uint8_t *out, int64_t tile_bytes, uint32_t v) {
    uint8_t *base = out + (int64_t)blockIdx.x * tile_bytes;
    for (int64_t o = (int64_t)threadIdx.x * LB; o + LB <= tile_bytes; o += (int64_t)blockDim.x * LB) {
#pragma unroll
        for (int k = 0; k < LB / 32; k++)
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(base + o + 32 * k), "r"(v) : "memory");
    }
}

extern "C" {
int fill_blocked_launch(void *out, int64_t n_bytes, int64_t tile_bytes, int lane_bytes, int block, void *stream) {
    const int grid = (int)(n_bytes / tile_bytes);
    if (lane_bytes == 64) fill_blocked_kernel<64><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, tile_bytes, 0x01000000u);
    else if (lane_bytes == 128) fill_blocked_kernel<128><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, tile_bytes, 0x01000000u);
    else fill_blocked_kernel<256><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, tile_bytes, 0x01000000u);
    return (int)cudaGetLastError();
}
int fill_launch(void *out, int64_t n_bytes, int width, int grid, int block, void *stream) {
    if (width == 16) fill_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, n_bytes, 0x01000000u);
    else fill_kernel<32><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, n_bytes, 0x01000000u);
    return (int)cudaGetLastError();
}
int fill_tiled_launch(void *out, int64_t n_bytes, int64_t tile_bytes, int width, int block, void *stream) {
    const int grid = (int)(n_bytes / tile_bytes);
    if (width == 16) fill_tiled_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, tile_bytes, 0x01000000u);
    else fill_tiled_kernel<32><<<grid, block, 0, (cudaStream_t)stream>>>((uint8_t *)out, tile_bytes, 0x01000000u);
    return (int)cudaGetLastError();
}
}
