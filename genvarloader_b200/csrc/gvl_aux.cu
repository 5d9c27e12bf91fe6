// gvl_aux.cu -- the small entries that sit either side of the haplotype kernels on the reference's boundary:
//   choose_exonic_variants  (src/ffi/mod.rs:229-238 -> src/genotypes/mod.rs:132-176)   keep mask for var_filter="exonic"
//   get_reference           (src/ffi/mod.rs:2402-2411 -> src/reference/mod.rs:9-120)   padded reference rows
//   ragged_to_padded        (src/ragged/mod.rs:7-23, seqpro-core Ragged::to_padded_into)  "variable" output shaping
// All three are bandwidth-trivial next to reconstruction; they exist so that the whole per-batch path stays on the
// device and in this library (no host round trip, no framework ops between plan and execute).
#include <cstring>

#include "gvl_internal.cuh"

using namespace gvl;

namespace {

constexpr int SCAN_THREADS = 1024;

// keep_offsets[k+1] = sum_{k' <= k} max(stop - start, 0) of the rows' genotype slices (genotypes/mod.rs:146-153).
// One CTA walks the rows in chunks of 1024 with a running carry (n_work is O(batch): a few thousand).
__global__ void __launch_bounds__(SCAN_THREADS) exonic_offsets_kernel(const int64_t *__restrict__ goi,
                                                                      const int64_t *__restrict__ g_starts,
                                                                      const int64_t *__restrict__ g_stops,
                                                                      int64_t n_work, int64_t *__restrict__ keep_offsets) {
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    __shared__ int64_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_carry = 0;
        keep_offsets[0] = 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < n_work; base += SCAN_THREADS) {
        const int64_t k = base + tid;
        int64_t v = 0;
        if (k < n_work) {
            const int64_t o = goi[k];
            v = imax64(g_stops[o] - g_starts[o], 0);
        }
        int64_t x = v;  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t incl = carry + (warp ? s_warp[warp - 1] : 0) + x;
        if (k < n_work) keep_offsets[k + 1] = incl;
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = incl;
        __syncthreads();
    }
}

// keep[keep_offsets[k] + i] = variant i of row k lies fully inside the query: v_pos >= start && v_ref_end <= end with
// v_ref_end = v_pos - min(ilen, 0) + 1 (genotypes/mod.rs:160-172).  One warp per row, lanes over its variants.
__global__ void __launch_bounds__(256) exonic_keep_kernel(gvl_sparse_tables tab, const int32_t *__restrict__ starts,
                                                          const int32_t *__restrict__ ends,
                                                          const int64_t *__restrict__ goi, int64_t n_work, int ploidy,
                                                          const int64_t *__restrict__ keep_offsets,
                                                          uint8_t *__restrict__ keep, int64_t keep_cap) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n_work; k += warps) {
        const int64_t q = k / ploidy, o = goi[k];
        const int64_t o_s = tab.geno_starts[o], n = imax64(tab.geno_stops[o] - o_s, 0), k_s = keep_offsets[k];
        const int64_t r_s = starts[q], r_e = ends[q];
        for (int64_t i = lane; i < n; i += 32) {
            if (k_s + i >= keep_cap) break;  // caller's buffer is too small: reported through keep_offsets[n_work]
            const int32_t vi = __ldg(tab.geno_v_idxs + o_s + i);
            const int64_t v_pos = __ldg(tab.v_starts + vi);
            const int64_t v_end = v_pos - imin64((int64_t)__ldg(tab.ilens + vi), 0) + 1;
            keep[k_s + i] = (uint8_t)(v_pos >= r_s && v_end <= r_e);
        }
    }
}

// Copy min(len, out_len) items of every ragged row into the caller's PRE-FILLED (n_rows, out_len) buffer.  One thread per
// 4 destination bytes (destination rows start 4-byte aligned whenever out_len * itemsize is a multiple of 4; otherwise,
// and at row tails, bytes go one at a time); the source is read as two aligned words + funnel shift.
__global__ void __launch_bounds__(256) ragged_to_padded_kernel(const uint8_t *__restrict__ data,
                                                               const int64_t *__restrict__ offsets, int64_t n_rows,
                                                               uint8_t *__restrict__ out, int64_t itemsize,
                                                               int64_t out_len, int64_t words_per_row, int fill,
                                                               uint64_t pad_bits) {
    // fill != 0: the kernel also writes the pad item (itemsize <= 8 bytes, little endian in pad_bits) behind every row, so the
    // caller needs no pre-fill pass over the whole output
    const int64_t n_units = n_rows * words_per_row;
    const int64_t row_bytes = out_len * itemsize;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = u / words_per_row, w = u - r * words_per_row;
        const int64_t o_s = offsets[r];
        const int64_t n_bytes = imin64(offsets[r + 1] - o_s, out_len) * itemsize;  // bytes of this row that are copied
        const int64_t b0 = w * 4;
        if (fill) {  // pad bytes of this unit: [max(b0, n_bytes), min(b0 + 4, row_bytes))
            uint8_t *row = out + r * row_bytes;
            for (int64_t b = imax64(b0, n_bytes); b < imin64(b0 + 4, row_bytes); b++)
                row[b] = (uint8_t)(pad_bits >> (8 * (b % itemsize)));
        }
        if (b0 >= n_bytes) continue;
        const uint8_t *src = data + o_s * itemsize + b0;
        uint8_t *dst = out + r * row_bytes + b0;
        if (b0 + 4 <= n_bytes && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(src);
            const uint32_t *p = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
            const unsigned sh = (unsigned)(a & 3) * 8u;
            const uint32_t lo = __ldg(p);
            const uint32_t hi = sh ? __ldg(p + 1) : 0u;  // (bytes src..src+3 reach into the next word only when unaligned)
            *reinterpret_cast<uint32_t *>(dst) = __funnelshift_r(lo, hi, sh);
        } else {
            const int64_t n = imin64(4, n_bytes - b0);
            for (int64_t i = 0; i < n; i++) dst[i] = __ldg(src + i);
        }
    }
}

}  // namespace

namespace gvl {

// Device zeros shared by entries that drive the haplotype kernels without genotypes (get_reference).
int ensure_zeros(gvl_ctx *ctx, int64_t bytes, cudaStream_t st) {
    if (bytes <= ctx->zeros_bytes) return GVL_OK;
    GVL_CUDA(cudaDeviceSynchronize());
    if (ctx->zeros) GVL_CUDA(cudaFree(ctx->zeros));
    ctx->zeros = nullptr;
    ctx->zeros_bytes = 0;
    const int64_t cap = ((bytes * 2 + 4095) / 4096) * 4096;
    GVL_CUDA(cudaMalloc(&ctx->zeros, (size_t)cap));
    GVL_CUDA(cudaMemsetAsync(ctx->zeros, 0, (size_t)cap, st));
    GVL_CUDA(cudaStreamSynchronize(st));  // other streams may use the buffer next
    ctx->zeros_bytes = cap;
    return GVL_OK;
}

}  // namespace gvl

extern "C" {

int gvl_dev_choose_exonic_variants(gvl_ctx *ctx, const gvl_sparse_tables *tab, const int32_t *starts, const int32_t *ends,
                                   const int64_t *geno_offset_idx, int64_t n_queries, int64_t ploidy, uint8_t *keep,
                                   int64_t keep_cap, int64_t *keep_offsets, gvl_stream stream) {
    if (!ctx || !tab || !keep_offsets) return fail(GVL_ERR_ARG, "gvl_dev_choose_exonic_variants: NULL argument");
    if (n_queries < 0 || ploidy < 1 || keep_cap < 0) return fail(GVL_ERR_ARG, "gvl_dev_choose_exonic_variants: bad sizes");
    const int64_t n_work = n_queries * ploidy;
    if (n_work && (!starts || !ends || !geno_offset_idx || !tab->geno_starts || !tab->geno_stops))
        return fail(GVL_ERR_ARG, "gvl_dev_choose_exonic_variants: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GVL_CUDA(cudaSetDevice(ctx->device));
    exonic_offsets_kernel<<<1, SCAN_THREADS, 0, st>>>(geno_offset_idx, tab->geno_starts, tab->geno_stops, n_work,
                                                      keep_offsets);
    GVL_LAUNCH_CHECK();
    if (n_work && keep && keep_cap) {
        if (!tab->geno_v_idxs || !tab->v_starts || !tab->ilens)
            return fail(GVL_ERR_ARG, "gvl_dev_choose_exonic_variants: variant table is NULL");
        const int64_t blocks = imin64((n_work + 7) / 8, 148 * 8);
        exonic_keep_kernel<<<(unsigned)blocks, 256, 0, st>>>(*tab, starts, ends, geno_offset_idx, n_work, (int)ploidy,
                                                            keep_offsets, keep, keep_cap);
        GVL_LAUNCH_CHECK();
    }
    return GVL_OK;
}

static int ragged_to_padded_impl(gvl_ctx *ctx, const void *data, const int64_t *offsets, int64_t n_rows, void *out, int64_t itemsize,
                                 int64_t out_len, int fill, uint64_t pad_bits, gvl_stream stream);

int gvl_dev_ragged_to_padded(gvl_ctx *ctx, const void *data, const int64_t *offsets, int64_t n_rows, void *out,
                             int64_t itemsize, int64_t out_len, gvl_stream stream) {
    return ragged_to_padded_impl(ctx, data, offsets, n_rows, out, itemsize, out_len, 0, 0, stream);
}

int gvl_dev_ragged_to_padded_fill(gvl_ctx *ctx, const void *data, const int64_t *offsets, int64_t n_rows, void *out,
                                  int64_t itemsize, int64_t out_len, const void *pad_item, gvl_stream stream) {
    if (!pad_item || itemsize < 1 || itemsize > 8) return fail(GVL_ERR_ARG, "gvl_dev_ragged_to_padded_fill: pad item of 1..8 bytes");
    uint64_t bits = 0;
    memcpy(&bits, pad_item, (size_t)itemsize);
    return ragged_to_padded_impl(ctx, data, offsets, n_rows, out, itemsize, out_len, 1, bits, stream);
}

static int ragged_to_padded_impl(gvl_ctx *ctx, const void *data, const int64_t *offsets, int64_t n_rows, void *out, int64_t itemsize,
                                 int64_t out_len, int fill, uint64_t pad_bits, gvl_stream stream) {
    if (!ctx) return fail(GVL_ERR_ARG, "gvl_dev_ragged_to_padded: ctx is NULL");
    if (n_rows < 0 || itemsize < 1 || out_len < 0) return fail(GVL_ERR_ARG, "gvl_dev_ragged_to_padded: bad sizes");
    if (n_rows == 0 || out_len == 0) return GVL_OK;
    if (!data || !offsets || !out) return fail(GVL_ERR_ARG, "gvl_dev_ragged_to_padded: NULL argument");
    GVL_CUDA(cudaSetDevice(ctx->device));
    const int64_t words_per_row = (out_len * itemsize + 3) / 4;
    const int64_t n_units = n_rows * words_per_row;
    const int64_t blocks = imin64((n_units + 255) / 256, 148 * 32);
    ragged_to_padded_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t *)data, offsets, n_rows, (uint8_t *)out, itemsize, out_len, words_per_row, fill, pad_bits);
    GVL_LAUNCH_CHECK();
    return GVL_OK;
}

int gvl_dev_get_reference(gvl_ctx *ctx, const gvl_sparse_tables *tab, const int32_t *regions, int64_t *out_offsets,
                          int64_t n_regions, int64_t row_length, const uint8_t *to_rc, int mode, uint8_t pad_char,
                          uint8_t *out, gvl_stream stream) {
    if (!ctx || !tab || !tab->ref || !tab->ref_offsets || !out_offsets)
        return fail(GVL_ERR_ARG, "gvl_dev_get_reference: NULL argument");
    if (mode != GVL_MODE_U8 && mode != GVL_MODE_ONEHOT)
        return fail(GVL_ERR_ARG, "gvl_dev_get_reference: mode must be GVL_MODE_U8 or GVL_MODE_ONEHOT");
    if (n_regions < 0 || row_length < -1) return fail(GVL_ERR_ARG, "gvl_dev_get_reference: bad sizes");
    if (n_regions && !regions) return fail(GVL_ERR_ARG, "gvl_dev_get_reference: regions is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    GVL_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // Rows without variants through the haplotype kernels: every row points at one empty genotype slice, shift 0.
    if ((rc = ensure_zeros(ctx, 8 * (n_regions + 2) + 64, st))) return rc;
    const int64_t *z64 = (const int64_t *)ctx->zeros;
    gvl_sparse_tables t = *tab;
    t.geno_starts = z64;
    t.geno_stops = z64;
    t.n_geno = 1;
    t.geno_v_idxs = (const int32_t *)ctx->zeros;
    t.v_starts = (const int32_t *)ctx->zeros;
    t.ilens = (const int32_t *)ctx->zeros;
    t.alt_offsets = z64;
    t.alt_alleles = tab->ref;
    t.n_variants = 0;
    t.alt_packed = tab->ref_packed;  // (never read: there are no ALT pieces; present so the packed one-hot kernel is eligible)
    if ((rc = gvl_dev_hap_plan(ctx, &t, regions, (const int32_t *)ctx->zeros, z64, n_regions, 1, nullptr, nullptr, to_rc,
                               row_length >= 0 ? row_length : -2, 0, out_offsets, nullptr, stream)))
        return rc;
    int64_t total = 0;  // (caller-sized rows: one stream sync, like every ragged plan)
    if ((rc = gvl_dev_hap_total(ctx, stream, &total))) return rc;
    if (total == 0) return GVL_OK;
    if (!out) return fail(GVL_ERR_ARG, "gvl_dev_get_reference: out is NULL");
    return gvl_dev_hap_exec(ctx, &t, mode, pad_char, out, nullptr, nullptr, stream);
}

}  // extern "C"
