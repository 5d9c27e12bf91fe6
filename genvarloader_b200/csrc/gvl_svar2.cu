// gvl_svar2.cu -- device merge of the svar2 two-channel variant source.
//
// Reference: merge_hap (src/svar2/mod.rs:45-66) concatenates a haplotype's own `var_key` entries with
// the presence-bit-selected entries of the query's shared `dense` window and stable-sorts by position
// (var_key first on ties), per row, into a fresh Vec.  Both channels are position-sorted, so on the GPU
// this is a rank computation instead of a sort: a var_key entry moves back by the number of PRESENT dense
// entries strictly before it, a present dense entry by the number of var_key entries at or before it.
#include "gvl_internal.cuh"

namespace gvl {

struct MergeParams {
    gvl_svar2_channels ch;
    int64_t n_work, ploidy, cap;
    int32_t *m_pos, *m_key;
    int64_t *m_off;
    int32_t *m_len;
    int64_t *words;
};

__device__ __forceinline__ int present_bit(const uint8_t *__restrict__ bits, int64_t bit) {  // src/svar2/mod.rs:35-38
    return (bits[bit >> 3] >> (bit & 7)) & 1;
}

// number of set bits in [bit0, bit0 + n)
__device__ __forceinline__ int64_t count_bits(const uint8_t *__restrict__ bits, int64_t bit0, int64_t n) {
    int64_t c = 0, b = bit0, e = bit0 + n;
    while (b < e && (b & 7)) c += present_bit(bits, b++);
    while (e - b >= 8) {
        c += __popc((unsigned)bits[b >> 3]);
        b += 8;
    }
    while (b < e) c += present_bit(bits, b++);
    return c;
}

constexpr int MERGE_WARPS = 4;
constexpr int MERGE_WORDS = 512;  // presence words (32 dense entries each) a warp keeps in shared memory: 16,384 dense entries
constexpr int MERGE_POS = 512;    // var_key positions a warp stages in shared memory (longer lists search global memory)
constexpr int MERGE_DPOS = 1024;  // dense-window positions (the window holds the whole cohort's dense variants of the region)

// 32 presence bits of a row starting at its bit j (bits beyond the row's window are garbage: callers mask)
__device__ __forceinline__ uint32_t present_word(const uint8_t *__restrict__ bits, int64_t bit) {
    const uint8_t *p = bits + (bit >> 3);
    const unsigned sh = (unsigned)(bit & 7);
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) v |= (uint64_t)p[i] << (8 * i);  // (the presence buffer carries 8 bytes of slack)
    return (uint32_t)(v >> sh);
}

__global__ void __launch_bounds__(MERGE_WARPS * 32) svar2_merge_kernel(MergeParams P) {
    // per warp: the row's presence bits as 32-bit words + the number of present entries before each word
    __shared__ uint32_t s_word[MERGE_WARPS][MERGE_WORDS];
    __shared__ int32_t s_pre[MERGE_WARPS][MERGE_WORDS];
    // both channels' positions, staged with coalesced loads: the rank searches below are then chains of shared-memory
    // reads instead of ~10 dependent global round trips per entry (33 -> 10 us for the 1,280 rows of a 20-batch cfg2 call)
    __shared__ int32_t s_vpos[MERGE_WARPS][MERGE_POS];
    __shared__ int32_t s_dpos[MERGE_WARPS][MERGE_DPOS];
    __shared__ uint16_t s_pj[MERGE_WARPS][MERGE_DPOS];  // dense index of the k-th PRESENT entry (compacted window)
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * MERGE_WARPS + wid;
    if (k >= P.n_work) return;
    const int64_t query = k / P.ploidy;
    const int64_t ks = P.ch.row_slot ? P.ch.row_slot[k] : k;  // entry of the per-row tables
    const int64_t vk_lo = P.ch.vk_off[ks], vk_hi = P.ch.vk_stop ? P.ch.vk_stop[ks] : P.ch.vk_off[ks + 1];
    const int64_t n_vk = imax64(vk_hi - vk_lo, 0);
    const int64_t qs = P.ch.query_div > 0 ? (P.ch.row_slot ? P.ch.row_slot[query * P.ploidy] : query * P.ploidy) / P.ch.query_div : query;
    const int64_t ds = P.ch.dense_range[qs * 2], de = P.ch.dense_range[qs * 2 + 1];
    const int64_t nd = imax64(de - ds, 0);
    const int64_t base_bit = P.ch.dense_present_off[ks];
    const int32_t *__restrict__ vpos = P.ch.vk_pos + vk_lo;
    const int32_t *__restrict__ vkey = P.ch.vk_key + vk_lo;
    const int32_t *__restrict__ dpos = P.ch.dense_pos + ds;
    const int32_t *__restrict__ dkey = P.ch.dense_key + ds;

    int64_t off = 0;
    if (lane == 0) off = (int64_t)atomicAdd((unsigned long long *)&P.words[W_MERGE_CURSOR], (unsigned long long)(n_vk + nd));
    off = __shfl_sync(0xffffffffu, off, 0);
    const bool overflow = off + n_vk + nd > P.cap;
    if (overflow) {
        if (lane == 0) {
            atomicMax((unsigned long long *)&P.words[W_STATUS], (unsigned long long)(off + n_vk + nd));
            P.m_off[k] = 0;
            P.m_len[k] = 0;
        }
        return;
    }
    const bool v_staged = n_vk <= MERGE_POS, d_staged = nd <= MERGE_DPOS;
    if (v_staged)
        for (int64_t i = lane; i < n_vk; i += 32) s_vpos[wid][i] = vpos[i];
    if (d_staged)
        for (int64_t i = lane; i < nd; i += 32) s_dpos[wid][i] = dpos[i];
    const int32_t *__restrict__ vsrch = v_staged ? s_vpos[wid] : vpos;  // (generic pointers: shared or global)
    const int32_t *__restrict__ dsrch = d_staged ? s_dpos[wid] : dpos;
    // presence words + exclusive prefix popcounts (a warp scan per 32 words) when the window fits shared memory
    const int64_t n_words = (nd + 31) >> 5;
    const bool staged = n_words <= MERGE_WORDS;
    if (staged) {
        int carry = 0;
        for (int64_t w0 = 0; w0 < n_words; w0 += 32) {
            const int64_t w = w0 + lane;
            uint32_t word = 0;
            if (w < n_words) {
                word = present_word(P.ch.dense_present, base_bit + 32 * w);
                const int64_t left = nd - 32 * w;
                if (left < 32) word &= (1u << left) - 1u;
            }
            int x = __popc(word);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (w < n_words) {
                s_word[wid][w] = word;
                s_pre[wid][w] = carry + x - __popc(word);
            }
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
    __syncwarp();  // staged positions and presence words are visible to the whole warp
    // present entries strictly before dense index j
    auto rank = [&](int64_t j) -> int64_t {
        if (!staged) return count_bits(P.ch.dense_present, base_bit, j);
        const int64_t w = j >> 5;
        if (w >= n_words) return s_pre[wid][n_words - 1] + __popc(s_word[wid][n_words - 1]);
        return s_pre[wid][w] + __popc(s_word[wid][w] & ((1u << (j & 31)) - 1u));
    };
    // dense entries that are present: rank among present + var_key entries at or before them
    int64_t n_present = 0;
    if (d_staged && staged) {
        // Compact the window to its PRESENT entries first (shared memory only: positions in place, dense indices beside
        // them), so that the loops that touch global memory run over ~n_present / 32 trips instead of nd / 32 -- a trip
        // is a dependent load -> store, and with one warp per row nothing else hides it.
        for (int64_t base = 0; base < nd; base += 32) {
            const int64_t j = base + lane;
            const bool pr = (j < nd) && ((s_word[wid][j >> 5] >> (j & 31)) & 1u) != 0;
            const int32_t p = j < nd ? s_dpos[wid][j] : 0;
            const unsigned mask = __ballot_sync(0xffffffffu, pr);
            __syncwarp();  // every lane has read its entry before the compacted prefix (indices <= j) is written
            if (pr) {
                const int64_t r = n_present + __popc(mask & ((1u << lane) - 1u));
                s_dpos[wid][r] = p;
                s_pj[wid][r] = (uint16_t)j;
            }
            n_present += __popc(mask);
        }
        __syncwarp();
        const int32_t *__restrict__ ppos = s_dpos[wid];
        for (int64_t i = lane; i < n_present; i += 32) {
            const int32_t p = ppos[i];
            int64_t lo = 0, hi = n_vk;  // number of var_key entries with pos <= p
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (vsrch[mid] <= p) lo = mid + 1; else hi = mid;
            }
            const int64_t w = off + i + lo;
            P.m_pos[w] = p;
            P.m_key[w] = dkey[s_pj[wid][i]];
        }
        for (int64_t i = lane; i < n_vk; i += 32) {
            const int32_t p = vsrch[i];
            int64_t lo = 0, hi = n_present;  // number of PRESENT dense entries with pos < p
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (ppos[mid] < p) lo = mid + 1; else hi = mid;
            }
            const int64_t w = off + i + lo;
            P.m_pos[w] = p;
            P.m_key[w] = vkey[i];
        }
        if (lane == 0) {
            P.m_off[k] = off;
            P.m_len[k] = (int32_t)(n_vk + n_present);
        }
        return;
    }
    for (int64_t base = 0; base < nd; base += 32) {
        const int64_t j = base + lane;
        const bool pr = (j < nd) && (staged ? ((s_word[wid][j >> 5] >> (j & 31)) & 1u) != 0 : present_bit(P.ch.dense_present, base_bit + j) != 0);
        const unsigned mask = __ballot_sync(0xffffffffu, pr);
        if (pr) {
            const int32_t p = dsrch[j];
            int64_t lo = 0, hi = n_vk;  // number of var_key entries with pos <= p
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (vsrch[mid] <= p) lo = mid + 1; else hi = mid;
            }
            const int64_t w = off + n_present + __popc(mask & ((1u << lane) - 1u)) + lo;
            P.m_pos[w] = p;
            P.m_key[w] = dkey[j];
        }
        n_present += __popc(mask);
    }
    // var_key entries: move back by the present dense entries strictly before them
    for (int64_t i = lane; i < n_vk; i += 32) {
        const int32_t p = vsrch[i];
        int64_t lo = 0, hi = nd;  // number of dense entries with pos < p
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (dsrch[mid] < p) lo = mid + 1; else hi = mid;
        }
        const int64_t w = off + i + (nd > 0 ? rank(lo) : 0);
        P.m_pos[w] = p;
        P.m_key[w] = vkey[i];
    }
    if (lane == 0) {
        P.m_off[k] = off;
        P.m_len[k] = (int32_t)(n_vk + n_present);
    }
}

}  // namespace gvl

using namespace gvl;

int gvl_svar2_merge_launch(gvl_ctx *ctx, gvl_workspace *ws, int64_t *words, const gvl_svar2_channels *ch, int64_t batch,
                           int64_t ploidy, int64_t max_merged, cudaStream_t st) {
    (void)ctx;
    const int64_t n_work = batch * ploidy;
    if (n_work == 0) return GVL_OK;
    MergeParams P;
    P.ch = *ch;
    P.n_work = n_work;
    P.ploidy = ploidy;
    P.cap = ws->m_cap;
    P.m_pos = ws->m_pos;
    P.m_key = ws->m_key;
    P.m_off = ws->m_off;
    P.m_len = ws->m_len;
    P.words = words;
    (void)max_merged;
    svar2_merge_kernel<<<(unsigned)((n_work + MERGE_WARPS - 1) / MERGE_WARPS), MERGE_WARPS * 32, 0, st>>>(P);
    GVL_LAUNCH_CHECK();
    return GVL_OK;
}
