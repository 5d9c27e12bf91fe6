// gvl_tracks.cu -- indel-aware track realignment on sm_100a (include/gvl_b200.h:
// gvl_dev_realign_tracks, gvl_dev_intervals_to_tracks).
//
// Reference path replaced: intervals_and_realign_track_fused (src/ffi/mod.rs:2553-2672) =
// intervals_to_tracks (src/intervals.rs:19-126) into a dense scratch -> shift_and_realign_tracks_sparse
// (src/tracks/mod.rs:224-406, 495-667) -> reverse_flat_rows_inplace (src/reverse.rs:25-38), called
// once per track from a Python loop (_dataset/_reconstruct.py:228-290).
//
// Here: ONE plan launch per batch (the variant state machine does not depend on the track) and ONE
// execute launch for all tracks.  Each execute CTA owns a tile of one (track, row): it paints the
// stored intervals that cover the tile's source window into shared memory (no dense scratch in
// HBM), then every thread produces 4 consecutive output values per step (source span copies,
// deletion/insertion fills, trailing zeros, reversal for negative strands) and writes them with one
// 16-byte store.
#include "gvl_internal.cuh"

namespace gvl {

#ifndef GVL_TRK_TILE
#define GVL_TRK_TILE 8192
#endif
#ifndef GVL_TRK_THREADS
#define GVL_TRK_THREADS 256
#endif
constexpr int TRK_TILE = GVL_TRK_TILE;             // output values per pass of an execute CTA (one staged source window)
constexpr int TRK_WIN = TRK_TILE + TRK_TILE / 8;   // source-window values staged in shared memory (room for net deletions)
#ifndef GVL_TRK_SEG_TILES
#define GVL_TRK_SEG_TILES 8
#endif
constexpr int TRK_SEG = GVL_TRK_SEG_TILES * TRK_TILE;       // output values per execute CTA: up to 8 tiles, walked in haplotype order
constexpr int TRK_MARGIN = 16;             // window starts a little before the first needed value
constexpr int TRK_THREADS = GVL_TRK_THREADS;
constexpr int TRK_REC_CAP = 128;

// =====================================================================================
// plan: shift_and_realign_track_core state machine (src/tracks/mod.rs:224-406), one warp per row
// =====================================================================================
struct TrkPlanParams {
    gvl_sparse_tables tab;
    MergedLists merged;
    const int32_t *regions;
    const int32_t *shifts;
    const int64_t *goi;
    const uint8_t *keep;
    const int64_t *keep_off;
    const uint8_t *to_rc;
    const int32_t *track_lengths;  // [batch]
    const int64_t *out_offsets;    // [n_work+1]
    int64_t n_work, ploidy, rec_cap;
    RowPlan *rows;
    RecArrays rec;
    int64_t *words;
    int32_t *row_len;
};

constexpr int TPLAN_WARPS = 4;

__global__ void __launch_bounds__(TPLAN_WARPS * 32) trk_plan_kernel(TrkPlanParams P) {
    const int lane = lane_id();
    const int64_t k = (int64_t)blockIdx.x * TPLAN_WARPS + (threadIdx.x >> 5);
    if (k >= P.n_work) return;
    const int64_t query = k / P.ploidy;
    const RowVars rv = row_vars(P.tab, P.merged, P.goi, k);
    const int64_t nvar = rv.nvar;
    const int64_t q_start = P.regions[query * 3 + 1];
    const int64_t shift = P.shifts[k];
    const bool has_keep = (P.keep && P.keep_off);
    const int64_t keep_base = has_keep ? P.keep_off[k] : 0;
    const int32_t *__restrict__ gv = rv.gv;
    const int64_t length = imax64(P.out_offsets[k + 1] - P.out_offsets[k], 0);
    const int64_t track_n = P.track_lengths[query];

    int64_t rec_off = 0;
    if (lane == 0) rec_off = (int64_t)atomicAdd((unsigned long long *)&P.words[W_CURSOR], (unsigned long long)(nvar + 1));
    rec_off = __shfl_sync(0xffffffffu, rec_off, 0);
    const bool overflow = rec_off + nvar + 1 > P.rec_cap;
    if (overflow && lane == 0) atomicMax((unsigned long long *)&P.words[W_STATUS], (unsigned long long)(rec_off + nvar + 1));

    TrkState ts;
    trk_init(ts, shift, length);
    int64_t n_emit = 0, track0 = 0, prev_resume = 0;
    bool done = false;
    // chunk loader: variant i = base + lane of the row (positions, ilens, keep flag)
    auto load_chunk = [&](int64_t base, int32_t &pos, int32_t &il, bool &kp) {
        pos = 0, il = 0, kp = false;
        const int64_t i = base + lane;
        if (i < nvar) {
            const int32_t vi = gv[i];
            pos = (int32_t)var_pos(P.tab, rv, i, vi);
            il = P.tab.ilens[vi];
            kp = has_keep ? (P.keep[keep_base + i] != 0) : true;
        }
    };
    int32_t pos, il, n_pos = 0, n_il = 0;
    bool kp, n_kp = false;
    load_chunk(0, pos, il, kp);
    for (int64_t base = 0; base < nvar && !done; base += 32) {
        if (base + 32 < nvar) load_chunk(base + 32, n_pos, n_il, n_kp);  // next chunk's gathers fly during this one
        // once the shift is consumed a SNP (ilen 0) changes nothing (src/tracks/mod.rs:277-314: skipped or
        // "writes nothing"), so only indels take part
        const bool part = kp && (il != 0 || ts.shifted < ts.shift);
        unsigned mask = __ballot_sync(0xffffffffu, part);
        const int64_t rel = (int64_t)pos - q_start;                        // v_rel_pos (:264)
        const int64_t v_end = rel - imin64(il, 0) + 1;                     // v_rel_end (:267)
        // ---- whole chunk at once: shift consumed, nothing left of the window, and every participating indel starts
        //      at or after the end of the previous one (no overlap -> every one is applied, :277-279) ----
        bool fast = mask != 0 && ts.shifted >= ts.shift && (n_emit == 0 || ts.track_idx == prev_resume) &&
                    !__any_sync(0xffffffffu, part && rel < 0);
        int64_t prev_end = ts.track_idx;
        if (fast) {
            const unsigned below = mask & ((1u << lane) - 1u);
            const int pl = below ? 31 - __clz(below) : 0;
            const int64_t pe = __shfl_sync(0xffffffffu, v_end, pl);
            if (below) prev_end = pe;
            fast = !__any_sync(0xffffffffu, part && rel < prev_end);
        }
        if (fast) {
            const int64_t v_len = imax64(il, 0) + 1;                         // :282
            const int64_t ref_len = part ? rel - prev_end : 0;               // track_len (:317)
            int64_t inc = part ? ref_len + v_len : 0, scan = inc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, scan, o);
                if (lane >= o) scan += y;
            }
            const int64_t a = ts.out_idx + (scan - inc) + ref_len;           // out_idx after the span copy
            const bool valid = part && a < ts.length;                        // :319-321 (positions grow: a prefix)
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const bool broke = __any_sync(0xffffffffu, part && !valid);
            const int64_t n = valid ? imin64(v_len, ts.length - a) : 0;      // writable_length (:329)
            if (vmask) {
                const int last = 31 - __clz(vmask);
                if (n_emit == 0) track0 = ts.track_idx;  // span_src of the first record
                if (valid && !overflow) {
                    const int64_t w = rec_off + n_emit + __popc(vmask & ((1u << lane) - 1u));
                    P.rec.a[w] = (int32_t)a;
                    P.rec.n[w] = (int32_t)n;
                    P.rec.src[w] = il;
                    P.rec.resume[w] = (int32_t)v_end;
                    P.rec.vidx[w] = (int32_t)v_len;
                    P.rec.vpos[w] = (int32_t)rel;
                }
                n_emit += __popc(vmask);
                ts.out_idx = __shfl_sync(0xffffffffu, a + n, last);
                ts.track_idx = __shfl_sync(0xffffffffu, v_end, last);
                prev_resume = ts.track_idx;
                if (ts.out_idx >= ts.length) done = true;  // :359-361
            }
            if (broke) done = true;
            mask = 0;
        }
        while (mask && !done) {
            int t = __ffs(mask) - 1;
            mask &= mask - 1;
            int64_t p = __shfl_sync(0xffffffffu, pos, t);
            int64_t l = __shfl_sync(0xffffffffu, il, t);
            TrkRec r;
            int act = trk_step(ts, p - q_start, l, r);  // v_rel_pos = v_start - query_start (:264)
            if (act == STEP_BREAK) {
                done = true;
            } else if (act == STEP_EMIT) {
                if (n_emit == 0) track0 = r.span_src;
                if (n_emit > 0 && r.span_src != prev_resume) {  // unsorted input: jump record
                    if (lane == t && !overflow) {
                        int64_t w = rec_off + n_emit;
                        P.rec.a[w] = (int32_t)(r.a - (r.v_rel_pos - r.span_src));
                        P.rec.n[w] = 0;
                        P.rec.src[w] = 0;
                        P.rec.resume[w] = (int32_t)r.span_src;
                        P.rec.vidx[w] = 1;
                        P.rec.vpos[w] = 0;
                    }
                    n_emit++;
                }
                if (lane == t && !overflow) {
                    int64_t w = rec_off + n_emit;
                    P.rec.a[w] = (int32_t)r.a;
                    P.rec.n[w] = (int32_t)r.n;
                    P.rec.src[w] = r.v_diff;
                    P.rec.resume[w] = (int32_t)r.resume;
                    P.rec.vidx[w] = (int32_t)r.v_len;
                    P.rec.vpos[w] = (int32_t)r.v_rel_pos;
                }
                n_emit++;
                prev_resume = r.resume;
                if (ts.out_idx >= ts.length) done = true;  // :359-361
            }
        }
        pos = n_pos, il = n_il, kp = n_kp;
    }
    if (nvar == 0) {
        track0 = 0;  // :240-246: an EMPTY variant list copies track[:length], whatever the shift
    } else {
        trk_finish(ts, track_n);
        if (n_emit == 0) {
            track0 = ts.track_idx;
        } else if (ts.track_idx != prev_resume) {
            if (lane == 0 && !overflow) {
                int64_t w = rec_off + n_emit;
                P.rec.a[w] = (int32_t)imin64(ts.out_idx, length);
                P.rec.n[w] = 0;
                P.rec.src[w] = 0;
                P.rec.resume[w] = (int32_t)ts.track_idx;
                P.rec.vidx[w] = 1;
                P.rec.vpos[w] = 0;
            }
            n_emit++;
        }
    }
    if (lane == 0) {
        RowPlan rp;
        rp.out_off = P.out_offsets[k];
        rp.ref_base = 0;
        rp.rec_off = rec_off;
        rp.length = (int32_t)length;
        rp.contig_len = (int32_t)track_n;
        rp.lead_pad = 0;
        rp.ref0 = (int32_t)track0;
        rp.n_rec = overflow ? 0 : (int32_t)n_emit;
        rp.rc = (P.to_rc && P.to_rc[k]) ? 1 : 0;
        rp.diff = 0;
        rp.q_start = (int32_t)q_start;
        P.rows[k] = rp;
        P.row_len[k] = (int32_t)length;
        plan_row_done(P.words, P.n_work);
    }
}

// identity plan for intervals_to_tracks: one row per query, no records, source window = row
__global__ void paint_plan_kernel(int64_t n, const int32_t *__restrict__ starts, const int64_t *__restrict__ out_offsets,
                                  const uint8_t *__restrict__ to_rc, RowPlan *rows, int32_t *row_len) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int64_t length = imax64(out_offsets[q + 1] - out_offsets[q], 0);
    RowPlan rp;
    rp.out_off = out_offsets[q];
    rp.ref_base = 0;
    rp.rec_off = 0;
    rp.length = (int32_t)length;
    rp.contig_len = (int32_t)length;
    rp.lead_pad = 0;
    rp.ref0 = 0;
    rp.n_rec = 0;
    rp.rc = (to_rc && to_rc[q]) ? 1 : 0;  // negative-strand rows come out reversed (src/reverse.rs:25-38)
    rp.diff = 0;
    rp.q_start = starts[q];
    rows[q] = rp;
    row_len[q] = (int32_t)length;
}

// tile map for TRK_SEG-sized segments (same scan as the haplotype path, different tile size)
__global__ void __launch_bounds__(1024) trk_tile_scan_kernel(int64_t n_work, const int32_t *__restrict__ row_len,
                                                             int64_t *tile_off) {
    __shared__ int64_t s_tile[32];
    __shared__ int64_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_work; base += 1024) {
        int64_t k = base + tid;
        int64_t til = (k < n_work) ? ((int64_t)row_len[k] + TRK_SEG - 1) / TRK_SEG : 0;
        int64_t x = til;
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_tile[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_tile[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_tile[lane] = w;
        }
        __syncthreads();
        if (k < n_work) tile_off[k] = carry + (warp ? s_tile[warp - 1] : 0) + x - til;
        __syncthreads();
        if (tid == 0) carry += s_tile[31];
        __syncthreads();
    }
    if (tid == 0) tile_off[n_work] = carry;
}

// =====================================================================================
// execute
// =====================================================================================
constexpr int MAX_TRACKS = 64;

struct TrkDesc {
    const int32_t *itv_starts;
    const int32_t *itv_ends;
    const float *itv_values;
    const int64_t *itv_offsets;
    const float *dense;            // non-NULL: dense f32 source windows instead of intervals
    const int64_t *dense_offsets;  //           i64[n_queries+1] offsets of each query's window
    int32_t strategy;
    double param;
};

struct TrkExecParams {
    const RowPlan *rows;
    RecArrays rec;
    const int64_t *tile_off;   // [n_work+1]
    int64_t n_work, ploidy;
    int64_t grid_per_track;    // CTAs per track (upper bound on tiles)
    int64_t total_per_track;   // output values per track
    int64_t n_tracks;
    int32_t layout_btp;        // 0: out is track-major (t, b, p, ~l); 1: (b, t, p, ~l), every query's tracks adjacent
    const int64_t *offset_idxs;  // [n_tracks * n_queries]
    int64_t n_queries;
    const int64_t *query_seed;   // optional [n_queries]
    uint64_t base_seed;
    const uint64_t *base_seed_dev;  // optional: the seed(s) live in device memory (CUDA-graph replays with new batches)
    int64_t sub_batch;              // > 0: queries per logical batch (seed and FlankSample row index are per logical batch)
    float *out;
    const TrkDesc *tracks;       // device array [n_tracks], or NULL: the descriptors travel in `inl`
    TrkDesc inl[8];              // (kernel parameters are captured by value: CUDA-graph safe)
};

struct TrkTileRecs {
    int32_t a[TRK_REC_CAP + 1];
    int32_t e[TRK_REC_CAP];
    int32_t resume[TRK_REC_CAP];
    int32_t vlen[TRK_REC_CAP];
    int32_t vrel[TRK_REC_CAP];
    int32_t vdiff[TRK_REC_CAP];
};

// value of the painted source track at relative position tp (0 <= tp < track_n), straight from the
// interval SoA: last interval with start <= q_start + tp, if it also ends after it (intervals are
// sorted and non-overlapping, src/intervals.rs contract).
__device__ __forceinline__ float track_at_global(const TrkDesc &T, int64_t lo, int64_t hi, int64_t q_start, int64_t tp) {
    if (T.dense) return T.dense[lo + tp];  // dense source: `lo` is the window's offset
    const int64_t g = q_start + tp;
    int64_t a = lo, b = hi;  // find last i in [lo,hi) with starts[i] <= g
    while (a < b) {
        int64_t mid = (a + b) >> 1;
        if ((int64_t)T.itv_starts[mid] <= g) a = mid + 1; else b = mid;
    }
    const int64_t i = a - 1;
    if (i < lo) return 0.0f;
    return ((int64_t)T.itv_ends[i] > g) ? T.itv_values[i] : 0.0f;
}

struct TrkSrc {
    const float *win;   // shared-memory window
    int64_t w0, w1;     // window covers source positions [w0, w1)
    int64_t track_n;
    const TrkDesc *T;
    int64_t itv_lo, itv_hi, q_start;
    __device__ __forceinline__ float at(int64_t tp) const {  // 0 <= tp < track_n expected
        if (tp >= w0 && tp < w1) return win[tp - w0];
        if (tp < 0 || tp >= track_n) return 0.0f;  // out of contract in the reference (index panic)
        return track_at_global(*T, itv_lo, itv_hi, q_start, tp);
    }
};

// Lagrange interpolation through K anchors on each side of the insertion (src/tracks/mod.rs:138-188), evaluated at
// index i of the written values; same operation order as the reference (term = y_a * prod_b (x - x_b) / (x_a - x_b)).
template <int K>
__device__ __forceinline__ float lagrange_fill(const TrkSrc &S, int64_t v_len, int64_t v_rel_pos, int64_t i) {
    double xs[2 * K], ys[2 * K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        xs[j] = -(double)j;
        ys[j] = (double)S.at(imax64(v_rel_pos - j, 0));
        xs[K + j] = (double)v_len + (double)j;
        ys[K + j] = (double)S.at(imin64(v_rel_pos + 1 + j, S.track_n - 1));
    }
    const double x = (double)i;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2 * K; a++) {
        double term = ys[a];
#pragma unroll
        for (int b = 0; b < 2 * K; b++) {
            if (b == a) continue;
            term = __dmul_rn(term, __ddiv_rn(__dsub_rn(x, xs[b]), __dsub_rn(xs[a], xs[b])));
        }
        acc = __dadd_rn(acc, term);
    }
    return (float)acc;
}

// apply_insertion_fill, src/tracks/mod.rs:87-190, for ONE written value (index i within the write).
__device__ float insertion_fill_value(const TrkSrc &S, int strategy, double param, int64_t v_len, int64_t v_rel_pos,
                                      int64_t i, int64_t out_pos, uint64_t base_seed, uint64_t query, uint64_t hap) {
    if (strategy == GVL_FILL_REPEAT_5P) {
        return S.at(v_rel_pos);
    } else if (strategy == GVL_FILL_REPEAT_5P_NORM) {
        return __fdiv_rn(S.at(v_rel_pos), (float)v_len);  // :115
    } else if (strategy == GVL_FILL_CONSTANT) {
        return (float)param;  // :121
    } else if (strategy == GVL_FILL_FLANK_SAMPLE) {  // :125-137
        int64_t width = (int64_t)param;
        int64_t pool_lo = imax64(v_rel_pos - width, 0);
        int64_t pool_hi = imin64(v_rel_pos + width, S.track_n - 1);
        uint64_t pool_size = (uint64_t)(pool_hi - pool_lo + 1);
        uint64_t seed = hash4(base_seed, query, hap, (uint64_t)out_pos);
        int64_t offset = (int64_t)(seed % pool_size);
        return S.at(pool_lo + offset);
    } else {  // GVL_FILL_INTERPOLATE :138-188
        const int64_t order = (int64_t)param;
        const int64_t k = (order + 1 + 1) / 2;
        // k anchors on each side: 2 anchors for order 1, 4 for orders 2 and 3 -- fixed-size instantiations keep the
        // anchors in registers and let the divisions overlap; the operation order is the reference's
        return k == 1 ? lagrange_fill<1>(S, v_len, v_rel_pos, i) : lagrange_fill<2>(S, v_len, v_rel_pos, i);
    }
}

#ifndef GVL_TRACE
#define GVL_TRACE 0
#endif
#if GVL_TRACE
__device__ unsigned long long *g_trk_trace = nullptr;  // [n_ctas][64]: per pass 6 globaltimer stamps (trace builds only)
#define TRK_TR(slot)                                                                                        \
    do {                                                                                                    \
        if (g_trk_trace && threadIdx.x == 0 && tr_pass < 10) {                                              \
            unsigned long long t_;                                                                          \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                           \
            g_trk_trace[(unsigned long long)blockIdx.x * 64 + tr_pass * 6 + (slot)] = t_;                     \
        }                                                                                                   \
    } while (0)
#else
#define TRK_TR(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(TRK_THREADS, 1024 / TRK_THREADS) trk_exec_kernel(TrkExecParams P) {
#if GVL_TRACE
    int tr_pass = 0;
#endif
    __shared__ TrkTileRecs S;
    extern __shared__ __align__(16) float s_win[];  // TRK_WIN floats (dynamic: larger than 48 KB in big-pass builds)
    __shared__ uint32_t s_flag[TRK_WIN / 32 + 4];  // positions of s_win that hold a run start (interval start / end)
    __shared__ float s_cval[TRK_THREADS / 32];
    __shared__ int32_t s_gt[TRK_TILE / 128 + 2];  // per group of the pass: window offset of a plain group, or -1
    __shared__ int s_chas[TRK_THREADS / 32];
    __shared__ int s_stop;
    __shared__ int64_t s_lo, s_hi, s_itv_first;

    const int64_t track = blockIdx.x / P.grid_per_track;
    const int64_t b = blockIdx.x % P.grid_per_track;
    if (b >= P.tile_off[P.n_work]) return;
#ifdef GVL_TRK_STAGGER_NS
    // CTAs that share an SM start GVL_TRK_STAGGER_NS apart, so that one CTA's preparation phases (latency) overlap
    // another's output phase (bandwidth) instead of all CTAs of the GPU marching in lock-step
    {
        unsigned nsm;
        asm volatile("mov.u32 %0, %nsmid;" : "=r"(nsm));
        const unsigned phase = (blockIdx.x / nsm) & 3u;
        if (phase) __nanosleep(phase * GVL_TRK_STAGGER_NS);
    }
#endif
    int64_t row;
    {
        int64_t lo = 0, hi = P.n_work;
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (P.tile_off[mid] <= b) lo = mid; else hi = mid;
        }
        row = lo;
    }
    const int64_t tile = b - P.tile_off[row];
    const RowPlan rp = P.rows[row];
    const int32_t L = rp.length;
    const int32_t t0 = (int32_t)(tile * TRK_SEG);
    if (t0 >= L) return;
    const int32_t t1 = (int32_t)imin64((int64_t)t0 + TRK_SEG, L);
    const bool rc = rp.rc != 0;
    const int32_t h0 = rc ? L - t1 : t0;
    const int32_t h1 = rc ? L - t0 : t1;
    const int64_t query = row / P.ploidy;
    const uint64_t hap = (uint64_t)(row % P.ploidy);
    const uint64_t qseed = P.query_seed ? (uint64_t)P.query_seed[query]
                                        : (uint64_t)(P.sub_batch > 0 ? query % P.sub_batch : query);
    const uint64_t base_seed = P.base_seed_dev ? P.base_seed_dev[P.sub_batch > 0 ? query / P.sub_batch : 0] : P.base_seed;
    const TrkDesc T = P.tracks ? P.tracks[track] : P.inl[track];
    int64_t itv_lo, itv_hi;
    if (T.dense) {
        itv_lo = T.dense_offsets[query];
        itv_hi = T.dense_offsets[query + 1];
    } else {
        const int64_t slot = P.offset_idxs[track * P.n_queries + query];
        itv_lo = T.itv_offsets[slot];
        itv_hi = T.itv_offsets[slot + 1];
    }
    const int64_t track_n = rp.contig_len;
    const int64_t q_start = rp.q_start;
    float *__restrict__ out = P.out;
    int64_t row_base = track * P.total_per_track + rp.out_off;  // flat index of the row's first value
    if (P.layout_btp) {  // all tracks of a query are adjacent: block of the query, then track, then the row inside the block
        const int64_t k0 = query * P.ploidy;
        const int64_t blk0 = P.rows[k0].out_off;
        const int64_t blk_len = P.rows[k0 + P.ploidy - 1].out_off + P.rows[k0 + P.ploidy - 1].length - blk0;
        row_base = P.n_tracks * blk0 + track * blk_len + (rp.out_off - blk0);
    }

    const int32_t *__restrict__ ra = P.rec.a + rp.rec_off;
    if (threadIdx.x < 32) {
        // r_lo = last record with a <= h0 (or -1), r_hi = first with a >= h1: counts over the sorted array,
        // 8 independent loads per lane and round trip
        const int lane = threadIdx.x;
        int64_t r_lo, r_hi;
        if (rp.n_rec <= 2048) {
            int c0 = 0, c1 = 0;
            for (int i0 = 0; i0 < rp.n_rec; i0 += 256) {
                int32_t a[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + 32 * u + lane;
                    a[u] = i < rp.n_rec ? ra[i] : INT32_MAX;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    c0 += (a[u] <= h0);
                    c1 += (a[u] < h1);
                }
            }
            r_lo = (int64_t)__reduce_add_sync(0xffffffffu, c0) - 1;
            r_hi = __reduce_add_sync(0xffffffffu, c1);
        } else {
            r_lo = warp_upper_le(ra, 0, rp.n_rec, h0);
            r_hi = warp_upper_le(ra, imax64(r_lo, 0), rp.n_rec, h1 - 1) + 1;
        }
        if (threadIdx.x == 0) {
            s_lo = r_lo;
            s_hi = r_hi;
        }
    }
    __syncthreads();
    const int64_t r_hi = s_hi;
    int64_t r = s_lo;
    int32_t cur = h0;
    int64_t itv_prev = -1;  // first interval of the previous pass's window (-1: none yet)
    int32_t tgt_prev = INT32_MIN;

    while (cur < h1) {
        TRK_TR(0);
        const int m_new = (int)imin64(TRK_REC_CAP - 1, r_hi - (r + 1));
        const int m = m_new + 1;
        const int32_t seg_end_rec = (r + 1 + m_new < r_hi) ? ra[r + 1 + m_new] : h1;
        // a pass ends at the first unstaged record, after TRK_TILE values (one source window), or at the segment end
        const int32_t seg_end = (int32_t)imin64(seg_end_rec, (int64_t)cur + TRK_TILE);
        if (seg_end <= cur) {  // (only with > 127 records at one position: skip them, nothing to write)
            r += m_new;
            continue;
        }
        // Later passes: the window only moves forward, so the intervals it needs start at the previous pass's
        // cursor.  Their loads are issued here, together with the record loads below (one round trip, not three).
        int32_t pre_st = INT32_MAX, pre_en = 0;
        float pre_v = 0.0f;
        const bool pre_ok = !T.dense && itv_prev >= 0;
        if (pre_ok && itv_prev + (int64_t)threadIdx.x < itv_hi) {
            const int64_t it = itv_prev + threadIdx.x;
            pre_st = T.itv_starts[it];
            pre_en = T.itv_ends[it];
            pre_v = T.itv_values[it];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += TRK_THREADS) {
            int64_t idx = r + i;
            if (idx < 0) {
                S.a[0] = 0;
                S.e[0] = 0;
                S.resume[0] = rp.ref0;
                S.vlen[0] = 1;
                S.vrel[0] = 0;
                S.vdiff[0] = 0;
            } else {
                int64_t g = rp.rec_off + idx;
                int32_t a = P.rec.a[g], n = P.rec.n[g];
                S.a[i] = a;
                S.e[i] = a + n;
                S.resume[i] = P.rec.resume[g];
                S.vlen[i] = P.rec.vidx[g];
                S.vrel[i] = P.rec.vpos[g];
                S.vdiff[i] = (int32_t)P.rec.src[g];
            }
        }
        if (threadIdx.x == 0) S.a[m] = INT32_MAX;
        __syncthreads();

        TRK_TR(1);
        // ---- source window: starts at the source position of `cur` (minus a margin) ----
        // source position feeding `cur`: inside the carry record's own values the reads go to
        // track[v_rel_pos] and then continue at its resume point; otherwise we are in its span.
        const int64_t src_cur = (cur < S.e[0]) ? imin64((int64_t)S.vrel[0], (int64_t)S.resume[0])
                                               : (int64_t)S.resume[0] + (cur - S.e[0]);
        const int64_t w0 = imax64(src_cur - TRK_MARGIN, 0);
        const int64_t w1 = imin64(w0 + TRK_WIN, track_n);
        if (T.dense) {
            for (int64_t i = threadIdx.x; i < w1 - w0; i += TRK_THREADS) s_win[i] = T.dense[itv_lo + w0 + i];
        } else {
            for (int i = threadIdx.x; i < TRK_WIN / 32 + 4; i += TRK_THREADS) s_flag[i] = 0u;
        }
        const int32_t target = (int32_t)imin64(q_start + w0, INT32_MAX);  // intervals ending at or before it are behind the window
        const bool spec = pre_ok && target >= tgt_prev;  // (unsorted lists can move the window backwards: search again)
        if (!T.dense && !spec && threadIdx.x < 32) {
            // first interval whose end is > q_start + w0 (ends are sorted: intervals do not overlap)
            const int64_t first = warp_upper_le(T.itv_ends, itv_lo, itv_hi, target) + 1;
            if (threadIdx.x == 0) s_itv_first = first;
        }
        __syncthreads();
        TRK_TR(2);
#ifndef GVL_TRK_EXP
#define GVL_TRK_EXP 0
#endif
        if (!(GVL_TRK_EXP & 1) && !T.dense && w1 > w0) {
            // Paint the window as a run-length expansion (src/intervals.rs:19-126 restated for one window):
            //  1. every thread holds ONE interval (coalesced loads) and drops two markers: 0 at its end, its value at
            //     its (clipped) start -- ends first, so that an adjacent interval's start wins;
            //  2. every thread then owns 36 consecutive window positions, finds the value in effect at its first
            //     position with a block-wide scan over "last marker" pairs, and fills.
            const int nwin = (int)(w1 - w0);
            int64_t base = spec ? itv_prev : s_itv_first;
            int64_t first = base;      // cursor for the next pass: intervals before it end at or before `target`
            bool counting = true;
            for (bool use_pre = spec;; use_pre = false) {
                const int64_t it = base + threadIdx.x;
                int32_t st_a = INT32_MAX, en_a = 0;
                float v_ = 0.0f;
                if (use_pre) {
                    st_a = pre_st, en_a = pre_en, v_ = pre_v;
                } else if (it < itv_hi) {
                    st_a = T.itv_starts[it];
                    en_a = T.itv_ends[it];
                    v_ = T.itv_values[it];
                }
                const int64_t st_ = (int64_t)st_a - q_start, en_ = (int64_t)en_a - q_start;
                const bool have = it < itv_hi;
                const bool live = have && st_ < w1 && en_ > w0 && en_ > st_;  // overlaps the window (also :72-76: start >= length)
                if (live && en_ < w1) {
                    const int x = (int)(en_ - w0);
                    s_win[x] = 0.0f;
                    atomicOr(&s_flag[x >> 5], 1u << (x & 31));
                }
                if (threadIdx.x == TRK_THREADS - 1) s_stop = (!have || st_ >= w1);  // sorted starts: the block's last interval decides
                const int behind = __syncthreads_count(have && en_a <= target);  // (also orders the two marker phases)
                if (counting) {
                    first += behind;
                    counting = behind == TRK_THREADS;
                }
                if (live) {
                    const int x = (int)(imax64(st_, w0) - w0);
                    s_win[x] = v_;
                    atomicOr(&s_flag[x >> 5], 1u << (x & 31));
                }
                if (s_stop) break;  // the block's LAST interval starts at or beyond the window end
                __syncthreads();    // (rare second block: s_stop is rewritten)
                base += TRK_THREADS;
            }
            if (threadIdx.x == 0) s_itv_first = first;
            __syncthreads();
            TRK_TR(3);
            constexpr int CH = 36;  // 9 float4 per thread: conflict-free 128-bit accesses, 36 * 256 = TRK_WIN
            static_assert(CH * TRK_THREADS >= TRK_WIN && CH % 4 == 0, "fill chunks must cover the window");
            const int b0 = CH * (int)threadIdx.x;
            const bool act = b0 < nwin;
            uint64_t mk = 0;  // marker bits of positions b0 .. b0 + 35
            if (act) {
                const int w = b0 >> 5, sft = b0 & 31;
                mk = (((uint64_t)s_flag[w + 1] << 32) | s_flag[w]) >> sft;
                if (sft) mk |= (uint64_t)s_flag[w + 2] << (64 - sft);
                mk &= (1ull << CH) - 1;
            }
            float x[CH];
            if (act) {
#pragma unroll
                for (int q = 0; q < CH / 4; q++) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(&s_win[b0 + 4 * q]);
                    x[4 * q] = v4.x, x[4 * q + 1] = v4.y, x[4 * q + 2] = v4.z, x[4 * q + 3] = v4.w;
                }
            }
            // (has, value) of the last marker in the chunk; scan with "right operand wins if it has one"
            const int has = mk != 0;
            const float val = has ? s_win[b0 + 63 - __clzll((long long)mk)] : 0.0f;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            int h_in = has;
            float v_in = val;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int h2 = __shfl_up_sync(0xffffffffu, h_in, o);
                const float v2 = __shfl_up_sync(0xffffffffu, v_in, o);
                if (lane >= o && !h_in) {
                    h_in = h2;
                    v_in = v2;
                }
            }
            if (lane == 31) {
                s_chas[warp] = h_in;
                s_cval[warp] = v_in;
            }
            int h_ex = __shfl_up_sync(0xffffffffu, h_in, 1);
            float v_ex = __shfl_up_sync(0xffffffffu, v_in, 1);
            if (lane == 0) h_ex = 0;
            __syncthreads();
            float carry = 0.0f;  // no marker before: nothing covers the window start
            int found = h_ex;
            if (found) carry = v_ex;
            for (int w = warp - 1; w >= 0 && !found; w--) {
                if (s_chas[w]) {
                    carry = s_cval[w];
                    found = 1;
                }
            }
            if (act) {
                float cur_v = carry;
                const uint32_t mk_lo = (uint32_t)mk, mk_hi = (uint32_t)(mk >> 32);
#pragma unroll
                for (int q = 0; q < CH; q++) {
                    const bool f = q < 32 ? ((mk_lo >> q) & 1u) : ((mk_hi >> (q - 32)) & 1u);
                    cur_v = f ? x[q] : cur_v;
                    x[q] = cur_v;
                }
#pragma unroll
                for (int q = 0; q < CH / 4; q++)
                    *reinterpret_cast<float4 *>(&s_win[b0 + 4 * q]) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            }
        }
        __syncthreads();
        TRK_TR(4);
        itv_prev = T.dense ? -1 : s_itv_first;
        tgt_prev = (int32_t)imin64(q_start + w0, INT32_MAX);
        TrkSrc src{s_win, w0, w1, track_n, &T, itv_lo, itv_hi, q_start};

        const int32_t jo_lo = rc ? L - seg_end : cur;
        const int32_t jo_hi = rc ? L - cur : seg_end;
        const int64_t g0 = (row_base + jo_lo) & ~(int64_t)3;
        const int32_t n_chunks = (int32_t)((row_base + jo_hi - g0 + 3) >> 2);
        // a GROUP is 32 chunks of 4 values (one chunk per lane); warp w owns groups w, w+8, ...  A group that lies
        // inside ONE reference span and inside the staged window is a straight (possibly reversed) copy out of
        // shared memory: one thread per group decides that up front (group table), so the copy loop itself is a
        // table read, four shared-memory reads and one 16-byte store per lane.
        const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;
        const int32_t n_groups = (n_chunks + 31) >> 5;
        if ((int)threadIdx.x < n_groups) {
            const int32_t jg = (int32_t)(g0 + 128 * (int64_t)threadIdx.x - row_base);
            int32_t off = -1;
            if (jg >= jo_lo && jg + 128 <= jo_hi) {
                const int32_t p_lo = rc ? (L - 128 - jg) : jg;  // lowest haplotype position of the group
                int lo = 0, hi = m;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (S.a[mid] <= p_lo) lo = mid; else hi = mid;
                }
                const int32_t e_i = S.e[lo];
                const int64_t tp_lo = (int64_t)S.resume[lo] + (p_lo - e_i);
                if (p_lo >= e_i && p_lo + 128 <= S.a[lo + 1] && tp_lo >= w0 && tp_lo + 128 <= w1) off = (int32_t)(tp_lo - w0);
            }
            s_gt[threadIdx.x] = off;
        }
        __syncthreads();
        // copy loop of the plain groups: lane l moves values l, l+32, l+64, l+96 of a group, so every shared-memory
        // read and every store is one contiguous 128-byte warp access (4 consecutive values per lane would be a 4-way
        // bank conflict); pointers are hoisted, the two directions are separate loops (constant offsets)
        {
            float *const og_lane = out + g0 + lane_;
            if (!rc) {
                const float *const ws_lane = s_win + lane_;
#pragma unroll 2
                for (int32_t grp = warp_; grp < n_groups; grp += TRK_THREADS / 32) {
                    const int32_t off = s_gt[grp];
                    if (off < 0) continue;
                    const float *ws = ws_lane + off;
                    const float v0 = ws[0], v1 = ws[32], v2 = ws[64], v3 = ws[96];
                    float *og = og_lane + 128 * grp;
                    if ((GVL_TRK_EXP & 2) && v0 != 123.456f) continue;
                    og[0] = v0, og[32] = v1, og[64] = v2, og[96] = v3;
                }
            } else {
                const float *const ws_lane = s_win + 127 - lane_;
#pragma unroll 2
                for (int32_t grp = warp_; grp < n_groups; grp += TRK_THREADS / 32) {
                    const int32_t off = s_gt[grp];
                    if (off < 0) continue;
                    const float *ws = ws_lane + off;
                    const float v0 = ws[0], v1 = ws[-32], v2 = ws[-64], v3 = ws[-96];
                    float *og = og_lane + 128 * grp;
                    if ((GVL_TRK_EXP & 2) && v0 != 123.456f) continue;
                    og[0] = v0, og[32] = v1, og[64] = v2, og[96] = v3;
                }
            }
        }
        for (int32_t grp = warp_; grp < n_groups; grp += TRK_THREADS / 32) {
            if (s_gt[grp] >= 0) continue;  // copied above
            const int32_t c = grp * 32 + lane_;
            const int64_t g = g0 + 4 * (int64_t)c;
            const int32_t j = (int32_t)(g - row_base);
            if (c >= n_chunks) continue;
            // lane-level fast path: the lane's 4 values lie inside one reference span and inside the window
            if (j >= jo_lo && j + 4 <= jo_hi) {
                const int32_t p4 = rc ? (L - 4 - j) : j;  // lowest haplotype position of the chunk
                int lo = 0, hi = m;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (S.a[mid] <= p4) lo = mid; else hi = mid;
                }
                const int32_t e_l = S.e[lo];
                const int64_t tp4 = (int64_t)S.resume[lo] + (p4 - e_l);
                if (p4 >= e_l && p4 + 4 <= S.a[lo + 1] && tp4 >= w0 && tp4 + 4 <= w1) {
                    const float *ws = s_win + (tp4 - w0);
                    *reinterpret_cast<float4 *>(out + g) =
                        rc ? make_float4(ws[3], ws[2], ws[1], ws[0]) : make_float4(ws[0], ws[1], ws[2], ws[3]);
                    continue;
                }
            }
            float vals[4];
            bool valid[4];
            int i = 0;
            bool have_i = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int32_t jj = j + q;
                valid[q] = (jj >= jo_lo) && (jj < jo_hi);
                vals[q] = 0.0f;
                if (!valid[q]) continue;
                const int32_t p = rc ? (L - 1 - jj) : jj;
                if (!have_i || p < S.a[i] || p >= S.a[i + 1]) {
                    int lo = 0, hi = m;
                    while (hi - lo > 1) {
                        int mid = (lo + hi) >> 1;
                        if (S.a[mid] <= p) lo = mid; else hi = mid;
                    }
                    i = lo;
                    have_i = true;
                }
                if (p < S.e[i]) {
                    // values written by the variant itself (:329-354)
                    const int64_t vrel = S.vrel[i];
                    if (S.vdiff[i] > 0 && T.strategy != GVL_FILL_REPEAT_5P) {
                        vals[q] = insertion_fill_value(src, T.strategy, T.param, S.vlen[i], vrel, p - S.a[i], p,
                                                       base_seed, qseed, hap);
                    } else {
                        vals[q] = src.at(vrel);
                    }
                } else {
                    const int64_t tp = (int64_t)S.resume[i] + (p - S.e[i]);
                    vals[q] = (tp < track_n) ? src.at(tp) : 0.0f;  // :381-404 trailing zeros
                }
            }
            if (valid[0] && valid[1] && valid[2] && valid[3]) {
                *reinterpret_cast<float4 *>(out + g) = make_float4(vals[0], vals[1], vals[2], vals[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (valid[q]) out[g + q] = vals[q];
            }
        }
#if GVL_TRACE
        __syncthreads();
        TRK_TR(5);
        tr_pass++;
#endif
        // records consumed by this pass: staged entries 1..m_new with a < seg_end (the rest are staged again)
        {
            int lo = 0, hi = m;  // last staged entry with a < seg_end
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (S.a[mid] < seg_end) lo = mid; else hi = mid;
            }
            r += lo;
        }
        cur = seg_end;
    }
}

__global__ void prng_kernel(uint64_t a, uint64_t b, uint64_t c, uint64_t d, int which, uint64_t *out) {
    *out = which ? hash4(a, b, c, d) : xorshift64(a);
}

}  // namespace gvl

using namespace gvl;

static int launch_trk_exec(gvl_ctx *ctx, int64_t n_work, int64_t ploidy, int64_t n_queries, int64_t n_tracks,
                           const TrkDesc *host_desc, const int64_t *offset_idxs, int64_t total_per_track,
                           const int64_t *query_seed, uint64_t base_seed, float *out, int layout_btp, cudaStream_t st,
                           const uint64_t *base_seed_dev = nullptr, int64_t sub_batch = 0) {
    if (n_tracks > MAX_TRACKS) return fail(GVL_ERR_ARG, "at most %d tracks per call", MAX_TRACKS);
    static_assert(sizeof(TrkDesc) * MAX_TRACKS <= GVL_TRK_DESC_BYTES, "descriptor buffer too small");
    TrkDesc *d_desc = nullptr;
    if (n_tracks > 8) {  // (not CUDA-graph safe: the copy reads the caller's host array at replay time)
        d_desc = reinterpret_cast<TrkDesc *>(ctx->trk_desc);
        GVL_CUDA(cudaMemcpyAsync(d_desc, host_desc, sizeof(TrkDesc) * (size_t)n_tracks, cudaMemcpyHostToDevice, st));
    }
    trk_tile_scan_kernel<<<1, 1024, 0, st>>>(n_work, ctx->trk.row_len, ctx->trk.tile_off);
    GVL_LAUNCH_CHECK();
    TrkExecParams P;
    P.rows = ctx->trk.rows;
    P.rec = ctx->trk.rec;
    P.tile_off = ctx->trk.tile_off;
    P.n_work = n_work;
    P.ploidy = ploidy;
    P.grid_per_track = total_per_track / TRK_SEG + n_work;
    P.total_per_track = total_per_track;
    P.n_tracks = n_tracks;
    P.layout_btp = layout_btp;
    P.offset_idxs = offset_idxs;
    P.n_queries = n_queries;
    P.query_seed = query_seed;
    P.base_seed = base_seed;
    P.base_seed_dev = base_seed_dev;
    P.sub_batch = sub_batch;
    P.out = out;
    P.tracks = d_desc;
    if (!d_desc)
        for (int64_t t = 0; t < n_tracks; t++) P.inl[t] = host_desc[t];
    const int64_t grid = P.grid_per_track * n_tracks;
    if (grid > INT32_MAX) return fail(GVL_ERR_ARG, "too many track tiles");
    static const bool smem_ok = [] {
        return cudaFuncSetAttribute(trk_exec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * TRK_WIN)) == cudaSuccess;
    }();
    if (!smem_ok) return fail(GVL_ERR_CUDA, "trk_exec_kernel: cannot reserve %d bytes of shared memory", (int)(sizeof(float) * TRK_WIN));
    trk_exec_kernel<<<(unsigned)grid, TRK_THREADS, sizeof(float) * TRK_WIN, st>>>(P);
    GVL_LAUNCH_CHECK();
    return GVL_OK;
}

// svar2 two-channel merge (gvl_svar2.cu)
int gvl_svar2_merge_launch(gvl_ctx *ctx, gvl_workspace *ws, int64_t *words, const gvl_svar2_channels *ch, int64_t batch,
                           int64_t ploidy, int64_t max_merged, cudaStream_t st);

extern "C" {

static int realign_impl(gvl_ctx *ctx, const gvl_sparse_tables *tab, const gvl_svar2_channels *svar2, const int32_t *regions,
                        const int32_t *shifts, const int64_t *geno_offset_idx, int64_t batch, int64_t ploidy,
                        const uint8_t *keep, const int64_t *keep_offsets, const uint8_t *to_rc,
                        int64_t n_tracks, const gvl_intervals *itv, const int64_t *offset_idxs,
                        const float *dense, const int64_t *dense_offsets,
                        const int32_t *track_lengths, const int64_t *out_offsets, int64_t total_per_track,
                        const int32_t *strategy_ids, const double *params, uint64_t base_seed,
                        const int64_t *query_seed, int64_t max_records, float *out, gvl_stream stream,
                        int layout_btp = 0, const uint64_t *base_seed_dev = nullptr, int64_t sub_batch = 0) {
    if (!ctx || !tab || !regions || !shifts || !(geno_offset_idx || svar2) || !(itv || dense) || !(offset_idxs || dense) ||
        !track_lengths || !out_offsets || !strategy_ids || !params)
        return fail(GVL_ERR_ARG, "gvl_dev_realign_tracks: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GVL_CUDA(cudaSetDevice(ctx->device));
    const int64_t n_work = batch * ploidy;
    if (n_work == 0 || n_tracks == 0 || total_per_track == 0) return GVL_OK;
    if (!out || ((uintptr_t)out & 15)) return fail(GVL_ERR_ARG, "gvl_dev_realign_tracks: out must be 16-byte aligned");
    int rc;
    if ((rc = ensure_rows(ctx, ctx->trk, n_work))) return rc;
    if ((rc = ensure_records(ctx, ctx->trk, max_records + n_work))) return rc;
    int64_t *words = ctx->dev_words + W_COUNT;  // (left at zero by the previous plan, see plan_row_done)
    TrkPlanParams PP;
    PP.tab = *tab;
    PP.merged = MergedLists{nullptr, nullptr, nullptr, nullptr};
    if (svar2) {  // merged two-channel variant lists (src/svar2/mod.rs:45-66), same merge as the haplotype path
        if ((rc = ensure_merged(ctx, ctx->trk, max_records))) return rc;
        if ((rc = gvl_svar2_merge_launch(ctx, &ctx->trk, words, svar2, batch, ploidy, max_records, st))) return rc;
        PP.merged = MergedLists{ctx->trk.m_pos, ctx->trk.m_key, ctx->trk.m_off, ctx->trk.m_len};
    }
    PP.regions = regions;
    PP.shifts = shifts;
    PP.goi = geno_offset_idx;
    PP.keep = keep;
    PP.keep_off = keep_offsets;
    PP.to_rc = to_rc;
    PP.track_lengths = track_lengths;
    PP.out_offsets = out_offsets;
    PP.n_work = n_work;
    PP.ploidy = ploidy;
    PP.rec_cap = ctx->trk.rec_cap;
    PP.rows = ctx->trk.rows;
    PP.rec = ctx->trk.rec;
    PP.words = words;
    PP.row_len = ctx->trk.row_len;
    trk_plan_kernel<<<(unsigned)((n_work + TPLAN_WARPS - 1) / TPLAN_WARPS), TPLAN_WARPS * 32, 0, st>>>(PP);
    GVL_LAUNCH_CHECK();
    TrkDesc desc[MAX_TRACKS];
    if (n_tracks > MAX_TRACKS) return fail(GVL_ERR_ARG, "at most %d tracks per call", MAX_TRACKS);
    for (int64_t t = 0; t < n_tracks; t++) {
        desc[t].itv_starts = itv ? itv[t].itv_starts : nullptr;
        desc[t].itv_ends = itv ? itv[t].itv_ends : nullptr;
        desc[t].itv_values = itv ? itv[t].itv_values : nullptr;
        desc[t].itv_offsets = itv ? itv[t].itv_offsets : nullptr;
        desc[t].dense = dense;
        desc[t].dense_offsets = dense_offsets;
        desc[t].strategy = strategy_ids[t];
        desc[t].param = params[t];
        if (strategy_ids[t] < 0 || strategy_ids[t] > GVL_FILL_INTERPOLATE)
            return fail(GVL_ERR_ARG, "unknown insertion-fill strategy %d", strategy_ids[t]);
        if (strategy_ids[t] == GVL_FILL_INTERPOLATE && !(params[t] >= 1 && params[t] <= 3))
            return fail(GVL_ERR_ARG, "Interpolate order must be 1, 2 or 3");
    }
    return launch_trk_exec(ctx, n_work, ploidy, batch, n_tracks, desc, offset_idxs, total_per_track, query_seed,
                           base_seed, out, layout_btp, st, base_seed_dev, sub_batch);
}

int gvl_dev_realign_tracks(gvl_ctx *ctx, const gvl_sparse_tables *tab, const int32_t *regions,
                           const int32_t *shifts, const int64_t *geno_offset_idx, int64_t batch, int64_t ploidy,
                           const uint8_t *keep, const int64_t *keep_offsets, const uint8_t *to_rc,
                           int64_t n_tracks, const gvl_intervals *itv, const int64_t *offset_idxs,
                           const int32_t *track_lengths, const int64_t *out_offsets, int64_t total_per_track,
                           const int32_t *strategy_ids, const double *params, uint64_t base_seed,
                           const uint64_t *base_seed_dev, int64_t sub_batch, const int64_t *query_seed, int64_t max_records,
                           float *out, gvl_stream stream) {
    if (!itv || !offset_idxs) return fail(GVL_ERR_ARG, "gvl_dev_realign_tracks: NULL argument");
    return realign_impl(ctx, tab, nullptr, regions, shifts, geno_offset_idx, batch, ploidy, keep, keep_offsets, to_rc, n_tracks,
                        itv, offset_idxs, nullptr, nullptr, track_lengths, out_offsets, total_per_track, strategy_ids,
                        params, base_seed, query_seed, max_records, out, stream, 0, base_seed_dev, sub_batch);
}

int gvl_dev_realign_tracks_btp(gvl_ctx *ctx, const gvl_sparse_tables *tab, const int32_t *regions,
                               const int32_t *shifts, const int64_t *geno_offset_idx, int64_t batch, int64_t ploidy,
                               const uint8_t *keep, const int64_t *keep_offsets, const uint8_t *to_rc,
                               int64_t n_tracks, const gvl_intervals *itv, const int64_t *offset_idxs,
                               const int32_t *track_lengths, const int64_t *out_offsets, int64_t total_per_track,
                               const int32_t *strategy_ids, const double *params, uint64_t base_seed,
                               const uint64_t *base_seed_dev, int64_t sub_batch, const int64_t *query_seed,
                               int64_t max_records, float *out, gvl_stream stream) {
    if (!itv || !offset_idxs) return fail(GVL_ERR_ARG, "gvl_dev_realign_tracks_btp: NULL argument");
    return realign_impl(ctx, tab, nullptr, regions, shifts, geno_offset_idx, batch, ploidy, keep, keep_offsets, to_rc, n_tracks,
                        itv, offset_idxs, nullptr, nullptr, track_lengths, out_offsets, total_per_track, strategy_ids,
                        params, base_seed, query_seed, max_records, out, stream, 1, base_seed_dev, sub_batch);
}

int gvl_dev_shift_and_realign_tracks(gvl_ctx *ctx, const gvl_sparse_tables *tab, const int32_t *regions,
                                     const int32_t *shifts, const int64_t *geno_offset_idx, int64_t batch,
                                     int64_t ploidy, const uint8_t *keep, const int64_t *keep_offsets,
                                     const uint8_t *to_rc, const float *tracks, const int64_t *track_offsets,
                                     const int32_t *track_lengths, const int64_t *out_offsets, int64_t total,
                                     int32_t strategy_id, double param, uint64_t base_seed,
                                     const int64_t *query_seed, int64_t max_records, float *out, gvl_stream stream) {
    if (!tracks || !track_offsets) return fail(GVL_ERR_ARG, "gvl_dev_shift_and_realign_tracks: NULL argument");
    return realign_impl(ctx, tab, nullptr, regions, shifts, geno_offset_idx, batch, ploidy, keep, keep_offsets, to_rc, 1, nullptr,
                        nullptr, tracks, track_offsets, track_lengths, out_offsets, total, &strategy_id, &param,
                        base_seed, query_seed, max_records, out, stream);
}

#if GVL_TRACE
__attribute__((visibility("default"))) int gvl_debug_set_trk_trace(void *dev_buf) {
    unsigned long long *p = (unsigned long long *)dev_buf;
    GVL_CUDA(cudaMemcpyToSymbol(g_trk_trace, &p, sizeof(p)));
    return GVL_OK;
}
#endif

int gvl_dev_shift_and_realign_tracks_svar2(gvl_ctx *ctx, const gvl_sparse_tables *tab, const gvl_svar2_channels *ch,
                                           const int32_t *regions, const int32_t *shifts, int64_t batch, int64_t ploidy,
                                           const uint8_t *to_rc, const float *tracks, const int64_t *track_offsets,
                                           const int32_t *track_lengths, const int64_t *out_offsets, int64_t total,
                                           int32_t strategy_id, double param, uint64_t base_seed,
                                           const int64_t *query_seed, int64_t max_merged, float *out, gvl_stream stream) {
    if (!ch || !ch->vk_off || !ch->dense_range || !ch->dense_present_off || !tracks || !track_offsets)
        return fail(GVL_ERR_ARG, "gvl_dev_shift_and_realign_tracks_svar2: NULL argument");
    return realign_impl(ctx, tab, ch, regions, shifts, nullptr, batch, ploidy, nullptr, nullptr, to_rc, 1, nullptr, nullptr,
                        tracks, track_offsets, track_lengths, out_offsets, total, &strategy_id, &param, base_seed,
                        query_seed, max_merged, out, stream);
}

int gvl_dev_intervals_to_tracks(gvl_ctx *ctx, const gvl_intervals *itv, const int64_t *offset_idxs,
                                const int32_t *starts, int64_t n_queries, const int64_t *out_offsets,
                                int64_t total, float *out, gvl_stream stream) {
    if (!ctx || !itv || !offset_idxs || !starts || !out_offsets)
        return fail(GVL_ERR_ARG, "gvl_dev_intervals_to_tracks: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GVL_CUDA(cudaSetDevice(ctx->device));
    if (n_queries == 0 || total == 0) return GVL_OK;
    if (!out || ((uintptr_t)out & 15)) return fail(GVL_ERR_ARG, "gvl_dev_intervals_to_tracks: out must be 16-byte aligned");
    int rc;
    if ((rc = ensure_rows(ctx, ctx->trk, n_queries))) return rc;
    if ((rc = ensure_records(ctx, ctx->trk, 1))) return rc;
    paint_plan_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, st>>>(n_queries, starts, out_offsets, nullptr,
                                                                            ctx->trk.rows, ctx->trk.row_len);
    GVL_LAUNCH_CHECK();
    TrkDesc desc;
    desc.itv_starts = itv->itv_starts;
    desc.itv_ends = itv->itv_ends;
    desc.itv_values = itv->itv_values;
    desc.itv_offsets = itv->itv_offsets;
    desc.dense = nullptr;
    desc.dense_offsets = nullptr;
    desc.strategy = GVL_FILL_REPEAT_5P;
    desc.param = 0.0;
    return launch_trk_exec(ctx, n_queries, 1, n_queries, 1, &desc, offset_idxs, total, nullptr, 0, out, 0, st);
}

int gvl_dev_paint_tracks(gvl_ctx *ctx, int64_t n_tracks, const gvl_intervals *itv, const int64_t *offset_idxs,
                         const int32_t *starts, int64_t n_queries, const int64_t *out_offsets, int64_t total_per_track,
                         const uint8_t *to_rc, float *out, gvl_stream stream) {
    if (!ctx || !itv || !offset_idxs || !starts || !out_offsets)
        return fail(GVL_ERR_ARG, "gvl_dev_paint_tracks: NULL argument");
    if (n_tracks < 0 || n_tracks > MAX_TRACKS) return fail(GVL_ERR_ARG, "gvl_dev_paint_tracks: 0..%d tracks per call", MAX_TRACKS);
    cudaStream_t st = (cudaStream_t)stream;
    GVL_CUDA(cudaSetDevice(ctx->device));
    if (n_queries == 0 || n_tracks == 0 || total_per_track == 0) return GVL_OK;
    if (!out || ((uintptr_t)out & 15)) return fail(GVL_ERR_ARG, "gvl_dev_paint_tracks: out must be 16-byte aligned");
    int rc;
    if ((rc = ensure_rows(ctx, ctx->trk, n_queries))) return rc;
    if ((rc = ensure_records(ctx, ctx->trk, 1))) return rc;
    paint_plan_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, st>>>(n_queries, starts, out_offsets, to_rc,
                                                                            ctx->trk.rows, ctx->trk.row_len);
    GVL_LAUNCH_CHECK();
    TrkDesc desc[MAX_TRACKS];
    for (int64_t t = 0; t < n_tracks; t++) {
        desc[t].itv_starts = itv[t].itv_starts;
        desc[t].itv_ends = itv[t].itv_ends;
        desc[t].itv_values = itv[t].itv_values;
        desc[t].itv_offsets = itv[t].itv_offsets;
        desc[t].dense = nullptr;
        desc[t].dense_offsets = nullptr;
        desc[t].strategy = GVL_FILL_REPEAT_5P;
        desc[t].param = 0.0;
    }
    return launch_trk_exec(ctx, n_queries, 1, n_queries, n_tracks, desc, offset_idxs, total_per_track, nullptr, 0, out, 1, st);
}

static int run_prng(gvl_ctx *ctx, uint64_t a, uint64_t b, uint64_t c, uint64_t d, int which, uint64_t *out) {
    if (!ctx || !out) return fail(GVL_ERR_ARG, "gvl_debug prng: NULL argument");
    GVL_CUDA(cudaSetDevice(ctx->device));
    uint64_t *dv = (uint64_t *)(ctx->dev_words + 2 * W_COUNT - 1);
    prng_kernel<<<1, 1, 0, ctx->own_stream>>>(a, b, c, d, which, dv);
    GVL_LAUNCH_CHECK();
    GVL_CUDA(cudaMemcpyAsync(out, dv, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->own_stream));
    GVL_CUDA(cudaStreamSynchronize(ctx->own_stream));
    return GVL_OK;
}

int gvl_debug_hash4(gvl_ctx *ctx, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t *out) {
    return run_prng(ctx, a, b, c, d, 1, out);
}

int gvl_debug_xorshift64(gvl_ctx *ctx, uint64_t x, uint64_t *out) { return run_prng(ctx, x, 0, 0, 0, 0, out); }

}  // extern "C"
