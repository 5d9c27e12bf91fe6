// gvl_tracks_exec.cuh -- execute kernel of the track path (included by gvl_tracks.cu, inside `namespace gvl`).
//
// Reference path replaced: intervals_to_tracks (src/intervals.rs:19-126) into a dense scratch, then
// shift_and_realign_track_core (src/tracks/mod.rs:224-406) + apply_insertion_fill (:87-190) + the reversal of
// negative-strand rows (src/reverse.rs:25-38).
//
// One CTA owns a segment of up to T2_SEG output values of one (track, row) and walks it in haplotype order in passes
// of T2_PASS values.  A pass never materialises the source window: the stored intervals are painted DIRECTLY IN
// OUTPUT COORDINATES as a run-length code --
//
//   inputs   the pass's records (32-byte AoS written by the plan) and its slice of the interval SoA arrive in shared
//            memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier) that thread 0 issued during the PREVIOUS pass,
//            so no thread waits on a global load;
//   markers  one thread per stored interval maps its two boundaries from source to haplotype coordinates (a
//            translation per reference span; boundaries inside a deletion vanish) and drops a marker -- the value in
//            effect from there on -- into `val[]` plus a bit into a two-level bitmap; one thread per record drops the
//            value in effect right after the variant (found by a search of the staged intervals) and flags the
//            positions an insertion fill has to compute;
//   output   every lane owns chunks of 8 consecutive output values: it finds the last marker before its chunk with
//            two bitmap probes (no block-wide scan), walks its 8 marker bits in registers, and writes the chunk with
//            ONE 256-bit store (a warp instruction writes 1 KiB of contiguous output).  Flagged positions (insertion
//            fills other than Repeat5p) are computed in place by the owning lane.
//
// Two CTA barriers per pass; the values themselves never travel through shared memory.
// Rows that need the reference's sequential semantics literally (variant lists that are not position-sorted leave
// "jump" records) and dense f32 sources (shift_and_realign_tracks_sparse) take `t2_generic_segment`: every value
// resolved on its own against the row's records and the source -- slow, exact for any input.
#pragma once

constexpr int T2_PASS = 8192;               // output values per pass
constexpr int T2_THREADS = 256;
constexpr int T2_SEG = TRK_SEG;             // output values per CTA (8 passes)
constexpr int T2_ITV = 384;                 // stored intervals staged per pass
constexpr int T2_REC = 96;                  // records staged per pass (carry + new ones + sentinel)
constexpr int T2_WORDS = T2_PASS / 32;
constexpr int FLAG_JUMPS = 1;               // RowPlan.lead_pad of a track row: the row has jump records

struct __align__(16) T2Smem {
    float val[T2_PASS];                 // marker values by pass-relative haplotype position
    TRec rec[2][T2_REC];              // staged records (double-buffered: pass n+1 loads while pass n reads)
    int32_t its[2][T2_ITV + 8];         // staged interval starts / ends / values (+ alignment slack)
    int32_t ite[2][T2_ITV + 8];
    float itv[2][T2_ITV + 8];
    uint32_t mk[2][T2_WORDS + 1];       // marker bitmap (+ one zero word so that a funnel shift may read past the end)
    uint32_t fl[2][T2_WORDS + 1];       // "insertion fill computes this position" bitmap
    uint32_t mk2[2][T2_WORDS / 32];     // second level of mk: bit w = word w is non-zero
    uint64_t bar[2];                    // TMA completion barriers
    // what thread 0 staged for each buffer
    int64_t d_r[2], d_it0[2];           // record cursor (index of the carry record, -1 = virtual) / first staged interval
    int32_t d_m[2], d_more_rec[2];      // staged records; 1 = more records follow (the last staged one is only a sentinel)
    int32_t d_cnt[2], d_off[2], d_more_itv[2];  // staged intervals, index of the first one inside its[] (alignment), more follow
};

__device__ __forceinline__ bool t2_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void stg_f8(float *p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e),
                 "f"(f), "f"(g), "f"(h)
                 : "memory");
}

// ---- sources ---------------------------------------------------------------------------------------------
// value of the painted source track at relative position tp, straight from the global arrays: dense windows, or the
// last interval with start <= q_start + tp if it also ends after it (intervals of a slot are sorted by start and do
// not overlap: overlapping slots are flattened when the track is uploaded, gvl_flatten_intervals)
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

// source seen by the generic path: global arrays only
struct SrcGlobal {
    const TrkDesc *T;
    int64_t itv_lo, itv_hi, q_start, track_n;
    __device__ __forceinline__ float at(int64_t tp) const {
        if (tp < 0 || tp >= track_n) return 0.0f;  // out of contract in the reference (index panic)
        return track_at_global(*T, itv_lo, itv_hi, q_start, tp);
    }
};

// source seen by the fast path: the pass's staged intervals first (shared memory), the global arrays otherwise
struct SrcStaged {
    const int32_t *its, *ite;  // staged arrays, already offset to the first staged interval
    const float *itv;
    int cnt;
    bool more, from_first;     // intervals follow the staged ones / the staged ones start at the slot's first interval
    SrcGlobal G;
    int64_t track_n;
    __device__ __forceinline__ float at(int64_t tp) const {
        if (tp < 0 || tp >= track_n) return 0.0f;
        const int64_t g = G.q_start + tp;
        if (cnt > 0 && g >= (int64_t)its[0]) {
            int lo = 0, hi = cnt;  // last staged interval with start <= g
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)its[mid] <= g) lo = mid; else hi = mid;
            }
            if ((int64_t)ite[lo] > g) return itv[lo];
            if (lo < cnt - 1 || !more) return 0.0f;  // the next interval starts beyond g, or there is none
        } else if (from_first && (cnt > 0 || !more)) {
            return 0.0f;  // left of the slot's first interval (or the slot is empty)
        }
        return G.at(tp);
    }
};

// Lagrange interpolation through K anchors on each side of the insertion (src/tracks/mod.rs:138-188), evaluated at
// index i of the written values; same operation order as the reference (term = y_a * prod_b (x - x_b) / (x_a - x_b)).
template <int K, class Src>
__device__ __forceinline__ float lagrange_fill(const Src &S, int64_t v_len, int64_t v_rel_pos, int64_t i) {
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
template <class Src>
__device__ __noinline__ float insertion_fill_value(const Src &S, int strategy, double param, int64_t v_len, int64_t v_rel_pos,
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

// ---- generic path ------------------------------------------------------------------------------------------
// output value at haplotype position p of a row, resolved against the row's records in GLOBAL memory.  `hint` caches
// the record found last (positions of one thread advance monotonically).
__device__ float t2_generic_value(const TRec *__restrict__ recs, int32_t n_rec, int32_t ref0, const SrcGlobal &src,
                                  const TrkDesc &T, int32_t p, uint64_t base_seed, uint64_t qseed, uint64_t hap, int32_t &hint) {
    // i = last record with a <= p, -1 = before every record (the virtual record: span from ref0)
    int32_t i = hint;
    if (!(i >= -1 && i < n_rec && (i < 0 || recs[i].a <= p) && (i + 1 >= n_rec || recs[i + 1].a > p))) {
        int32_t lo = -1, hi = n_rec;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (recs[mid].a <= p) lo = mid; else hi = mid;
        }
        i = lo;
    }
    hint = i;
    if (i < 0) return src.at((int64_t)ref0 + p);
    const TRec r = recs[i];
    if (p < r.e) {  // values written by the variant itself (src/tracks/mod.rs:329-354)
        if (r.vdiff > 0 && T.strategy != GVL_FILL_REPEAT_5P)
            return insertion_fill_value(src, T.strategy, T.param, r.vlen, r.vrel, p - r.a, p, base_seed, qseed, hap);
        return src.at(r.vrel);
    }
    const int64_t tp = (int64_t)r.resume + (p - r.e);
    return tp < src.track_n ? src.at(tp) : 0.0f;  // :381-404 trailing zeros
}

__device__ void t2_generic_segment(const TRec *__restrict__ recs, int32_t n_rec, int32_t ref0, const SrcGlobal &src,
                                   const TrkDesc &T, float *__restrict__ out_row, int32_t L, bool rc, int32_t t0, int32_t t1,
                                   uint64_t base_seed, uint64_t qseed, uint64_t hap) {
    int32_t hint = -2;
    for (int32_t j = t0 + (int32_t)threadIdx.x; j < t1; j += T2_THREADS) {
        const int32_t p = rc ? (L - 1 - j) : j;
        out_row[j] = t2_generic_value(recs, n_rec, ref0, src, T, p, base_seed, qseed, hap, hint);
    }
}

// ---- fast path helpers -------------------------------------------------------------------------------------
// haplotype position at which relative source position x appears, given the staged records R[0..m) (R[0] = carry).
// dropped = x is not copied (inside a deletion, or behind the carry span); the returned position is then the first
// one whose source lies beyond x.
__device__ __forceinline__ int32_t t2_map(const TRec *R, int m, int64_t x, bool &dropped) {
    if (x < (int64_t)R[0].resume) {
        dropped = true;
        return R[0].e;
    }
    int lo = 0, hi = m;  // last staged record with resume <= x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)R[mid].resume <= x) lo = mid; else hi = mid;
    }
    if (lo + 1 < m && x > (int64_t)R[lo + 1].vrel) {  // deleted by the next record
        dropped = true;
        return R[lo + 1].e;
    }
    dropped = false;
    return R[lo].e + (int32_t)(x - (int64_t)R[lo].resume);
}

__device__ __forceinline__ void t2_set_bit(uint32_t *mk, uint32_t *mk2, int u) {
    const uint32_t old = atomicOr(&mk[u >> 5], 1u << (u & 31));
    if (old == 0) atomicOr(&mk2[u >> 10], 1u << ((u >> 5) & 31));
}

// set bits [lo, hi) of a bitmap (hi > lo)
__device__ __forceinline__ void t2_set_range(uint32_t *bm, int lo, int hi) {
    for (int w = lo >> 5; w <= (hi - 1) >> 5; w++) {
        const int b0 = max(lo - 32 * w, 0), b1 = min(hi - 32 * w, 32);
        const uint32_t m = (b1 >= 32 ? 0xffffffffu : ((1u << b1) - 1u)) & ~((1u << b0) - 1u);
        atomicOr(&bm[w], m);
    }
}

__global__ void __launch_bounds__(T2_THREADS, 4) trk_exec2_kernel(TrkExecParams P) {
    extern __shared__ __align__(16) unsigned char t2_raw[];
    T2Smem &S = *reinterpret_cast<T2Smem *>(t2_raw);
    const int tid = threadIdx.x;

    // ---- CTA -> (track, row, segment) ----
    const int64_t track = blockIdx.x / P.grid_per_track;
    const int64_t b = blockIdx.x % P.grid_per_track;
    if (b >= P.tile_off[P.n_work]) return;
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
    const int32_t t0 = (int32_t)(tile * T2_SEG);
    if (t0 >= L) return;
    const int32_t t1 = (int32_t)imin64((int64_t)t0 + T2_SEG, L);
    const bool rc = rp.rc != 0;
    const int32_t h0 = rc ? L - t1 : t0;  // the segment in haplotype coordinates
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
    int64_t row_base = track * P.total_per_track + rp.out_off;  // flat index of the row's first value
    if (P.layout_btp) {  // all tracks of a query are adjacent: block of the query, then track, then the row inside the block
        const int64_t k0 = query * P.ploidy;
        const int64_t blk0 = P.rows[k0].out_off;
        const int64_t blk_len = P.rows[k0 + P.ploidy - 1].out_off + P.rows[k0 + P.ploidy - 1].length - blk0;
        row_base = P.n_tracks * blk0 + track * blk_len + (rp.out_off - blk0);
    }
    const TRec *__restrict__ recs = P.trecs + rp.rec_off;
    const int32_t n_rec = rp.n_rec;
    const SrcGlobal srcg{&T, itv_lo, itv_hi, q_start, track_n};

    if (T.dense || (rp.lead_pad & FLAG_JUMPS)) {
        t2_generic_segment(recs, n_rec, rp.ref0, srcg, T, P.out + row_base, L, rc, t0, t1, base_seed, qseed, hap);
        return;
    }

    // ---- thread 0 stages a pass: carry record + following records, the interval slice from it0 on ----
    auto issue_loads = [&](int buf, int64_t r, int64_t it0) {
        int dst0 = 0;
        int64_t src0 = r;
        if (r < 0) {  // virtual carry: nothing written yet, the span starts at ref0
            TRec v;
            v.a = 0, v.e = 0, v.resume = rp.ref0, v.vrel = rp.ref0, v.vlen = 1, v.vdiff = 0, v.pad0 = 0, v.pad1 = 0;
            S.rec[buf][0] = v;
            dst0 = 1;
            src0 = 0;
        }
        const int n = (int)imin64((int64_t)(T2_REC - dst0), imax64((int64_t)n_rec - src0, 0));
        S.d_m[buf] = dst0 + n;
        S.d_more_rec[buf] = (src0 + n < (int64_t)n_rec) ? 1 : 0;
        const int64_t it_al = it0 & ~(int64_t)3;
        const int cnt = (int)imin64((int64_t)T2_ITV, imax64(itv_hi - it0, 0));
        const int n_al = ((int)(it0 - it_al) + cnt + 3) & ~3;
        S.d_cnt[buf] = cnt;
        S.d_off[buf] = (int)(it0 - it_al);
        S.d_more_itv[buf] = (it0 + cnt < itv_hi) ? 1 : 0;
        S.d_r[buf] = r;
        S.d_it0[buf] = it0;
        const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(TRec) + (cnt > 0 ? 3u * (uint32_t)n_al * 4u : 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy accesses of the buffer come first
        mbar_expect_tx(&S.bar[buf], bytes);
        if (n > 0) bulk_g2s(&S.rec[buf][dst0], recs + src0, (uint32_t)n * (uint32_t)sizeof(TRec), &S.bar[buf]);
        if (cnt > 0) {
            bulk_g2s(S.its[buf], T.itv_starts + it_al, (uint32_t)n_al * 4u, &S.bar[buf]);
            bulk_g2s(S.ite[buf], T.itv_ends + it_al, (uint32_t)n_al * 4u, &S.bar[buf]);
            bulk_g2s(S.itv[buf], T.itv_values + it_al, (uint32_t)n_al * 4u, &S.bar[buf]);
        }
    };
    // first interval of [itv_lo, itv_hi) whose end lies beyond relative source position x (ends are sorted)
    auto first_itv_after = [&](int64_t from, int64_t x) -> int64_t {
        int64_t a = from, bb = itv_hi;
        const int64_t g = q_start + x;
        while (a < bb) {
            const int64_t mid = (a + bb) >> 1;
            if ((int64_t)T.itv_ends[mid] <= g) a = mid + 1; else bb = mid;
        }
        return a;
    };

    // ---- prologue: barriers, bitmaps, cursors of the first pass ----
    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
    }
    for (int i = tid; i < T2_WORDS + 1; i += T2_THREADS) S.mk[0][i] = S.mk[1][i] = S.fl[0][i] = S.fl[1][i] = 0u;
    if (tid < T2_WORDS / 32) S.mk2[0][tid] = S.mk2[1][tid] = 0u;
    if (tid < 32) {
        // r = last record with a <= h0 (-1: none): a count over the sorted array (independent loads), a search when long
        int c = 0;
        if (n_rec <= 4096) {
            for (int i = tid; i < n_rec; i += 32) c += (recs[i].a <= h0);
            c = __reduce_add_sync(0xffffffffu, c);
        } else {
            int lo = -1, hi = n_rec;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (recs[mid].a <= h0) lo = mid; else hi = mid;
            }
            c = lo + 1;
        }
        const int64_t r = (int64_t)c - 1;
        int64_t src = (int64_t)rp.ref0 + h0;
        if (r >= 0) {
            const TRec cr = recs[r];
            src = h0 < cr.e ? (int64_t)cr.vrel : (int64_t)cr.resume + (h0 - cr.e);
        }
        // first interval whose end lies beyond the source position that feeds h0: warp-wide 32-ary search (ends are sorted)
        const int64_t first = warp_upper_le(T.itv_ends, itv_lo, itv_hi, (int32_t)imin64(q_start + src, INT32_MAX)) + 1;
        if (tid == 0) issue_loads(0, r, first);
    }

    float *__restrict__ out_row = P.out + row_base;
    const int64_t base_elems = (int64_t)(reinterpret_cast<uintptr_t>(P.out) >> 2) + row_base;  // address of out_row[0] in floats
    int32_t cur = h0;
    for (int pass = 0; cur < h1; pass++) {
        const int buf = pass & 1;
        __syncthreads();  // the previous pass's readers are done with val[]; thread 0's descriptors are visible
        while (!t2_try_wait(&S.bar[buf], (uint32_t)(pass >> 1) & 1u)) {
        }
        const TRec *R = S.rec[buf];
        const int m = S.d_m[buf], cnt = S.d_cnt[buf], off = S.d_off[buf];
        const bool more_itv = S.d_more_itv[buf] != 0;
        const int64_t it0 = S.d_it0[buf];
        uint32_t *mk = S.mk[buf], *fl = S.fl[buf], *mk2 = S.mk2[buf];
        const int32_t *its = S.its[buf] + off, *ite = S.ite[buf] + off;
        const float *itv = S.itv[buf] + off;

        // ---- extent of the pass ----
        int32_t pass_end = (int32_t)imin64((int64_t)cur + T2_PASS, h1);
        int m_eff = m;
        if (S.d_more_rec[buf]) {  // the last staged record is a sentinel: the pass stops where it starts
            m_eff = m - 1;
            pass_end = min(pass_end, R[m - 1].a);
        }
        const int32_t e0 = R[0].e;
        const int64_t src_lo = cur < e0 ? (int64_t)R[0].vrel : (int64_t)R[0].resume + (cur - e0);
        if (more_itv) {  // likewise the last staged interval: nothing beyond its start is known yet
            bool dr;
            const int32_t hl = t2_map(R, m_eff, imax64((int64_t)its[cnt - 1] - q_start, 0), dr);
            pass_end = min(pass_end, max(hl, cur + 1));
        }
        const int32_t plen = pass_end - cur;
        const SrcStaged src{its, ite, itv, cnt, more_itv, it0 == itv_lo, srcg, track_n};

        // ---- markers: intervals ----
        auto place = [&](int64_t x, float v) {
            bool dr;
            const int32_t h = t2_map(R, m_eff, x, dr);
            const int32_t u = h - cur;
            if (dr || u <= 0 || u >= plen) return;
            S.val[u] = v;
            t2_set_bit(mk, mk2, u);
        };
        for (int k = tid; k < cnt; k += T2_THREADS) {
            const int64_t s = imax64((int64_t)its[k] - q_start, 0);
            const int64_t e = imin64((int64_t)ite[k] - q_start, track_n);
            if (e <= s || e <= src_lo) continue;  // empty after clipping, or wholly behind the pass
            if (s > src_lo) place(s, itv[k]);     // (an interval that covers the pass start is the pass-start marker's)
            const bool adjacent = (k + 1 < cnt) && ((int64_t)its[k + 1] - q_start == e);  // the next interval starts right there
            if (!adjacent) place(e, 0.0f);
        }
        // ---- markers: records (threads from the top of the CTA) + the pass start ----
        {
            const int i = T2_THREADS - 1 - tid;
            if (i < m_eff) {
                const TRec r = R[i];
                const bool live = i == 0 ? (cur < r.e) : (r.a < pass_end);
                if (live) {
                    if (r.vdiff > 0 && T.strategy != GVL_FILL_REPEAT_5P)
                        t2_set_range(fl, max(r.a, cur) - cur, min(r.e, pass_end) - cur);
                    if (r.e < pass_end && r.e > cur) {  // value in effect right after the variant
                        S.val[r.e - cur] = src.at(r.resume);
                        t2_set_bit(mk, mk2, r.e - cur);
                    }
                }
            } else if (i == m_eff) {
                S.val[0] = src.at(cur < e0 ? (int64_t)R[0].vrel : src_lo);
                t2_set_bit(mk, mk2, 0);
            }
        }
        __syncthreads();

        // ---- thread 0: cursors + loads of the next pass (they land while this pass is written out) ----
        if (tid == 0 && pass_end < h1) {
            int lo = 0, hi = m;  // next carry: last staged record with a <= pass_end
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (R[mid].a <= pass_end) lo = mid; else hi = mid;
            }
            const TRec cr = R[lo];
            const int64_t src_next = pass_end < cr.e ? (int64_t)cr.vrel : (int64_t)cr.resume + (pass_end - cr.e);
            int a = 0, bb = cnt;  // staged intervals that end at or before src_next are history
            while (a < bb) {
                const int mid = (a + bb) >> 1;
                if ((int64_t)ite[mid] - q_start <= src_next) a = mid + 1; else bb = mid;
            }
            int64_t it_next = it0 + a;
            if (a == cnt && more_itv) it_next = first_itv_after(it_next, src_next);  // (a long deletion skipped them all)
            issue_loads(buf ^ 1, S.d_r[buf] + lo, it_next);
        }
        // the other buffer's bitmaps were last read in the previous pass: clear them for the next one
        S.mk[buf ^ 1][tid] = 0u;
        S.fl[buf ^ 1][tid] = 0u;
        if (tid < T2_WORDS / 32) S.mk2[buf ^ 1][tid] = 0u;

        // ---- output: chunks of 8 values on 32-byte aligned addresses ----
        const int32_t jo_lo = rc ? L - pass_end : cur;
        const int32_t jo_hi = rc ? L - cur : pass_end;
        const int32_t j0 = (int32_t)(((base_elems + jo_lo) & ~(int64_t)7) - base_elems);  // may be < jo_lo
        const int32_t n_chunks = (jo_hi - j0 + 7) >> 3;
        for (int32_t c = tid; c < n_chunks; c += T2_THREADS) {
            const int32_t j = j0 + 8 * c;
            const int32_t u_lo = (rc ? (L - 8 - j) : j) - cur;  // lowest pass-relative haplotype position of the chunk
            uint32_t bits, fbits;
            if (u_lo >= 0) {
                const int w = u_lo >> 5, sh = u_lo & 31;
                bits = __funnelshift_r(mk[w], mk[w + 1], sh) & 0xffu;
                fbits = __funnelshift_r(fl[w], fl[w + 1], sh) & 0xffu;
            } else {
                bits = (mk[0] << (-u_lo)) & 0xffu;
                fbits = (fl[0] << (-u_lo)) & 0xffu;
            }
            float cv = 0.0f;
            const int q = max(u_lo, 0) - 1;  // last position before the chunk (position 0 always holds a marker)
            if (q >= 0) {
                int w = q >> 5;
                uint32_t mm = mk[w] & (0xffffffffu >> (31 - (q & 31)));
                if (mm == 0u) {
                    int w2 = w >> 5;
                    uint32_t m2 = mk2[w2] & ((1u << (w & 31)) - 1u);
                    while (m2 == 0u) m2 = mk2[--w2];
                    w = 32 * w2 + 31 - __clz(m2);
                    mm = mk[w];
                }
                cv = S.val[32 * w + 31 - __clz(mm)];
            }
            float x[8];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                if ((bits >> t) & 1u) cv = S.val[u_lo + t];
                x[t] = cv;
            }
            if (fbits) {  // insertion fills other than Repeat5p: computed by the owning lane
#pragma unroll
                for (int t = 0; t < 8; t++) {  // (unrolled: x[] stays in registers)
                    if (!((fbits >> t) & 1u)) continue;
                    const int32_t p = cur + u_lo + t;
                    int lo = 0, hi = m_eff;  // the staged record whose values cover p
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (R[mid].a <= p) lo = mid; else hi = mid;
                    }
                    const TRec r = R[lo];
                    x[t] = insertion_fill_value(src, T.strategy, T.param, r.vlen, r.vrel, p - r.a, p, base_seed, qseed, hap);
                }
            }
            float *dst = out_row + j;
            if (j >= jo_lo && j + 8 <= jo_hi) {
                if (rc) stg_f8(dst, x[7], x[6], x[5], x[4], x[3], x[2], x[1], x[0]);
                else stg_f8(dst, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            } else {  // chunk cut by the pass / row ends
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const int32_t jj = j + t;
                    if (jj >= jo_lo && jj < jo_hi) dst[t] = rc ? x[7 - t] : x[t];
                }
            }
        }
        cur = pass_end;
    }
}
