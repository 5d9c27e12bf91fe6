// gvl_tracks_direct.cuh -- track execute kernel that expands the intervals directly in OUTPUT coordinates.
// Included by gvl_tracks.cu (needs TrkExecParams, TrkDesc, TrkSrc, insertion_fill_value).
//
// The windowed kernel (trk_exec_kernel) paints every pass's source window into shared memory and copies it out: about one
// warp-instruction per value and ~9 CTA barriers per 8,192 values (profiles/r1_packed.md).  Here every WARP owns a
// contiguous stretch of the CTA's segment and walks it in steps of 128 values (one float4 per lane) with two register
// resident batches that it reads with shuffles:
//   * 32 records  (lane k: virtual record vb + k; virtual record 0 is the span before the first variant)
//   * 32 intervals (lane k: interval cb + k of the row's slice)
// A step that lies inside one reference span (the common case: ~1 variant per kb) needs no per-lane record search; each
// lane finds the interval in effect at its first source position with a 5-round shuffle binary search over the batch's
// starts and walks its 4 values from there (the warp fetches a lane's next interval whenever one starts).  No shared memory, no CTA barriers.
// Lanes whose 4 values cross a record boundary, hold variant-written values, straddle the row's ends, or whose step the
// batches do not span take the per-value path (the same arithmetic as the windowed kernel's
// generic path, reading the interval SoA through TrkSrc::at near the cursor).
#pragma once

constexpr int TRKD_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

// number of batch entries <= key, for a batch sorted over the lanes whose lane 31 is known to be > key: 0..31
__device__ __forceinline__ int batch_count_le(int32_t mine, int32_t key) {
    int lo = 0;
#pragma unroll
    for (int half = 16; half; half >>= 1) {
        const int32_t s = __shfl_sync(FULL, mine, lo + half - 1);
        if (s <= key) lo += half;
    }
    return lo;
}

// the rare variant-written values: kept out of line so that the Lagrange fill's doubles do not set the kernel's register count
__device__ __noinline__ float trkd_fill_value(const TrkSrc &S, int strategy, double param, int64_t v_len, int64_t v_rel_pos,
                                              int64_t i, int64_t out_pos, uint64_t base_seed, uint64_t query, uint64_t hap) {
    return insertion_fill_value(S, strategy, param, v_len, v_rel_pos, i, out_pos, base_seed, query, hap);
}

__global__ void __launch_bounds__(TRKD_THREADS, 4) trk_exec_direct_kernel(TrkExecParams P) {
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
    const int32_t t0 = (int32_t)(tile * TRK_SEG);
    if (t0 >= L) return;
    const int32_t t1 = (int32_t)imin64((int64_t)t0 + TRK_SEG, L);
    const bool rc = rp.rc != 0;
    const int32_t h0 = rc ? L - t1 : t0;
    const int32_t h1 = rc ? L - t0 : t1;
    const int64_t query = row / P.ploidy;
    const uint64_t hap = (uint64_t)(row % P.ploidy);
    const uint64_t qseed = P.query_seed ? (uint64_t)P.query_seed[query] : (uint64_t)query;
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
    const int32_t track_n = rp.contig_len;
    const int32_t q_start = rp.q_start;
    float *__restrict__ out = P.out;
    const int64_t row_base = track * P.total_per_track + rp.out_off;  // flat index of the row's first value
    const int32_t *__restrict__ ra = P.rec.a + rp.rec_off;
    const int32_t n_rec = rp.n_rec;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t jo_lo = rc ? L - h1 : h0, jo_hi = rc ? L - h0 : h1;  // output positions of the CTA (== [t0, t1))
    const int64_t G0 = (row_base + jo_lo) & ~(int64_t)31;              // steps are 128-byte... 512-byte aligned stores
    const int32_t n_steps = (int32_t)((row_base + jo_hi - G0 + 127) >> 7);
    const int32_t spw = (n_steps + TRKD_THREADS / 32 - 1) / (TRKD_THREADS / 32);
    const int32_t s_lo = warp * spw, s_hi = min(n_steps, s_lo + spw);
    if (s_lo >= s_hi) return;

    // ---- record batch: lane k holds virtual record vb + k ----
    int64_t vb;
    int32_t ba, be, bres;
    auto load_recs = [&]() {
        const int64_t v = vb + lane;
        if (v == 0) {
            ba = 0, be = 0, bres = rp.ref0;
        } else if (v <= n_rec) {
            const int64_t g = rp.rec_off + v - 1;
            ba = P.rec.a[g];
            be = ba + P.rec.n[g];
            bres = P.rec.resume[g];
        } else {
            ba = INT32_MAX, be = INT32_MAX, bres = 0;
        }
    };
    {
        // lowest haplotype position of the warp's stretch
        const int32_t jw_lo = max(jo_lo, (int32_t)(G0 + 128 * (int64_t)s_lo - row_base));
        const int32_t jw_hi = min(jo_hi, (int32_t)(G0 + 128 * (int64_t)s_hi - row_base));
        const int32_t p_first = rc ? L - jw_hi : jw_lo;
        vb = warp_upper_le(ra, 0, n_rec, p_first) + 1;  // last real record with a <= p_first (or -1), as a virtual index
        load_recs();
    }
    // ---- interval batch: lane k holds interval cb + k ----
    int64_t cb = -1;
    int32_t x_cursor = INT32_MIN;
    int32_t ist = INT32_MAX, ien = INT32_MAX;
    float iv = 0.0f;
    auto load_itvs = [&]() {
        const int64_t i = cb + lane;
        if (i < itv_hi) {
            ist = T.itv_starts[i];
            ien = T.itv_ends[i];
            iv = T.itv_values[i];
        } else {
            ist = INT32_MAX, ien = INT32_MAX, iv = 0.0f;
        }
    };

    TrkSrc src{nullptr, 0, 0, (int64_t)track_n, &T, itv_lo, itv_hi, (int64_t)q_start, 0, 0};

    for (int32_t it = 0; it < s_hi - s_lo; it++) {
        const int32_t s = rc ? (s_hi - 1 - it) : (s_lo + it);  // haplotype positions ascend either way
        const int64_t gs = G0 + 128 * (int64_t)s;
        const int32_t js = (int32_t)(gs - row_base);
        const bool full = js >= jo_lo && js + 128 <= jo_hi;
        const int32_t jl = max(js, jo_lo), jh = min(js + 128, jo_hi);
        const int32_t p_lo = rc ? L - jh : jl, p_hi = rc ? L - jl : jh;  // haplotype positions [p_lo, p_hi) of the step

        // record cursor: the last record with a <= p_lo moves to the front part of the batch
        int cnt = __popc(__ballot_sync(FULL, ba <= p_lo));
        while (cnt > 16) {
            vb += cnt - 1;
            load_recs();
            cnt = __popc(__ballot_sync(FULL, ba <= p_lo));
        }
        const int vi = cnt - 1;
        const int32_t e_i = __shfl_sync(FULL, be, vi), res_i = __shfl_sync(FULL, bres, vi);
        const int32_t a_n = __shfl_sync(FULL, ba, vi + 1);
        const bool rec_cover = __shfl_sync(FULL, ba, 31) >= p_hi;  // the batch spans the step
        const int32_t tp_lo = res_i + (p_lo - e_i);
        const bool uniform = full && p_lo >= e_i && p_hi <= a_n && tp_lo + 128 <= track_n;

        // the lane's chunk: output positions j0 .. j0+3 = haplotype positions p4 .. p4+3 (reversed when rc)
        const int32_t j0 = js + 4 * lane;
        const int32_t p4 = rc ? L - 4 - j0 : j0;
        const bool chunk_ok = j0 >= jo_lo && j0 + 4 <= jo_hi;
        bool plain = uniform;
        int32_t tp4 = res_i + (p4 - e_i);
        if (!uniform) {
            plain = false;
            if (rec_cover) {  // per-lane record (warp-uniform branch: shuffles inside)
                const int32_t pq = chunk_ok ? p4 : p_lo;
                const int li = batch_count_le(ba, pq) - 1;
                const int32_t e_l = __shfl_sync(FULL, be, li), res_l = __shfl_sync(FULL, bres, li);
                const int32_t a_nl = __shfl_sync(FULL, ba, li + 1);
                tp4 = res_l + (pq - e_l);
                plain = chunk_ok && pq >= e_l && pq + 4 <= a_nl && tp4 + 4 <= track_n;
            }
        }
        float val[4];
        bool done = false;
        if (T.dense) {
            if (plain) {
                const float *d = T.dense + itv_lo + tp4;
                val[0] = d[0], val[1] = d[1], val[2] = d[2], val[3] = d[3];
                done = true;
            }
        } else if (__any_sync(FULL, plain)) {
            const int32_t x = q_start + tp4;  // source coordinate of the chunk's lowest position
            const int32_t Xlo = uniform ? q_start + tp_lo : __reduce_min_sync(FULL, plain ? x : INT32_MAX);
            const int32_t Xhi = uniform ? Xlo + 128 : __reduce_max_sync(FULL, plain ? x + 4 : INT32_MIN);
            if (cb < 0 || Xlo < x_cursor) {  // first use (or a backward move: unsorted input): search the slice
                cb = warp_upper_le(T.itv_ends, itv_lo, itv_hi, Xlo) + 1;  // first interval that ends after Xlo
                load_itvs();
            }
            x_cursor = Xlo;
            bool cover;
            for (;;) {
                cover = __shfl_sync(FULL, ist, 31) >= Xhi;  // the batch spans the step's source range
                if (cover) break;
                const int behind = __popc(__ballot_sync(FULL, ien <= Xlo));
                if (behind == 0) break;  // 32 live intervals inside one step: per-value path
                cb += behind;
                load_itvs();
            }
            if (cover) {
                // interval in effect at the chunk's first position (shuffle search over the batch's starts), then walk
                // the 4 positions; whenever ANY lane reaches the start of its next interval the warp fetches one more
                // (warp-uniform loop: a few rounds per step even when intervals are only a few positions long)
                const int32_t xk = plain ? x : Xlo;
                int nxt = batch_count_le(ist, xk);  // intervals of the batch that start at or before xk: 0..31
                const int k0 = max(nxt - 1, 0);
                int32_t cur_en = __shfl_sync(FULL, ien, k0);
                float cur_v = __shfl_sync(FULL, iv, k0);
                if (nxt == 0) cur_en = INT32_MIN;  // nothing of the batch starts before: earlier intervals ended already
                int32_t nst = __shfl_sync(FULL, ist, nxt);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int32_t pos = xk + q;
                    while (__any_sync(FULL, plain && pos >= nst)) {  // (lane 31's start is beyond the step: nxt stays < 31)
                        const bool adv = plain && pos >= nst;
                        const int32_t en_t = __shfl_sync(FULL, ien, nxt);
                        const float v_t = __shfl_sync(FULL, iv, nxt);
                        const int32_t st_t = __shfl_sync(FULL, ist, min(nxt + 1, 31));
                        if (adv) {
                            cur_en = en_t, cur_v = v_t, nst = st_t;
                            nxt++;
                        }
                    }
                    val[q] = pos < cur_en ? cur_v : 0.0f;
                }
                done = plain;
            }
        }
        const int64_t g = gs + 4 * lane;
        if (done) {
            *reinterpret_cast<float4 *>(out + g) =
                rc ? make_float4(val[3], val[2], val[1], val[0]) : make_float4(val[0], val[1], val[2], val[3]);
            continue;
        }
        if (j0 + 4 <= jo_lo || j0 >= jo_hi) continue;  // chunk entirely outside the CTA's range
        // ---- per-value path (src/tracks/mod.rs:224-406 value by value) ----
        if (!T.dense) {  // the interval lookups start near the cursor
            src.hint_lo = cb >= 0 ? imax64(itv_lo, cb - 8) : 0;
            src.hint_hi = cb >= 0 ? imin64(itv_hi, cb + 40) : 0;
        }
        // virtual records to search: the batch's span when it covers the step, else the whole row
        const int64_t v_lo = vb, v_hi = rec_cover ? imin64(vb + 32, (int64_t)n_rec + 1) : (int64_t)n_rec + 1;
        bool valid[4];
        int64_t vcur = -1;
        int32_t c_a = 0, c_an = 0, c_e = 0, c_res = 0, c_vlen = 0, c_vrel = 0, c_vdiff = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int32_t jj = j0 + q;
            valid[q] = jj >= jo_lo && jj < jo_hi;
            val[q] = 0.0f;
            if (!valid[q]) continue;
            const int32_t p = rc ? (L - 1 - jj) : jj;
            if (vcur < 0 || p < c_a || p >= c_an) {
                int64_t lo = v_lo, hi = v_hi;  // last virtual record with a <= p (a of virtual v is ra[v-1]; v = 0: 0)
                while (hi - lo > 1) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (ra[mid - 1] <= p) lo = mid; else hi = mid;
                }
                vcur = lo;
                if (lo == 0) {
                    c_a = 0, c_e = 0, c_res = rp.ref0, c_vlen = 1, c_vrel = 0, c_vdiff = 0;
                } else {
                    const int64_t gi = rp.rec_off + lo - 1;
                    c_a = P.rec.a[gi];
                    c_e = c_a + P.rec.n[gi];
                    c_res = P.rec.resume[gi];
                    c_vlen = P.rec.vidx[gi];
                    c_vrel = P.rec.vpos[gi];
                    c_vdiff = (int32_t)P.rec.src[gi];
                }
                c_an = lo < n_rec ? ra[lo] : INT32_MAX;
            }
            if (p < c_e) {
                // values written by the variant itself (:329-354)
                if (c_vdiff > 0 && T.strategy != GVL_FILL_REPEAT_5P) {
                    val[q] = trkd_fill_value(src, T.strategy, T.param, c_vlen, c_vrel, p - c_a, p, P.base_seed, qseed, hap);
                } else {
                    val[q] = src.at(c_vrel);
                }
            } else {
                const int64_t tp = (int64_t)c_res + (p - c_e);
                val[q] = (tp < track_n) ? src.at(tp) : 0.0f;  // :381-404 trailing zeros
            }
        }
        if (valid[0] && valid[1] && valid[2] && valid[3]) {
            *reinterpret_cast<float4 *>(out + g) = make_float4(val[0], val[1], val[2], val[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (valid[q]) out[g + q] = val[q];
        }
    }
}
