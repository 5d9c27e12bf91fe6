// gvl_trk_plan.cuh -- lock-step (one variant per step) plan of ONE track row by one warp: the reference's loop
// (shift_and_realign_track_core, src/tracks/mod.rs:224-406) driven through gvl_plan.cuh's trk_step.  Exact for any input
// order; the scan-based kernel (gvl_plan_par.cuh, TRK = true) falls back to it for rows whose variant list is not
// position-sorted (they leave "jump" records and take the execute kernel's generic path).
// (included by gvl_hap.cu inside `namespace gvl`, after HapPlanParams)
#pragma once

__device__ __forceinline__ void put_trec(TRec *t, int64_t a, int64_t n, int64_t resume, int64_t vrel, int64_t vlen, int64_t vdiff) {
    TRec r;
    r.a = (int32_t)a, r.e = (int32_t)(a + n), r.resume = (int32_t)resume, r.vrel = (int32_t)vrel, r.vlen = (int32_t)vlen;
    r.vdiff = (int32_t)vdiff, r.pad0 = 0, r.pad1 = 0;
    *t = r;
}

__device__ void trk_plan_row_serial(const HapPlanParams &P, const int64_t k, const int64_t rec_off_in) {
    const int lane = lane_id();
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

    const int64_t rec_off = rec_off_in;  // (taken by the caller)
    const bool overflow = rec_off + nvar + 1 > P.rec_cap;
    if (overflow && lane == 0) atomicMax((unsigned long long *)&P.words[W_STATUS], (unsigned long long)(rec_off + nvar + 1));

    TrkState ts;
    trk_init(ts, shift, length);
    int64_t n_emit = 0, track0 = 0, prev_resume = 0;
    bool done = false, jumps = false;  // jumps: an unsorted list moved the source cursor between emissions
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
                    put_trec(P.trecs + w, a, n, v_end, rel, v_len, il);
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
                    if (lane == t && !overflow) put_trec(P.trecs + rec_off + n_emit, r.a - (r.v_rel_pos - r.span_src), 0, r.span_src, 0, 1, 0);
                    n_emit++;
                    jumps = true;
                }
                if (lane == t && !overflow) put_trec(P.trecs + rec_off + n_emit, r.a, r.n, r.resume, r.v_rel_pos, r.v_len, r.v_diff);
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
            if (lane == 0 && !overflow) put_trec(P.trecs + rec_off + n_emit, imin64(ts.out_idx, length), 0, ts.track_idx, 0, 1, 0);
            n_emit++;
            jumps = true;
        }
    }
    if (lane == 0) {
        RowPlan rp;
        rp.out_off = P.out_offsets[k];
        rp.ref_base = 0;
        rp.rec_off = rec_off;
        rp.length = (int32_t)length;
        rp.contig_len = (int32_t)track_n;
        rp.lead_pad = jumps ? FLAG_JUMPS : 0;  // (track rows carry flags here: see gvl_tracks_exec.cuh)
        rp.ref0 = (int32_t)track0;
        rp.n_rec = overflow ? 0 : (int32_t)n_emit;
        rp.rc = (P.to_rc && P.to_rc[k]) ? 1 : 0;
        rp.diff = 0;
        rp.q_start = (int32_t)q_start;
        P.rows[k] = rp;
        P.row_len[k] = (int32_t)length;
    }
}

