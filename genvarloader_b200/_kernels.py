"""Host-buffer kernels with the names and argument order of the reference's PyO3 module
``genvarloader.genvarloader`` (src/ffi/mod.rs) -- numpy arrays in, numpy arrays out, the work
done by the CUDA library through the gvl_* host layer of include/gvl_b200.h.

A maintainer swaps ``from ..genvarloader import reconstruct_haplotypes_fused`` for
``from genvarloader_b200._kernels import reconstruct_haplotypes_fused``.
`parallel` is accepted for signature compatibility and ignored (the GPU is always parallel).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import MODE_ANNOTATED, MODE_ONEHOT, MODE_ONEHOT_CF, MODE_U8, c_i64, c_u8, c_u64, c_vp, check, lib

_MODES = {"u8": MODE_U8, "bytes": MODE_U8, "onehot": MODE_ONEHOT, "onehot_cf": MODE_ONEHOT_CF,
          "annotated": MODE_ANNOTATED}

_ctx: dict[int, _ffi.Ctx] = {}


def default_ctx(device: int = 0) -> _ffi.Ctx:
    if device not in _ctx:
        _ctx[device] = _ffi.Ctx(device)
    return _ctx[device]


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def _p(a):
    return c_vp(0) if a is None else c_vp(a.ctypes.data)


def _starts_stops(geno_offsets) -> np.ndarray:
    """(2, n) starts/stops; 1-D offsets normalised like _dataset/_genotypes.py:13-22."""
    go = np.asarray(geno_offsets)
    if go.ndim == 1:
        go = np.stack([go[:-1], go[1:]])
    return np.ascontiguousarray(go, np.int64)


def pin_static(*arrays, ctx: _ffi.Ctx | None = None) -> None:
    """Upload sample-scale arrays once; later calls that pass the SAME array objects reuse the
    device copies (the zero-copy memmap crossing of _dataset/_utils.py:13-35, GPU edition)."""
    ctx = ctx or default_ctx()
    for a in arrays:
        if a is None:
            continue
        assert a.flags.c_contiguous
        check(lib.gvl_pin_static(ctx.handle, _p(a), c_i64(a.nbytes)))


def unpin_static(*arrays, ctx: _ffi.Ctx | None = None) -> None:
    ctx = ctx or default_ctx()
    for a in arrays:
        if a is not None:
            check(lib.gvl_unpin_static(ctx.handle, _p(a)))


class PinnedBuffer:
    """Page-locked host array (gvl_host_alloc) -- outputs copied into it travel at full PCIe rate."""

    def __init__(self, nbytes: int):
        self._p = c_vp(0)
        check(lib.gvl_host_alloc(c_i64(int(nbytes)), C.byref(self._p)))
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(max(self.nbytes, 1),))[: self.nbytes]

    def close(self):
        if self._p:
            lib.gvl_host_free(self._p)
            self._p = c_vp(0)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _begin(ctx, regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs, v_starts, ilens, alt_alleles,
           alt_offsets, ref_, ref_offsets, output_length, keep, keep_offsets, to_rc):
    rg, sh = _c(regions, np.int32), _c(shifts, np.int32)
    goi = _c(geno_offset_idx, np.int64)
    batch, ploidy = goi.shape
    go = geno_offsets if (isinstance(geno_offsets, np.ndarray) and geno_offsets.ndim == 2 and
                          geno_offsets.dtype == np.int64 and geno_offsets.flags.c_contiguous) else _starts_stops(geno_offsets)
    gv, vs, il = _c(geno_v_idxs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    aa, ao = _c(alt_alleles, np.uint8), _c(alt_offsets, np.int64)
    rf, ro = _c(ref_, np.uint8), _c(ref_offsets, np.int64)
    kp, ko, rc = _c(keep, np.bool_), _c(keep_offsets, np.int64), _c(to_rc, np.bool_)
    out_offsets = np.empty(batch * ploidy + 1, np.int64)
    total = c_i64(0)
    keepalive = (rg, sh, goi, go, gv, vs, il, aa, ao, rf, ro, kp, ko, rc)
    check(lib.gvl_reconstruct_haplotypes_fused_begin(
        ctx.handle, _p(rg), _p(sh), _p(goi), c_i64(batch), c_i64(ploidy), _p(go), c_i64(go.shape[1]), _p(gv),
        c_i64(gv.size), _p(vs), _p(il), c_i64(vs.size), _p(aa), _p(ao), _p(rf), _p(ro), c_i64(ro.size - 1),
        c_i64(int(output_length)), _p(kp), _p(ko), _p(rc), _p(out_offsets), C.byref(total)))
    del keepalive
    return out_offsets, int(total.value)


def reconstruct_haplotypes_fused(regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs, v_starts, ilens,
                                 alt_alleles, alt_offsets, ref_, ref_offsets, pad_char, output_length, keep=None,
                                 keep_offsets=None, to_rc=None, parallel=True, *, mode="u8", out=None, ctx=None):
    """src/ffi/mod.rs:724-860.  Returns ``(out_data, out_offsets)``.

    ``mode="u8"`` is the reference's return value; ``"onehot"`` / ``"onehot_cf"`` return the fused
    one-hot encoding instead ((total, 4) uint8, resp. (n_rows, 4, L) for a fixed length).
    ``out`` may be a preallocated (e.g. pinned) uint8 buffer of the right size."""
    ctx = ctx or default_ctx()
    m = _MODES[mode]
    if m == MODE_ANNOTATED:
        raise ValueError("use reconstruct_annotated_haplotypes_fused")
    out_offsets, total = _begin(ctx, regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs, v_starts, ilens,
                                alt_alleles, alt_offsets, ref_, ref_offsets, output_length, keep, keep_offsets, to_rc)
    nbytes = total * (4 if m in (MODE_ONEHOT, MODE_ONEHOT_CF) else 1)
    if out is None:
        out = np.empty(nbytes, np.uint8)
    assert out.dtype == np.uint8 and out.size >= nbytes and out.flags.c_contiguous
    check(lib.gvl_reconstruct_haplotypes_fused_finish(ctx.handle, C.c_int(m), c_u8(int(pad_char)), _p(out), c_vp(0), c_vp(0)))
    data = out[:nbytes]
    if m == MODE_ONEHOT:
        data = data.reshape(total, 4)
    elif m == MODE_ONEHOT_CF:
        n_rows = out_offsets.size - 1
        data = data.reshape(n_rows, 4, int(output_length)) if n_rows else data.reshape(0, 4, 0)
    return data, out_offsets


def reconstruct_annotated_haplotypes_fused(regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs, v_starts,
                                           ilens, alt_alleles, alt_offsets, ref_, ref_offsets, pad_char,
                                           output_length, keep=None, keep_offsets=None, to_rc=None, parallel=True,
                                           *, ctx=None):
    """src/ffi/mod.rs:2239-2397.  Returns ``(out_data, annot_v, annot_pos, out_offsets)``."""
    ctx = ctx or default_ctx()
    out_offsets, total = _begin(ctx, regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs, v_starts, ilens,
                                alt_alleles, alt_offsets, ref_, ref_offsets, output_length, keep, keep_offsets, to_rc)
    out = np.empty(total, np.uint8)
    av, ap = np.empty(total, np.int32), np.empty(total, np.int32)
    check(lib.gvl_reconstruct_haplotypes_fused_finish(ctx.handle, C.c_int(MODE_ANNOTATED), c_u8(int(pad_char)),
                                                      _p(out), _p(av), _p(ap)))
    return out, av, ap, out_offsets


def reconstruct_haplotypes_from_sparse(out, out_offsets, regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs,
                                       v_starts, ilens, alt_alleles, alt_offsets, ref_, ref_offsets, pad_char,
                                       keep=None, keep_offsets=None, annot_v_idxs=None, annot_ref_pos=None,
                                       parallel=True, *, ctx=None):
    """src/ffi/mod.rs:634-655.  Caller-sized rows; writes ``out`` (and the annotation buffers) in place."""
    ctx = ctx or default_ctx()
    assert out.dtype == np.uint8 and out.flags.c_contiguous
    oo = _c(out_offsets, np.int64)
    rg, sh = _c(regions, np.int32), _c(shifts, np.int32)
    goi = _c(geno_offset_idx, np.int64)
    batch, ploidy = goi.shape
    go = _starts_stops(geno_offsets)
    gv, vs, il = _c(geno_v_idxs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    aa, ao = _c(alt_alleles, np.uint8), _c(alt_offsets, np.int64)
    rf, ro = _c(ref_, np.uint8), _c(ref_offsets, np.int64)
    kp, ko = _c(keep, np.bool_), _c(keep_offsets, np.int64)
    for a in (annot_v_idxs, annot_ref_pos):
        assert a is None or (a.dtype == np.int32 and a.flags.c_contiguous)
    check(lib.gvl_reconstruct_haplotypes_from_sparse(
        ctx.handle, _p(out), _p(oo), _p(rg), _p(sh), _p(goi), c_i64(batch), c_i64(ploidy), _p(go), c_i64(go.shape[1]),
        _p(gv), c_i64(gv.size), _p(vs), _p(il), c_i64(vs.size), _p(aa), _p(ao), _p(rf), _p(ro), c_i64(ro.size - 1),
        c_u8(int(pad_char)), _p(kp), _p(ko), _p(annot_v_idxs), _p(annot_ref_pos)))


def reconstruct_haplotypes_spliced_fused(permuted_regions, flat_shifts, flat_geno_offset_idx, out_offsets, geno_offsets,
                                         geno_v_idxs, v_starts, ilens, alt_alleles, alt_offsets, ref_, ref_offsets, pad_char,
                                         keep=None, keep_offsets=None, to_rc=None, parallel=True, *, _annotated=False,
                                         ctx=None):
    """src/ffi/mod.rs:1983-2071.  The splice plan's permuted elements as rows of ploidy 1 sized by the caller's
    ``out_offsets``; returns ``out_data`` only (the caller holds the offsets)."""
    ctx = ctx or default_ctx()
    oo = _c(out_offsets, np.int64)
    rg, sh = _c(permuted_regions, np.int32), _c(flat_shifts, np.int32).reshape(-1)
    goi = _c(flat_geno_offset_idx, np.int64).reshape(-1)
    n_perm = goi.size
    assert oo.size == n_perm + 1 and rg.shape == (n_perm, 3) and sh.size == n_perm
    go = _starts_stops(geno_offsets)
    gv, vs, il = _c(geno_v_idxs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    aa, ao = _c(alt_alleles, np.uint8), _c(alt_offsets, np.int64)
    rf, ro = _c(ref_, np.uint8), _c(ref_offsets, np.int64)
    kp, ko, rc = _c(keep, np.bool_), _c(keep_offsets, np.int64), _c(to_rc, np.bool_)
    total = int(oo[-1])
    out = np.empty(total, np.uint8)
    av = np.empty(total, np.int32) if _annotated else None
    ap = np.empty(total, np.int32) if _annotated else None
    check(lib.gvl_reconstruct_haplotypes_spliced_fused(
        ctx.handle, _p(out), _p(av), _p(ap), _p(rg), _p(sh), _p(goi), c_i64(n_perm), _p(oo), _p(go), c_i64(go.shape[1]),
        _p(gv), c_i64(gv.size), _p(vs), _p(il), c_i64(vs.size), _p(aa), _p(ao), _p(rf), _p(ro), c_i64(ro.size - 1),
        c_u8(int(pad_char)), _p(kp), _p(ko), _p(rc)))
    return (out, av, ap) if _annotated else out


def reconstruct_annotated_haplotypes_spliced_fused(permuted_regions, flat_shifts, flat_geno_offset_idx, out_offsets,
                                                   geno_offsets, geno_v_idxs, v_starts, ilens, alt_alleles, alt_offsets,
                                                   ref_, ref_offsets, pad_char, keep=None, keep_offsets=None, to_rc=None,
                                                   parallel=True, *, ctx=None):
    """src/ffi/mod.rs:2097-2211.  Returns ``(out_data, annot_v, annot_pos)``."""
    return reconstruct_haplotypes_spliced_fused(permuted_regions, flat_shifts, flat_geno_offset_idx, out_offsets,
                                                geno_offsets, geno_v_idxs, v_starts, ilens, alt_alleles, alt_offsets, ref_,
                                                ref_offsets, pad_char, keep, keep_offsets, to_rc, parallel,
                                                _annotated=True, ctx=ctx)


def choose_exonic_variants(starts, ends, geno_offset_idx, geno_v_idxs, geno_offsets, v_starts, ilens, *, ctx=None):
    """src/ffi/mod.rs:229-238.  Returns ``(keep bool[n], keep_offsets i64[n_rows + 1])``."""
    ctx = ctx or default_ctx()
    goi = _c(geno_offset_idx, np.int64)
    n_q, ploidy = goi.shape
    go = _starts_stops(geno_offsets)
    st, en = _c(starts, np.int32), _c(ends, np.int32)
    gv, vs, il = _c(geno_v_idxs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    flat = goi.reshape(-1)
    n = int(np.maximum(go[1, flat] - go[0, flat], 0).sum()) if flat.size else 0  # the allocation size (O(batch) host work)
    keep = np.zeros(n, np.bool_)
    koff = np.zeros(n_q * ploidy + 1, np.int64)
    check(lib.gvl_choose_exonic_variants(ctx.handle, _p(st), _p(en), _p(goi), c_i64(n_q), c_i64(ploidy), _p(gv),
                                         c_i64(gv.size), _p(go), c_i64(go.shape[1]), _p(vs), _p(il), c_i64(vs.size),
                                         _p(keep), c_i64(n), _p(koff)))
    return keep, koff


def get_reference(regions, out_offsets, reference, ref_offsets, pad_char, parallel=True, to_rc=None, *, mode="u8",
                  ctx=None):
    """src/ffi/mod.rs:2402-2411.  Returns the flat padded reference rows (``mode="onehot"``: (total, 4) uint8)."""
    ctx = ctx or default_ctx()
    m = _MODES[mode]
    if m not in (MODE_U8, MODE_ONEHOT):
        raise ValueError("get_reference: mode must be 'u8' or 'onehot'")
    rg, oo = _c(regions, np.int32), _c(out_offsets, np.int64)
    rf, ro, rc = _c(reference, np.uint8), _c(ref_offsets, np.int64), _c(to_rc, np.bool_)
    n = rg.shape[0]
    assert oo.size == n + 1
    total = int(oo[-1])
    out = np.zeros(total * (4 if m == MODE_ONEHOT else 1), np.uint8)  # (zeros like the reference's allocation)
    check(lib.gvl_get_reference(ctx.handle, _p(rg), _p(oo), c_i64(n), _p(rf), _p(ro), c_i64(ro.size - 1),
                                c_u8(int(pad_char)), _p(rc), C.c_int(m), _p(out)))
    return out.reshape(total, 4) if m == MODE_ONEHOT else out


def ragged_to_padded(data, offsets, out, itemsize, out_len, *, ctx=None):
    """src/ragged/mod.rs:7-23.  Copies each ragged row into the PRE-FILLED ``(n_rows, out_len)`` buffer ``out`` in place."""
    ctx = ctx or default_ctx()
    d, oo = np.ascontiguousarray(data), _c(offsets, np.int64)
    if not out.flags.c_contiguous:
        raise ValueError("out must be contiguous")  # (PyValueError at src/ragged/mod.rs:18)
    n_rows = oo.size - 1
    if out.nbytes != n_rows * int(out_len) * int(itemsize):
        raise ValueError("out holds %d bytes, (n_rows, out_len) x itemsize needs %d" % (out.nbytes, n_rows * out_len * itemsize))
    check(lib.gvl_ragged_to_padded(ctx.handle, _p(d), _p(oo), c_i64(n_rows), _p(out), c_i64(int(itemsize)),
                                   c_i64(int(out_len))))


def get_diffs_sparse(geno_offset_idx, geno_v_idxs, geno_offsets, ilens, keep=None, keep_offsets=None, q_starts=None,
                     q_ends=None, v_starts=None, parallel=True, *, ctx=None):
    """src/ffi/mod.rs:145-157.  Returns int32 ``(n_queries, ploidy)``."""
    ctx = ctx or default_ctx()
    goi = _c(geno_offset_idx, np.int64)
    n_q, ploidy = goi.shape
    go = _starts_stops(geno_offsets)
    gv, il = _c(geno_v_idxs, np.int32), _c(ilens, np.int32)
    kp, ko = _c(keep, np.bool_), _c(keep_offsets, np.int64)
    qs, qe, vs = _c(q_starts, np.int32), _c(q_ends, np.int32), _c(v_starts, np.int32)
    if not (qs is not None and qe is not None and vs is not None):
        qs = qe = vs = None  # src/genotypes/mod.rs:35: the clipped branch needs all three
    diffs = np.zeros((n_q, ploidy), np.int32)
    check(lib.gvl_get_diffs_sparse(ctx.handle, _p(goi), c_i64(n_q), c_i64(ploidy), _p(gv), c_i64(gv.size), _p(go),
                                   c_i64(go.shape[1]), _p(il), c_i64(il.size), _p(kp), _p(ko), _p(qs), _p(qe), _p(vs),
                                   _p(diffs)))
    return diffs


# ---------------------------------------------------------------------------------- tracks
def intervals_to_tracks(offset_idxs, starts, itv_starts, itv_ends, itv_values, itv_offsets, out, out_offsets,
                        parallel=True, *, ctx=None):
    """src/ffi/mod.rs:190-201.  Paints intervals into ``out`` (f32, fully overwritten) in place."""
    ctx = ctx or default_ctx()
    assert out.dtype == np.float32 and out.flags.c_contiguous
    oi, st = _c(offset_idxs, np.int64), _c(starts, np.int32)
    s, e, v = _c(itv_starts, np.int32), _c(itv_ends, np.int32), _c(itv_values, np.float32)
    io, oo = _c(itv_offsets, np.int64), _c(out_offsets, np.int64)
    check(lib.gvl_intervals_to_tracks(ctx.handle, _p(oi), _p(st), c_i64(st.size), _p(s), _p(e), _p(v), c_i64(s.size),
                                      _p(io), c_i64(io.size - 1), _p(out), _p(oo)))


def shift_and_realign_tracks_sparse(out, out_offsets, regions, shifts, geno_offset_idx, geno_v_idxs, geno_offsets,
                                    v_starts, ilens, tracks, track_offsets, params, keep=None, keep_offsets=None,
                                    strategy_id=0, base_seed=0, parallel=True, *, ctx=None):
    """src/ffi/mod.rs:2439-2458.  Dense f32 source windows; writes ``out`` in place."""
    ctx = ctx or default_ctx()
    assert out.dtype == np.float32 and out.flags.c_contiguous
    oo, rg, sh = _c(out_offsets, np.int64), _c(regions, np.int32), _c(shifts, np.int32)
    goi = _c(geno_offset_idx, np.int64)
    batch, ploidy = goi.shape
    go = _starts_stops(geno_offsets)
    gv, vs, il = _c(geno_v_idxs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    tr, to, pa = _c(tracks, np.float32), _c(track_offsets, np.int64), _c(params, np.float64)
    kp, ko = _c(keep, np.bool_), _c(keep_offsets, np.int64)
    check(lib.gvl_shift_and_realign_tracks_sparse(
        ctx.handle, _p(out), _p(oo), _p(rg), _p(sh), _p(goi), c_i64(batch), c_i64(ploidy), _p(gv), c_i64(gv.size), _p(go),
        c_i64(go.shape[1]), _p(vs), _p(il), c_i64(vs.size), _p(tr), _p(to), _p(pa), _p(kp), _p(ko),
        c_i64(int(strategy_id)), c_u64(int(base_seed))))


def shift_and_realign_tracks_from_svar2(out, out_offsets, regions, shifts, vk_pos, vk_key, vk_off, dense_pos, dense_key,
                                        dense_range, dense_present, dense_present_off, key_ilen, tracks, track_offsets,
                                        params, strategy_id=0, base_seed=0, query_seed=None, parallel=True, *, ctx=None):
    """src/tracks/mod.rs:705-856 (the core behind src/ffi/mod.rs:1836-1960): dense f32 source windows realigned with the
    svar2 two-channel variant source; ``key_ilen`` is the decoded key table.  Writes ``out`` in place; rows are sized
    by ``out_offsets`` (the offsets `reconstruct_haplotypes_from_svar2(..., output_length=-1)` returns)."""
    ctx = ctx or default_ctx()
    assert out.dtype == np.float32 and out.flags.c_contiguous
    oo, rg, sh = _c(out_offsets, np.int64), _c(regions, np.int32), _c(shifts, np.int32)
    batch, ploidy = sh.shape
    vp, vk, vo = _c(vk_pos, np.int32), _c(vk_key, np.int32), _c(vk_off, np.int64)
    dp, dk, dr = _c(dense_pos, np.int32), _c(dense_key, np.int32), _c(dense_range, np.int32)
    db, do_, ki = _c(dense_present, np.uint8), _c(dense_present_off, np.int64), _c(key_ilen, np.int32)
    tr, to, pa = _c(tracks, np.float32), _c(track_offsets, np.int64), _c(params, np.float64)
    qs = _c(query_seed, np.int64)
    check(lib.gvl_shift_and_realign_tracks_from_svar2(
        ctx.handle, _p(out), _p(oo), _p(rg), _p(sh), c_i64(batch), c_i64(ploidy), _p(vp), _p(vk), _p(vo), _p(dp), _p(dk),
        c_i64(dp.size), _p(dr), _p(db), _p(do_), _p(ki), c_i64(ki.size), _p(tr), _p(to), _p(pa), c_i64(int(strategy_id)),
        c_u64(int(base_seed)), _p(qs)))


def intervals_and_realign_track_fused(out, out_offsets, regions, shifts, geno_offset_idx, geno_v_idxs, geno_offsets,
                                      v_starts, ilens, offset_idxs, itv_starts, itv_ends, itv_values, itv_offsets,
                                      track_offsets, params, strategy_id, base_seed, keep=None, keep_offsets=None,
                                      to_rc=None, parallel=True, *, ctx=None):
    """src/ffi/mod.rs:2553-2672.  Paint + realign (+ reverse for negative strands) of ONE track into
    the caller's ``out`` slice, exactly the call the reference makes per track (_reconstruct.py:257-290)."""
    ctx = ctx or default_ctx()
    assert out.dtype == np.float32 and out.flags.c_contiguous
    oo, rg, sh = _c(out_offsets, np.int64), _c(regions, np.int32), _c(shifts, np.int32)
    goi = _c(geno_offset_idx, np.int64)
    batch, ploidy = goi.shape
    go = geno_offsets if (isinstance(geno_offsets, np.ndarray) and geno_offsets.ndim == 2 and
                          geno_offsets.dtype == np.int64 and geno_offsets.flags.c_contiguous) else _starts_stops(geno_offsets)
    gv, vs, il = _c(geno_v_idxs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    oi = _c(offset_idxs, np.int64)
    s, e, v = _c(itv_starts, np.int32), _c(itv_ends, np.int32), _c(itv_values, np.float32)
    io, to, pa = _c(itv_offsets, np.int64), _c(track_offsets, np.int64), _c(params, np.float64)
    kp, ko, rc = _c(keep, np.bool_), _c(keep_offsets, np.int64), _c(to_rc, np.bool_)
    check(lib.gvl_intervals_and_realign_track_fused(
        ctx.handle, _p(out), _p(oo), _p(rg), _p(sh), _p(goi), c_i64(batch), c_i64(ploidy), _p(gv), c_i64(gv.size), _p(go),
        c_i64(go.shape[1]), _p(vs), _p(il), c_i64(vs.size), _p(oi), _p(s), _p(e), _p(v), c_i64(s.size), _p(io),
        c_i64(io.size - 1), _p(to), _p(pa), c_i64(int(strategy_id)), c_u64(int(base_seed)), _p(kp), _p(ko), _p(rc)))


def _debug_xorshift64(x: int, *, ctx=None) -> int:
    """src/ffi/mod.rs:2824 -- evaluated by a device kernel."""
    ctx = ctx or default_ctx()
    out = c_u64(0)
    check(lib.gvl_debug_xorshift64(ctx.handle, c_u64(int(x)), C.byref(out)))
    return int(out.value)


def _debug_hash4(a: int, b: int, c: int, d: int, *, ctx=None) -> int:
    """src/ffi/mod.rs:2830 -- evaluated by a device kernel."""
    ctx = ctx or default_ctx()
    out = c_u64(0)
    check(lib.gvl_debug_hash4(ctx.handle, c_u64(int(a)), c_u64(int(b)), c_u64(int(c)), c_u64(int(d)), C.byref(out)))
    return int(out.value)


# ---------------------------------------------------------------------------------- svar2 source
def hap_diffs_svar2(regions, ploidy, vk_pos, vk_key, vk_off, dense_pos, dense_key, dense_range, dense_present,
                    dense_present_off, key_ilen, *, ctx=None):
    """src/svar2/mod.rs:73-146 (the core of hap_diffs_from_svar2_readbound, src/ffi/mod.rs:1414-1427) over the decoded key
    table's ILEN column.  Returns int32 ``(batch, ploidy)``."""
    ctx = ctx or default_ctx()
    rg = _c(regions, np.int32)
    batch = rg.shape[0]
    vp, vk, vo = _c(vk_pos, np.int32), _c(vk_key, np.int32), _c(vk_off, np.int64)
    dp, dk, dr = _c(dense_pos, np.int32), _c(dense_key, np.int32), _c(dense_range, np.int32)
    db, do, ki = _c(dense_present, np.uint8), _c(dense_present_off, np.int64), _c(key_ilen, np.int32)
    diffs = np.zeros((batch, int(ploidy)), np.int32)
    check(lib.gvl_hap_diffs_svar2(ctx.handle, _p(rg), c_i64(batch), c_i64(int(ploidy)), _p(vp), _p(vk), _p(vo), _p(dp), _p(dk),
                                  c_i64(dp.size), _p(dr), _p(db), _p(do), _p(ki), c_i64(ki.size), _p(diffs)))
    return diffs


def reconstruct_haplotypes_from_svar2(regions, shifts, vk_pos, vk_key, vk_off, dense_pos, dense_key, dense_range,
                                      dense_present, dense_present_off, key_ilen, key_alt, key_alt_off, ref_,
                                      ref_offsets, pad_char, output_length, parallel=True, *, to_rc=None, mode="u8",
                                      out=None, ctx=None):
    """src/ffi/mod.rs:874-893 with the codec keys + LUT replaced by the DECODED key table
    (key_ilen, key_alt, key_alt_off; see gvl_svar2_channels in include/gvl_b200.h).  Returns (out, out_offsets)."""
    ctx = ctx or default_ctx()
    m = _MODES[mode]
    rg, sh = _c(regions, np.int32), _c(shifts, np.int32)
    batch, ploidy = sh.shape
    vp, vk, vo = _c(vk_pos, np.int32), _c(vk_key, np.int32), _c(vk_off, np.int64)
    dp, dk, dr = _c(dense_pos, np.int32), _c(dense_key, np.int32), _c(dense_range, np.int32)
    db, do = _c(dense_present, np.uint8), _c(dense_present_off, np.int64)
    ki, ka, ko = _c(key_ilen, np.int32), _c(key_alt, np.uint8), _c(key_alt_off, np.int64)
    rf, ro, rc = _c(ref_, np.uint8), _c(ref_offsets, np.int64), _c(to_rc, np.bool_)
    out_offsets = np.empty(batch * ploidy + 1, np.int64)
    total = c_i64(0)
    check(lib.gvl_reconstruct_haplotypes_from_svar2_begin(
        ctx.handle, _p(rg), _p(sh), c_i64(batch), c_i64(ploidy), _p(vp), _p(vk), _p(vo), _p(dp), _p(dk), c_i64(dp.size),
        _p(dr), _p(db), _p(do), _p(ki), _p(ka), _p(ko), c_i64(ki.size), _p(rf), _p(ro), c_i64(ro.size - 1),
        c_i64(int(output_length)), _p(rc), _p(out_offsets), C.byref(total)))
    n = int(total.value)
    if m == MODE_ANNOTATED:
        out, av, ap = np.empty(n, np.uint8), np.empty(n, np.int32), np.empty(n, np.int32)
        check(lib.gvl_reconstruct_haplotypes_fused_finish(ctx.handle, C.c_int(m), c_u8(int(pad_char)), _p(out), _p(av), _p(ap)))
        return out, av, ap, out_offsets
    nbytes = n * (4 if m in (MODE_ONEHOT, MODE_ONEHOT_CF) else 1)
    if out is None:
        out = np.empty(nbytes, np.uint8)
    assert out.dtype == np.uint8 and out.size >= nbytes and out.flags.c_contiguous
    check(lib.gvl_reconstruct_haplotypes_fused_finish(ctx.handle, C.c_int(m), c_u8(int(pad_char)), _p(out), c_vp(0), c_vp(0)))
    out = out[:nbytes]
    return (out.reshape(n, 4) if m == MODE_ONEHOT else out), out_offsets
