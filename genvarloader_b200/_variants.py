"""`variants` / `variant-windows` kernels with the names and argument order of the reference's PyO3 module
(src/ffi/mod.rs:255-630, 2808): numpy in, numpy out, the work done by the CUDA library through the gvl_* host layer
(include/gvl_b200.h, csrc/gvl_variants.cu + gvl_host_variants.cuh).  Drop-in for the imports at
python/genvarloader/_dataset/_flat_variants.py:15-32.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import c_i32, c_i64, c_u8, c_vp, check, lib

_FIELDS = ("alt", "ref", "flank_tokens", "ref_window", "alt_window")  # field_kind of gvl_assemble_variant_buffers


def _ctx(ctx):
    from ._kernels import default_ctx

    return ctx or default_ctx()


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def _p(a):
    return c_vp(0) if a is None else c_vp(a.ctypes.data)


def _fetch(ctx, which: int, n_items: int, dtype) -> np.ndarray:
    out = np.empty(int(n_items), dtype)
    check(lib.gvl_variants_fetch(ctx.handle, C.c_int(which), _p(out), c_i64(out.nbytes)))
    return out


def _word_dtype(a: np.ndarray):
    if a.dtype.itemsize != 4:
        raise TypeError(f"4-byte items expected (int32 / float32), got {a.dtype}")
    return a.dtype


def gather_rows(geno_offset_idx, geno_offsets, data, ctx=None):
    """gather_rows_i32 / gather_rows_f32 (src/ffi/mod.rs:255-288): -> (values, out_offsets)."""
    ctx = _ctx(ctx)
    goi = _c(np.asarray(geno_offset_idx).reshape(-1), np.int64)
    go = _c(geno_offsets, np.int64)
    assert go.ndim == 2 and go.shape[0] == 2, "geno_offsets is the (2, n) starts / stops array"
    data = np.ascontiguousarray(data)
    dt = _word_dtype(data)
    off = np.empty(len(goi) + 1, np.int64)
    total = c_i64(0)
    check(lib.gvl_gather_rows(ctx.handle, _p(goi), c_i64(len(goi)), _p(go), c_i64(go.shape[1]), _p(data), c_i64(data.size), _p(off),
                              C.byref(total)))
    return _fetch(ctx, 0, total.value, dt), off


gather_rows_i32 = gather_rows_f32 = gather_rows


def gather_alleles(v_idxs, allele_bytes, allele_offsets, ctx=None):
    """src/ffi/mod.rs:291-304: -> (bytes, seq_offsets)."""
    ctx = _ctx(ctx)
    v = _c(v_idxs, np.int32)
    ab, ao = _c(allele_bytes, np.uint8), _c(allele_offsets, np.int64)
    off = np.empty(len(v) + 1, np.int64)
    total = c_i64(0)
    check(lib.gvl_gather_alleles(ctx.handle, _p(v), c_i64(len(v)), _p(ab), c_i64(ab.size), _p(ao), c_i64(len(ao) - 1), _p(off),
                                 C.byref(total)))
    return _fetch(ctx, 0, total.value, np.uint8), off


def rc_alleles(byte_data, seq_offsets, var_offsets, to_rc_row, ctx=None) -> None:
    """src/ffi/mod.rs:2808: reverse-complements the alleles of the masked rows IN PLACE."""
    ctx = _ctx(ctx)
    assert byte_data.dtype == np.uint8 and byte_data.flags.c_contiguous, "rc_alleles requires a contiguous uint8 byte_data"
    so, vo = _c(seq_offsets, np.int64), _c(var_offsets, np.int64)
    m = _c(np.asarray(to_rc_row).reshape(-1), np.bool_)
    check(lib.gvl_rc_alleles(ctx.handle, _p(byte_data), c_i64(byte_data.size), _p(so), c_i64(len(so) - 1), _p(vo),
                             c_i64(len(vo) - 1), _p(m)))


def compact_keep(values, row_offsets, keep, ctx=None):
    """compact_keep_i32 / compact_keep_f32 (src/ffi/mod.rs:308-333): -> (kept values, new row offsets)."""
    ctx = _ctx(ctx)
    values = np.ascontiguousarray(values)
    dt = _word_dtype(values)
    ro, k = _c(row_offsets, np.int64), _c(keep, np.bool_)
    assert len(k) == len(values)
    new_off = np.empty(len(ro), np.int64)
    total = c_i64(0)
    check(lib.gvl_compact_keep(ctx.handle, _p(values), c_i64(values.size), _p(ro), c_i64(len(ro) - 1), _p(k), _p(new_off),
                               C.byref(total)))
    return _fetch(ctx, 0, total.value, dt), new_off


compact_keep_i32 = compact_keep_f32 = compact_keep


def fill_empty_fixed(data, offsets, inner, fill, ctx=None):
    """fill_empty_fixed_i32 / _f32 (src/ffi/mod.rs:362-388): -> (data, new offsets)."""
    ctx = _ctx(ctx)
    data = np.ascontiguousarray(data)
    dt = _word_dtype(data)
    off = _c(offsets, np.int64)
    bits = int(np.array(fill, dt).view(np.uint32))
    new_off = np.empty(len(off), np.int64)
    total = c_i64(0)
    check(lib.gvl_fill_empty_fixed(ctx.handle, _p(data), c_i64(data.size), _p(off), c_i64(len(off) - 1), c_i64(int(inner)),
                                   C.c_uint32(bits), _p(new_off), C.byref(total)))
    return _fetch(ctx, 0, total.value * int(inner), dt), new_off


fill_empty_fixed_i32 = fill_empty_fixed_f32 = fill_empty_fixed


def fill_empty_scalar(data, offsets, fill, ctx=None):
    """fill_empty_scalar_i32 / _f32 (src/ffi/mod.rs:336-359)."""
    return fill_empty_fixed(data, offsets, 1, fill, ctx)


fill_empty_scalar_i32 = fill_empty_scalar_f32 = fill_empty_scalar


def fill_empty_seq(data, var_offsets, seq_offsets, dummy, ctx=None):
    """fill_empty_seq_u8 / _i32 (src/ffi/mod.rs:391-445): -> (data, new_var_offsets, new_seq_offsets)."""
    ctx = _ctx(ctx)
    data = np.ascontiguousarray(data)
    if data.dtype.itemsize not in (1, 4):
        raise TypeError(f"1- or 4-byte items expected, got {data.dtype}")
    vo, so = _c(var_offsets, np.int64), _c(seq_offsets, np.int64)
    dummy = np.ascontiguousarray(dummy, data.dtype)
    new_var = np.empty(len(vo), np.int64)
    n_new, total = c_i64(0), c_i64(0)
    check(lib.gvl_fill_empty_seq(ctx.handle, _p(data), C.c_int(data.dtype.itemsize), c_i64(data.size), _p(vo), c_i64(len(vo) - 1),
                                 _p(so), c_i64(len(so) - 1), _p(dummy), c_i64(dummy.size), _p(new_var), C.byref(n_new),
                                 C.byref(total)))
    return _fetch(ctx, 0, total.value, data.dtype), new_var, _fetch(ctx, 1, n_new.value + 1, np.int64)


fill_empty_seq_u8 = fill_empty_seq_i32 = fill_empty_seq


def assemble_variant_buffers(mode, v_idxs, row_offsets, alt_global, alt_off_global, ref_global, ref_off_global, want_ref_bytes,
                             want_flank, ref_mode, alt_mode, flank_len, lut, v_contigs, v_starts, ilens, reference, ref_offsets,
                             pad_char, ctx=None) -> dict:
    """assemble_variant_buffers_u8 / _i32 (src/ffi/mod.rs:460-630; the dtype of `lut` selects the token type as in
    _flat_variants.py:773-832): -> {field: (data, seq_offsets)} in the reference's order."""
    ctx = _ctx(ctx)
    v = _c(v_idxs, np.int32)
    ro = _c(row_offsets, np.int64)
    ag, ao = _c(alt_global, np.uint8), _c(alt_off_global, np.int64)
    rg, rgo = _c(ref_global, np.uint8), _c(ref_off_global, np.int64)
    tok_dt = np.dtype(np.uint8)
    if lut is not None:
        lut = np.asarray(lut)
        tok_dt = np.dtype(np.uint8) if lut.dtype == np.uint8 else np.dtype(np.int32)
        lut = np.ascontiguousarray(lut, tok_dt)
        assert lut.size >= 256, "tokenize: lut must have >= 256 entries"
    vc, vs, il = _c(v_contigs, np.int32), _c(v_starts, np.int32), _c(ilens, np.int32)
    rf, rfo = _c(reference, np.uint8), _c(ref_offsets, np.int64)
    n_fields = c_i32(0)
    kind = (c_i32 * 4)()
    items = (c_i64 * 4)()
    tok = (c_i32 * 4)()
    check(lib.gvl_assemble_variant_buffers(
        ctx.handle, c_i64(int(mode)), _p(v), c_i64(len(v)), _p(ag), _p(ao), _p(rg), _p(rgo), c_i64(len(ao) - 1),
        C.c_int(bool(want_ref_bytes)), C.c_int(bool(want_flank)), c_i64(int(ref_mode)), c_i64(int(alt_mode)), c_i64(int(flank_len)),
        _p(lut), C.c_int(tok_dt.itemsize), _p(vc), _p(vs), _p(il), _p(rf), _p(rfo), c_i64(len(rfo) - 1), c_u8(int(pad_char)),
        C.byref(n_fields), kind, items, tok))
    out = {}
    for j in range(n_fields.value):
        name = _FIELDS[kind[j]]
        dt = np.uint8 if tok[j] == 1 else np.int32
        data = _fetch(ctx, 2 * j, items[j], dt)
        off = ro.copy() if name == "flank_tokens" else _fetch(ctx, 2 * j + 1, len(v) + 1, np.int64)
        out[name] = (data, off)
    return out


assemble_variant_buffers_u8 = assemble_variant_buffers_i32 = assemble_variant_buffers
