"""CPU restatement (numpy) of the reference's `variants` / `variant-windows` flat-buffer cores.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under genvarloader_b200/).  Every function follows the
Rust core it names in /root/reference/src/variants/{mod.rs,windows.rs} and takes the argument list of the matching
#[pyfunction] in src/ffi/mod.rs:255-630, 2808.  Pinned by the reference's own frozen goldens
(tests/parity/golden/{gather_rows_*,gather_alleles,rc_alleles,compact_keep_*,fill_empty_*,assemble_variant_buffers}.npz,
copied value-for-value into tests/golden/ref_*.npz by tests/golden/make_golden.py): tests/test_oracle_variants.py.
"""
from __future__ import annotations

import numpy as np


def _excl(lengths) -> np.ndarray:
    off = np.zeros(len(lengths) + 1, np.int64)
    np.cumsum(lengths, out=off[1:])
    return off


def _seg_index(off: np.ndarray):
    """(segment, position inside the segment) of every element of a ragged layout with offsets `off`."""
    n = int(off[-1])
    seg = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
    return seg, np.arange(n, dtype=np.int64) - off[seg]


def gather_rows(geno_offset_idx, geno_offsets, data):
    """src/variants/mod.rs:6-29 (gather_rows_impl), entries src/ffi/mod.rs:255-288: row i of the output is
    data[o_starts[goi[i]] : o_stops[goi[i]]]."""
    goi = np.asarray(geno_offset_idx, np.int64)
    go = np.asarray(geno_offsets, np.int64)
    data = np.asarray(data)
    starts, stops = go[0][goi], go[1][goi]
    off = _excl(stops - starts)
    seg, k = _seg_index(off)
    return data[starts[seg] + k], off


def gather_alleles(v_idxs, allele_bytes, allele_offsets):
    """src/variants/mod.rs:52-78: the byte strings of the selected variants, back to back."""
    v = np.asarray(v_idxs, np.int64)
    ao = np.asarray(allele_offsets, np.int64)
    off = _excl(ao[v + 1] - ao[v])
    seg, k = _seg_index(off)
    return np.asarray(allele_bytes, np.uint8)[ao[v][seg] + k], off


def _comp(b: np.ndarray) -> np.ndarray:
    """src/reverse.rs:45-53: A<->T, C<->G, every other byte unchanged (lowercase included)."""
    at = (b == 65) | (b == 84)
    cg = (b == 67) | (b == 71)
    return (b ^ np.where(at, 21, 0) ^ np.where(cg, 4, 0)).astype(np.uint8)


def rc_alleles(byte_data, seq_offsets, var_offsets, to_rc_row):
    """src/variants/mod.rs:90-108 (rc_alleles_inplace), entry src/ffi/mod.rs:2808: reverse-complement every allele of
    the masked (b*p) rows; returns the new bytes (the reference mutates in place)."""
    out = np.array(byte_data, np.uint8, copy=True)
    so = np.asarray(seq_offsets, np.int64)
    vo = np.asarray(var_offsets, np.int64)
    mask = np.asarray(to_rc_row, np.bool_).reshape(-1)
    for g in np.nonzero(mask)[0]:
        for a in range(int(vo[g]), int(vo[g + 1])):
            s, e = int(so[a]), int(so[a + 1])
            out[s:e] = _comp(out[s:e][::-1])
    return out


def compact_keep(values, row_offsets, keep):
    """src/variants/mod.rs:112-135: drop values with keep == False, rebuild the row offsets."""
    values = np.asarray(values)
    keep = np.asarray(keep, np.bool_)
    ro = np.asarray(row_offsets, np.int64)
    pos = _excl(keep.astype(np.int64))
    return values[keep], pos[ro] - pos[ro[0]]  # new_offsets[i] = kept values of the rows before row i


def fill_empty_scalar(data, offsets, fill):
    """src/variants/mod.rs:157-183: every empty row receives one `fill` element."""
    return fill_empty_fixed(data, offsets, 1, fill)


def fill_empty_fixed(data, offsets, inner, fill):
    """src/variants/mod.rs:205-234: every empty row receives `inner` copies of `fill`; rows with n variants copy their
    n * inner elements."""
    data = np.asarray(data)
    off = np.asarray(offsets, np.int64)
    ln = np.diff(off)
    new_off = _excl(np.where(ln > 0, ln, 1))
    seg, k = _seg_index(new_off)
    out = np.full(int(new_off[-1]) * int(inner), fill, data.dtype)
    src_var = off[seg] + k  # source variant of every new variant (unused where the row is empty)
    has = ln[seg] > 0
    out2 = out.reshape(-1, int(inner)) if inner else out.reshape(len(seg), 0)
    if has.any():
        out2[has] = data.reshape(-1, int(inner))[src_var[has]]
    return out2.reshape(-1), new_off


def fill_empty_seq(data, var_offsets, seq_offsets, dummy):
    """src/variants/mod.rs:259-309: two-level ragged: every empty (b*p) row receives one dummy sequence."""
    data = np.asarray(data)
    vo = np.asarray(var_offsets, np.int64)
    so = np.asarray(seq_offsets, np.int64)
    dummy = np.asarray(dummy, data.dtype)
    nv = np.diff(vo)
    new_var = _excl(np.where(nv > 0, nv, 1))
    row, k = _seg_index(new_var)
    has = nv[row] > 0
    src_var = np.where(has, vo[row] + k, 0)
    if len(so) > 1:
        lens = np.where(has, so[np.minimum(src_var + 1, len(so) - 1)] - so[np.minimum(src_var, len(so) - 1)], len(dummy))
    else:
        lens = np.full(len(row), len(dummy), np.int64)
    new_seq = _excl(lens)
    seg, j = _seg_index(new_seq)
    out = np.empty(int(new_seq[-1]), data.dtype)
    h = has[seg]
    if h.any():
        out[h] = data[so[src_var[seg[h]]] + j[h]]
    if (~h).any():
        out[~h] = dummy[j[~h]]
    return out, new_var, new_seq


def fetch_windows(v_contigs, starts_v, ilens_v, flank_len, reference, ref_offsets, pad_char):
    """src/variants/windows.rs:98-134: the reference window [start - L, end + L) of every variant with
    end = start - min(ilen, 0) + 1; positions outside the contig read as pad_char (src/reference/mod.rs:9-53)."""
    s = np.asarray(starts_v, np.int64)
    il = np.asarray(ilens_v, np.int64)
    c = np.asarray(v_contigs, np.int64)
    ro = np.asarray(ref_offsets, np.int64)
    ref = np.asarray(reference, np.uint8)
    rstart = s - flank_len
    rend = s - np.minimum(il, 0) + 1 + flank_len
    rw_off = _excl(rend - rstart)
    seg, k = _seg_index(rw_off)
    pos = rstart[seg] + k
    c_s, c_len = ro[c][seg], (ro[c + 1] - ro[c])[seg]
    ok = (pos >= 0) & (pos < c_len)
    out = np.full(len(seg), pad_char, np.uint8)
    out[ok] = ref[c_s[ok] + pos[ok]]
    return out, rw_off


def assemble_variant_buffers(mode, v_idxs, row_offsets, alt_global, alt_off_global, ref_global, ref_off_global,
                             want_ref_bytes, want_flank, ref_mode, alt_mode, flank_len, lut, v_contigs, v_starts, ilens,
                             reference, ref_offsets, pad_char):
    """src/ffi/mod.rs:460-529 -> src/variants/windows.rs:162-296.  mode 0 = `variants` tail (raw alt bytes, optional raw
    ref bytes, optional flank tokens), mode 1 = `variant-windows` tail (token buffers only).  Returns the reference's dict
    {field: (data, seq_offsets)} in its insertion order."""
    v = np.asarray(v_idxs, np.int64)
    L = int(flank_len)
    out: dict = {}
    alt_data, alt_off = gather_alleles(v, alt_global, alt_off_global)
    sv, iv = np.asarray(v_starts)[v], np.asarray(ilens)[v]
    if int(mode) == 0:
        out["alt"] = (alt_data, alt_off)
        if want_ref_bytes and ref_global is not None and ref_off_global is not None:
            out["ref"] = gather_alleles(v, ref_global, ref_off_global)
        if want_flank:
            lut = np.asarray(lut)
            rw, rw_off = fetch_windows(v_contigs, sv, iv, L, reference, ref_offsets, pad_char)
            cols = np.arange(L, dtype=np.int64)
            f5 = rw[(rw_off[:-1, None] + cols).reshape(-1)].reshape(len(v), L)  # windows.rs:27-52
            f3 = rw[(rw_off[1:, None] - L + cols).reshape(-1)].reshape(len(v), L)
            out["flank_tokens"] = (lut[np.concatenate([f5, f3], axis=1).reshape(-1)], np.asarray(row_offsets, np.int64).copy())
        return out
    lut = np.asarray(lut)
    fetched = None
    if ref_mode == 1 or alt_mode == 1:
        fetched = fetch_windows(v_contigs, sv, iv, L, reference, ref_offsets, pad_char)
    if ref_mode == 1:
        out["ref_window"] = (lut[fetched[0]], fetched[1])
    elif ref_mode == 2:
        rd, ro_ = gather_alleles(v, ref_global, ref_off_global)
        out["ref"] = (lut[rd], ro_)
    if alt_mode == 1:
        rw, rw_off = fetched
        a_len = np.diff(alt_off)
        w_off = _excl(2 * L + a_len)  # windows.rs:55-90: flank5 . alt . flank3
        seg, k = _seg_index(w_off)
        b = np.empty(len(seg), np.uint8)
        in5, in3 = k < L, k >= L + a_len[seg]
        mid = ~in5 & ~in3
        b[in5] = rw[rw_off[seg[in5]] + k[in5]]
        b[mid] = alt_data[alt_off[seg[mid]] + k[mid] - L]
        b[in3] = rw[rw_off[seg[in3] + 1] - L + (k[in3] - L - a_len[seg[in3]])]
        out["alt_window"] = (lut[b], w_off)
    elif alt_mode == 2:
        out["alt"] = (lut[alt_data], alt_off)
    return out
