"""GPU parity of the boundary's small entries (gvl_aux.cu + the spliced fused entry) through the C ABI host layer:
get_reference, choose_exonic_variants, ragged_to_padded, reconstruct_(annotated_)haplotypes_spliced_fused --
against the reference's frozen goldens (tests/parity/golden/{get_reference,choose_exonic_variants}.npz) and the CPU
oracle, bit-exact.  Mirrors tests/parity/test_get_reference_parity.py, test_choose_exonic_variants_parity.py,
tests/unit/test_ragged_to_padded_rust.py and tests/parity/test_fused_haps_parity.py (spliced cases)."""
import numpy as np
import pytest

from tests import _golden

pytestmark = pytest.mark.gpu

N = ord("N")


@pytest.fixture(scope="module")
def K(cuda_device):
    from genvarloader_b200 import _kernels

    return _kernels


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    return oracle


def test_get_reference_golden(K):
    cases = _golden.load_golden("get_reference")
    assert len(cases) == 200
    _golden.replay_return(K.get_reference, "get_reference", cases)


def test_choose_exonic_variants_golden(K):
    cases = _golden.load_golden("choose_exonic_variants")
    assert len(cases) == 200
    _golden.replay_tuple(K.choose_exonic_variants, "choose_exonic_variants", cases)


@pytest.mark.parametrize("fixed", [True, False])
@pytest.mark.parametrize("pin", [False, True])
def test_get_reference_rc_onehot_vs_oracle(K, O, fixed, pin):
    """Rows over both contig ends, reverse-complemented rows, u8 and one-hot; with the reference registered
    (gvl_pin_static) equal-length rows take the packed one-hot kernel."""
    from genvarloader_b200 import synth
    from genvarloader_b200._ffi import lib

    d = synth.make_dataset(5, 150_000, 2, 4, 1000, 1.0)
    rng = np.random.default_rng(17 + fixed)
    n = 24
    c = rng.integers(0, len(d.ref_offsets) - 1, n)
    clen = np.diff(d.ref_offsets)[c]
    lens = np.full(n, 4096) if fixed else rng.integers(0, 6000, n)
    starts = rng.integers(-3000, clen - 1)  # (a row entirely right of the contig panics in the reference)
    starts[:3] = [-5000, -10, 0]
    starts[3] = clen[3] - 100
    regions = np.stack([c, starts, starts + lens], 1).astype(np.int32)
    oo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    to_rc = rng.random(n) < 0.5
    if pin:
        K.pin_static(d.reference)
    try:
        for rc in (None, to_rc):
            exp = O.get_reference(regions, oo, d.reference, d.ref_offsets, N, False, rc)
            got = K.get_reference(regions, oo, d.reference, d.ref_offsets, N, True, rc)
            _golden.eq("get_reference.u8", 0, got, exp)
            oh = K.get_reference(regions, oo, d.reference, d.ref_offsets, N, True, rc, mode="onehot")
            _golden.eq("get_reference.onehot", 0, oh, O.onehot(exp))
            if pin and fixed:
                assert int(lib.gvl_debug_last_exec_kernel(K.default_ctx().handle)) == 1, "packed kernel did not run"
    finally:
        if pin:
            K.unpin_static(d.reference)
    # empty batch
    assert K.get_reference(np.zeros((0, 3), np.int32), np.zeros(1, np.int64), d.reference, d.ref_offsets, N).size == 0


def test_choose_exonic_variants_vs_oracle_large(K, O):
    """More rows than one scan chunk (1024) and long genotype slices."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(23, 200_000, 4, 40, 3000, 12.0, max_indel=30)
    rng = np.random.default_rng(3)
    b = 1500
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, _, _ = synth.batch_args(d, r_idx, s_idx)
    starts, ends = regions[:, 1] + 200, regions[:, 2] - 300  # exons strictly inside the stored regions
    e_keep, e_ko = O.choose_exonic_variants(starts, ends, goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens)
    g_keep, g_ko = K.choose_exonic_variants(starts, ends, goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens)
    _golden.eq("exonic.offsets", 0, g_ko, e_ko)
    _golden.eq("exonic.keep", 0, g_keep, e_keep)
    assert 0 < int(e_keep.sum()) < e_keep.size
    # no rows
    k0, o0 = K.choose_exonic_variants(starts[:0], ends[:0], goi[:0], d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens)
    assert k0.size == 0 and o0.tolist() == [0]


@pytest.mark.parametrize("dtype,out_len", [(np.uint8, 37), (np.uint8, 64), (np.int32, 50), (np.float32, 9),
                                           (np.dtype("V3"), 21), (np.int64, 16)])
def test_ragged_to_padded_vs_oracle(K, O, dtype, out_len):
    rng = np.random.default_rng(out_len)
    dtype = np.dtype(dtype)
    n_rows = 57
    lens = rng.integers(0, 2 * out_len, n_rows)
    lens[[0, 5, 9]] = [0, out_len, out_len + 1]
    lead = 11  # offsets[0] > 0: a slice of a larger ragged buffer
    oo = (lead + np.concatenate([[0], np.cumsum(lens)])).astype(np.int64)
    raw = rng.integers(0, 256, int(oo[-1] + 5) * dtype.itemsize, dtype=np.uint8)
    data = raw.view(dtype)
    pad = rng.integers(0, 256, dtype.itemsize, dtype=np.uint8)
    exp = np.tile(pad, n_rows * out_len).view(dtype).reshape(n_rows, out_len).copy()
    got = exp.copy()
    O.ragged_to_padded(data, oo, exp, dtype.itemsize, out_len)
    K.ragged_to_padded(data, oo, got, dtype.itemsize, out_len)
    assert got.tobytes() == exp.tobytes()
    with pytest.raises(ValueError):
        K.ragged_to_padded(data, oo, got[:, ::2], dtype.itemsize, out_len)


@pytest.mark.parametrize("dtype", ["u1", "<u2", "<i4", "<f4", "<i8"])
def test_ragged_to_padded_fill_writes_the_padding_itself(O, dtype):
    """gvl_dev_ragged_to_padded_fill: rows AND padding in one pass over an UNINITIALISED output (what `Ragged.to_padded`
    calls), against the oracle's pre-fill + copy, for every item size up to 8 bytes and unaligned row starts."""
    import ctypes as C

    import torch

    from genvarloader_b200._engine import _stream
    from genvarloader_b200._ffi import check, lib, ptr
    from genvarloader_b200._kernels import default_ctx

    dtype = np.dtype(dtype)
    rng = np.random.default_rng(dtype.itemsize)
    for out_len in (1, 7, 33, 130):
        n_rows = 41
        lens = rng.integers(0, 2 * out_len, n_rows)
        lens[[0, 3, 8]] = [0, out_len, out_len + 1]
        oo = (5 + np.concatenate([[0], np.cumsum(lens)])).astype(np.int64)
        data = rng.integers(0, 256, int(oo[-1] + 3) * dtype.itemsize, dtype=np.uint8).view(dtype)
        pad = rng.integers(0, 256, dtype.itemsize, dtype=np.uint8)
        exp = np.tile(pad, n_rows * out_len).view(dtype).reshape(n_rows, out_len).copy()
        O.ragged_to_padded(data, oo, exp, dtype.itemsize, out_len)
        dev = torch.device("cuda", 0)
        d_data = torch.from_numpy(data.view(np.uint8).copy()).to(dev)
        d_oo = torch.from_numpy(oo).to(dev)
        d_out = torch.full((n_rows * out_len * dtype.itemsize,), 0xAB, dtype=torch.uint8, device=dev)  # (garbage, not the pad)
        check(lib.gvl_dev_ragged_to_padded_fill(default_ctx(0).handle, ptr(d_data), ptr(d_oo), C.c_int64(n_rows), ptr(d_out),
                                                C.c_int64(dtype.itemsize), C.c_int64(out_len), C.c_char_p(pad.tobytes()), _stream()))
        assert d_out.cpu().numpy().tobytes() == exp.tobytes(), (dtype, out_len)


@pytest.mark.parametrize("vkb,use_keep", [(2.0, False), (15.0, True)])
def test_spliced_fused_vs_oracle(K, O, vkb, use_keep):
    """The splice entries: permuted exon elements (ploidy-1 rows) with caller-sized rows, per-element RC;
    the annotated twin reverses the annotation rows of masked elements (src/ffi/mod.rs:2180-2200)."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(31, 120_000, 3, 12, 1500, vkb, max_indel=25)
    rng = np.random.default_rng(int(vkb))
    n_perm = 40
    r_idx, s_idx = rng.integers(0, d.n_regions, n_perm), rng.integers(0, d.n_samples, n_perm)
    regions, goi, _, _ = synth.batch_args(d, r_idx, s_idx)
    hap = rng.integers(0, d.ploidy, n_perm)
    flat_goi = goi[np.arange(n_perm), hap].reshape(-1, 1)
    flat_shifts = np.zeros((n_perm, 1), np.int32)
    keep = ko = None
    if use_keep:
        keep, ko = O.choose_exonic_variants(regions[:, 1], regions[:, 2], flat_goi, d.geno_v_idxs, d.geno_offsets,
                                            d.v_starts, d.ilens)
    diffs = O.get_diffs_sparse(flat_goi, d.geno_v_idxs, d.geno_offsets, d.ilens, keep, ko, regions[:, 1], regions[:, 2],
                               d.v_starts)
    lens = np.maximum(regions[:, 2] - regions[:, 1] + diffs.reshape(-1), 0)
    oo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    to_rc = rng.random(n_perm) < 0.5
    tabs = (d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.reference, d.ref_offsets)
    for rc in (None, to_rc):
        total = int(oo[-1])
        e_out, e_av, e_ap = np.zeros(total, np.uint8), np.zeros(total, np.int32), np.zeros(total, np.int32)
        O.reconstruct_haplotypes_from_sparse(e_out, oo, regions, flat_shifts, flat_goi, *tabs, N, keep, ko, e_av, e_ap)
        if rc is not None:
            O.rc_flat_rows_inplace(e_out, oo, rc)
            O.reverse_flat_rows_inplace(e_av, oo, rc)
            O.reverse_flat_rows_inplace(e_ap, oo, rc)
        got = K.reconstruct_haplotypes_spliced_fused(regions, flat_shifts, flat_goi, oo, *tabs, N, keep, ko, rc)
        _golden.eq("spliced.out", 0, got, e_out)
        g_out, g_av, g_ap = K.reconstruct_annotated_haplotypes_spliced_fused(regions, flat_shifts, flat_goi, oo, *tabs, N,
                                                                             keep, ko, rc)
        _golden.eq("spliced.annot.out", 0, g_out, e_out)
        _golden.eq("spliced.annot.v", 0, g_av, e_av)
        _golden.eq("spliced.annot.pos", 0, g_ap, e_ap)
