"""GPU parity of the PACKED one-hot execute kernel (gvl_hap_oh.cuh: 4-bit reference codes, 8 positions per
lane, 256-bit stores) against the oracle -- bit-exact -- and against the byte-oriented kernel it replaces for
GVL_MODE_ONEHOT.  The host layer packs a reference that was registered with gvl_pin_static; every test asserts
through gvl_debug_last_exec_kernel that the packed kernel is the one that ran."""
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


def _last_kernel(K):
    from genvarloader_b200._ffi import lib

    return int(lib.gvl_debug_last_exec_kernel(K.default_ctx().handle))


def _args(d, regions, sh, goi, pad, out_len):
    return (regions, sh, goi, d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
            d.reference, d.ref_offsets, pad, out_len)


def _check(K, O, d, args, rc, tag):
    e_out, e_oo = O.reconstruct_haplotypes_fused(*args, None, None, rc)
    exp = O.onehot(e_out)
    g_b, g_oo = K.reconstruct_haplotypes_fused(*args, None, None, rc, mode="onehot")  # reference not pinned: bytes kernel
    assert _last_kernel(K) == 0
    _golden.eq(tag + ".bytes_kernel", 0, g_b, exp)
    # reference + ALT alleles registered as static arrays: the host layer packs both and runs the packed kernel
    pinned = (d.reference, d.alt_alleles)
    K.pin_static(*pinned)
    try:
        g_p, g_oo2 = K.reconstruct_haplotypes_fused(*args, None, None, rc, mode="onehot")
        assert _last_kernel(K) == 1, "packed one-hot kernel did not run"
    finally:
        K.unpin_static(*pinned)
    _golden.eq(tag + ".offsets", 0, g_oo2, e_oo)
    _golden.eq(tag + ".packed_kernel", 0, g_p, exp)


@pytest.mark.parametrize("vkb,L,out_len,shifts", [(1.0, 2000, 2000, False), (10.0, 3000, -1, False),
                                                  (5.0, 5000, 4096, True), (0.1, 9000, 9000, False),
                                                  (20.0, 700, -1, False), (3.0, 16384, 16384, True),
                                                  (2.0, 3001, 2999, True), (50.0, 1237, 1237, False)])
def test_packed_onehot_vs_oracle(K, O, vkb, L, out_len, shifts):
    """fixed (aligned and odd lengths) + ragged rows, RC rows, shifts, windows over both contig ends."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(int(vkb * 10 + L), 200_000, 4, 24, L, vkb, neg_strand_frac=0.5, max_indel=30)
    rng = np.random.default_rng(L)
    b = 16
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, _ = synth.batch_args(d, r_idx, s_idx)
    sh = rng.integers(0, 40, (b, d.ploidy)).astype(np.int32) if shifts else np.zeros((b, d.ploidy), np.int32)
    for rc in (None, to_rc, np.ones_like(to_rc)):
        _check(K, O, d, _args(d, regions, sh, goi, N, out_len), rc, f"packed[{vkb},{L},{out_len}]")


@pytest.mark.parametrize("pad", [ord("A"), ord("G"), ord("n"), 0])
def test_packed_onehot_pad_chars_and_contig_ends(K, O, pad):
    """a pad char that IS a base must be complemented in reverse-complemented rows, like the reference does
    by running rc over the finished row (src/ffi/mod.rs:819-823)."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(5, 30_000, 3, 12, 4000, 4.0, neg_strand_frac=0.5, max_indel=25)
    regions = np.array([[0, -700, 3300], [0, 28_500, 32_500], [0, -10_000, -6_000], [0, 29_990, 33_990],
                        [0, -3, 3997], [0, 40_000, 44_000]], np.int32)
    goi = (np.arange(6)[:, None] * d.ploidy + np.arange(d.ploidy)[None, :]).astype(np.int64)
    sh = np.zeros(goi.shape, np.int32)
    to_rc = np.tile(np.array([False, True, True]), 4)[: goi.size].copy()
    for out_len in (4000, -1, 4100):
        _check(K, O, d, _args(d, regions, sh, goi, pad, out_len), to_rc, f"pad[{pad},{out_len}]")


def test_packed_onehot_dense_variants_many_passes(K, O):
    """more records per tile than one shared-memory pass holds, overlapping variants, every group mixed."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(11, 60_000, 2, 8, 20_000, 400.0, snp_frac=0.6, max_indel=3, dense_af=0.5)
    rng = np.random.default_rng(2)
    r_idx, s_idx = rng.integers(0, d.n_regions, 6), rng.integers(0, d.n_samples, 6)
    regions, goi, to_rc, _ = synth.batch_args(d, r_idx, s_idx)
    to_rc[::2] = True
    sh = np.zeros(goi.shape, np.int32)
    for out_len in (-1, 20_000):
        _check(K, O, d, _args(d, regions, sh, goi, N, out_len), to_rc, f"dense[{out_len}]")


def test_packed_onehot_non_acgt_bytes(K, O):
    """soft-masked / IUPAC / arbitrary bytes in the reference and in ALT alleles encode to all-zero rows; only
    upper-case ACGT are hot (definition of the one-hot)."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(21, 50_000, 3, 10, 6000, 6.0, neg_strand_frac=0.5, max_indel=12)
    rng = np.random.default_rng(9)
    ref = d.reference.copy()
    noise = rng.integers(0, 256, ref.size).astype(np.uint8)
    sel = rng.random(ref.size) < 0.3
    ref[sel] = noise[sel]
    low = rng.random(ref.size) < 0.2
    ref[low] = np.char.lower(ref[low].view("S1")).view(np.uint8)
    alt = d.alt_alleles.copy()
    sel = rng.random(alt.size) < 0.3
    alt[sel] = rng.integers(0, 256, alt.size).astype(np.uint8)[sel]
    import dataclasses

    d2 = dataclasses.replace(d, reference=ref, alt_alleles=alt)
    r_idx, s_idx = rng.integers(0, d.n_regions, 10), rng.integers(0, d.n_samples, 10)
    regions, goi, to_rc, _ = synth.batch_args(d2, r_idx, s_idx)
    sh = rng.integers(0, 30, goi.shape).astype(np.int32)
    for out_len in (-1, 5800):
        _check(K, O, d2, _args(d2, regions, sh, goi, N, out_len), to_rc, f"bytes[{out_len}]")


@pytest.mark.parametrize("L,b", [(131_072, 8), (40_000, 24), (524_288, 2)])
def test_packed_onehot_long_rows_many_tiles(K, O, L, b):
    from genvarloader_b200 import synth

    d = synth.make_dataset(L % 1000 + 3, 2_000_000, 4, 12, L, 1.0, neg_strand_frac=0.5, max_indel=20)
    rng = np.random.default_rng(L)
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, _ = synth.batch_args(d, r_idx, s_idx)
    sh = rng.integers(0, 20, goi.shape).astype(np.int32)
    for out_len in (L, -1):
        _check(K, O, d, _args(d, regions, sh, goi, N, out_len), to_rc, f"long[{L},{out_len}]")


def test_packed_onehot_svar2_pure_deletion_anchor(K, O):
    """svar2 rows: pure-DEL anchors come from the ASCII reference inside the packed kernel's slow path."""
    from genvarloader_b200 import synth
    from tests.test_gpu_svar2 import _oracle

    L = 5000
    d = synth.make_dataset(31, 150_000, 3, 12, L, 20.0, max_indel=14, snp_frac=0.4, neg_strand_frac=0.5)
    rng = np.random.default_rng(4)
    r_idx, s_idx = rng.integers(0, d.n_regions, 9), rng.integers(0, d.n_samples, 9)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
    ch = synth.to_svar2_channels(d, regions, ds_idx, dense_frac=0.5, seed=3)
    shifts = np.zeros(goi.shape, np.int32)
    K.pin_static(d.reference, ch["key_alt"])
    try:
        for out_len in (-1, L - 100):
            exp, _ = _oracle(O, d, ch, regions, shifts, out_len, to_rc)
            oh, _ = K.reconstruct_haplotypes_from_svar2(
                regions, shifts, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"],
                ch["dense_range"], ch["dense_present"], ch["dense_present_off"], ch["key_ilen"], ch["key_alt"],
                ch["key_alt_off"], d.reference, d.ref_offsets, N, out_len, to_rc=to_rc, mode="onehot")
            assert _last_kernel(K) == 1
            assert (oh == O.onehot(exp)).all()
    finally:
        K.unpin_static(d.reference, ch["key_alt"])


def test_engine_packs_reference_by_default(cuda_device):
    """the Dataset/Engine path (device pointers) runs the packed kernel for one-hot output."""
    import torch

    from genvarloader_b200 import synth
    from genvarloader_b200._dataset import Dataset
    from genvarloader_b200._ffi import lib
    from oracle import oracle as O

    d = synth.make_dataset(3, 300_000, 4, 10, 4096, 3.0, neg_strand_frac=0.5)
    ds = Dataset.from_synth(cuda_device, d, rng=1).with_tracks(False).with_len(4096).with_encoding("onehot")
    out = ds[:6, :2]
    torch.cuda.synchronize()
    assert int(lib.gvl_debug_last_exec_kernel(ds.engine.ctx.handle)) == 1
    plain = ds.with_encoding("bytes")[:6, :2]
    assert (out.cpu().numpy().reshape(-1, 4) == O.onehot(plain.cpu().numpy().ravel())).all()
