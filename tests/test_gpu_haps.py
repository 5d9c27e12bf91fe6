"""GPU parity: the CUDA haplotype path (through the C ABI host layer) vs the reference's frozen
goldens, the pyref goldens and the CPU oracle -- bit-exact.  Mirrors the reference's
tests/parity/test_reconstruct_haplotypes_parity.py, test_get_diffs_sparse_parity.py,
test_fused_haps_parity.py."""
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


def test_reconstruct_haplotypes_from_sparse_golden(K):
    """reference: tests/parity/test_reconstruct_haplotypes_parity.py:14-21"""
    cases = _golden.load_golden("reconstruct_haplotypes_from_sparse")
    assert len(cases) == 200
    _golden.replay_inplace(K.reconstruct_haplotypes_from_sparse, "reconstruct_haplotypes_from_sparse", cases,
                           out_factory=lambda inputs: np.zeros(int(np.asarray(inputs[0])[-1]), np.uint8), out_index=0)


def test_pyref_haps_annotated_golden(K):
    cases = _golden.load_golden("pyref_haps")
    for ci, (inputs, (g_out, g_av, g_ap)) in enumerate(cases):
        n = int(inputs[0][-1])
        out, av, ap = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
        K.reconstruct_haplotypes_from_sparse(out, *inputs, av, ap)
        _golden.eq("pyref_haps.out", ci, out, g_out)
        _golden.eq("pyref_haps.annot_v", ci, av, g_av)
        _golden.eq("pyref_haps.annot_pos", ci, ap, g_ap)


def test_get_diffs_sparse_golden(K):
    cases = _golden.load_golden("get_diffs_sparse")
    assert len(cases) == 200
    _golden.replay_tuple(K.get_diffs_sparse, "get_diffs_sparse", cases)


def test_get_reference_golden_via_zero_variant_rows(K):
    """get_reference (src/reference/mod.rs:56-120) equals reconstruction of rows without variants;
    its 200-case golden (arbitrary bytes, pads, both contig ends) is an extra check of pad/ref logic."""
    cases = _golden.load_golden("get_reference")
    for ci, (inputs, golden) in enumerate(cases):
        regions, out_offsets, reference, ref_offsets, pad_char, _parallel = inputs
        n = regions.shape[0]
        out = np.zeros(int(out_offsets[-1]), np.uint8)
        K.reconstruct_haplotypes_from_sparse(
            out, out_offsets, regions, np.zeros((n, 1), np.int32), np.zeros((n, 1), np.int64),
            np.zeros((2, 1), np.int64), np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32),
            np.zeros(1, np.uint8), np.array([0, 1]), reference, ref_offsets, pad_char)
        _golden.eq("get_reference", ci, out, golden)


def _rand_case(rng, d, b, output_length, synth, shifts=False):
    r_idx = rng.integers(0, d.n_regions, b)
    s_idx = rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, _ = synth.batch_args(d, r_idx, s_idx)
    sh = np.zeros((b, d.ploidy), np.int32)
    if shifts:
        sh = rng.integers(0, 40, (b, d.ploidy)).astype(np.int32)
    return (regions, sh, goi, d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
            d.reference, d.ref_offsets, N, output_length), to_rc


@pytest.mark.parametrize("vkb,L,out_len,shifts", [(1.0, 2000, 2000, False), (10.0, 3000, -1, False),
                                                  (5.0, 5000, 4096, True), (0.1, 9000, 9000, False),
                                                  (20.0, 700, -1, False), (3.0, 16384, 16384, True)])
def test_fused_vs_oracle(K, O, vkb, L, out_len, shifts):
    """reconstruct_haplotypes_fused / annotated vs the oracle: fixed + ragged, RC rows, shifts, contig ends."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(int(vkb * 10 + L), 200_000, 4, 24, L, vkb, neg_strand_frac=0.5, max_indel=30)
    rng = np.random.default_rng(L)
    args, to_rc = _rand_case(rng, d, 16, out_len, synth, shifts)
    for rc in (None, to_rc):
        e_out, e_oo = O.reconstruct_haplotypes_fused(*args, None, None, rc)
        g_out, g_oo = K.reconstruct_haplotypes_fused(*args, None, None, rc)
        _golden.eq("fused.offsets", 0, g_oo, e_oo)
        _golden.eq("fused.out", 0, g_out, e_out)
        g_oh, _ = K.reconstruct_haplotypes_fused(*args, None, None, rc, mode="onehot")
        _golden.eq("fused.onehot", 0, g_oh, O.onehot(e_out))
        e = O.reconstruct_annotated_haplotypes_fused(*args, None, None, rc)
        g = K.reconstruct_annotated_haplotypes_fused(*args, None, None, rc)
        for j, nm in enumerate(["out", "annot_v", "annot_pos", "offsets"]):
            _golden.eq(f"annot.{nm}", 0, g[j], e[j])
        if out_len > 0 and out_len % 4 == 0:
            g_cf, _ = K.reconstruct_haplotypes_fused(*args, None, None, rc, mode="onehot_cf")
            exp = O.onehot(e_out).reshape(-1, out_len, 4).transpose(0, 2, 1)
            _golden.eq("fused.onehot_cf", 0, g_cf, np.ascontiguousarray(exp))


def test_fused_exonic_keep_mask(K, O):
    from genvarloader_b200 import synth

    d = synth.make_dataset(7, 100_000, 4, 16, 3000, 8.0, max_indel=40)
    rng = np.random.default_rng(1)
    args, _ = _rand_case(rng, d, 12, -1, synth)
    regions, goi = args[0], args[2]
    keep, ko = O.choose_exonic_variants(regions[:, 1], regions[:, 2], goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens)
    e_out, e_oo = O.reconstruct_haplotypes_fused(*args, keep, ko, None)
    g_out, g_oo = K.reconstruct_haplotypes_fused(*args, keep, ko, None)
    _golden.eq("keep.offsets", 0, g_oo, e_oo)
    _golden.eq("keep.out", 0, g_out, e_out)


def test_dense_variants_many_passes(K, O):
    """More records per tile than one shared-memory pass holds (REC_CAP) and overlapping variants."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(11, 60_000, 2, 8, 20_000, 400.0, snp_frac=0.6, max_indel=3, dense_af=0.5)
    rng = np.random.default_rng(2)
    args, to_rc = _rand_case(rng, d, 6, -1, synth)
    to_rc[::2] = True
    e_out, e_oo = O.reconstruct_haplotypes_fused(*args, None, None, to_rc)
    g_out, g_oo = K.reconstruct_haplotypes_fused(*args, None, None, to_rc)
    _golden.eq("dense.offsets", 0, g_oo, e_oo)
    _golden.eq("dense.out", 0, g_out, e_out)


def test_empty_batch_and_zero_length_rows(K, O):
    from genvarloader_b200 import synth

    d = synth.make_dataset(3, 50_000, 2, 4, 500, 2.0)
    args, _ = _rand_case(np.random.default_rng(0), d, 0, 100, synth)
    out, oo = K.reconstruct_haplotypes_fused(*args)
    assert out.size == 0 and oo.tolist() == [0]
    # rows whose ragged length collapses to 0 (region of length 0)
    args, _ = _rand_case(np.random.default_rng(0), d, 3, -1, synth)
    args[0][1, 2] = args[0][1, 1]
    e_out, e_oo = O.reconstruct_haplotypes_fused(*args)
    g_out, g_oo = K.reconstruct_haplotypes_fused(*args)
    _golden.eq("zero.offsets", 0, g_oo, e_oo)
    _golden.eq("zero.out", 0, g_out, e_out)


def test_config1_full_size_properties(K, O):
    """BASELINE config 1 at full size: 1,000 regions x 8 samples x 16,384 bp.  A slice is checked
    against the oracle byte for byte; the whole set through size-independent properties:
    one-hot rows sum to (base in ACGT), RC(RC(x)) == x, fixed offsets are k*L."""
    from genvarloader_b200 import synth

    d = synth.cfg1()
    rng = np.random.default_rng(5)
    L = 16_384
    r_idx = np.repeat(np.arange(d.n_regions), d.n_samples)
    s_idx = np.tile(np.arange(d.n_samples), d.n_regions)
    regions, goi, _, _ = synth.batch_args(d, r_idx, s_idx)
    sh = np.zeros(goi.shape, np.int32)
    K.pin_static(d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.reference, d.ref_offsets)
    a = (regions, sh, goi, d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
         d.reference, d.ref_offsets, N, L)
    out, oo = K.reconstruct_haplotypes_fused(*a)
    assert (oo == np.arange(goi.size + 1) * L).all()
    sl = slice(0, 64 * L)
    e_out, _ = O.reconstruct_haplotypes_fused(regions[:32], sh[:32], goi[:32], *a[3:])
    _golden.eq("cfg1.slice", 0, out[sl], e_out)
    tail, _ = O.reconstruct_haplotypes_fused(regions[-32:], sh[-32:], goi[-32:], *a[3:])
    _golden.eq("cfg1.tail", 0, out[-64 * L:], tail)
    oh, _ = K.reconstruct_haplotypes_fused(*a, mode="onehot")
    acgt = np.isin(out, np.frombuffer(b"ACGT", np.uint8))
    assert (oh.sum(1) == acgt).all()
    assert (oh.argmax(1)[acgt] == np.searchsorted(np.frombuffer(b"ACGT", np.uint8), out[acgt])).all()
    rc_all = np.ones(goi.size, np.bool_)
    rc_out, _ = K.reconstruct_haplotypes_fused(*a, None, None, rc_all)
    O.rc_flat_rows_inplace(rc_out, oo, rc_all)
    _golden.eq("cfg1.rc_roundtrip", 0, rc_out, out)
    K.unpin_static(d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.reference, d.ref_offsets)


def test_unsorted_variant_lists_fall_back_to_exact_serial_plan(K, O):
    """Position-unsorted genotype lists are outside the reference writers' contract but legal kernel
    inputs: the scan-based plan must detect them and reproduce the sequential semantics exactly."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(21, 80_000, 3, 10, 2500, 12.0, max_indel=15, snp_frac=0.5)
    rng = np.random.default_rng(9)
    gv = d.geno_v_idxs.copy()
    for s, e in zip(d.geno_offsets[0], d.geno_offsets[1]):
        if e - s > 2 and rng.random() < 0.5:
            gv[s:e] = rng.permutation(gv[s:e])
    args, to_rc = _rand_case(rng, d, 14, -1, synth, shifts=False)
    args = list(args)
    args[4] = gv
    for out_len, sh in ((-1, False), (2000, True)):
        args[12] = out_len
        args[1] = rng.integers(0, 50, args[1].shape).astype(np.int32) if sh else np.zeros_like(args[1])
        e = O.reconstruct_annotated_haplotypes_fused(*args, None, None, to_rc)
        g = K.reconstruct_annotated_haplotypes_fused(*args, None, None, to_rc)
        for j, nm in enumerate(["out", "annot_v", "annot_pos", "offsets"]):
            _golden.eq(f"unsorted.{nm}", out_len, g[j], e[j])


@pytest.mark.parametrize("vkb,L", [(0.5, 40_000), (40.0, 30_000), (150.0, 9_000)])
def test_long_rows_many_chunks_with_shifts(K, O, vkb, L):
    """Rows with far more variants than one plan chunk (256), shifts that skip many variants,
    overlapping variants (dense_af) and windows running over the contig end."""
    from genvarloader_b200 import synth

    d = synth.make_dataset(int(L + vkb), 150_000, 2, 12, L, vkb, max_indel=12, snp_frac=0.4, dense_af=0.5,
                           neg_strand_frac=0.3)
    rng = np.random.default_rng(int(vkb))
    for out_len in (-1, L - 1000, L + 64):
        args, to_rc = _rand_case(rng, d, 8, out_len, synth)
        args = list(args)
        if out_len > 0:
            args[1] = rng.integers(0, 900, args[1].shape).astype(np.int32)
        e = O.reconstruct_annotated_haplotypes_fused(*args, None, None, to_rc)
        g = K.reconstruct_annotated_haplotypes_fused(*args, None, None, to_rc)
        for j, nm in enumerate(["out", "annot_v", "annot_pos", "offsets"]):
            _golden.eq(f"long.{nm}", out_len, g[j], e[j])
