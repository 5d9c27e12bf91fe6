"""svar2 as a DATASET source (reference `Svar2Haps`, python/genvarloader/_dataset/_svar2_haps.py:183-672):
the two channels stay resident in HBM as range tables over all (region, sample, ploid) slots, every read merges them on the
device.  Checked against the oracle's svar2 restatement fed with the per-call flat channels the reference would gather
(`synth.svar2_batch_channels`), and against the SVAR1 form of the same variants (identical bytes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = ord("N")


@pytest.fixture(scope="module")
def env(cuda_device):
    from genvarloader_b200 import synth
    from genvarloader_b200._dataset import Dataset
    from oracle import oracle as O

    d = synth.make_dataset(23, 400_000, 4, 10, 3000 + 2 * 8, 6.0, max_jitter=8, neg_strand_frac=0.5, straddle_ends=False,
                           n_tracks=1, max_indel=20, snp_frac=0.5)
    sv = synth.to_svar2_dataset(d, dense_frac=0.5, seed=4)
    ds2 = Dataset.from_synth(cuda_device, d, rng=9, svar2=sv)
    ds1 = Dataset.from_synth(cuda_device, d, rng=9)
    return d, sv, ds1, ds2, O, synth


def _oracle_haps(O, synth, d, sv, r_idx, s_idx, L, shifts=None):
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
    ch = synth.svar2_batch_channels(sv, ds_idx, d.ploidy, d.n_samples)
    p = d.ploidy
    sh = np.zeros(goi.shape, np.int32) if shifts is None else shifts
    cargs = (ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"], ch["dense_range"], ch["dense_present"],
             ch["dense_present_off"])
    diffs = O.hap_diffs_svar2(regions, p, *cargs, ch["key_ilen"])
    lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    out_len = np.full(goi.shape, L, np.int64) if L >= 0 else lengths[:, None] + diffs
    oo = np.concatenate([[0], np.cumsum(out_len.ravel())]).astype(np.int64)
    out = np.zeros(int(oo[-1]), np.uint8)
    O.reconstruct_haplotypes_from_svar2(out, np.stack([oo[:-1], oo[1:]], 1), regions, sh, *cargs, ch["key_ilen"], ch["key_alt"],
                                        ch["key_alt_off"], d.reference, d.ref_offsets, N)
    O.rc_flat_rows_inplace(out, oo, to_rc)
    return out, oo, diffs, regions, goi, to_rc, ds_idx, ch


def test_fixed_and_ragged_haplotypes_vs_oracle_and_svar1(env):
    d, sv, ds1, ds2, O, synth = env
    r_idx, s_idx = np.array([0, 3, 7, 9, 3]), np.array([1, 0, 3, 2, 2])
    for L in (2048, -1):
        a = ds2.with_tracks(False)
        b = ds1.with_tracks(False)
        if L > 0:
            a, b = a.with_len(L), b.with_len(L)
        got, ref = a[r_idx, s_idx], b[r_idx, s_idx]
        exp, oo, *_ = _oracle_haps(O, synth, d, sv, r_idx, s_idx, L)
        g = got.cpu().numpy().ravel() if L > 0 else got.data.cpu().numpy()
        assert (g == exp).all(), L
        r = ref.cpu().numpy().ravel() if L > 0 else ref.data.cpu().numpy()
        assert (g == r).all()  # the same variants through the SVAR1 CSR give the same bytes
        if L < 0:
            assert (got.offsets.cpu().numpy() == oo).all()
    oh = ds2.with_tracks(False).with_len(1024).with_encoding("onehot")[r_idx, s_idx]
    exp, *_ = _oracle_haps(O, synth, d, sv, r_idx, s_idx, 1024)
    assert (oh.cpu().numpy().reshape(-1, 4) == O.onehot(exp)).all()


def test_random_shifts_and_reference_rows(env):
    d, sv, ds1, ds2, O, synth = env
    L, J = 2500, 8
    dsj = ds2.with_tracks(False).with_len(L).with_settings(jitter=J, deterministic=False, rng=31)
    out = dsj[2:6, 1:3]
    r_idx, s_idx = np.repeat(np.arange(2, 6), 2), np.tile(np.arange(1, 3), 4)
    rng = np.random.default_rng(31)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx, rng, jitter=J)
    ch = synth.svar2_batch_channels(sv, ds_idx, d.ploidy, d.n_samples)
    cargs = (ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"], ch["dense_range"], ch["dense_present"],
             ch["dense_present_off"])
    diffs = O.hap_diffs_svar2(regions, d.ploidy, *cargs, ch["key_ilen"])
    lengths = regions[:, 2] - regions[:, 1]
    max_shift = diffs.clip(min=0) + (lengths - L).clip(min=0)[:, None]
    shifts = rng.integers(0, max_shift + 1, dtype=np.int32)
    oo = (np.arange(goi.size + 1) * L).astype(np.int64)
    exp = np.zeros(goi.size * L, np.uint8)
    O.reconstruct_haplotypes_from_svar2(exp, np.stack([oo[:-1], oo[1:]], 1), regions, shifts, *cargs, ch["key_ilen"],
                                        ch["key_alt"], ch["key_alt_off"], d.reference, d.ref_offsets, N)
    O.rc_flat_rows_inplace(exp, oo, to_rc)
    assert (out.cpu().numpy().ravel() == exp).all()
    ref2 = ds2.with_seqs("reference").with_tracks(False).with_len(900)[[1, 4], [0, 0]]
    ref1 = ds1.with_seqs("reference").with_tracks(False).with_len(900)[[1, 4], [0, 0]]
    assert (ref2 == ref1).all()


def test_tracks_and_loader_on_the_svar2_source(env):
    from genvarloader_b200 import FlankSample

    d, sv, ds1, ds2, O, synth = env
    L = 2200
    a = ds2.with_len(L).with_insertion_fill(FlankSample(5))
    b = ds1.with_len(L).with_insertion_fill(FlankSample(5))
    idx = ([1, 5, 8], [3, 0, 2])
    (h2, t2), (h1, t1) = a[idx], b[idx]
    assert (h2 == h1).all() and (t2.view(__import__("torch").int32) == t1.view(__import__("torch").int32)).all()
    # ragged rows too (general path), and the read-ahead loader over the svar2 source
    (h2r, t2r), (h1r, t1r) = ds2[idx], ds1[idx]
    assert (h2r.data == h1r.data).all() and (t2r.offsets == t1r.offsets).all()
    assert (t2r.data.view(__import__("torch").int32) == t1r.data.view(__import__("torch").int32)).all()
    xs2 = [x for x, _ in a.to_dataloader(batch_size=6, mode="double_buffered", ring=2, copy=True)]
    xs1 = [x for x, _ in b.to_dataloader(batch_size=6, mode="double_buffered", ring=2, copy=True)]
    assert len(xs2) == len(xs1) and all((p == q).all() for p, q in zip(xs2, xs1))
    with pytest.raises(NotImplementedError):
        ds2.with_settings(var_filter="exonic")
