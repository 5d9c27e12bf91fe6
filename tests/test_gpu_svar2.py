"""GPU parity for the svar2 two-channel source (device merge + the shared plan/execute kernels) against the
oracle's restatement of merge_hap / hap_diffs_svar2 / reconstruct_haplotypes_from_svar2
(src/svar2/mod.rs:45-146, src/reconstruct/mod.rs:620-826) at the decoded-key level."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = ord("N")


def _oracle(O, d, ch, regions, shifts, out_len, to_rc):
    p = shifts.shape[1]
    if out_len < 0:
        diffs = O.hap_diffs_svar2(regions, p, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"],
                                  ch["dense_range"], ch["dense_present"], ch["dense_present_off"], ch["key_ilen"])
        lens = np.maximum((regions[:, 2] - regions[:, 1])[:, None] + diffs, 0).ravel()
    else:
        lens = np.full(shifts.size, out_len)
    oo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = np.zeros(int(oo[-1]), np.uint8)
    O.reconstruct_haplotypes_from_svar2(out, np.stack([oo[:-1], oo[1:]], 1), regions, shifts, ch["vk_pos"], ch["vk_key"],
                                        ch["vk_off"], ch["dense_pos"], ch["dense_key"], ch["dense_range"],
                                        ch["dense_present"], ch["dense_present_off"], ch["key_ilen"], ch["key_alt"],
                                        ch["key_alt_off"], d.reference, d.ref_offsets, N)
    if to_rc is not None:
        O.rc_flat_rows_inplace(out, oo, to_rc)
    return out, oo


@pytest.mark.parametrize("vkb,L,dense_frac", [(4.0, 3000, 0.5), (30.0, 5000, 0.9), (1.0, 20000, 0.1), (60.0, 1500, 0.0),
                                              (15.0, 2500, 1.0)])
def test_svar2_source_vs_oracle(cuda_device, vkb, L, dense_frac):
    from genvarloader_b200 import _kernels as K
    from genvarloader_b200 import synth
    from oracle import oracle as O

    d = synth.make_dataset(int(L + vkb), 150_000, 3, 12, L, vkb, max_indel=14, snp_frac=0.5, neg_strand_frac=0.5)
    rng = np.random.default_rng(int(L))
    r_idx, s_idx = rng.integers(0, d.n_regions, 9), rng.integers(0, d.n_samples, 9)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
    ch = synth.to_svar2_channels(d, regions, ds_idx, dense_frac=dense_frac, seed=3)
    cargs = (regions, d.ploidy, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"], ch["dense_range"],
             ch["dense_present"], ch["dense_present_off"], ch["key_ilen"])
    e_diffs, g_diffs = O.hap_diffs_svar2(*cargs), K.hap_diffs_svar2(*cargs)  # src/svar2/mod.rs:73-146
    assert g_diffs.dtype == e_diffs.dtype and g_diffs.shape == e_diffs.shape and (g_diffs == e_diffs).all()
    for out_len, use_shift in ((-1, False), (L - 500, True), (L + 40, False)):
        shifts = rng.integers(0, 60, goi.shape).astype(np.int32) if use_shift else np.zeros(goi.shape, np.int32)
        for rc in (None, to_rc):
            exp, eoo = _oracle(O, d, ch, regions, shifts, out_len, rc)
            got, goo = K.reconstruct_haplotypes_from_svar2(
                regions, shifts, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"],
                ch["dense_range"], ch["dense_present"], ch["dense_present_off"], ch["key_ilen"], ch["key_alt"],
                ch["key_alt_off"], d.reference, d.ref_offsets, N, out_len, to_rc=rc)
            assert (goo == eoo).all()
            assert (got == exp).all()
        oh, _ = K.reconstruct_haplotypes_from_svar2(
            regions, shifts, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"], ch["dense_range"],
            ch["dense_present"], ch["dense_present_off"], ch["key_ilen"], ch["key_alt"], ch["key_alt_off"], d.reference,
            d.ref_offsets, N, out_len, to_rc=to_rc, mode="onehot")
        assert (oh == O.onehot(exp)).all()


def test_svar2_tie_order_known_answer(cuda_device):
    """src/svar2/mod.rs:616-649: on equal positions the var_key entry precedes the dense entry."""
    from genvarloader_b200 import _kernels as K

    ref = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", np.uint8)
    key_ilen = np.array([0, 0, -2, 2, 0], np.int32)
    key_alt = np.frombuffer(b"TGCAAA", np.uint8)
    key_alt_off = np.array([0, 1, 2, 2, 5, 6], np.int64)
    out, oo = K.reconstruct_haplotypes_from_svar2(
        np.array([[0, 0, 40]], np.int32), np.zeros((1, 2), np.int32), np.array([10, 20, 5], np.int32),
        np.array([0, 1, 3], np.int32), np.array([0, 2, 3]), np.array([15, 20, 30], np.int32), np.array([2, 4, 0], np.int32),
        np.array([[0, 3]], np.int32), np.array([0b010111], np.uint8), np.array([0, 3, 6]), key_ilen, key_alt, key_alt_off,
        ref, np.array([0, 40]), N, -1)
    assert oo.tolist() == [0, 38, 80]
    h0 = out[:38].tobytes()
    assert h0[10] == ord("T") and h0[18] == ord("G")  # SNP T@10; after the 2-bp deletion the vk SNP 'G' wins at pos 20
    assert h0[15] == ref[15]                          # pure-DEL anchor comes from the reference
    assert out[38:].tobytes()[5:8] == b"CAA"          # hap1: insertion at 5


@pytest.mark.parametrize("strategy,param", [(0, 0.0), (1, 0.0), (2, 3.5), (3, 4.0), (4, 1.0), (4, 3.0)])
def test_svar2_tracks_vs_oracle(cuda_device, strategy, param):
    """shift_and_realign_tracks_from_svar2 (src/tracks/mod.rs:705-856): dense source windows + the svar2 variant
    source, all five insertion fills, shifts, the global-query seed of FlankSample -- bit-exact vs the oracle."""
    from genvarloader_b200 import _kernels as K
    from genvarloader_b200 import synth
    from oracle import oracle as O

    L = 4000
    d = synth.make_dataset(91, 150_000, 3, 12, L, 12.0, max_indel=14, snp_frac=0.5)
    rng = np.random.default_rng(strategy * 7 + 1)
    b = 7
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, _, ds_idx = synth.batch_args(d, r_idx, s_idx)
    ch = synth.to_svar2_channels(d, regions, ds_idx, dense_frac=0.5, seed=5)
    p = goi.shape[1]
    diffs = O.hap_diffs_svar2(regions, p, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"],
                              ch["dense_range"], ch["dense_present"], ch["dense_present_off"], ch["key_ilen"])
    lens = np.maximum((regions[:, 2] - regions[:, 1])[:, None] + diffs, 0).ravel()
    oo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    # source windows: the region, room for net deletions (_reconstruct.py:191) and for the largest shift used below
    tlen = (regions[:, 2] - regions[:, 1]) - np.minimum(diffs.min(1), 0) + 40
    to = np.concatenate([[0], np.cumsum(tlen)]).astype(np.int64)
    tracks = rng.standard_normal(int(to[-1])).astype(np.float32)
    qseed = rng.integers(0, 1000, b).astype(np.int64)
    for shifts in (np.zeros(goi.shape, np.int32), rng.integers(0, 30, goi.shape).astype(np.int32)):
        for qs in (None, qseed):
            args = (oo, regions, shifts, ch["vk_pos"], ch["vk_key"], ch["vk_off"], ch["dense_pos"], ch["dense_key"],
                    ch["dense_range"], ch["dense_present"], ch["dense_present_off"], ch["key_ilen"], tracks, to,
                    np.array([param]), strategy, 1234, qs)
            exp = np.zeros(int(oo[-1]), np.float32)
            got = np.full(int(oo[-1]), 7.0, np.float32)
            O.shift_and_realign_tracks_from_svar2(exp, *args)
            K.shift_and_realign_tracks_from_svar2(got, *args)
            assert (exp.view(np.uint32) == got.view(np.uint32)).all()
