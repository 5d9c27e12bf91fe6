"""Randomised parity sweep: many small random datasets (window lengths 1..6,000, 0.5..300 variants/kb, long indels,
windows over both contig ends, shifts, reverse-complement, odd fixed lengths, ragged rows) through every output mode of
the haplotype path -- byte kernel, packed one-hot kernel, annotated -- and one realigned track with a random fill,
bit-for-bit against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = ord("N")


@pytest.mark.parametrize("seed", range(150))
def test_random_case(cuda_device, seed):
    from genvarloader_b200 import _kernels as K
    from genvarloader_b200 import synth
    from oracle import oracle as O

    rng = np.random.default_rng(1000 + seed)
    L = int(rng.choice([1, 7, 8, 9, 31, 255, 256, 257, 1000, 2048, 4097, 6000]))
    vkb = float(rng.choice([0.5, 3.0, 20.0, 100.0, 300.0]))
    contig = int(rng.choice([max(4 * L, 600), 50_000]))
    max_indel = int(rng.choice([1, 5, 40, 200]))
    d = synth.make_dataset(seed, contig, int(rng.integers(1, 4)), int(rng.integers(1, 9)), L, vkb, neg_strand_frac=0.5,
                           straddle_ends=bool(rng.integers(0, 2)), max_indel=max_indel, snp_frac=float(rng.choice([0.2, 0.8])),
                           n_tracks=1)
    b = int(rng.integers(1, 9))
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
    if rng.integers(0, 2):
        regions[:, 1] += rng.integers(-L, L + 1, b).astype(np.int32)  # also far outside the stored window / contig
        regions[:, 2] = regions[:, 1] + L
    out_len = int(rng.choice([-1, L, max(1, L - 3), L + 5]))
    sh = rng.integers(0, 12, goi.shape).astype(np.int32) if out_len > 0 and rng.integers(0, 2) else np.zeros(goi.shape, np.int32)
    pad = int(rng.choice([N, ord("A"), ord("T"), 0]))
    rc = [None, to_rc, np.ones_like(to_rc)][int(rng.integers(0, 3))]
    a = (regions, sh, goi, d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.reference,
         d.ref_offsets, pad, out_len, None, None, rc)
    e_out, e_oo = O.reconstruct_haplotypes_fused(*a)
    g_out, g_oo = K.reconstruct_haplotypes_fused(*a)
    assert (g_oo == e_oo).all() and (g_out == e_out).all()
    exp_oh = O.onehot(e_out)
    assert (K.reconstruct_haplotypes_fused(*a, mode="onehot")[0] == exp_oh).all()
    K.pin_static(d.reference, d.alt_alleles)
    try:
        assert (K.reconstruct_haplotypes_fused(*a, mode="onehot")[0] == exp_oh).all()
    finally:
        K.unpin_static(d.reference, d.alt_alleles)
    e = O.reconstruct_annotated_haplotypes_fused(*a)
    g = K.reconstruct_annotated_haplotypes_fused(*a)
    assert all((x == y).all() for x, y in zip(g, e))
    # one realigned track over the same rows (ragged sizing from the diffs, like the reference's HapsTracks)
    diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
    lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    oo = np.concatenate([[0], np.cumsum(np.maximum(lengths[:, None] + diffs, 0).ravel())]).astype(np.int64)
    to = np.concatenate([[0], np.cumsum(lengths - diffs.clip(max=0).min(1))]).astype(np.int64)
    s, en, v, io = d.tracks["track0"]
    strat = int(rng.integers(0, 5))
    par = {0: 0.0, 1: 0.0, 2: 2.5, 3: 3.0, 4: float(rng.integers(1, 4))}[strat]
    ta = (oo, regions, np.zeros(goi.shape, np.int32), goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, ds_idx, s, en, v,
          io, to, np.array([par]), strat, 99 + seed, None, None, rc)
    et = np.zeros(int(oo[-1]), np.float32)
    gt = np.ones(int(oo[-1]), np.float32)
    O.intervals_and_realign_track_fused(et, *ta)
    K.intervals_and_realign_track_fused(gt, *ta)
    assert (et.view(np.uint32) == gt.view(np.uint32)).all()
