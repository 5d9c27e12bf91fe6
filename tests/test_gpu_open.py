"""`Dataset.open` over the on-disk GenVarLoader layout vs the same data as an in-memory dataset: identical
haplotypes / annotations / one-hot / tracks, regions addressed in input-BED order (r_idx_map)."""
import numpy as np
import pytest

from tests._gvl_disk import write_fasta, write_gvl_dataset

pytestmark = pytest.mark.gpu


def test_open_matches_in_memory_dataset(cuda_device, tmp_path):
    from genvarloader_b200 import Dataset, synth

    d = synth.make_dataset(23, 150_000, 3, 10, 2000 + 2 * 8, 6.0, max_jitter=8, neg_strand_frac=0.5, straddle_ends=False,
                           n_tracks=2, max_indel=9)
    order = np.random.default_rng(1).permutation(d.n_regions)
    write_gvl_dataset(tmp_path / "ds", d, ["chr1"], ["a", "b", "c"], order)
    write_fasta(tmp_path / "ref.fa", d.reference, d.ref_offsets, ["chr1"], lower_every=53)
    mem = Dataset.from_synth(cuda_device, d, rng=5)
    dsk = Dataset.open(tmp_path / "ds", tmp_path / "ref.fa", device=cuda_device, rng=5)
    assert dsk.shape == mem.shape and dsk.samples == ["a", "b", "c"] and dsk.max_jitter == 8
    assert (dsk.regions == d.regions[order]).all()
    L = 1536
    for cfg in (dict(seqs="haplotypes", enc="bytes"), dict(seqs="haplotypes", enc="onehot"), dict(seqs="annotated", enc="bytes")):
        a = mem.with_len(L).with_seqs(cfg["seqs"]).with_encoding(cfg["enc"]).with_tracks(False)
        b = dsk.with_len(L).with_seqs(cfg["seqs"]).with_encoding(cfg["enc"]).with_tracks(False)
        got = b[:4, :2]
        exp = a[order[:4], :2]  # input row i of the opened dataset is storage region order[i]
        if cfg["seqs"] == "annotated":
            for f in ("haps", "var_idxs", "ref_coords"):
                assert (getattr(got, f) == getattr(exp, f)).all()
        else:
            assert (got == exp).all()
    # ragged haplotypes + both tracks, fancy indices
    rg = dsk.with_seqs("haplotypes")[[0, 7, 3], [2, 0, 1]]
    re_ = mem.with_seqs("haplotypes")[order[[0, 7, 3]], [2, 0, 1]]
    assert (rg[0].data == re_[0].data).all() and (rg[0].offsets == re_[0].offsets).all()
    assert (rg[1].data.view(dtype=rg[1].data.dtype) == re_[1].data).all()
    # subset by input-order region indices
    sub = dsk.subset_to(regions=[2, 5], samples=["c"]).with_len(L).with_tracks(False)
    assert (sub[:, :] == mem.with_len(L).with_tracks(False)[order[[2, 5]], [2]].reshape(sub[:, :].shape)).all()
