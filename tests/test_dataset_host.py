"""Host logic of the GPU Dataset (no GPU needed): index parsing follows numpy semantics over the
(regions, samples) grid exactly like the reference's DatasetIndexer.parse_idx
(python/genvarloader/_dataset/_indexing.py:208-264), and the with_* validators raise like the reference."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from genvarloader_b200._dataset import Dataset  # noqa: E402
from genvarloader_b200 import FlankSample, Interpolate  # noqa: E402


def _ds(R=7, S=5, L=100, max_jitter=4):
    regions = np.stack([np.zeros(R), np.arange(R) * 1000, np.arange(R) * 1000 + L, np.ones(R)], 1).astype(np.int32)
    return Dataset(engine=None, full_regions=regions, sample_names=tuple(f"s{i}" for i in range(S)), ploidy=2,
                   max_jitter=max_jitter, track_kinds={"t0": "sample", "t1": "annot"})


@pytest.mark.parametrize("idx", [3, (3, 2), (slice(None, 4), slice(1, 3)), slice(2, 5), (slice(None), 0),
                                 ([0, 2, 5], [1, 1, 4]), ([1, 3], slice(None)), (slice(0, 3), [4, 0]),
                                 (np.array([[0, 1], [2, 3]]), np.array([[0, 1], [2, 3]])), ([6], 3),
                                 (np.array([True, False, True, False, False, False, True]), slice(None))])
def test_parse_idx_matches_numpy_grid(idx):
    ds = _ds()
    R, S = ds.full_shape
    grid = np.arange(R * S).reshape(R, S)
    flat, squeeze, out_reshape = ds._parse_idx(idx)
    exp = grid[idx]
    assert flat.tolist() == np.asarray(exp).ravel().tolist()
    assert squeeze == (np.ndim(exp) == 0)
    if out_reshape is not None:
        assert int(np.prod(out_reshape)) == flat.size
    # two slices keep both axes, (n, m)
    if isinstance(idx, tuple) and all(isinstance(i, slice) for i in idx):
        assert out_reshape == exp.shape


def test_subset_to_maps_into_full_grid():
    ds = _ds().subset_to(regions=[5, 1, 2], samples=["s4", "s0"])
    assert ds.shape == (3, 2) and ds.full_shape == (7, 5) and ds.samples == ["s4", "s0"]
    flat, _, _ = ds._parse_idx((slice(None), slice(None)))
    assert flat.tolist() == [5 * 5 + 4, 5 * 5 + 0, 1 * 5 + 4, 1 * 5 + 0, 2 * 5 + 4, 2 * 5 + 0]
    assert ds.to_full_dataset().shape == (7, 5)
    with pytest.raises(KeyError):
        _ds().subset_to(samples=["nope"])


def test_validation_like_reference():
    ds = _ds()
    with pytest.raises(ValueError, match="positive integer"):
        ds.with_len(0)
    with pytest.raises(ValueError, match="maximum output length"):
        ds.with_len(100 + 2 * 4 + 1)
    assert ds.with_len(108).output_length == 108
    with pytest.raises(ValueError, match="maximum jitter"):
        ds.with_settings(jitter=5)
    with pytest.raises(ValueError, match="Effective length"):
        ds.with_len(104).with_settings(jitter=4)
    with pytest.raises(ValueError, match="not found"):
        ds.with_tracks(["missing"])
    with pytest.raises(ValueError, match="window_opt"):
        ds.with_seqs("variant-windows")
    with pytest.raises(ValueError, match="set together"):
        ds.with_settings(token_alphabet="ACGT")
    with pytest.raises(ValueError, match="token LUT"):
        ds.with_settings(flank_length=4)
    assert ds.with_seqs("variants").sequence_type == "variants"
    with pytest.raises(ValueError, match="unphased_union"):
        ds.with_settings(unphased_union=True)  # haplotype output
    assert ds.with_seqs("variants").with_settings(unphased_union=True).unphased_union
    with pytest.raises(NotImplementedError, match="AF"):
        ds.with_settings(min_af=0.1)  # haplotype output (_haps.py:695-698); allowed with with_seqs("variants")
    with pytest.raises(NotImplementedError):
        ds.with_settings(min_af=0.1)
    with pytest.raises(ValueError):
        Interpolate(order=4)
    with pytest.raises(ValueError):
        FlankSample(flank_width=-1)
    assert ds.with_tracks(False).active_tracks == ()
    assert ds.with_tracks("t1").active_tracks == ("t1",)
    # with_insertion_fill mirrors the reference's guards (_impl.py:856-873): tracks active, haplotypes active, realign on
    with pytest.raises(ValueError, match="tracks are active"):
        ds.with_tracks(False).with_insertion_fill(Interpolate(2))
    with pytest.raises(ValueError, match="haplotypes"):
        ds.with_tracks(["t0", "t1"]).with_seqs("reference").with_insertion_fill(Interpolate(2))
    with pytest.raises(ValueError, match="realign_tracks=False"):
        ds.with_tracks(["t0", "t1"]).with_settings(realign_tracks=False).with_insertion_fill(Interpolate(2))
    d2 = ds.with_tracks(["t0", "t1"]).with_insertion_fill(Interpolate(2))
    assert set(d2.insertion_fill) == {"t0", "t1"}
    with pytest.raises(ValueError):
        ds.with_encoding("onehot_cf")  # ragged length
    assert ds.with_len(64).with_encoding("onehot_cf").encoding == "onehot_cf"
    with pytest.raises(FileNotFoundError):
        Dataset.open("/nowhere")


def test_batch_loader_host_logic():
    """BatchLoader: lengths, ordering, drop_last, sampler exclusivity (no GPU: a stub dataset records the indices)."""
    from genvarloader_b200._dataset import BatchLoader

    class Stub:
        n_samples = 3
        _r_idx = np.arange(10, 14)
        _s_idx = np.array([2, 0, 1])

        def __len__(self):
            return 12

        def __getitem__(self, idx):
            return idx

    ds = Stub()
    bl = BatchLoader(ds, 5, False, None, False, None, False, None)
    got = list(bl)
    assert len(bl) == 3 and [len(b[0]) for b in got] == [5, 5, 2]
    assert (np.concatenate([b[0] * 3 + b[1] for b in got]) == np.arange(12)).all()
    assert len(BatchLoader(ds, 5, False, None, True, None, False, None)) == 2
    sh = BatchLoader(ds, 4, True, None, False, 7, True, None)
    a, b = list(sh), list(sh)
    flat = np.sort(np.concatenate([x[0] * 3 + x[1] for x in a]))  # (the stub's batch is the (r, s) tuple; indices are appended)
    assert (flat == np.arange(12)).all()                            # epoch covers everything
    # return_indices appends the (region, sample) indices the dataset was indexed with (_torch.py:293-300)
    assert all(len(x) == 4 and (x[2] == x[0]).all() and (x[3] == x[1]).all() for x in a)
    assert any((x[0] != y[0]).any() for x, y in zip(a, b))          # reshuffled every epoch
    # the transform receives the batch's elements as separate arguments (_torch.py:302-303)
    smp = BatchLoader(ds, 2, False, [3, 1, 0], False, None, False, lambda r, s: ("x", r, s))
    out = list(smp)
    assert len(smp) == 2 and out[0][0] == "x" and (out[0][1] == np.array([1, 0])).all() and (out[0][2] == np.array([0, 1])).all()
    smp = BatchLoader(ds, 2, False, np.array([3, 1, 0]), False, None, True, lambda r, s, ri, si: (ri, si))
    assert (list(smp)[1][0] == np.array([0])).all()
    with pytest.raises(ValueError):
        BatchLoader(ds, 0, False, None, False, None, False, None)
