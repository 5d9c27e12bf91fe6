"""GPU parity for the track path: paint (src/intervals.rs), realign (src/tracks/mod.rs), the fused
per-track entry (src/ffi/mod.rs:2553-2672) and the PRNG -- bit-exact vs goldens and oracle."""
import numpy as np
import pytest

from tests import _golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(cuda_device):
    from genvarloader_b200 import _kernels

    return _kernels


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    return oracle


def test_intervals_to_tracks_golden(K):
    # reference: tests/parity/test_intervals_to_tracks_parity.py (out inserted at index 6)
    cases = _golden.load_golden("intervals_to_tracks")
    assert len(cases) == 200
    _golden.replay_inplace(K.intervals_to_tracks, "intervals_to_tracks", cases,
                           out_factory=lambda inputs: np.full(int(np.asarray(inputs[-1])[-1]), 7.0, np.float32),
                           out_index=6)


def test_shift_and_realign_tracks_sparse_golden(K):
    # reference: tests/parity/test_shift_and_realign_tracks_parity.py
    cases = _golden.load_golden("shift_and_realign_tracks_sparse")
    assert len(cases) == 200
    _golden.replay_inplace(K.shift_and_realign_tracks_sparse, "shift_and_realign_tracks_sparse", cases,
                           out_factory=lambda inputs: np.full(int(np.asarray(inputs[0])[-1]), 7.0, np.float32),
                           out_index=0)


def test_pyref_tracks_golden(K):
    cases = _golden.load_golden("pyref_tracks")
    for ci, (inputs, g_out) in enumerate(cases):
        out = np.full(int(inputs[0][-1]), 7.0, np.float32)
        args = list(inputs)
        args[13] = int(args[13])
        K.shift_and_realign_tracks_sparse(out, *args)
        _golden.eq("pyref_tracks", ci, out, g_out)


def test_prng_goldens(K):
    # reference: tests/parity/test_prng_parity.py -- only a sample goes through the device
    for ci, (inputs, golden) in enumerate(_golden.load_golden("prng_xorshift64")[::6]):
        assert K._debug_xorshift64(int(inputs[0])) == int(golden), ci
    for ci, (inputs, golden) in enumerate(_golden.load_golden("prng_hash4")[::4]):
        assert K._debug_hash4(*(int(x) for x in inputs)) == int(golden), ci
    assert K._debug_hash4(1, 2, 3, 4) == 11_323_120_931_611_735_037


def _fused_args(d, synth, rng, b, out_len, name, jitter=0, shifts=False):
    r_idx = rng.integers(0, d.n_regions, b)
    s_idx = rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx, rng, jitter)
    return regions, goi, to_rc, ds_idx


@pytest.mark.parametrize("strategy,param", [(0, 0.0), (1, 0.0), (2, 1.5), (2, float("nan")), (3, 5.0), (4, 1.0),
                                            (4, 2.0), (4, 3.0)])
@pytest.mark.parametrize("fixed", [True, False])
def test_fused_track_vs_oracle(K, O, strategy, param, fixed):
    """HapsTracks-style call (python/genvarloader/_dataset/_reconstruct.py:168-290): diffs -> lengths ->
    track windows -> fused paint+realign, with jitter and negative strands."""
    from genvarloader_b200 import synth

    L = 6000
    d = synth.make_dataset(31 + strategy, 300_000, 3, 12, L + 64, 6.0, max_jitter=32, neg_strand_frac=0.5,
                           straddle_ends=False, n_tracks=1, max_indel=25, snp_frac=0.5)
    rng = np.random.default_rng(strategy * 7 + int(fixed))
    regions, goi, to_rc, ds_idx = _fused_args(d, synth, rng, 10, None, "track0", jitter=32)
    diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
    lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    out_len = np.full(goi.shape, L - 100, np.int64) if fixed else lengths[:, None] + diffs
    out_offsets = np.concatenate([[0], np.cumsum(out_len.ravel())]).astype(np.int64)
    track_lengths = lengths - diffs.clip(max=0).min(1)
    track_offsets = np.concatenate([[0], np.cumsum(track_lengths)]).astype(np.int64)
    shifts = np.zeros(goi.shape, np.int32)
    if fixed:
        shifts = rng.integers(0, 60, goi.shape).astype(np.int32)
    s, e, v, io = d.tracks["track0"]
    params = np.array([param], np.float64)
    seed = int(rng.integers(0, 2**63)) * 2 + 1
    a = (out_offsets, regions, shifts, goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, ds_idx, s, e, v, io,
         track_offsets, params, strategy, seed, None, None, to_rc)
    exp = np.zeros(int(out_offsets[-1]), np.float32)
    O.intervals_and_realign_track_fused(exp, *a)
    got = np.full(int(out_offsets[-1]), 7.0, np.float32)
    K.intervals_and_realign_track_fused(got, *a)
    _golden.eq(f"fused_track[{strategy},{param},{fixed}]", 0, got, exp)


def _overlapping_track(rng, d, n_slots, dense=False):
    """Interval SoA with overlaps, identical starts, containment and empty intervals (legal data for the reference: it
    paints in stored order, later intervals win, src/intervals.rs:64-85); sorted by start inside every slot."""
    ss, ee, vv, off = [], [], [], [0]
    for k in range(n_slots):
        r = k // d.n_samples
        lo, hi = int(d.regions[r, 1]) - 100, int(d.regions[r, 2]) + 100
        n = int(rng.integers(0, 600 if dense else 60))
        st = np.sort(rng.integers(lo, hi, n)).astype(np.int32)
        ln = rng.integers(0, 40 if dense else 400, n).astype(np.int32)
        ss.append(st), ee.append(st + ln), vv.append(rng.normal(size=n).astype(np.float32))
        off.append(off[-1] + n)
    return np.concatenate(ss), np.concatenate(ee), np.concatenate(vv), np.array(off, np.int64)


@pytest.mark.parametrize("dense", [False, True])
def test_overlapping_intervals_last_write_wins(K, O, cuda_device, dense):
    from genvarloader_b200 import FlankSample, synth
    from genvarloader_b200._dataset import Dataset

    L = 3000
    d = synth.make_dataset(5, 200_000, 3, 8, L + 40, 5.0, max_jitter=0, neg_strand_frac=0.5, straddle_ends=False, n_tracks=0,
                           max_indel=30, snp_frac=0.4)
    rng = np.random.default_rng(11 + int(dense))
    s, e, v, io = _overlapping_track(rng, d, d.n_regions * d.n_samples, dense)
    # host entry: intervals_to_tracks paints the touched slots through a private flattened copy
    q = rng.integers(0, d.n_regions * d.n_samples, 12)
    starts = d.regions[q // d.n_samples, 1].astype(np.int32) + rng.integers(-50, 50, 12).astype(np.int32)
    oo = np.concatenate([[0], np.cumsum(rng.integers(1, L, 12))]).astype(np.int64)
    exp, got = np.zeros(int(oo[-1]), np.float32), np.full(int(oo[-1]), 7.0, np.float32)
    O.intervals_to_tracks(q, starts, s, e, v, io, exp, oo)
    K.intervals_to_tracks(q, starts, s, e, v, io, got, oo)
    _golden.eq("overlap.paint", 0, got, exp)
    # fused realign entry + Dataset (Engine.add_track flattens at upload)
    r_idx, s_idx = rng.integers(0, d.n_regions, 6), rng.integers(0, d.n_samples, 6)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
    diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
    lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    out_offsets = (np.arange(goi.size + 1) * L).astype(np.int64)
    track_offsets = np.concatenate([[0], np.cumsum(lengths - diffs.clip(max=0).min(1))]).astype(np.int64)
    seed = int(np.bitwise_xor.reduce(ds_idx.astype(np.uint64)))
    a = (out_offsets, regions, np.zeros(goi.shape, np.int32), goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, ds_idx,
         s, e, v, io, track_offsets, np.array([7.0]), 3, seed, None, None, to_rc)
    exp = np.zeros(int(out_offsets[-1]), np.float32)
    O.intervals_and_realign_track_fused(exp, *a)
    got = np.full(int(out_offsets[-1]), 7.0, np.float32)
    K.intervals_and_realign_track_fused(got, *a)
    _golden.eq("overlap.fused", 0, got, exp)
    d.tracks["ov"] = (s, e, v, io)
    ds = Dataset.from_synth(cuda_device, d).with_len(L).with_insertion_fill(FlankSample(7))
    _, trk = ds[r_idx, s_idx]
    _golden.eq("overlap.dataset", 0, trk.cpu().numpy().reshape(-1), exp)


def test_one_bp_runs_and_long_deletions(K, O):
    """Intervals of 1-3 bp (more stored intervals per pass than the kernel stages at once) under deletions long enough to
    skip whole staged slices; every fill; ragged and fixed rows."""
    from genvarloader_b200 import synth

    L = 20_000
    d = synth.make_dataset(9, 400_000, 2, 6, L + 64, 2.0, max_jitter=0, neg_strand_frac=0.5, straddle_ends=False, n_tracks=0,
                           max_indel=900, snp_frac=0.3)
    rng = np.random.default_rng(3)
    ss, ee, vv, off = [], [], [], [0]
    for k in range(d.n_regions * d.n_samples):
        r = k // d.n_samples
        lo, hi = int(d.regions[r, 1]) - 10, int(d.regions[r, 2]) + 1000
        cuts = np.unique(rng.integers(lo, hi, (hi - lo) // 2))
        keep = rng.random(cuts.size - 1) > 0.2
        ss.append(cuts[:-1][keep]), ee.append(cuts[1:][keep]), vv.append(rng.normal(size=int(keep.sum())).astype(np.float32))
        off.append(off[-1] + int(keep.sum()))
    s, e, v = np.concatenate(ss).astype(np.int32), np.concatenate(ee).astype(np.int32), np.concatenate(vv)
    io = np.array(off, np.int64)
    for fixed, (strategy, param) in [(True, (0, 0.0)), (False, (4, 2.0)), (True, (3, 9.0)), (False, (1, 0.0))]:
        r_idx, s_idx = rng.integers(0, d.n_regions, 5), rng.integers(0, d.n_samples, 5)
        regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
        diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
        lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
        out_len = np.full(goi.shape, L - 37, np.int64) if fixed else lengths[:, None] + diffs
        out_offsets = np.concatenate([[0], np.cumsum(out_len.ravel())]).astype(np.int64)
        track_offsets = np.concatenate([[0], np.cumsum(lengths - diffs.clip(max=0).min(1))]).astype(np.int64)
        shifts = rng.integers(0, 40, goi.shape).astype(np.int32) if fixed else np.zeros(goi.shape, np.int32)
        a = (out_offsets, regions, shifts, goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, ds_idx, s, e, v, io,
             track_offsets, np.array([param]), strategy, 77, None, None, to_rc)
        exp = np.zeros(int(out_offsets[-1]), np.float32)
        O.intervals_and_realign_track_fused(exp, *a)
        got = np.full(int(out_offsets[-1]), 7.0, np.float32)
        K.intervals_and_realign_track_fused(got, *a)
        _golden.eq(f"one_bp[{fixed},{strategy}]", 0, got, exp)
