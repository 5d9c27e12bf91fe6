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
