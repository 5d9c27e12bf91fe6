"""Pin the CPU oracle (oracle/gvl_oracle.c) against the reference's own frozen goldens.

Each test mirrors the reference test of the same kernel (tests/parity/test_*_parity.py):
same golden file, same replay helper, same out_factory / out_index.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import _golden


def test_reconstruct_haplotypes_from_sparse_golden():
    # reference: tests/parity/test_reconstruct_haplotypes_parity.py:14-21
    cases = _golden.load_golden("reconstruct_haplotypes_from_sparse")
    assert len(cases) == 200
    _golden.replay_inplace(O.reconstruct_haplotypes_from_sparse, "reconstruct_haplotypes_from_sparse", cases,
                           out_factory=lambda inputs: np.zeros(int(np.asarray(inputs[0])[-1]), np.uint8),
                           out_index=0)


def test_reconstruct_golden_parallel_equals_serial():
    # reference: tests/parity/test_rayon_equivalence.py:31-58
    O.set_threads(4)
    cases = _golden.load_golden("reconstruct_haplotypes_from_sparse")
    for ci, (inputs, golden) in enumerate(cases):
        out = np.zeros(int(np.asarray(inputs[0])[-1]), np.uint8)
        O.reconstruct_haplotypes_from_sparse(out, *inputs, parallel=True)
        _golden.eq("recon-par", ci, out, golden)
    O.set_threads(1)


def test_shift_and_realign_tracks_sparse_golden():
    # reference: tests/parity/test_shift_and_realign_tracks_parity.py
    cases = _golden.load_golden("shift_and_realign_tracks_sparse")
    assert len(cases) == 200
    strategies = {int(c[0][13]) for c in cases}
    assert strategies == {0, 1, 2, 3, 4}
    _golden.replay_inplace(O.shift_and_realign_tracks_sparse, "shift_and_realign_tracks_sparse", cases,
                           out_factory=lambda inputs: np.zeros(int(np.asarray(inputs[0])[-1]), np.float32),
                           out_index=0)


def test_get_diffs_sparse_golden():
    cases = _golden.load_golden("get_diffs_sparse")
    assert len(cases) == 200
    _golden.replay_tuple(O.get_diffs_sparse, "get_diffs_sparse", cases)


def test_get_reference_golden():
    cases = _golden.load_golden("get_reference")
    assert len(cases) == 200
    _golden.replay_return(O.get_reference, "get_reference", cases)


def test_intervals_to_tracks_golden():
    # reference: tests/parity/test_intervals_to_tracks_parity.py (out inserted at index 6)
    cases = _golden.load_golden("intervals_to_tracks")
    assert len(cases) == 200
    _golden.replay_inplace(O.intervals_to_tracks, "intervals_to_tracks", cases,
                           out_factory=lambda inputs: np.zeros(int(np.asarray(inputs[-1])[-1]), np.float32),
                           out_index=6)


def test_choose_exonic_variants_golden():
    cases = _golden.load_golden("choose_exonic_variants")
    assert len(cases) == 200
    _golden.replay_tuple(O.choose_exonic_variants, "choose_exonic_variants", cases)


def test_prng_goldens():
    # reference: tests/parity/test_prng_parity.py:30-56
    for ci, (inputs, golden) in enumerate(_golden.load_golden("prng_xorshift64")):
        assert O._debug_xorshift64(int(inputs[0])) == int(golden), ci
    for ci, (inputs, golden) in enumerate(_golden.load_golden("prng_hash4")):
        assert O._debug_hash4(*(int(x) for x in inputs)) == int(golden), ci


def test_prng_known_vectors():
    # reference: tests/parity/test_prng_parity.py:62-77
    assert O._debug_xorshift64(1) == 1_082_269_761
    assert O._debug_xorshift64(2) == 2_164_539_522
    assert O._debug_xorshift64(42) == 45_454_805_674
    assert O._debug_xorshift64(0xDEADBEEF) == 4_018_790_486_776_397_394
    assert O._debug_xorshift64(2**64 - 1) == 1_065_361_344
    assert O._debug_hash4(1, 2, 3, 4) == 11_323_120_931_611_735_037
    assert O._debug_hash4(0, 0, 0, 0) == 0
    assert O._debug_hash4(0xDEADBEEF, 0xCAFE, 0xBABE, 1) == 5_244_362_157_944_750_963


def test_pyref_haps_annotated():
    """Oracle == the reference's own pure-Python reconstruct_haplotype_from_sparse
    (python/genvarloader/_dataset/_genotypes.py:125-248) on 150 richer cases incl. annotations,
    negative starts, contig overshoot, shifts, keep masks (fixtures: tests/golden/make_pyref_golden.py)."""
    cases = _golden.load_golden("pyref_haps")
    assert len(cases) == 150
    for ci, (inputs, (g_out, g_av, g_ap)) in enumerate(cases):
        n = int(inputs[0][-1])
        out, av, ap = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
        O.reconstruct_haplotypes_from_sparse(out, *inputs, av, ap)
        _golden.eq("pyref_haps.out", ci, out, g_out)
        _golden.eq("pyref_haps.annot_v", ci, av, g_av)
        _golden.eq("pyref_haps.annot_pos", ci, ap, g_ap)


def test_pyref_tracks():
    """Oracle == the reference's pure-Python shift_and_realign_track_sparse
    (python/genvarloader/_dataset/_tracks.py:706-824), all five fills, bit-exact."""
    cases = _golden.load_golden("pyref_tracks")
    assert len(cases) == 150
    for ci, (inputs, g_out) in enumerate(cases):
        out = np.zeros(int(inputs[0][-1]), np.float32)
        args = list(inputs)
        args[13] = int(args[13])
        O.shift_and_realign_tracks_sparse(out, *args)
        _golden.eq("pyref_tracks", ci, out, g_out)
