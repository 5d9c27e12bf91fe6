"""GPU parity at BASELINE.json's FULL batch sizes (the workloads bench.py times): one whole batch of cfg2 / cfg2d / cfg3t /
cfg4 through the C ABI host layer against the oracle (it finishes these in well under a second each), plus the
size-independent properties the domain offers -- RC rows are the reverse complement of the plain rows, fixed-length
rows and ragged rows share their common prefix, one-hot rows sum to (base is ACGT)."""
import numpy as np
import pytest

from tests import _golden

pytestmark = pytest.mark.gpu

N = ord("N")
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


@pytest.fixture(scope="module")
def K(cuda_device):
    from genvarloader_b200 import _kernels

    return _kernels


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    return oracle


def _workload(name):
    import bench

    w, d = bench.build_workload(name, 2)
    b = bench.host_batches(d, w, 1, 3)[0]
    args = (b["regions"], b["shifts"], b["goi"], d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles,
            d.alt_offsets, d.reference, d.ref_offsets, N)
    return w, d, b, args


@pytest.mark.parametrize("name", ["cfg2", "cfg2d", "cfg3", "cfg4"])
def test_full_batch_haplotypes_vs_oracle(K, O, name):
    from genvarloader_b200._ffi import lib

    w, d, b, args = _workload(name)
    L, rows = w["window"], b["goi"].size
    K.pin_static(d.reference, d.alt_alleles, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, d.alt_offsets)
    try:
        to_rc = b["to_rc"]
        e_out, e_oo = O.reconstruct_haplotypes_fused(*args, L, None, None, to_rc, parallel=True)
        assert e_out.size == rows * L
        g_out, g_oo = K.reconstruct_haplotypes_fused(*args, L, None, None, to_rc)
        _golden.eq(f"{name}.offsets", 0, g_oo, e_oo)
        _golden.eq(f"{name}.u8", 0, g_out, e_out)
        g_oh, _ = K.reconstruct_haplotypes_fused(*args, L, None, None, to_rc, mode="onehot")
        assert int(lib.gvl_debug_last_exec_kernel(K.default_ctx().handle)) == 1, "packed one-hot kernel did not run"
        _golden.eq(f"{name}.onehot", 0, g_oh, O.onehot(e_out, parallel=True))
        # property: one-hot rows sum to 1 exactly where the byte is A/C/G/T
        assert (g_oh.sum(1) == np.isin(g_out, np.frombuffer(b"ACGT", np.uint8))).all()
        # property: flipping every row's strand gives the reverse complement of each row
        f_out, _ = K.reconstruct_haplotypes_fused(*args, L, None, None, ~to_rc)
        assert (f_out.reshape(rows, L) == _COMP[g_out.reshape(rows, L)[:, ::-1]]).all()
        # property: fixed rows and ragged rows agree on their common prefix (a fixed row shorter than L reads on into the
        # reference past the region end, a longer one is cut)
        r_out, r_oo = K.reconstruct_haplotypes_fused(*args, -1, None, None, None)
        p_out, _ = K.reconstruct_haplotypes_fused(*args, L, None, None, None)
        p_out = p_out.reshape(rows, L)
        for k in range(0, rows, max(1, rows // 64)):
            row = r_out[r_oo[k]:r_oo[k + 1]]
            n = min(row.size, L)
            assert (p_out[k, :n] == row[:n]).all()
        # annotated at full size
        e = O.reconstruct_annotated_haplotypes_fused(*args, L, None, None, to_rc, parallel=True)
        g = K.reconstruct_annotated_haplotypes_fused(*args, L, None, None, to_rc)
        for j, nm in enumerate(["out", "annot_v", "annot_pos", "offsets"]):
            _golden.eq(f"{name}.annot.{nm}", 0, g[j], e[j])
    finally:
        K.unpin_static(d.reference, d.alt_alleles, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, d.alt_offsets)


def test_full_batch_tracks_vs_oracle(K, O):
    """cfg3t: 32 haplotypes x 524,288 values x 2 tracks (Repeat5p and Interpolate(1)), negative strands reversed."""
    w, d, b, args = _workload("cfg3t")
    L = w["window"]
    regions, goi, to_rc = b["regions"], b["goi"], b["to_rc"]
    diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
    lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    out_offsets = (np.arange(goi.size + 1) * L).astype(np.int64)
    track_lengths = lengths - diffs.clip(max=0).min(1)
    track_offsets = np.concatenate([[0], np.cumsum(track_lengths)]).astype(np.int64)
    for ti, (name, strategy, param) in enumerate([("track0", 0, 0.0), ("track1", 4, 1.0)]):
        s, e, v, io = d.tracks[name]
        a = (out_offsets, regions, b["shifts"], goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens, b["ds_idx"], s, e,
             v, io, track_offsets, np.array([param], np.float64), strategy, 12345 + ti, None, None, to_rc)
        exp = np.zeros(int(out_offsets[-1]), np.float32)
        O.intervals_and_realign_track_fused(exp, *a)
        got = np.full(int(out_offsets[-1]), 7.0, np.float32)
        K.intervals_and_realign_track_fused(got, *a)
        _golden.eq(f"cfg3t.{name}", 0, got, exp)
        assert np.count_nonzero(exp) > exp.size // 4
