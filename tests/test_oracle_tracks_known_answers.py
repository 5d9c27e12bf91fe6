"""Known-answer vectors for the track core, transcribed from the reference's in-file Rust unit tests
(src/tracks/mod.rs:1208-2178: `run_fill` and `run_singular` cases) and replayed on the oracle -- a third pin of the
track path besides the 200-case golden and the 150 pure-Python-twin cases.  Only inputs and expected outputs are
taken from the reference; the Lagrange expectations are recomputed here from the anchors its doc comments list."""
import numpy as np

from oracle import oracle as O

REPEAT_5P, REPEAT_5P_NORM, CONSTANT, FLANK_SAMPLE, INTERPOLATE = 0, 1, 2, 3, 4


def _singular(track, v_starts, ilens, shift, out_len, strategy=REPEAT_5P, params=(0.0,), query_start=0):
    """One (query, hap) row through shift_and_realign_tracks_sparse (run_singular, src/tracks/mod.rs:1705-1762)."""
    track = np.asarray(track, np.float32)
    n = len(v_starts)
    out = np.full(out_len, -7.0, np.float32)
    O.shift_and_realign_tracks_sparse(
        out, np.array([0, out_len]), np.array([[0, query_start, query_start + track.size]], np.int32),
        np.array([[shift]], np.int32), np.zeros((1, 1), np.int64), np.arange(n, dtype=np.int32), np.array([0, n]),
        np.asarray(v_starts, np.int32), np.asarray(ilens, np.int32), track, np.array([0, track.size]),
        np.asarray(params, np.float64), None, None, strategy, 0)
    return out


def test_singular_cases():
    # :1764 no variants -> track[:length]
    assert _singular([1, 2, 3, 4, 5], [], [], 0, 4).tolist() == [1, 2, 3, 4]
    # :1810 deletion: track[v_rel_pos] once, then the track resumes after the deleted span, trailing zero
    assert _singular([10, 20, 30, 40, 50], [1], [-2], 0, 4).tolist() == [10, 20, 50, 0]
    # :1871 / :1934 deletion running past the track end: zero pad starts at out_idx = 4
    assert _singular([1, 2, 3, 4, 5], [3], [-3], 0, 8).tolist() == [1, 2, 3, 4, 0, 0, 0, 0]
    # :1972 SNPs write nothing
    assert _singular([1, 2, 3, 4], [2], [0], 0, 4).tolist() == [1, 2, 3, 4]
    # :2014 insertion, REPEAT_5P: ilen + 1 copies of track[v_rel_pos]
    assert _singular([5, 10, 15, 20, 25], [1], [2], 0, 6).tolist() == [5, 10, 10, 10, 15, 20]
    # :2048 insertion, CONSTANT
    assert _singular([5, 10, 15, 20], [1], [1], 0, 5, CONSTANT, (99.0,)).tolist() == [5, 99, 99, 15, 20]
    # :2086 no variants, shift 0
    assert _singular([0, 1, 2, 3, 4, 5], [], [], 0, 4).tolist() == [0, 1, 2, 3]
    # :2140 shift of 2 consumed inside an insertion of 4 written values
    assert _singular(np.arange(7), [1], [3], 2, 4).tolist() == [1, 1, 1, 2]


def _fill(track, v_rel_pos, v_len, strategy, params):
    """apply_insertion_fill for a whole insertion (run_fill, :1166-1199) = the values an insertion at v_rel_pos with
    ilen = v_len - 1 writes."""
    out = _singular(track, [v_rel_pos], [v_len - 1], 0, v_rel_pos + v_len, strategy, params)
    assert out[:v_rel_pos].tolist() == np.asarray(track, np.float32)[:v_rel_pos].tolist()
    return out[v_rel_pos:]


def _lagrange(xs, ys, n):
    """term = y_a * prod_b (x - x_b) / (x_a - x_b), accumulated in anchor order in float64, stored as float32."""
    res = []
    for i in range(n):
        x, acc = float(i), 0.0
        for a in range(len(xs)):
            term = ys[a]
            for b in range(len(xs)):
                if b != a:
                    term *= (x - xs[b]) / (xs[a] - xs[b])
            acc += term
        res.append(np.float32(acc))
    return np.array(res, np.float32)


def test_fill_cases():
    # :1208 REPEAT_5P_NORM: 6 / 3 = 2 (sum preserving)
    got = _fill([1, 6, 2], 1, 3, REPEAT_5P_NORM, (0.0,))
    assert got.tolist() == [2, 2, 2] and got.sum() == 6
    # :1246 f32 / f32 precision
    got = _fill([0, 1, 0], 1, 3, REPEAT_5P_NORM, (0.0,))
    assert (got == np.float32(1.0) / np.float32(3.0)).all()
    # :1276, :1289 CONSTANT, NaN default
    assert (_fill([0, 0, 0, 0, 0], 0, 4, CONSTANT, (3.14,)) == np.float32(3.14)).all()  # (params[0] as f32)
    got = _fill([0], 0, 3, CONSTANT, (float("nan"),))
    assert got.size == 3 and np.isnan(got).all()
    # :1435 INTERPOLATE order 1: anchors xs = [0, 3], ys = [track[1], track[2]]
    got = _fill([0, 4, 8], 1, 3, INTERPOLATE, (1.0,))
    assert got[0] == 4.0 and (got.view(np.uint32) == _lagrange([0.0, 3.0], [4.0, 8.0], 3).view(np.uint32)).all()
    # :1511 order 2: k = 2 anchors per side, xs = [0, -1, 2, 3], ys = [4, 2, 8, 16]
    got = _fill([1, 2, 4, 8, 16], 2, 2, INTERPOLATE, (2.0,))
    assert got[0] == 4.0 and (got.view(np.uint32) == _lagrange([0.0, -1.0, 2.0, 3.0], [4.0, 2.0, 8.0, 16.0], 2).view(np.uint32)).all()
    # :1581 order 3 (same anchors as order 2): xs = [0, -1, 4, 5], ys = [5, 1, 9, 2]
    got = _fill([3, 1, 5, 9, 2, 6], 2, 4, INTERPOLATE, (3.0,))
    assert got[0] == 5.0 and (got.view(np.uint32) == _lagrange([0.0, -1.0, 4.0, 5.0], [5.0, 1.0, 9.0, 2.0], 4).view(np.uint32)).all()
    # :1645 order 1 reaches for the 3' anchor: xs = [0, 2], ys = [10, 6] -> [10, 8]
    assert _fill([2, 10, 6], 1, 2, INTERPOLATE, (1.0,)).tolist() == [10, 8]
    # :1691 REPEAT_5P
    assert _fill([5, 11, 7], 1, 4, REPEAT_5P, (0.0,)).tolist() == [11, 11, 11, 11]
