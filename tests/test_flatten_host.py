"""Overlapping intervals: `gvl_flatten_intervals` (host code of the library, no GPU needed) rewrites a slot whose intervals
overlap into the equivalent disjoint list -- "equivalent" = the reference's paint order, later intervals overwrite earlier
ones (src/intervals.rs:64-85), which the oracle restates.  Checked by painting both forms with the oracle."""
import ctypes as C

import numpy as np


def _flatten(s, e, v, off):
    from genvarloader_b200._ffi import check, lib

    cap = 2 * len(s) + 16
    o_s, o_e, o_v = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.float32)
    o_off, n = np.empty(len(off), np.int64), C.c_int64(0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    check(lib.gvl_flatten_intervals(p(s), p(e), p(v), p(off), C.c_int64(len(off) - 1), p(o_s), p(o_e), p(o_v), p(o_off),
                                    C.c_int64(cap), C.byref(n)))
    return o_s[: n.value], o_e[: n.value], o_v[: n.value], o_off


def _n_overlapping(s, e, off):
    from genvarloader_b200._ffi import check, lib

    n = C.c_int64(0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    check(lib.gvl_intervals_overlap(p(s), p(e), p(off), C.c_int64(len(off) - 1), C.byref(n)))
    return n.value


def test_known_cases():
    # [0,10)=1 then [5,8)=2: 1 1 1 1 1 2 2 2 1 1 (the ADVICE example); identical starts; an empty interval; containment
    s = np.array([0, 5, 20, 20, 30, 31, 40, 40], np.int32)
    e = np.array([10, 8, 25, 22, 30, 35, 50, 50], np.int32)
    v = np.array([1, 2, 3, 4, 5, 6, 7, 8], np.float32)
    off = np.array([0, 2, 4, 6, 8], np.int64)
    assert _n_overlapping(s, e, off) == 3  # slot 2 ([30,30) empty + [31,35)) does not overlap
    fs, fe, fv, fo = _flatten(s, e, v, off)
    got = [list(zip(fs[fo[k]:fo[k + 1]].tolist(), fe[fo[k]:fo[k + 1]].tolist(), fv[fo[k]:fo[k + 1]].tolist())) for k in range(4)]
    assert got[0] == [(0, 5, 1.0), (5, 8, 2.0), (8, 10, 1.0)]
    assert got[1] == [(20, 22, 4.0), (22, 25, 3.0)]
    assert got[2] == [(30, 30, 5.0), (31, 35, 6.0)]  # untouched slot: copied verbatim
    assert got[3] == [(40, 50, 8.0)]


def test_random_slots_paint_identically():
    from oracle import oracle as O

    rng = np.random.default_rng(5)
    for trial in range(40):
        n_slots = int(rng.integers(1, 6))
        ss, ee, vv, off = [], [], [], [0]
        for _ in range(n_slots):
            n = int(rng.integers(0, 25))
            st = np.sort(rng.integers(-20, 300, n)).astype(np.int32)
            ln = rng.integers(0, 60, n).astype(np.int32)
            ss.append(st), ee.append(st + ln), vv.append(rng.normal(size=n).astype(np.float32))
            off.append(off[-1] + n)
        s, e, v = np.concatenate(ss), np.concatenate(ee), np.concatenate(vv)
        off = np.array(off, np.int64)
        fs, fe, fv, fo = _flatten(s, e, v, off)
        for k in range(n_slots):  # disjoint and sorted
            a, b = fs[fo[k]:fo[k + 1]], fe[fo[k]:fo[k + 1]]
            keep = b > a
            assert (a[keep][1:] >= b[keep][:-1]).all()
        assert _n_overlapping(fs, fe, fo) == 0
        q = np.arange(n_slots, dtype=np.int64)
        starts = rng.integers(-30, 50, n_slots).astype(np.int32)
        oo = np.concatenate([[0], np.cumsum(rng.integers(1, 320, n_slots))]).astype(np.int64)
        exp, got = np.full(int(oo[-1]), 9, np.float32), np.full(int(oo[-1]), 7, np.float32)
        O.intervals_to_tracks(q, starts, s, e, v, off, exp, oo)
        O.intervals_to_tracks(q, starts, fs, fe, fv, fo, got, oo)
        assert (exp.view(np.uint32) == got.view(np.uint32)).all(), trial
