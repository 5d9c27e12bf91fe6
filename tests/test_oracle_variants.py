"""The numpy oracle of the `variants` / `variant-windows` cores against the reference's frozen goldens
(/root/reference/tests/parity/golden/*.npz, copied value-for-value by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import variants_oracle as vo
from tests._golden import eq, load_golden


def _tuple_cases(name, fn):
    cases = load_golden(name)
    assert len(cases) >= 100
    for ci, (inputs, gold) in enumerate(cases):
        got = fn(*inputs)
        gold = gold if isinstance(gold, tuple) else (gold,)
        got = got if isinstance(got, tuple) else (got,)
        assert len(got) == len(gold)
        for j, (a, b) in enumerate(zip(got, gold)):
            eq(f"{name}#{ci}", j, a, b)


@pytest.mark.parametrize("name", ["gather_rows_i32", "gather_rows_f32"])
def test_gather_rows(name):
    _tuple_cases(name, vo.gather_rows)


def test_gather_alleles():
    _tuple_cases("gather_alleles", vo.gather_alleles)


def test_rc_alleles():
    _tuple_cases("rc_alleles", vo.rc_alleles)


@pytest.mark.parametrize("name", ["compact_keep_i32", "compact_keep_f32"])
def test_compact_keep(name):
    _tuple_cases(name, vo.compact_keep)


@pytest.mark.parametrize("name", ["fill_empty_scalar_i32", "fill_empty_scalar_f32"])
def test_fill_empty_scalar(name):
    dt = np.int32 if name.endswith("i32") else np.float32
    _tuple_cases(name, lambda d, o, f: vo.fill_empty_scalar(d, o, dt(f)))


@pytest.mark.parametrize("name", ["fill_empty_fixed_i32", "fill_empty_fixed_f32"])
def test_fill_empty_fixed(name):
    dt = np.int32 if name.endswith("i32") else np.float32
    _tuple_cases(name, lambda d, o, inner, f: vo.fill_empty_fixed(d, o, int(inner), dt(f)))


@pytest.mark.parametrize("name", ["fill_empty_seq_u8", "fill_empty_seq_i32"])
def test_fill_empty_seq(name):
    _tuple_cases(name, vo.fill_empty_seq)


def test_assemble_variant_buffers():
    cases = load_golden("assemble_variant_buffers")
    assert len(cases) == 18
    for ci, (inputs, gold) in enumerate(cases):
        got = vo.assemble_variant_buffers(*inputs)
        assert list(got) == list(gold), (ci, list(got), list(gold))  # same fields, same order
        for nm in gold:
            eq(f"assemble#{ci}.{nm}", 0, got[nm][0], gold[nm][0])
            eq(f"assemble#{ci}.{nm}", 1, got[nm][1], gold[nm][1])


# ---- vectors produced by executing the reference's own numpy twins (tests/golden/make_pyref_variants_golden.py) ----
def _pyref(name, fn):
    cases = load_golden(f"pyref_{name}")
    assert len(cases) >= 40
    for ci, (inputs, gold) in enumerate(cases):
        got = fn(*inputs)
        gold = gold if isinstance(gold, tuple) else (gold,)
        assert len(got) == len(gold)
        for j, (a, b) in enumerate(zip(got, gold)):
            eq(f"pyref_{name}#{ci}", j, a, b)


def alt_window_args(rw, rw_off, a_data, a_off, L):
    """`assemble_variant_buffers` inputs whose per-variant reference windows are exactly the fixture's `rw` rows: the
    windows laid end to end form the contig, variant i is a deletion spanning its window minus the flanks."""
    n = len(rw_off) - 1
    L = int(L)
    span = np.diff(rw_off) - 2 * L
    v_starts = (rw_off[:-1] + L).astype(np.int32)
    ilens = (1 - span).astype(np.int32)
    v = np.arange(n, dtype=np.int32)
    lut = np.arange(256, dtype=np.uint8)
    return (1, v, np.array([0, n], np.int64), a_data, a_off, None, None, False, False, 1, 1, L, lut, np.zeros(n, np.int32),
            v_starts, ilens, rw, np.array([0, len(rw)], np.int64), ord("N"))


def test_pyref_twins():
    _pyref("gather_rows", vo.gather_rows)
    _pyref("compact_keep", vo.compact_keep)
    _pyref("fill_empty_scalar", lambda d, o, f: vo.fill_empty_scalar(d, o, d.dtype.type(f)))
    _pyref("fill_empty_fixed", lambda d, o, inner, f: vo.fill_empty_fixed(d, o, int(inner), d.dtype.type(f)))
    _pyref("fill_empty_seq", vo.fill_empty_seq)

    def windows(rw, rw_off, a_data, a_off, L):
        out = vo.assemble_variant_buffers(*alt_window_args(rw, rw_off, a_data, a_off, L))
        eq("ref_window", 0, out["ref_window"][0], rw)  # identity token table: the window bytes themselves
        eq("ref_window", 1, out["ref_window"][1], rw_off)
        return out["alt_window"]

    _pyref("alt_windows", windows)


def test_pyref_token_lut():
    from genvarloader_b200._types import build_token_lut

    for (alphabet, unk), lut in load_golden("pyref_token_lut"):
        got, dt = build_token_lut(alphabet.tobytes(), int(unk))
        eq("token_lut", 0, got, lut)
        assert dt == lut.dtype
