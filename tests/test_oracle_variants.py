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
