"""Loader for the pickle-free golden fixtures under tests/golden/ (see make_golden.py).

Mirrors the replay helpers of the reference's tests/parity/_golden.py:139-197 so the
parity tests read the same way: ``load_golden(name)`` returns ``[(inputs_tuple, golden), ...]``.
"""
from __future__ import annotations

from functools import lru_cache
from pathlib import Path

import numpy as np

GOLDEN_DIR = Path(__file__).parent / "golden"


def _unbox(a: np.ndarray):
    return a[()] if a.ndim == 0 else a


@lru_cache(maxsize=None)
def load_golden(name: str) -> list:
    z = np.load(GOLDEN_DIR / f"ref_{name}.npz")
    n_cases, n_args, n_gold = int(z["n_cases"]), int(z["n_args"]), int(z["n_gold"])
    cases = []
    for ci in range(n_cases):
        none = set(z[f"c{ci}_none"].tolist())
        inputs = tuple(None if j in none else _unbox(z[f"c{ci}_a{j}"]) for j in range(n_args))
        if n_gold < 0:  # dict-valued golden (assemble_variant_buffers): {field: (data, seq_offsets)} in the reference's order
            names = [str(x) for x in z[f"c{ci}_gnames"]]
            cases.append((inputs, {nm: (z[f"c{ci}_g{2 * j}"], z[f"c{ci}_g{2 * j + 1}"]) for j, nm in enumerate(names)}))
            continue
        gold = tuple(_unbox(z[f"c{ci}_g{j}"]) for j in range(n_gold))
        cases.append((inputs, gold if n_gold > 1 else gold[0]))
    return cases


def eq(name: str, i: int, got, exp) -> None:
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.dtype == exp.dtype, f"{name}[{i}]: dtype {got.dtype} != {exp.dtype}"
    assert got.shape == exp.shape, f"{name}[{i}]: shape {got.shape} != {exp.shape}"
    if got.dtype.kind == "f":  # bit-exact, NaN-safe
        np.testing.assert_array_equal(got.view(np.uint32), exp.view(np.uint32), err_msg=f"{name}[{i}] bits differ")
    else:
        np.testing.assert_array_equal(got, exp, err_msg=f"{name}[{i}] value mismatch")


def replay_return(fn, name, cases):
    for ci, (inputs, golden) in enumerate(cases):
        eq(f"{name}#{ci}", 0, fn(*inputs), golden)


def replay_tuple(fn, name, cases):
    for ci, (inputs, golden) in enumerate(cases):
        got = fn(*inputs)
        got = got if isinstance(got, tuple) else (got,)
        gold = golden if isinstance(golden, tuple) else (golden,)
        assert len(got) == len(gold)
        for j, (a, b) in enumerate(zip(got, gold)):
            eq(f"{name}#{ci}", j, a, b)


def replay_inplace(fn, name, cases, out_factory, out_index):
    for ci, (inputs, golden) in enumerate(cases):
        out = out_factory(inputs)
        args = list(inputs)
        args.insert(out_index, out)
        fn(*args)
        eq(f"{name}#{ci}", 0, out, golden)
