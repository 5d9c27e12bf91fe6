"""Insertion-fill strategies for realigned tracks -- same names, parameters and ids as the
reference's python/genvarloader/_dataset/_insertion_fill.py (ids :9-13, lowering :89-121)."""
from __future__ import annotations

import math
from dataclasses import dataclass


class InsertionFill:
    """How the values of bases INSERTED by a haplotype are filled in a realigned track."""

    strategy_id: int = -1

    def param(self) -> float:
        return 0.0


@dataclass(frozen=True)
class Repeat5p(InsertionFill):
    """Repeat the value at the variant position over the whole insertion (default)."""

    strategy_id = 0


@dataclass(frozen=True)
class Repeat5pNormalized(InsertionFill):
    """Repeat value / (inserted length + 1): the written values sum to the original value."""

    strategy_id = 1


@dataclass(frozen=True)
class Constant(InsertionFill):
    """Write a fixed value (default NaN) at every inserted position."""

    value: float = math.nan
    strategy_id = 2

    def param(self) -> float:
        return float(self.value)


@dataclass(frozen=True)
class FlankSample(InsertionFill):
    """Sample each inserted value (with replacement, hash-seeded) from the 2*flank_width+1 reference
    values centred on the variant position."""

    flank_width: int = 5
    strategy_id = 3

    def __post_init__(self):
        if self.flank_width < 0:
            raise ValueError(f"flank_width must be >= 0, got {self.flank_width}")

    def param(self) -> float:
        return float(self.flank_width)


@dataclass(frozen=True)
class Interpolate(InsertionFill):
    """Lagrange interpolation of order 1, 2 or 3 across the insertion (f64 arithmetic)."""

    order: int = 1
    strategy_id = 4

    def __post_init__(self):
        if self.order not in (1, 2, 3):
            raise ValueError(f"Interpolate order must be 1, 2, or 3 (got {self.order})")

    def param(self) -> float:
        return float(self.order)


def lower(strategies) -> tuple[list[int], list[float]]:
    """(strategy ids, one f64 parameter each) in the order given."""
    ids, params = [], []
    for s in strategies:
        if not isinstance(s, InsertionFill) or s.strategy_id < 0:
            raise TypeError(f"Unknown InsertionFill: {type(s).__name__}")
        ids.append(int(s.strategy_id))
        params.append(s.param())
    return ids, params
