"""Host-side pieces of the `variants` / `variant-windows` outputs (no GPU): the dummy-variant defaults, the byte -> token
table, option validation and the ragged containers (reference: _flat_variants.py:39-68, 291-321, _flat_flanks.py:23-40)."""
import numpy as np
import pytest
import torch

from genvarloader_b200._types import DummyVariant, Ragged, RaggedAlleles, RaggedVariants, VarWindowOpt, build_token_lut


def test_dummy_variant_scalars():
    d = DummyVariant(start=-3, ilen=2, dosage=0.5, info={"AF": 0.25, "AC": 7})
    assert d.scalar_for("start", np.int32) == np.int32(-3) and d.scalar_for("ilen", np.int32) == np.int32(2)
    assert d.scalar_for("dosage", np.float32) == np.float32(0.5)
    assert d.scalar_for("AF", np.float32) == np.float32(0.25) and d.scalar_for("AC", np.int32) == 7
    assert np.isnan(DummyVariant().scalar_for("QUAL", np.float32)) and DummyVariant().scalar_for("DP", np.int32) == 0
    assert DummyVariant().alt == b"N" and DummyVariant().ref == b"N" and DummyVariant().start == -1


def test_build_token_lut():
    lut, dt = build_token_lut("ACGT", 4)
    assert dt == np.uint8 and lut.shape == (256,) and [lut[b] for b in b"ACGTN"] == [0, 1, 2, 3, 4] and lut[ord("a")] == 4
    lut, dt = build_token_lut(b"ACGT", 300)  # an unknown token beyond uint8 widens the table
    assert dt == np.int32 and lut[ord("N")] == 300 and lut[ord("T")] == 3
    lut, dt = build_token_lut(bytes(range(256)), 0)
    assert dt == np.uint8 and (lut == np.arange(256)).all()


def test_var_window_opt():
    o = VarWindowOpt(8, "ACGT", 4)
    assert o.token_alphabet == b"ACGT" and o.ref == "window" and o.alt == "window"
    assert VarWindowOpt(0, b"AC", 2, ref="allele", alt="allele").flank_length == 0
    with pytest.raises(ValueError):
        VarWindowOpt(8, "ACGT", 4, ref="flank")


def test_ragged_variants_container():
    off = torch.tensor([0, 2, 2, 3, 5])
    start = Ragged(torch.arange(5, dtype=torch.int32), off, (2, 2, None))
    alt = RaggedAlleles(torch.frombuffer(bytearray(b"ACGGTTA"), dtype=torch.uint8), torch.tensor([0, 1, 3, 4, 6, 7]), off, (2, 2, None))
    rv = RaggedVariants({"start": start, "alt": alt}, off, (2, 2, None))
    assert rv.start is start and rv["alt"] is alt and rv.lengths.tolist() == [[2, 0], [1, 2]]
    assert alt.to_list() == [[b"A", b"CG"], [], [b"G"], [b"TT", b"A"]]
    flat = rv.reshape(4)
    assert flat.shape == (4, None) and flat["alt"].shape == (4, None) and flat["start"].shape == (4, None)
    one = RaggedVariants({"start": Ragged(start.data[:2], off[:3], (1, 2, None))}, off[:3], (1, 2, None)).squeeze(0)
    assert one.shape == (2, None)
    with pytest.raises(AttributeError):
        rv.dosage
