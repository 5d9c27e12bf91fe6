"""GPU parity of the `variants` / `variant-windows` outputs (csrc/gvl_variants.cu) through the C ABI: the reference's 13
frozen golden files (tests/parity/golden/{gather_rows_*,gather_alleles,rc_alleles,compact_keep_*,fill_empty_*,
assemble_variant_buffers}.npz, 1,318 cases) replayed through the host layer, seeded cases large enough to cross several
scan CTAs and the spine loop against the numpy oracle, and `Dataset.with_seqs("variants")` against the oracle's
composition of `get_variants_flat` (python/genvarloader/_dataset/_flat_variants.py:869-1112).  Bit-exact.
Mirrors tests/parity/test_{gather_rows,gather_alleles,rc_alleles,compact_keep,fill_empty_*,assemble_variant_buffers}_parity.py."""
import numpy as np
import pytest

from tests._golden import eq, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V(cuda_device):
    from genvarloader_b200 import _variants

    return _variants


@pytest.fixture(scope="module")
def O():
    from oracle import variants_oracle

    return variants_oracle


def _replay(name, fn, min_cases=100):
    cases = load_golden(name)
    assert len(cases) >= min_cases
    for ci, (inputs, gold) in enumerate(cases):
        got = fn(*inputs)
        gold = gold if isinstance(gold, tuple) else (gold,)
        got = got if isinstance(got, tuple) else (got,)
        assert len(got) == len(gold)
        for j, (a, b) in enumerate(zip(got, gold)):
            eq(f"{name}#{ci}", j, a, b)


@pytest.mark.parametrize("name", ["gather_rows_i32", "gather_rows_f32"])
def test_gather_rows_golden(V, name):
    _replay(name, getattr(V, name))


def test_gather_alleles_golden(V):
    _replay("gather_alleles", V.gather_alleles)


def test_rc_alleles_golden(V):
    def fn(byte_data, seq_offsets, var_offsets, to_rc_row):
        out = np.array(byte_data, np.uint8, copy=True)
        V.rc_alleles(out, seq_offsets, var_offsets, to_rc_row)
        return out

    _replay("rc_alleles", fn, 200)


@pytest.mark.parametrize("name", ["compact_keep_i32", "compact_keep_f32"])
def test_compact_keep_golden(V, name):
    _replay(name, getattr(V, name))


@pytest.mark.parametrize("name", ["fill_empty_scalar_i32", "fill_empty_scalar_f32"])
def test_fill_empty_scalar_golden(V, name):
    _replay(name, getattr(V, name))


@pytest.mark.parametrize("name", ["fill_empty_fixed_i32", "fill_empty_fixed_f32"])
def test_fill_empty_fixed_golden(V, name):
    _replay(name, getattr(V, name))


@pytest.mark.parametrize("name", ["fill_empty_seq_u8", "fill_empty_seq_i32"])
def test_fill_empty_seq_golden(V, name):
    _replay(name, getattr(V, name))


def _eq_bufs(tag, got, gold):
    assert list(got) == list(gold), (tag, list(got), list(gold))  # same fields, same order
    for nm in gold:
        eq(f"{tag}.{nm}", 0, got[nm][0], gold[nm][0])
        eq(f"{tag}.{nm}", 1, got[nm][1], gold[nm][1])


def test_assemble_variant_buffers_golden(V):
    cases = load_golden("assemble_variant_buffers")
    assert len(cases) == 18
    for ci, (inputs, gold) in enumerate(cases):
        _eq_bufs(f"assemble#{ci}", V.assemble_variant_buffers(*inputs), gold)


def test_pyref_twins(V):
    """Vectors produced by executing the reference's own numpy twins of these cores (tests/golden/make_pyref_variants_golden.py:
    bigger, more ragged cases than the frozen goldens), through the C ABI."""
    from tests.test_oracle_variants import alt_window_args

    def windows(rw, rw_off, a_data, a_off, L):
        out = V.assemble_variant_buffers(*alt_window_args(rw, rw_off, a_data, a_off, L))
        eq("ref_window", 0, out["ref_window"][0], rw)
        eq("ref_window", 1, out["ref_window"][1], rw_off)
        return out["alt_window"]

    for name, fn in (("gather_rows", V.gather_rows), ("compact_keep", V.compact_keep), ("fill_empty_scalar", V.fill_empty_scalar),
                     ("fill_empty_fixed", V.fill_empty_fixed), ("fill_empty_seq", V.fill_empty_seq), ("alt_windows", windows)):
        _replay(f"pyref_{name}", fn, 40)


# ---------------------------------------------------------------- seeded cases vs the oracle
def _ragged_offsets(rng, n, max_len, empty_frac=0.3):
    ln = rng.integers(1, max_len + 1, n)
    ln[rng.random(n) < empty_frac] = 0
    return np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)


@pytest.mark.parametrize("n_rows,max_len", [(3, 5), (700, 40), (5000, 900)])
def test_gather_rows_vs_oracle(V, O, n_rows, max_len):
    rng = np.random.default_rng(n_rows)
    n_slots = n_rows * 2 + 3
    slot_off = _ragged_offsets(rng, n_slots, max_len)
    go = np.stack([slot_off[:-1], slot_off[1:]])
    goi = rng.integers(0, n_slots, n_rows)
    for data in (rng.integers(-2**31, 2**31 - 1, slot_off[-1]).astype(np.int32), rng.standard_normal(slot_off[-1]).astype(np.float32)):
        got, exp = V.gather_rows(goi, go, data), O.gather_rows(goi, go, data)
        eq("gather_rows", 0, got[0], exp[0])
        eq("gather_rows", 1, got[1], exp[1])


def _allele_table(rng, n_table, max_len=30):
    off = np.concatenate([[0], np.cumsum(rng.integers(0, max_len + 1, n_table))]).astype(np.int64)
    return rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), off[-1]), off


@pytest.mark.parametrize("n", [1, 3000, 2_300_000])
def test_gather_alleles_vs_oracle(V, O, n):
    """2.3 M selected variants = 1,124 scan CTAs: more than one chunk of the spine kernel."""
    rng = np.random.default_rng(n)
    ab, ao = _allele_table(rng, 5000, 6 if n > 10**6 else 30)
    v = rng.integers(0, 5000, n).astype(np.int32)
    got, exp = V.gather_alleles(v, ab, ao), O.gather_alleles(v, ab, ao)
    eq("gather_alleles", 0, got[0], exp[0])
    eq("gather_alleles", 1, got[1], exp[1])


def test_rc_alleles_vs_oracle(V, O):
    rng = np.random.default_rng(5)
    n_rows = 400
    var_off = _ragged_offsets(rng, n_rows, 12)
    seq_off = np.concatenate([[0], np.cumsum(rng.integers(0, 40, var_off[-1]))]).astype(np.int64)  # (empty alleles included)
    data = rng.choice(np.frombuffer(b"ACGTNacgtRY", np.uint8), seq_off[-1])
    mask = rng.random(n_rows) < 0.5
    got = data.copy()
    V.rc_alleles(got, seq_off, var_off, mask)
    eq("rc_alleles", 0, got, O.rc_alleles(data, seq_off, var_off, mask))
    twice = got.copy()
    V.rc_alleles(twice, seq_off, var_off, mask)  # reverse-complement is an involution
    eq("rc_alleles twice", 0, twice, data)


@pytest.mark.parametrize("n_rows", [1, 50, 9000])
def test_compact_and_fill_vs_oracle(V, O, n_rows):
    rng = np.random.default_rng(n_rows + 1)
    off = _ragged_offsets(rng, n_rows, 60)
    n = int(off[-1])
    keep = rng.random(n) < 0.6
    for vals in (rng.integers(-9, 9, n).astype(np.int32), rng.standard_normal(n).astype(np.float32)):
        got, exp = V.compact_keep(vals, off, keep), O.compact_keep(vals, off, keep)
        eq("compact_keep", 0, got[0], exp[0])
        eq("compact_keep", 1, got[1], exp[1])
        fill = vals.dtype.type(-7)
        got, exp = V.fill_empty_scalar(vals, off, fill), O.fill_empty_scalar(vals, off, fill)
        eq("fill_empty_scalar", 0, got[0], exp[0])
        eq("fill_empty_scalar", 1, got[1], exp[1])
    inner = 6
    fixed = rng.integers(0, 100, n * inner).astype(np.int32)
    got, exp = V.fill_empty_fixed(fixed, off, inner, 4), O.fill_empty_fixed(fixed, off, inner, np.int32(4))
    eq("fill_empty_fixed", 0, got[0], exp[0])
    eq("fill_empty_fixed", 1, got[1], exp[1])
    seq_off = np.concatenate([[0], np.cumsum(rng.integers(0, 25, n))]).astype(np.int64)
    for data, dummy in ((rng.integers(65, 90, seq_off[-1]).astype(np.uint8), np.frombuffer(b"NNN", np.uint8)),
                        (rng.integers(0, 9, seq_off[-1]).astype(np.int32), np.array([7, 7], np.int32))):
        got, exp = V.fill_empty_seq(data, off, seq_off, dummy), O.fill_empty_seq(data, off, seq_off, dummy)
        for j in range(3):
            eq("fill_empty_seq", j, got[j], exp[j])


@pytest.mark.parametrize("tok", [np.uint8, np.int32])
@pytest.mark.parametrize("flank", [0, 3, 64])
def test_assemble_variant_buffers_vs_oracle(V, O, tok, flank):
    """Two contigs, variants next to both contig ends (flanks reach outside: pad), deletions, long insertions; every
    ref / alt mode of the windows tail and every option of the variants tail."""
    rng = np.random.default_rng(flank + 11)
    contig_lens = [5000, 3000]
    ref_off = np.concatenate([[0], np.cumsum(contig_lens)]).astype(np.int64)
    reference = rng.choice(np.frombuffer(b"ACGTN", np.uint8), ref_off[-1])
    n_var = 600
    v_starts = np.sort(rng.integers(0, 3000, n_var)).astype(np.int32)
    v_starts[:3], v_starts[-3:] = [0, 1, 2], [2997, 2998, 2999]
    ilens = rng.integers(-30, 31, n_var).astype(np.int32)
    alt_len = np.where(ilens > 0, ilens + 1, 1)
    ref_len = np.where(ilens < 0, 1 - ilens, 1)
    alt_off = np.concatenate([[0], np.cumsum(alt_len)]).astype(np.int64)
    rfo = np.concatenate([[0], np.cumsum(ref_len)]).astype(np.int64)
    alt = rng.choice(np.frombuffer(b"ACGT", np.uint8), alt_off[-1])
    rfa = rng.choice(np.frombuffer(b"ACGT", np.uint8), rfo[-1])
    n = 900
    v = rng.integers(0, n_var, n).astype(np.int32)
    v_contigs = rng.integers(0, 2, n).astype(np.int32)
    row_off = np.concatenate([[0], np.sort(rng.integers(0, n + 1, 19)), [n]]).astype(np.int64)
    lut = (np.arange(256) * 7 % 251).astype(tok)
    common = (lut, v_contigs, v_starts, ilens, reference, ref_off, ord("N"))
    for ref_mode in (1, 2):
        for alt_mode in (1, 2):
            args = (1, v, row_off, alt, alt_off, rfa, rfo, False, False, ref_mode, alt_mode, flank) + common
            _eq_bufs(f"windows r{ref_mode} a{alt_mode}", V.assemble_variant_buffers(*args), O.assemble_variant_buffers(*args))
    for want_ref in (False, True):
        for want_flank in (False, True):
            args = (0, v, row_off, alt, alt_off, rfa, rfo, want_ref, want_flank, 0, 0, flank) + common
            _eq_bufs(f"variants ref{want_ref} flank{want_flank}", V.assemble_variant_buffers(*args),
                     O.assemble_variant_buffers(*args))


# ---------------------------------------------------------------- Dataset.with_seqs("variants")
def _expected_variants(O, d, ds_idx, to_rc_q, fields, dummy, ref_alleles=None, af=None, min_af=None, max_af=None, fold=False,
                       dosages=None):
    """get_variants_flat (_flat_variants.py:869-1112) + the strand pass of _query.py:485-529, from oracle primitives."""
    p = d.ploidy
    goi = (ds_idx[:, None] * p + np.arange(p)[None, :]).reshape(-1)
    go = np.stack([d.geno_offsets[:-1], d.geno_offsets[1:]]) if np.asarray(d.geno_offsets).ndim == 1 else d.geno_offsets
    v, off = O.gather_rows(goi, go, d.geno_v_idxs)
    dos = O.gather_rows(goi, go, dosages)[0] if dosages is not None else None
    if min_af is not None or max_af is not None:
        keep = np.ones(len(v), bool)
        if min_af is not None:
            keep &= af[v] >= np.float32(min_af)
        if max_af is not None:
            keep &= af[v] <= np.float32(max_af)
        if dos is not None:
            dos, _ = O.compact_keep(dos, off, keep)
        v, off = O.compact_keep(v, off, keep)
    to_rc = np.repeat(to_rc_q, p)
    if fold:
        off, to_rc = off[::p].copy(), to_rc[::p]
    out = {}
    new_off = off
    for name in fields:
        if name in ("alt", "ref"):
            ab, ao = (d.alt_alleles, d.alt_offsets) if name == "alt" else ref_alleles
            data, so = O.gather_alleles(v, ab, ao)
            vo = off
            if dummy is not None:
                data, vo, so = O.fill_empty_seq(data, off, so, np.frombuffer(dummy.alt if name == "alt" else dummy.ref, np.uint8))
                new_off = vo
            out[name] = (O.rc_alleles(data, so, vo, to_rc), so)
        else:
            col = dos if name == "dosage" else {"start": d.v_starts, "ilen": d.ilens, "AF": af}[name][v]
            if dummy is not None:
                col, new_off = O.fill_empty_scalar(col, off, dummy.scalar_for(name, col.dtype))
            out[name] = col
    return out, new_off


def _check_variants(got, exp, exp_off, shape):
    assert got.shape == shape
    eq("variants.offsets", 0, got.offsets.cpu().numpy(), exp_off)
    for name, e in exp.items():
        f = got[name]
        if isinstance(e, tuple):
            eq(f"variants.{name}.data", 0, f.data.cpu().numpy(), e[0])
            eq(f"variants.{name}.seq_offsets", 0, f.seq_offsets.cpu().numpy(), e[1])
            eq(f"variants.{name}.var_offsets", 0, f.var_offsets.cpu().numpy(), exp_off)
        else:
            eq(f"variants.{name}", 0, f.data.cpu().numpy(), e)
            eq(f"variants.{name}.offsets", 0, f.offsets.cpu().numpy(), exp_off)


def test_dataset_variants_vs_oracle(cuda_device, O):
    from genvarloader_b200 import DummyVariant, synth
    from genvarloader_b200._dataset import Dataset

    d = synth.make_dataset(31, 300_000, 6, 12, 2500, 1.5, neg_strand_frac=0.5, straddle_ends=False, max_indel=15, snp_frac=0.5)
    rng = np.random.default_rng(3)
    n_var = len(d.v_starts)
    ref_len = np.where(d.ilens < 0, 1 - d.ilens, 1)
    rfo = np.concatenate([[0], np.cumsum(ref_len)]).astype(np.int64)
    rfa = rng.choice(np.frombuffer(b"ACGT", np.uint8), rfo[-1])
    af = rng.random(n_var).astype(np.float32)
    ds = Dataset.from_arrays(cuda_device, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
                             d.geno_v_idxs, d.geno_offsets, d.regions, d.n_samples, d.ploidy, ref_alleles=(rfa, rfo),
                             variant_info={"AF": af}, dosages=(dz := rng.random(len(d.geno_v_idxs)).astype(np.float32)))
    assert ds.available_var_fields == ["alt", "ilen", "start", "ref", "dosage", "AF"]
    S = d.n_samples
    dsv = ds.with_seqs("variants")
    r, s = np.arange(12).repeat(S), np.tile(np.arange(S), 12)
    ds_idx = r * S + s
    to_rc_q = d.regions[r, 3] == -1
    assert to_rc_q.any() and not to_rc_q.all()

    # default fields, whole dataset, no dummy: empty groups stay empty
    got = dsv[:, :]
    exp, off = _expected_variants(O, d, ds_idx, to_rc_q, ("alt", "ilen", "start"), None)
    assert (np.diff(off) == 0).any() and (np.diff(off) > 3).any()
    _check_variants(got, exp, off, (12, S, d.ploidy, None))
    assert got.alt.to_list()[0] == [bytes(exp["alt"][0][exp["alt"][1][a]:exp["alt"][1][a + 1]]) for a in range(off[0], off[1])]

    # every field + dummy variant + rc_neg off
    dummy = DummyVariant(start=-5, ilen=9, alt=b"AC", ref=b"G", dosage=0.5, info={"AF": 0.25})
    every = ("alt", "start", "ref", "ilen", "dosage", "AF")
    full = dsv.with_settings(var_fields=list(every), dummy_variant=dummy)
    sel_r, sel_s = np.array([0, 3, 3, 11, 7]), np.array([5, 0, 1, 2, 2])
    got = full[sel_r, sel_s]
    exp, off = _expected_variants(O, d, sel_r * S + sel_s, d.regions[sel_r, 3] == -1, every, dummy, (rfa, rfo), af, dosages=dz)
    assert (np.diff(off) >= 1).all()
    _check_variants(got, exp, off, (5, d.ploidy, None))
    got = full.with_settings(rc_neg=False)[sel_r, sel_s]
    exp, off = _expected_variants(O, d, sel_r * S + sel_s, np.zeros(5, bool), every, dummy, (rfa, rfo), af, dosages=dz)
    _check_variants(got, exp, off, (5, d.ploidy, None))

    # AF filter + unphased union
    flt = dsv.with_settings(min_af=0.2, max_af=0.9, unphased_union=True, dummy_variant=DummyVariant(),
                            var_fields=["alt", "ilen", "start", "dosage"])
    got = flt[:, :]
    exp, off = _expected_variants(O, d, ds_idx, to_rc_q, ("alt", "ilen", "start", "dosage"), DummyVariant(), None, af, 0.2, 0.9,
                                  fold=True, dosages=dz)
    _check_variants(got, exp, off, (12, S, 1, None))
    with pytest.raises(ValueError, match="Missing variant fields"):
        dsv.with_settings(var_fields=["alt", "QUAL"])


@pytest.mark.parametrize("alphabet,unk", [("ACGT", 4), ("ACGT", 300)])
def test_dataset_variant_windows_and_flank_tokens_vs_oracle(cuda_device, O, alphabet, unk):
    """with_seqs("variant-windows", VarWindowOpt(...)) and with_settings(flank_length=...) against the oracle's
    assemble_variant_buffers + fill_empty_* (get_variants_flat, _flat_variants.py:1006-1110); unk = 300 makes the tokens int32
    (build_token_lut, _flat_flanks.py:23-40).  Two contigs, so the per-variant contig expansion matters."""
    from genvarloader_b200 import DummyVariant, VarWindowOpt
    from genvarloader_b200._dataset import Dataset
    from genvarloader_b200._types import build_token_lut

    rng = np.random.default_rng(unk)
    contig_lens = [4000, 2500]
    ref_off = np.concatenate([[0], np.cumsum(contig_lens)]).astype(np.int64)
    reference = rng.choice(np.frombuffer(b"ACGTN", np.uint8), ref_off[-1])
    n_var = 300
    v_starts = np.sort(rng.integers(0, 2500, n_var)).astype(np.int32)
    v_starts[:2], v_starts[-2:] = [0, 1], [2498, 2499]
    ilens = rng.integers(-12, 13, n_var).astype(np.int32)
    alt_off = np.concatenate([[0], np.cumsum(np.where(ilens > 0, ilens + 1, 1))]).astype(np.int64)
    rfo = np.concatenate([[0], np.cumsum(np.where(ilens < 0, 1 - ilens, 1))]).astype(np.int64)
    alt = rng.choice(np.frombuffer(b"ACGT", np.uint8), alt_off[-1])
    rfa = rng.choice(np.frombuffer(b"ACGT", np.uint8), rfo[-1])
    R, S, P = 6, 3, 2
    regions = np.stack([rng.integers(0, 2, R), rng.integers(0, 2000, R), np.zeros(R, np.int64), np.where(rng.random(R) < 0.5, -1, 1)], 1)
    regions[:, 2] = regions[:, 1] + 400
    regions = regions.astype(np.int32)
    glen = rng.integers(0, 9, R * S * P)
    glen[rng.random(R * S * P) < 0.3] = 0
    geno_off = np.concatenate([[0], np.cumsum(glen)]).astype(np.int64)
    geno_v = np.concatenate([np.sort(rng.choice(n_var, k, replace=False)) for k in glen] + [np.zeros(0, np.int64)]).astype(np.int32)
    ds = Dataset.from_arrays(cuda_device, reference, ref_off, v_starts, ilens, alt, alt_off, geno_v, geno_off, regions, S, P,
                             ref_alleles=(rfa, rfo))
    lut, tok_dt = build_token_lut(alphabet, unk)
    assert tok_dt == (np.uint8 if unk < 256 else np.int32)
    r, s = np.arange(R).repeat(S), np.tile(np.arange(S), R)
    goi = ((r * S + s)[:, None] * P + np.arange(P)[None, :]).reshape(-1)
    go = np.stack([geno_off[:-1], geno_off[1:]])
    v, off = O.gather_rows(goi, go, geno_v)
    v_contigs = np.repeat(np.repeat(regions[r, 0], P), np.diff(off)).astype(np.int32)
    L = 5
    dummy = DummyVariant(alt=b"NN", ref=b"N")
    for ref_kind, alt_kind in (("window", "window"), ("allele", "window"), ("window", "allele"), ("allele", "allele")):
        opt = VarWindowOpt(L, alphabet, unk, ref=ref_kind, alt=alt_kind)
        for dv in (None, dummy):
            got = ds.with_seqs("variant-windows", opt).with_settings(dummy_variant=dv if dv is not None else False)[:, :]
            exp = O.assemble_variant_buffers(1, v, off, alt, alt_off, rfa, rfo, False, False, 1 if ref_kind == "window" else 2,
                                             1 if alt_kind == "window" else 2, L, lut, v_contigs, v_starts, ilens, reference, ref_off,
                                             ord("N"))
            exp_off = off
            assert sorted(got.fields) == sorted(["ilen", "start"] + list(exp))
            for name, (data, so) in exp.items():
                if dv is not None:
                    base = len(dv.alt if name in ("alt", "alt_window") else dv.ref)
                    wl = 2 * L + base if name.endswith("_window") else base
                    data, exp_off, so = O.fill_empty_seq(data, off, so, np.full(wl, unk, tok_dt))
                f = got[name]
                eq(f"windows.{name}.data", 0, f.data.cpu().numpy(), data)
                eq(f"windows.{name}.seq_offsets", 0, f.seq_offsets.cpu().numpy(), so)
                eq(f"windows.{name}.var_offsets", 0, f.var_offsets.cpu().numpy(), exp_off)
            st = v_starts[v]
            if dv is not None:
                st, _ = O.fill_empty_scalar(st, off, np.int32(dv.start))
            eq("windows.start", 0, got["start"].data.cpu().numpy(), st)
            assert got.shape == (R, S, P, None)
    # variants + flank tokens ride-along (negative-strand alleles reverse-complemented, tokens reference-oriented)
    fl = ds.with_seqs("variants").with_settings(token_alphabet=alphabet, unknown_token=unk, flank_length=L, dummy_variant=dummy)
    got = fl[:, :]
    exp = O.assemble_variant_buffers(0, v, off, alt, alt_off, None, None, False, True, 0, 0, L, lut, v_contigs, v_starts, ilens,
                                     reference, ref_off, ord("N"))
    tok, new_off = O.fill_empty_fixed(exp["flank_tokens"][0], off, 2 * L, tok_dt.type(unk))
    eq("flank_tokens", 0, got["flank_tokens"].data.cpu().numpy(), tok.reshape(-1, 2 * L))
    eq("flank_tokens.offsets", 0, got.offsets.cpu().numpy(), new_off)
    a_data, a_vo, a_so = O.fill_empty_seq(exp["alt"][0], off, exp["alt"][1], np.frombuffer(b"NN", np.uint8))
    to_rc = np.repeat(regions[r, 3] == -1, P)
    eq("alt", 0, got["alt"].data.cpu().numpy(), O.rc_alleles(a_data, a_so, a_vo, to_rc))
