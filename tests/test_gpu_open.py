"""`Dataset.open` over the on-disk GenVarLoader layout vs the same data as an in-memory dataset: identical
haplotypes / annotations / one-hot / tracks, regions addressed in input-BED order (r_idx_map)."""
import numpy as np
import pytest

from tests._gvl_disk import write_fasta, write_gvl_dataset

pytestmark = pytest.mark.gpu


def test_open_matches_in_memory_dataset(cuda_device, tmp_path):
    from genvarloader_b200 import Dataset, synth

    d = synth.make_dataset(23, 150_000, 3, 10, 2000 + 2 * 8, 6.0, max_jitter=8, neg_strand_frac=0.5, straddle_ends=False,
                           n_tracks=2, max_indel=9)
    order = np.random.default_rng(1).permutation(d.n_regions)
    write_gvl_dataset(tmp_path / "ds", d, ["chr1"], ["a", "b", "c"], order)
    write_fasta(tmp_path / "ref.fa", d.reference, d.ref_offsets, ["chr1"], lower_every=53)
    mem = Dataset.from_synth(cuda_device, d, rng=5)
    dsk = Dataset.open(tmp_path / "ds", tmp_path / "ref.fa", device=cuda_device, rng=5)
    assert dsk.shape == mem.shape and dsk.samples == ["a", "b", "c"] and dsk.max_jitter == 8
    assert (dsk.regions == d.regions[order]).all()
    L = 1536
    for cfg in (dict(seqs="haplotypes", enc="bytes"), dict(seqs="haplotypes", enc="onehot"), dict(seqs="annotated", enc="bytes")):
        a = mem.with_len(L).with_seqs(cfg["seqs"]).with_encoding(cfg["enc"]).with_tracks(False)
        b = dsk.with_len(L).with_seqs(cfg["seqs"]).with_encoding(cfg["enc"]).with_tracks(False)
        got = b[:4, :2]
        exp = a[order[:4], :2]  # input row i of the opened dataset is storage region order[i]
        if cfg["seqs"] == "annotated":
            for f in ("haps", "var_idxs", "ref_coords"):
                assert (getattr(got, f) == getattr(exp, f)).all()
        else:
            assert (got == exp).all()
    # ragged haplotypes + both tracks, fancy indices
    rg = dsk.with_seqs("haplotypes")[[0, 7, 3], [2, 0, 1]]
    re_ = mem.with_seqs("haplotypes")[order[[0, 7, 3]], [2, 0, 1]]
    assert (rg[0].data == re_[0].data).all() and (rg[0].offsets == re_[0].offsets).all()
    assert (rg[1].data.view(dtype=rg[1].data.dtype) == re_[1].data).all()
    # the "variants" output of the opened dataset; its AF column (float64 on disk) drives the AF filter
    af = np.linspace(0, 1, d.v_starts.size)  # what tests/_gvl_disk.py writes
    memv = Dataset.from_arrays(cuda_device, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
                               d.geno_v_idxs, d.geno_offsets, d.regions, d.n_samples, d.ploidy, variant_info={"AF": af})
    assert "AF" not in memv.available_var_fields and "AF" not in dsk.available_var_fields  # (8-byte column: filter only)
    for kw in (dict(), dict(min_af=0.3, max_af=0.8)):
        gv = dsk.with_tracks(False).with_seqs("variants").with_settings(**kw)[:5, :]
        ev = memv.with_seqs("variants").with_settings(**kw)[order[:5], :]
        assert gv.shape == ev.shape == (5, 3, d.ploidy, None)
        assert (gv.offsets == ev.offsets).all() and (gv["start"].data == ev["start"].data).all()
        assert (gv["alt"].data == ev["alt"].data).all() and (gv["alt"].seq_offsets == ev["alt"].seq_offsets).all()
        assert (gv["ilen"].data == ev["ilen"].data).all()
    assert int(gv.offsets[-1]) < int(dsk.with_tracks(False).with_seqs("variants")[:5, :].offsets[-1])  # the filter removed variants
    # subset by input-order region indices
    sub = dsk.subset_to(regions=[2, 5], samples=["c"]).with_len(L).with_tracks(False)
    assert (sub[:, :] == mem.with_len(L).with_tracks(False)[order[[2, 5]], [2]].reshape(sub[:, :].shape)).all()


def test_spliced_haplotypes_are_concatenated_elements(cuda_device, tmp_path):
    """Spliced output (reference _splice.py / _query.py:206-330): the haplotype of a splice row is the concatenation of
    its elements' haplotypes, element by element reverse-complemented on negative strands -- checked against the
    unspliced (oracle-verified) output; splice rows from an explicit mapping and from input-BED columns."""
    import pyarrow as pa
    import pyarrow.ipc as ipc

    from genvarloader_b200 import Dataset, synth

    d = synth.make_dataset(29, 150_000, 3, 12, 900, 8.0, neg_strand_frac=0.5, straddle_ends=False, max_indel=9)
    mem = Dataset.from_synth(cuda_device, d, rng=5).with_tracks(False)
    rows = {"tA": [3, 0, 7], "tB": [5], "tC": [11, 10, 9, 8]}
    sp = mem.with_settings(splice_info=rows)
    assert sp.is_spliced and sp.n_regions == 3 and len(sp) == 3 * mem.n_samples

    def expect(ds_plain, row, s, e):
        parts = []
        for r in row:
            x = ds_plain[r, s]          # Ragged (ploidy, ~len)
            o = x.offsets.cpu().numpy()
            parts.append(x.data[o[e]: o[e + 1]].cpu().numpy())
        return np.concatenate(parts)

    for enc in ("bytes", "onehot"):
        a = sp.with_encoding(enc)
        plain = mem.with_encoding(enc)
        out = a[:, :]
        assert out.shape == (3, mem.n_samples, d.ploidy, None)
        o = out.offsets.cpu().numpy()
        k = 0
        for ri, row in enumerate(rows.values()):
            for s in range(mem.n_samples):
                for e in range(d.ploidy):
                    got = out.data[o[k]: o[k + 1]].cpu().numpy()
                    assert (got == expect(plain, row, s, e)).all(), (enc, ri, s, e)
                    k += 1
    one = sp[1, 2]
    assert one.shape == (d.ploidy, None)
    assert (one.data.cpu().numpy() == np.concatenate([expect(mem, rows["tB"], 2, e) for e in range(d.ploidy)])).all()
    pair = sp[[2, 0], [1, 1]]
    assert pair.shape == (2, d.ploidy, None)
    ann = sp.with_seqs("annotated")[0, 0]
    assert (ann.haps.data.cpu().numpy() == sp[0, 0].data.cpu().numpy()).all()
    # edge cases: a tuple of ragged index lists is a mapping (not column names); "variable" pads like the unspliced path;
    # region subsets do not mix with splice rows
    sp_t = mem.with_settings(splice_info=([3, 0, 7], [5]))
    assert sp_t.n_regions == 2 and (sp_t[0, 1].data.cpu().numpy() == sp[0, 1].data.cpu().numpy()).all()
    var = sp.with_len("variable")[:, :]
    rag = sp[:, :]
    lens = (rag.offsets[1:] - rag.offsets[:-1]).cpu().numpy()
    assert var.shape == (3, mem.n_samples, d.ploidy, int(lens.max()))
    flat = var.reshape(-1, var.shape[-1]).cpu().numpy()
    ro = rag.offsets.cpu().numpy()
    for k in (0, 4, len(lens) - 1):
        assert (flat[k, : lens[k]] == rag.data[ro[k]: ro[k + 1]].cpu().numpy()).all() and (flat[k, lens[k]:] == ord("N")).all()
    with pytest.raises(ValueError, match="spliced"):
        sp.subset_to(regions=[0, 1])
    # the same rows from BED columns of an opened dataset: (id column, order column)
    order = np.arange(d.n_regions)
    write_gvl_dataset(tmp_path / "ds", d, ["chr1"], ["a", "b", "c"], order)
    with pa.memory_map(str(tmp_path / "ds" / "input_regions.arrow"), "r") as src:
        bed = ipc.open_file(src).read_all()
    tid = ["x"] * d.n_regions
    rank = [0] * d.n_regions
    for name, idxs in rows.items():
        for k, r in enumerate(idxs):
            tid[r], rank[r] = name, k
    bed = bed.append_column("transcript", pa.array(tid)).append_column("exon", pa.array(rank))
    with pa.OSFile(str(tmp_path / "bed.new"), "wb") as f, ipc.new_file(f, bed.schema) as w:  # (bed is a view of the mapped file)
        w.write_table(bed)
    del bed
    (tmp_path / "bed.new").replace(tmp_path / "ds" / "input_regions.arrow")
    write_fasta(tmp_path / "ref.fa", d.reference, d.ref_offsets, ["chr1"])
    dsk = Dataset.open(tmp_path / "ds", tmp_path / "ref.fa", device=cuda_device).with_tracks(False)
    spd = dsk.with_settings(splice_info=("transcript", "exon"))
    names = list(spd.splice_names)
    assert set(names) == {"tA", "tB", "tC", "x"}
    got = spd[names.index("tC"), 1]
    assert (got.data.cpu().numpy() == sp[2, 1].data.cpu().numpy()).all() and (got.offsets.cpu().numpy() == sp[2, 1].offsets.cpu().numpy()).all()
    # BED-column splice rows of a region SUBSET: elements outside the subset drop out, rows left empty disappear
    sub = dsk.subset_to(regions=[11, 10, 3, 0, 7]).with_settings(splice_info=("transcript", "exon"))
    sn = list(sub.splice_names)
    assert set(sn) == {"tA", "tC"}
    g = sub[sn.index("tA"), 2]
    assert (g.data.cpu().numpy() == sp[0, 2].data.cpu().numpy()).all()
    g = sub[sn.index("tC"), 0]   # elements 11, 10 of [11, 10, 9, 8] are left, in exon order
    assert (g.data.cpu().numpy() == mem.with_settings(splice_info=[[11, 10]])[0, 0].data.cpu().numpy()).all()


def test_open_svar_linked_dataset(cuda_device, tmp_path):
    """A dataset whose genotypes live in a linked .svar store (strided (2, r, s, p) offsets into the store's sparse genotype
    array) gives the same batches as the self-contained dataset."""
    from genvarloader_b200 import Dataset, synth
    from tests._gvl_disk import write_svar_store

    d = synth.make_dataset(29, 150_000, 3, 10, 2000 + 2 * 8, 6.0, max_jitter=8, neg_strand_frac=0.5, straddle_ends=False,
                           n_tracks=1, max_indel=9)
    order = np.random.default_rng(2).permutation(d.n_regions)
    fp = write_svar_store(tmp_path / "cohort.svar", d, with_ilen=True)
    write_gvl_dataset(tmp_path / "linked", d, ["chr1"], ["a", "b", "c"], order, svar_dir=tmp_path / "cohort.svar", svar_fingerprint=fp)
    write_gvl_dataset(tmp_path / "plain", d, ["chr1"], ["a", "b", "c"], order)
    write_fasta(tmp_path / "ref.fa", d.reference, d.ref_offsets, ["chr1"])
    a = Dataset.open(tmp_path / "plain", tmp_path / "ref.fa", device=cuda_device)
    b = Dataset.open(tmp_path / "linked", tmp_path / "ref.fa", device=cuda_device, svar=tmp_path / "cohort.svar")
    L = 1800
    for seqs in ("haplotypes", "annotated"):
        x, y = a.with_len(L).with_seqs(seqs).with_tracks(False)[:, :], b.with_len(L).with_seqs(seqs).with_tracks(False)[:, :]
        if seqs == "annotated":
            # variant indices are GLOBAL table indices in both datasets (same table, same order)
            assert (x.haps == y.haps).all() and (x.var_idxs == y.var_idxs).all() and (x.ref_coords == y.ref_coords).all()
        else:
            assert (x == y).all()
    hx, tx = a.with_seqs("haplotypes")[[1, 4], [0, 2]]
    hy, ty = b.with_seqs("haplotypes")[[1, 4], [0, 2]]
    assert (hx.data == hy.data).all() and (hx.offsets == hy.offsets).all() and (tx.data == ty.data).all()
