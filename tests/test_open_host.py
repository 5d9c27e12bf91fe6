"""Host-side parsing of the on-disk GenVarLoader dataset layout (no GPU): `_open.read_dataset_arrays`, `Reference`."""
import numpy as np
import pytest

from tests._gvl_disk import write_fasta, write_gvl_dataset


@pytest.fixture(scope="module")
def disk(tmp_path_factory):
    from genvarloader_b200 import synth

    d = synth.make_dataset(17, 120_000, 3, 9, 1500 + 2 * 8, 6.0, max_jitter=8, neg_strand_frac=0.5, straddle_ends=False,
                           n_tracks=2, max_indel=9)
    root = tmp_path_factory.mktemp("gvl")
    order = np.random.default_rng(0).permutation(d.n_regions)
    write_gvl_dataset(root / "ds", d, ["chr1"], ["s2", "s0", "s1"], order)
    write_fasta(root / "ref.fa", d.reference, d.ref_offsets, ["1"], lower_every=37)
    return d, root, order


def test_read_dataset_arrays_roundtrip(disk):
    from genvarloader_b200._open import Reference, read_dataset_arrays

    d, root, order = disk
    a = read_dataset_arrays(root / "ds", root / "ref.fa")
    assert a["samples"] == ["s2", "s0", "s1"] and a["ploidy"] == d.ploidy and a["max_jitter"] == 8
    assert (a["full_regions"] == d.regions).all()          # storage order restored through r_idx_map
    assert (a["region_map"] == order).all()
    assert (a["v_starts"] == d.v_starts).all() and (a["ilens"] == d.ilens).all()  # 1-based POS from version 0.18.0
    assert (a["alt_alleles"] == d.alt_alleles).all() and (a["alt_offsets"] == d.alt_offsets).all()
    assert list(a["variant_info"]) == ["AF"] and (a["variant_info"]["AF"] == np.linspace(0, 1, d.v_starts.size)).all()
    assert "ref_alleles" not in a  # (this table has no REF column; the .svar store's index.arrow has one)
    go = np.asarray(d.geno_offsets)
    assert a["geno_offsets"].size == go.shape[1] + 1
    for k in (0, 5, go.shape[1] - 1):
        got = a["geno_v_idxs"][a["geno_offsets"][k]: a["geno_offsets"][k + 1]]
        assert (got == d.geno_v_idxs[go[0, k]: go[1, k]]).all()
    assert sorted(a["tracks"]) == sorted(d.tracks) and set(a["track_kinds"].values()) == {"sample"}
    ref = a["reference"]
    assert isinstance(ref, Reference) and ref.contigs == ["chr1"]
    assert (ref.reference == d.reference).all() and (ref.offsets == d.ref_offsets).all()  # upper-cased, "1" == "chr1"


def test_version_and_errors(disk, tmp_path):
    import json

    from genvarloader_b200._open import _version_tuple, read_dataset_arrays

    d, root, order = disk
    assert _version_tuple("0.18.0") == (0, 18, 0) and _version_tuple("0.21.3+dev") == (0, 21, 3)
    assert _version_tuple({"major": 1, "minor": 2, "patch": 3}) == (1, 2, 3) and _version_tuple(None) is None
    old = tmp_path / "old"
    write_gvl_dataset(old, d, ["chr1"], ["a", "b", "c"], order, version="0.17.2", pos_one_based=False, strand_as_str=False)
    a = read_dataset_arrays(old)
    assert (a["v_starts"] == d.v_starts).all() and a["reference"] is None and (a["full_regions"] == d.regions).all()
    meta = json.loads((old / "metadata.json").read_text())
    meta["svar2_link"] = {"relative_path": "../x.svar2", "absolute_path": "/x.svar2", "fingerprint": {}}
    (old / "metadata.json").write_text(json.dumps(meta))
    with pytest.raises(NotImplementedError, match="svar2"):
        read_dataset_arrays(old)
    with pytest.raises(ValueError, match="not present in reference"):
        read_dataset_arrays(root / "ds", root / "ref.fa").get  # ok
        from genvarloader_b200._open import Reference

        Reference.from_path(root / "ref.fa", ["chr7"])


def test_svar_linked_dataset(disk, tmp_path):
    """Genotypes in a linked .svar store (reference _haps.py:389-446, _svar_link.py): resolution order, fingerprint,
    ILEN derived from the allele lengths, (2, r, s, p) offsets into the store's variant_idxs.npy."""
    import json
    import shutil

    from genvarloader_b200._open import read_dataset_arrays
    from tests._gvl_disk import write_svar_store

    d, root, order = disk
    fp = write_svar_store(tmp_path / "cohort.svar", d)
    write_gvl_dataset(tmp_path / "ds", d, ["chr1"], ["s2", "s0", "s1"], order, svar_dir=tmp_path / "cohort.svar", svar_fingerprint=fp)
    a = read_dataset_arrays(tmp_path / "ds")
    assert (a["v_starts"] == d.v_starts).all() and (a["ilens"] == d.ilens).all()
    assert (a["alt_alleles"] == d.alt_alleles).all() and (a["alt_offsets"] == d.alt_offsets).all()
    assert a["variant_info"] == {}  # the store's index.arrow: POS, REF, ALT[, ILEN]
    rb, ro = a["ref_alleles"]
    assert (np.diff(ro) == np.diff(d.alt_offsets) - d.ilens).all() and (rb == ord("N")).all()  # len(ALT) - len(REF) = ILEN
    go = np.asarray(d.geno_offsets)
    assert a["geno_offsets"].shape == go.shape and (np.asarray(a["geno_offsets"]) == go).all()
    assert (np.asarray(a["geno_v_idxs"]) == d.geno_v_idxs).all() and a["svar_path"] == (tmp_path / "cohort.svar").resolve()
    # moved store: the stored paths dangle, a unique sibling *.svar is found; two siblings are ambiguous; svar= overrides
    moved = tmp_path / "moved"
    moved.mkdir()
    shutil.move(str(tmp_path / "ds"), str(moved / "ds"))
    shutil.move(str(tmp_path / "cohort.svar"), str(moved / "renamed.svar"))
    meta = json.loads((moved / "ds" / "metadata.json").read_text())
    meta["svar_link"]["absolute_path"] = str(tmp_path / "gone.svar")
    (moved / "ds" / "metadata.json").write_text(json.dumps(meta))
    assert read_dataset_arrays(moved / "ds")["svar_path"] == moved / "renamed.svar"
    (moved / "other.svar").mkdir()
    with pytest.raises(FileNotFoundError, match="gone.svar"):
        read_dataset_arrays(moved / "ds")
    assert (read_dataset_arrays(moved / "ds", svar=moved / "renamed.svar")["ilens"] == d.ilens).all()
    with pytest.raises(FileNotFoundError, match="override"):
        read_dataset_arrays(moved / "ds", svar=moved / "nope.svar")
    # fingerprint mismatch
    meta["svar_link"]["fingerprint"]["n_variants"] += 1
    (moved / "ds" / "metadata.json").write_text(json.dumps(meta))
    with pytest.raises(ValueError, match="fingerprint mismatch"):
        read_dataset_arrays(moved / "ds", svar=moved / "renamed.svar")


def test_linked_store_ranges_are_compacted():
    """A dataset linked to a large `.svar` store uploads only the genotype slices it references (`_compact_ranges`): same
    slice contents, rebased (2, n) starts / stops, empty and inverted ranges kept empty; a mostly-referenced store is left as
    it is."""
    from genvarloader_b200._open import _compact_ranges

    rng = np.random.default_rng(0)
    store = rng.integers(0, 1000, size=10_000).astype(np.int32)
    st = np.sort(rng.integers(0, 9_000, size=50))
    ln = rng.integers(0, 20, size=50)
    ln[3] = 0
    go = np.stack([st, st + ln]).astype(np.int64)
    go[1, 7] = go[0, 7] - 5  # stop < start: an empty slot
    g2, s2 = _compact_ranges(go, store)
    assert s2.size < store.size and s2.size == int(np.maximum(go[1] - go[0], 0).sum())
    for k in range(50):
        assert (store[go[0, k]: max(go[1, k], go[0, k])] == s2[g2[0, k]: g2[1, k]]).all(), k
    whole = np.stack([np.arange(0, 10_000, 100), np.arange(100, 10_100, 100)]).astype(np.int64)
    g3, s3 = _compact_ranges(whole, store)
    assert s3 is store and g3 is whole
