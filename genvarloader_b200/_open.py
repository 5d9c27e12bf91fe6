"""Read a GenVarLoader dataset directory (written by `gvl.write`) into the GPU-resident `Dataset`.

Layout followed (reference `docs/source/format.md:8-49`):

    metadata.json                      samples, contigs, n_regions, ploidy, max_jitter, version, svar_link, svar2_link
    input_regions.arrow                the caller's BED rows + `r_idx_map` (input row -> sorted storage row)
    genotypes/variants.arrow           POS (1-based from 0.18.0, _haps.py:462-468), ILEN, ALT[, REF]
    genotypes/variant_idxs.npy         raw int32 memmap (no .npy header, _haps.py:469-473)
    genotypes/offsets.npy              raw int64 memmap, R*S*P + 1 offsets (_haps.py:474-478)
    intervals/<track>/{starts,ends,values,offsets}.npy         raw int32 / int32 / float32 / int64, R*S + 1 offsets
    annot_intervals/<track>/...                                 same, R + 1 offsets (_tracks.py:327-339)

    genotypes/svar_meta.json           present when the genotypes live in a linked `.svar` store (SVAR1): offsets.npy is
                                       then a raw (2, R, S, P) starts/stops memmap into the store's variant_idxs.npy and
                                       the variant table is the store's index.arrow (_haps.py:389-446)

Only numpy + pyarrow are used.  `.svar` (SVAR1) links are followed: the store is resolved like the reference does
(`svar=` override, stored relative path, stored absolute path, a unique sibling `*.svar`; legacy `link.svar` symlink) and its
fingerprint is verified (_svar_link.py:24-103).  Datasets that back-reference a `.svar2` store need the third-party genoray
store format and are refused with a clear error.
"""
from __future__ import annotations

import gzip
import json
from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class Reference:
    """In-memory reference genome: contigs concatenated, upper-cased (reference `Reference`,
    python/genvarloader/_dataset/_reference.py:32-127; the FASTA cache upper-cases, _fasta_cache.py:115)."""

    reference: np.ndarray  # uint8
    offsets: np.ndarray    # int64 (n_contigs + 1)
    contigs: list
    pad_char: int = ord("N")

    @classmethod
    def from_arrays(cls, reference, offsets, contigs, pad_char: int = ord("N")) -> "Reference":
        return cls(np.ascontiguousarray(reference, np.uint8), np.ascontiguousarray(offsets, np.int64), list(contigs), pad_char)

    @classmethod
    def from_path(cls, fasta, contigs=None) -> "Reference":
        """Plain, gzip or bgzip FASTA.  `contigs` selects and orders contigs ("chr1" and "1" match either way)."""
        path = Path(fasta)
        opener = gzip.open if path.suffix in (".gz", ".bgz") else open
        seqs: dict = {}
        name, chunks = None, []
        with opener(path, "rb") as f:
            for line in f:
                if line.startswith(b">"):
                    if name is not None:
                        seqs[name] = b"".join(chunks)
                    name, chunks = line[1:].split()[0].decode(), []
                else:
                    chunks.append(line.strip().upper())
        if name is not None:
            seqs[name] = b"".join(chunks)
        if contigs is None:
            contigs = list(seqs)
        norm = _ContigMap(list(seqs))
        missing = [c for c in contigs if norm.get(c) is None]
        if missing:
            raise ValueError(f"Some of the given contig names are not present in reference file: {missing}")
        parts = [np.frombuffer(seqs[norm.get(c)], np.uint8) for c in contigs]
        offsets = np.concatenate([[0], np.cumsum([p.size for p in parts])]).astype(np.int64)
        return cls(np.concatenate(parts) if parts else np.zeros(0, np.uint8), offsets, list(contigs))


class _ContigMap:
    """UCSC / Ensembl tolerant contig lookup (reference `ContigNormalizer`, python/genvarloader/_utils.py)."""

    def __init__(self, contigs):
        self.contigs = list(contigs)
        self._m = {}
        for c in self.contigs:
            self._m[c] = c
            self._m[c[3:] if c.startswith("chr") else "chr" + c] = c

    def get(self, name):
        return self._m.get(name)

    def index(self, name) -> int:
        c = self.get(name)
        if c is None:
            raise ValueError(f"contig {name!r} is not among the dataset's contigs {self.contigs}")
        return self.contigs.index(c)


def _version_tuple(v) -> tuple | None:
    if v is None:
        return None
    if isinstance(v, dict):
        return (int(v.get("major", 0)), int(v.get("minor", 0)), int(v.get("patch", 0)))
    core = str(v).split("+")[0].split("-")[0]
    parts = (core.split(".") + ["0", "0"])[:3]
    return tuple(int("".join(ch for ch in p if ch.isdigit()) or 0) for p in parts)


def _first_of_lists(col):
    import pyarrow as pa
    import pyarrow.compute as pc

    col = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    if pa.types.is_list(col.type) or pa.types.is_large_list(col.type):
        col = pc.list_element(col, 0)
    return col


def _utf8_to_bytes_offsets(col) -> tuple:
    """Arrow (large_)utf8 array -> (uint8 bytes, int64 offsets) without a Python loop."""
    import pyarrow as pa

    col = _first_of_lists(col)
    if pa.types.is_dictionary(col.type):
        col = col.dictionary_decode()
    if not (pa.types.is_string(col.type) or pa.types.is_large_string(col.type)):
        col = col.cast(pa.large_utf8())
    if col.null_count:
        raise ValueError("ALT contains nulls")
    bufs = col.buffers()
    odt = np.int64 if pa.types.is_large_string(col.type) else np.int32
    offs = np.frombuffer(bufs[1], odt)[col.offset: col.offset + len(col) + 1].astype(np.int64)
    data = np.frombuffer(bufs[2], np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
    lo, hi = int(offs[0]), int(offs[-1])
    return np.ascontiguousarray(data[lo:hi]), offs - lo


def _compact_ranges(go: np.ndarray, store_idxs: np.ndarray, keep_below: float = 0.5):
    """A dataset linked to a cohort-scale `.svar` store references only the slices of the store's `variant_idxs.npy` that
    fall into its regions.  When those are less than `keep_below` of the store, gather them into a compact array (in slot
    order) and rebase the `(2, n)` starts / stops, so that only referenced genotypes are uploaded to HBM; otherwise the
    store's array is used as it is (zero-copy memmap).  Ranges may be empty (stop <= start)."""
    starts, lens = go[0], np.maximum(go[1] - go[0], 0)
    total = int(lens.sum())
    if total >= keep_below * store_idxs.size:
        return go, store_idxs
    new_starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64) if lens.size else np.zeros(0, np.int64)
    src = np.repeat(starts - new_starts, lens) + np.arange(total, dtype=np.int64)  # store index of every kept entry
    return np.stack([new_starts, new_starts + lens]), np.ascontiguousarray(store_idxs[src])


def _resolve_svar(gvl_path: Path, link, override) -> Path:
    """The `.svar` directory a dataset points at (reference `_resolve_svar`, _svar_link.py:24-61): override, stored
    relative path, stored absolute path, a unique sibling `*.svar`; legacy datasets carry a `genotypes/link.svar` symlink."""
    if override is not None:
        p = Path(override)
        if not p.is_dir():
            raise FileNotFoundError(f"svar override path does not exist or is not a directory: {p}")
        return p
    if link:
        rel = (gvl_path / link["relative_path"]).resolve()
        if rel.is_dir():
            return rel
        absp = Path(link["absolute_path"])
        if absp.is_dir():
            return absp
    else:
        legacy = gvl_path / "genotypes" / "link.svar"
        if legacy.exists():
            return legacy.resolve()
    siblings = sorted(gvl_path.parent.glob("*.svar"))
    if len(siblings) == 1:
        return siblings[0]
    expected = Path(link["absolute_path"]).name if link else "<unknown>.svar"
    raise FileNotFoundError(f"Could not locate svar '{expected}' for GVL dataset at {gvl_path}. Tried: stored relative path, "
                            "stored absolute path, sibling *.svar. Pass `svar=` to `Dataset.open(...)` to override.")


def _verify_svar_fingerprint(svar_path: Path, link, n_variants: int) -> None:
    """reference `_verify_fingerprint`, _svar_link.py:64-103 (no-op for legacy datasets without a link record)."""
    if not link:
        return
    vi = svar_path / "variant_idxs.npy"
    if not vi.exists():
        raise FileNotFoundError(f"Expected variant_idxs.npy at {vi}; resolved svar is malformed.")
    exp = link.get("fingerprint") or {}
    bad = []
    if "n_variants" in exp and int(exp["n_variants"]) != n_variants:
        bad.append(f"n_variants: expected {exp['n_variants']}, observed {n_variants}")
    if "variant_idxs_bytes" in exp and int(exp["variant_idxs_bytes"]) != vi.stat().st_size:
        bad.append(f"variant_idxs_bytes: expected {exp['variant_idxs_bytes']}, observed {vi.stat().st_size}")
    if bad:
        raise ValueError(f"svar fingerprint mismatch at {svar_path}: " + "; ".join(bad))


def read_dataset_arrays(path, reference=None, svar=None) -> dict:
    """Parse the directory into the arrays `Dataset.from_arrays` takes (host side, numpy memmaps where possible)."""
    import pyarrow as pa
    import pyarrow.ipc as ipc

    path = Path(path)
    meta = json.loads((path / "metadata.json").read_text())
    if meta.get("svar2_link") or (path / "genotypes" / "svar2_ranges").exists():
        raise NotImplementedError(
            "this dataset back-references a .svar2 store; reading genoray's svar2 store format is outside the current scope "
            "(SURVEY.md 8f-2).  Datasets written from VCF/PGEN or linked to a .svar (SVAR1) store can be opened.")
    samples, contigs = list(meta["samples"]), list(meta["contigs"])
    ploidy = meta.get("ploidy")
    max_jitter = int(meta.get("max_jitter") or 0)
    cmap = _ContigMap(contigs)

    # ---- regions: the un-padded BED rows, stored in sorted order (r_idx_map: input row -> storage row) ----
    with pa.memory_map(str(path / "input_regions.arrow"), "r") as src:
        bed = ipc.open_file(src).read_all()
    cols = {n: bed.column(n) for n in bed.column_names}
    chrom = [str(c) for c in _first_of_lists(cols["chrom"]).cast(pa.large_utf8()).to_pylist()] \
        if not pa.types.is_dictionary(cols["chrom"].type) else [str(c) for c in cols["chrom"].to_pylist()]
    c_idx = np.array([cmap.index(c) for c in chrom], np.int32)
    starts = cols["chromStart"].to_numpy().astype(np.int32)
    ends = cols["chromEnd"].to_numpy().astype(np.int32)
    if "strand" in cols:
        st = cols["strand"]
        if pa.types.is_integer(st.type):
            strand = st.to_numpy().astype(np.int32)
        else:
            # write side maps {"+": 1, "-": -1, ".": 1} (_write.py:566-571); open side is strict (_utils.py:105-109)
            smap = {"+": 1, "-": -1, ".": 1}
            vals = st.to_pylist()
            bad = sorted({repr(x) for x in vals if x is None or str(x) not in smap})
            if bad:
                raise ValueError(f"input_regions.arrow: unknown strand value(s) {', '.join(bad)} (expected '+', '-' or '.')")
            strand = np.array([smap[str(x)] for x in vals], np.int32)
    else:
        strand = np.ones(len(starts), np.int32)
    r_idx_map = cols["r_idx_map"].to_numpy().astype(np.int64)
    n_regions = len(starts)
    if sorted(r_idx_map.tolist()) != list(range(n_regions)):
        raise ValueError("input_regions.arrow: r_idx_map is not a permutation of the region indices")
    full_regions = np.empty((n_regions, 4), np.int32)
    full_regions[r_idx_map] = np.stack([c_idx, starts, ends, strand], 1)
    extra = {n: np.asarray(c.to_pylist()) for n, c in cols.items()
             if n not in ("chrom", "chromStart", "chromEnd", "strand", "r_idx_map")}  # e.g. transcript ids for splicing
    out = dict(samples=samples, contigs=contigs, ploidy=ploidy, max_jitter=max_jitter, full_regions=full_regions,
               region_map=r_idx_map, tracks={}, track_kinds={}, bed_columns=extra)

    # ---- genotypes ----
    gdir = path / "genotypes"
    if gdir.exists():
        if ploidy is None:
            raise ValueError("metadata.json has genotypes but no ploidy")
        svar_meta = gdir / "svar_meta.json"
        linked = svar_meta.exists()
        if linked:  # genotypes live in a .svar store: its variant table, its variant_idxs.npy, our (2, R, S, P) offsets
            svar_path = _resolve_svar(path, meta.get("svar_link"), svar)
            table_path, one_based = svar_path / "index.arrow", True  # (`_Variants.from_table` default, _haps.py:106-110)
        else:
            ver = _version_tuple(meta.get("version"))
            table_path, one_based = gdir / "variants.arrow", ver is not None and ver >= (0, 18, 0)
        with pa.memory_map(str(table_path), "r") as src:
            vt = ipc.open_file(src).read_all()
        pos = _first_of_lists(vt.column("POS")).to_numpy(zero_copy_only=False).astype(np.int64) - int(one_based)
        alt, alt_off = _utf8_to_bytes_offsets(vt.column("ALT"))
        if "ILEN" in vt.column_names:
            ilen = _first_of_lists(vt.column("ILEN")).to_numpy(zero_copy_only=False).astype(np.int32)
        else:  # ALT length - REF length (_haps.py:127-134)
            _, ref_off = _utf8_to_bytes_offsets(vt.column("REF"))
            ilen = (np.diff(alt_off) - np.diff(ref_off)).astype(np.int32)
        # columns of the "variants" output (`_Variants.from_table`, _haps.py:139-157): REF strings, numeric INFO columns
        import pyarrow.types as pat

        if "REF" in vt.column_names:
            out["ref_alleles"] = _utf8_to_bytes_offsets(vt.column("REF"))
        out["variant_info"] = {
            n: vt.column(n).to_numpy(zero_copy_only=False) for n in vt.column_names
            if n not in ("POS", "ILEN") and (pat.is_integer(vt.schema.field(n).type) or pat.is_floating(vt.schema.field(n).type))
            and vt.column(n).null_count == 0}
        n_slots = n_regions * len(samples) * int(ploidy)
        if linked:
            _verify_svar_fingerprint(svar_path, meta.get("svar_link"), len(pos))
            sm = json.loads(svar_meta.read_text())
            shape = tuple(int(x) for x in sm["shape"])  # (2, r, s, p)
            if len(shape) != 4 or shape[0] != 2 or int(np.prod(shape[1:])) != n_slots:
                raise ValueError(f"genotypes/svar_meta.json: shape {shape} does not match (2, {n_regions}, {len(samples)}, {ploidy})")
            go = np.memmap(gdir / "offsets.npy", shape=shape, dtype=np.dtype(sm["dtype"]), mode="r").reshape(2, -1)
            store_idxs = np.memmap(svar_path / "variant_idxs.npy", dtype=np.int32, mode="r")
            go, store_idxs = _compact_ranges(np.asarray(go, np.int64), store_idxs)
            out.update(v_starts=pos.astype(np.int32), ilens=ilen, alt_alleles=alt, alt_offsets=alt_off, geno_offsets=go,
                       geno_v_idxs=store_idxs, svar_path=svar_path)
        else:
            out.update(v_starts=pos.astype(np.int32), ilens=ilen, alt_alleles=alt, alt_offsets=alt_off,
                       geno_v_idxs=np.memmap(gdir / "variant_idxs.npy", dtype=np.int32, mode="r"),
                       geno_offsets=np.memmap(gdir / "offsets.npy", dtype=np.int64, mode="r"))
            if out["geno_offsets"].size != n_slots + 1:
                raise ValueError(f"genotypes/offsets.npy holds {out['geno_offsets'].size} offsets, expected {n_slots + 1} "
                                 f"(regions x samples x ploidy + 1)")
    # ---- tracks ----
    for sub, kind, n_slots in (("intervals", "sample", n_regions * len(samples)), ("annot_intervals", "annot", n_regions)):
        tdir = path / sub
        if not tdir.exists():
            continue
        for p in sorted(tdir.iterdir()):
            if not p.is_dir() or ".tmp." in p.name or ".old." in p.name or p.name.endswith(".lock") or not any(p.iterdir()):
                continue
            if p.name in out["tracks"]:
                raise ValueError(f"Found sample and annotation tracks with the same name: {{{p.name!r}}}")
            offs = np.memmap(p / "offsets.npy", dtype=np.int64, mode="r")
            if offs.size != n_slots + 1:
                raise ValueError(f"{sub}/{p.name}/offsets.npy holds {offs.size} offsets, expected {n_slots + 1}")
            out["tracks"][p.name] = (np.memmap(p / "starts.npy", dtype=np.int32, mode="r"),
                                     np.memmap(p / "ends.npy", dtype=np.int32, mode="r"),
                                     np.memmap(p / "values.npy", dtype=np.float32, mode="r"), offs)
            out["track_kinds"][p.name] = kind
    # ---- reference ----
    if reference is not None:
        if not isinstance(reference, Reference):
            reference = Reference.from_path(reference, contigs)
        else:
            rmap = _ContigMap(reference.contigs)
            if [rmap.get(c) for c in contigs] != list(reference.contigs[: len(contigs)]) or len(reference.contigs) != len(contigs):
                # re-order / subset to the dataset's contig order
                parts = []
                for c in contigs:
                    rc = rmap.get(c)
                    if rc is None:
                        raise ValueError(f"Some of the given contig names are not present in reference file: [{c!r}]")
                    i = reference.contigs.index(rc)
                    parts.append(reference.reference[reference.offsets[i]: reference.offsets[i + 1]])
                reference = Reference.from_arrays(np.concatenate(parts), np.concatenate([[0], np.cumsum([p.size for p in parts])]),
                                                  contigs, reference.pad_char)
    out["reference"] = reference
    return out
