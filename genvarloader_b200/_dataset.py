"""GPU-resident `Dataset`: the reference's `gvl.Dataset` surface for the haplotype / track hot path.

Mirrors (names, argument meaning, error behaviour, return shapes):
  * `Dataset` state + `with_settings / with_len / with_seqs / with_tracks / with_insertion_fill /
    with_output_format / subset_to / __getitem__`      python/genvarloader/_dataset/_impl.py:78-2121
  * index parsing (`DatasetIndexer.parse_idx`)          python/genvarloader/_dataset/_indexing.py:208-264
  * read-time prep: jitter, strand mask, shifts, seeds   _dataset/_query.py:153-204, _haps.py:678-768,
                                                         _reconstruct.py:168-226
  * output shaping (to_fixed / pad / ragged, squeeze)    _dataset/_query.py:94-127

Host work per call is O(batch) numpy (index maths, RNG draws in the reference's order); everything
sample-scale lives on the GPU inside `Engine`.  Results are torch CUDA tensors.
Not supported (outside the hot-path scope, raise): splicing, `variants` / `variant-windows`, AF filters.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Literal

import numpy as np
import torch

from ._engine import Engine
from ._insertion_fill import InsertionFill, Repeat5p, lower
from ._types import (AnnotatedHaps, DummyVariant, Ragged, RaggedAlleles, RaggedAnnotatedHaps, RaggedVariants, VarWindowOpt,
                     build_token_lut)

SeqKind = Literal["reference", "haplotypes", "annotated"]
N_CHAR = ord("N")


def _idx_to_array(idx, n: int) -> np.ndarray:
    if isinstance(idx, slice):
        return np.arange(n, dtype=np.int64)[idx]
    a = np.asarray(idx)
    if a.dtype == np.bool_:
        if a.shape != (n,):
            raise IndexError(f"boolean index of shape {a.shape} does not match axis of size {n}")
        return np.flatnonzero(a).astype(np.int64)
    a = a.astype(np.int64)
    if ((a < -n) | (a >= n)).any():
        raise IndexError(f"index out of bounds for axis of size {n}")
    return np.where(a < 0, a + n, a)


@dataclass(frozen=True)
class Dataset:
    """Effective shape `(n_regions, n_samples, [tracks], [ploidy], output_length)`; indexable on the
    first two axes like a 2-D array (reference docstring, _impl.py:80-115)."""

    engine: Engine
    full_regions: np.ndarray            # int32 (R, 4): contig_idx, start, end, strand
    sample_names: tuple
    ploidy: int
    max_jitter: int = 0
    track_kinds: dict = field(default_factory=dict)   # name -> "sample" | "annot"
    # ---- settings (reference defaults: _impl.py:119-162) ----
    output_length: object = "ragged"                  # "ragged" | "variable" | int
    sequence_type: object = "haplotypes"              # "reference" | "haplotypes" | "annotated" | None
    active_tracks: tuple = ()
    insertion_fill: dict = field(default_factory=dict)
    jitter: int = 0
    deterministic: bool = True
    rc_neg: bool = True
    var_filter: object = None                         # None | "exonic"
    realign_tracks: bool = True
    output_format: str = "ragged"                     # "ragged" | "flat"  (with_output_format)
    encoding: str = "bytes"                           # "bytes" | "onehot" | "onehot_cf"  (this build's fused one-hot)
    var_fields: tuple = ("alt", "ilen", "start")      # fields of the "variants" output (_haps.py:268)
    dummy_variant: object = None                      # DummyVariant for empty (region, sample, ploid) groups, or None
    unphased_union: bool = False                      # fold the ploidy rows of a (region, sample) into one (_flat_variants.py:925-938)
    min_af: object = None                             # AF filter of the "variants" output (_flat_variants.py:899-923)
    max_af: object = None
    window_opt: object = None                         # VarWindowOpt of with_seqs("variant-windows", ...)
    flank_length: int = 0                             # "variants": flank tokens around every variant (with a token alphabet)
    token_alphabet: object = None                     # bytes; with unknown_token defines the byte -> token table
    unknown_token: object = None
    rng: np.random.Generator = field(default_factory=np.random.default_rng)
    region_subset: object = None
    sample_subset: object = None
    region_map: object = None                         # input-BED row -> storage row (datasets opened from disk)
    splice_rows: object = None                        # spliced: tuple of int64 arrays of STORAGE region indices, one per splice row
    splice_names: object = None                       # spliced: names of the splice rows (or None)
    bed_columns: object = None                        # opened datasets: extra input-BED columns (name -> array, input order)
    _cache: dict = field(default_factory=dict, compare=False, repr=False)  # per-state lazies (fixed-length pipeline)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, device, reference, ref_offsets, v_starts, ilens, alt_alleles, alt_offsets, geno_v_idxs,
                    geno_offsets, regions, n_samples: int, ploidy: int, max_jitter: int = 0, tracks: dict | None = None,
                    track_kinds: dict | None = None, sample_names=None, rng=None, ref_alleles=None,
                    variant_info: dict | None = None, dosages=None) -> "Dataset":
        """In-memory dataset (the GPU counterpart of `get_dummy_dataset`, python/genvarloader/_dummy.py).
        `tracks`: name -> (itv_starts, itv_ends, itv_values, itv_offsets); SAMPLE tracks have one interval slot
        per (region, sample), ANNOT tracks one per region (_reconstruct.py:233-236).  `ref_alleles` = (bytes, offsets) of
        the REF strings, `variant_info` = {name: numeric column, one value per variant} and `dosages` (float32, parallel to
        `geno_v_idxs`) feed the "variants" output (`var_fields`, AF filter)."""
        eng = Engine(device, reference, ref_offsets, v_starts, ilens, alt_alleles, alt_offsets, geno_v_idxs, geno_offsets)
        if ref_alleles is not None or variant_info or dosages is not None:
            eng.set_variant_fields(ref_alleles, variant_info, dosages)
        kinds = {}
        for name, t in (tracks or {}).items():
            eng.add_track(name, *t)
            kinds[name] = (track_kinds or {}).get(name, "sample")
        regions = np.ascontiguousarray(regions, np.int32)
        if regions.shape[1] == 3:
            regions = np.concatenate([regions, np.ones((len(regions), 1), np.int32)], 1)
        names = tuple(sample_names) if sample_names is not None else tuple(f"s{i}" for i in range(n_samples))
        return cls(engine=eng, full_regions=regions, sample_names=names, ploidy=int(ploidy), max_jitter=int(max_jitter),
                   track_kinds=kinds, active_tracks=tuple(kinds), rng=np.random.default_rng(rng))

    @classmethod
    def from_synth(cls, device, d, rng=None, svar2=None) -> "Dataset":
        """From a `genvarloader_b200.synth.SynthData`; `svar2`: a `synth.to_svar2_dataset(d)` dict = keep the variants as
        a resident svar2 two-channel source instead of the SVAR1 CSR."""
        if svar2 is not None:
            return cls.from_svar2_arrays(device, d.reference, d.ref_offsets, svar2, d.regions, d.n_samples, d.ploidy,
                                         d.max_jitter, d.tracks, rng=rng)
        return cls.from_arrays(device, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
                               d.geno_v_idxs, d.geno_offsets, d.regions, d.n_samples, d.ploidy, d.max_jitter, d.tracks,
                               rng=rng)

    @classmethod
    def from_svar2_arrays(cls, device, reference, ref_offsets, sv: dict, regions, n_samples: int, ploidy: int,
                          max_jitter: int = 0, tracks: dict | None = None, track_kinds: dict | None = None,
                          sample_names=None, rng=None) -> "Dataset":
        """In-memory dataset over a resident svar2 two-channel variant source -- the GPU counterpart of `Svar2Haps`
        (python/genvarloader/_dataset/_svar2_haps.py:183): per-(region, sample, ploid) var_key ranges, per-region dense
        windows and presence bits (docs/source/format.md:88-96) stay in HBM, every read merges the two channels on the
        device (src/svar2/mod.rs:45-66).  `sv`: see `synth.to_svar2_dataset` for the arrays."""
        regions = np.ascontiguousarray(regions, np.int32)
        eng = Engine.from_svar2(device, reference, ref_offsets, sv, len(regions) * int(n_samples) * int(ploidy))
        kinds = {}
        for name, t in (tracks or {}).items():
            eng.add_track(name, *t)
            kinds[name] = (track_kinds or {}).get(name, "sample")
        if regions.shape[1] == 3:
            regions = np.concatenate([regions, np.ones((len(regions), 1), np.int32)], 1)
        names = tuple(sample_names) if sample_names is not None else tuple(f"s{i}" for i in range(n_samples))
        return cls(engine=eng, full_regions=regions, sample_names=names, ploidy=int(ploidy), max_jitter=int(max_jitter),
                   track_kinds=kinds, active_tracks=tuple(kinds), rng=np.random.default_rng(rng))

    @classmethod
    def open(cls, path, reference=None, device="cuda", jitter: int = 0, rng=None, deterministic: bool = True,
             rc_neg: bool = True, svar=None) -> "Dataset":
        """Open a dataset directory written by `gvl.write` (reference `Dataset.open`, _impl.py:164-226 ->
        `_open.py:62-345`; layout: docs/source/format.md:8-49) and upload it to `device`.

        `reference`: FASTA path (plain / gzip / bgzip) or a `genvarloader_b200.Reference`; required when the dataset
        has genotypes.  Regions are addressed in the ORDER OF THE INPUT BED, like the reference (`r_idx_map`).
        Datasets linked to a `.svar` (SVAR1) store are followed (`svar=` overrides the stored path, like the reference);
        `.svar2`-backed datasets are refused (third-party store format)."""
        from ._open import Reference, read_dataset_arrays

        a = read_dataset_arrays(path, reference, svar)
        ref = a["reference"]
        n_regions, samples = len(a["full_regions"]), a["samples"]
        has_geno = "geno_v_idxs" in a
        ploidy = int(a["ploidy"]) if has_geno else 1
        if has_geno and ref is None:
            raise ValueError("this dataset has genotypes: pass `reference=` (FASTA path or Reference) to reconstruct haplotypes")
        if ref is None:  # tracks only: contig extents are never read
            ref = Reference.from_arrays(np.zeros(0, np.uint8), np.zeros(len(a["contigs"]) + 1, np.int64), a["contigs"])
        if has_geno:
            eng = Engine(device, ref.reference, ref.offsets, a["v_starts"], a["ilens"], a["alt_alleles"], a["alt_offsets"],
                         a["geno_v_idxs"], a["geno_offsets"], pad_char=ref.pad_char)
        else:
            eng = Engine(device, ref.reference, ref.offsets, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8),
                         np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(n_regions * len(samples) * ploidy + 1, np.int64),
                         pad_char=ref.pad_char)
        if has_geno and (a.get("ref_alleles") is not None or a.get("variant_info")):
            eng.set_variant_fields(a.get("ref_alleles"), a.get("variant_info"))  # "variants" output: REF strings, INFO columns
        for name, t in a["tracks"].items():
            eng.add_track(name, *t)
        ds = cls(engine=eng, full_regions=a["full_regions"], sample_names=tuple(samples), ploidy=ploidy,
                 max_jitter=a["max_jitter"], track_kinds=dict(a["track_kinds"]), active_tracks=tuple(a["track_kinds"]),
                 sequence_type="haplotypes" if has_geno else None, rng=np.random.default_rng(rng),
                 region_map=np.ascontiguousarray(a["region_map"], np.int64), bed_columns=a.get("bed_columns"))
        return ds.with_settings(jitter=jitter, deterministic=deterministic, rc_neg=rc_neg)

    # ------------------------------------------------------------------ properties (reference: _impl.py:954-1110)
    @property
    def n_regions(self) -> int:
        """Number of regions -- of SPLICE ROWS when the dataset is spliced (reference _impl.py:1001-1005)."""
        return len(self.splice_rows) if self.splice_rows is not None else len(self._r_idx)

    @property
    def n_samples(self) -> int:
        return len(self._s_idx)

    @property
    def samples(self) -> list:
        return [self.sample_names[i] for i in self._s_idx]

    @property
    def shape(self) -> tuple:
        return (self.n_regions, self.n_samples)

    @property
    def full_shape(self) -> tuple:
        return (len(self.full_regions), len(self.sample_names))

    @property
    def available_var_fields(self) -> list:
        """Reference: `Haps.available_var_fields`, _haps.py:313-319."""
        eng = self.engine
        return (["alt", "ilen", "start"] + (["ref"] if eng.ref_alleles is not None else [])
                + (["dosage"] if eng.dosages is not None else []) + sorted(eng.var_info))

    @property
    def available_tracks(self) -> list:
        return list(self.track_kinds)

    @property
    def is_subset(self) -> bool:
        return self.region_subset is not None or self.sample_subset is not None

    # ---- what the dataset holds (reference: _impl.py:954-1110, 1234-1340, 1848-1895) ----
    @property
    def has_reference(self) -> bool:
        return True  # (an Engine always holds a reference; datasets without one are not opened here)

    @property
    def has_genotypes(self) -> bool:
        return self.engine.svar2 is not None or int(self.engine.alt_offsets.numel()) > 1  # (a variant table with entries)

    @property
    def has_intervals(self) -> bool:
        return bool(self.track_kinds)

    def __repr__(self) -> str:
        return (f"GVL dataset (B200-resident) with shape {self.shape}: sequences={self.sequence_type!r}, tracks={list(self.active_tracks)}, "
                f"output_length={self.output_length!r}, jitter={self.jitter} (max {self.max_jitter}), deterministic={self.deterministic}, "
                f"rc_neg={self.rc_neg}, is_subset={self.is_subset}, is_spliced={self.is_spliced}")

    def _shape_counts(self, x: np.ndarray, squeeze: bool, out_reshape):
        if squeeze:
            x = x.squeeze(0)
        if out_reshape is not None:
            x = x.reshape(*out_reshape, x.shape[-1])
        return x

    def n_variants(self, regions=None, samples=None) -> np.ndarray:
        """`Dataset.n_variants`, _impl.py:1293-1340: variants per (region, sample, ploid) -> int32 (..., ploidy)."""
        ds_idx, squeeze, out_reshape = self._parse_idx((slice(None) if regions is None else regions,
                                                        slice(None) if samples is None else samples))
        p = self.ploidy
        goi = ds_idx[:, None] * p + np.arange(p, dtype=np.int64)[None, :]
        if self.engine.svar2 is not None:
            raise NotImplementedError("n_variants of an svar2-backed dataset needs the presence-bit popcounts (not kept on the host)")
        go = self.engine.geno_offsets_host
        n = np.maximum(go[1, goi] - go[0, goi], 0).astype(np.int32)
        return self._shape_counts(n, squeeze, out_reshape)

    def n_intervals(self, regions=None, samples=None) -> np.ndarray:
        """`Dataset.n_intervals`, _impl.py:1848-1895: stored intervals per (region, sample) and ACTIVE track -> int32 (..., tracks)."""
        ds_idx, squeeze, out_reshape = self._parse_idx((slice(None) if regions is None else regions,
                                                        slice(None) if samples is None else samples))
        if not self.active_tracks:
            return self._shape_counts(np.zeros((len(ds_idx), 0), np.int32), squeeze, out_reshape)
        r_idx = ds_idx // len(self.sample_names)
        cols = []
        for name in self.active_tracks:
            off = self._cache.get(("itv_off", name))
            if off is None:
                off = self._cache[("itv_off", name)] = self.engine.tracks[name][3].cpu().numpy()
            slot = r_idx if self.track_kinds[name] == "annot" else ds_idx
            cols.append((off[slot + 1] - off[slot]).astype(np.int32))
        return self._shape_counts(np.stack(cols, -1), squeeze, out_reshape)

    def haplotype_lengths(self, regions=None, samples=None):
        """`Dataset.haplotype_lengths`, _impl.py:1234-1291: lengths of the JITTER-EXTENDED haplotypes -> int32 (..., ploidy);
        None when the lengths are not a fixed property of the dataset (no genotypes, random shifts)."""
        if self.splice_rows is not None:
            raise NotImplementedError("Haplotype lengths are not yet implemented for spliced datasets.")
        if not self.has_genotypes or not self.deterministic:
            return None
        ds_idx, squeeze, out_reshape = self._parse_idx((slice(None) if regions is None else regions,
                                                        slice(None) if samples is None else samples))
        eng, dev, p = self.engine, self.engine.device, self.ploidy
        reg = self.full_regions[ds_idx // len(self.sample_names)].copy()
        reg[:, 1] -= self.jitter
        reg[:, 2] += self.jitter
        goi = ds_idx[:, None] * p + np.arange(p, dtype=np.int64)[None, :]
        t_reg = torch.from_numpy(np.ascontiguousarray(reg[:, :3])).to(dev)
        t_goi = torch.from_numpy(goi).to(dev)
        keep = keep_off = None
        if self.var_filter == "exonic":
            keep, keep_off = self._exonic_keep(t_goi, t_reg, goi)
        d = eng.get_diffs(t_goi, t_reg[:, 1].contiguous(), t_reg[:, 2].contiguous(), keep, keep_off, regions=t_reg,
                          max_records=eng.max_records(goi)).cpu().numpy()
        lens = ((reg[:, 2] - reg[:, 1])[:, None] + d).astype(np.int32)
        return self._shape_counts(lens, squeeze, out_reshape)

    def to_torch_dataset(self, return_indices: bool = False, transform=None):
        """`Dataset.to_torch_dataset`, _impl.py:1940-1961 -> `TorchDataset`, _torch.py:272-307: a map-style dataset over the flat
        (region, sample) index; an int or a list of ints selects a batch (the batches are CUDA tensors)."""
        return _MapDataset(self, bool(return_indices), transform)

    @property
    def regions(self) -> np.ndarray:
        return self.full_regions[self._r_idx]

    def __len__(self) -> int:
        return self.n_regions * self.n_samples

    @property
    def _r_idx(self) -> np.ndarray:
        v = self._cache.get("r_idx")
        if v is None:
            if self.region_subset is not None:
                v = self.region_subset
            else:
                v = np.arange(len(self.full_regions)) if self.region_map is None else self.region_map
            v = self._cache["r_idx"] = np.ascontiguousarray(v, np.int64)
        return v

    @property
    def _s_idx(self) -> np.ndarray:
        v = self._cache.get("s_idx")
        if v is None:
            v = np.arange(len(self.sample_names)) if self.sample_subset is None else self.sample_subset
            v = self._cache["s_idx"] = np.ascontiguousarray(v, np.int64)
        return v

    # ------------------------------------------------------------------ with_* (immutable evolution)
    def _check_valid_state(self) -> None:
        """Same checks and messages as the reference's `_check_valid_state` (_impl.py:501-568) for this scope."""
        if self.jitter < 0:
            raise ValueError(f"Jitter ({self.jitter}) must be a non-negative integer.")
        if self.jitter > self.max_jitter:
            raise ValueError(f"Jitter ({self.jitter}) must be less than or equal to the maximum jitter of the dataset ({self.max_jitter}).")
        if isinstance(self.output_length, (int, np.integer)) and not isinstance(self.output_length, bool):
            if self.output_length < 1:
                raise ValueError(f"Output length ({self.output_length}) must be a positive integer.")
            min_r_len = int((self.full_regions[:, 2] - self.full_regions[:, 1]).min())
            max_output_length = min_r_len + 2 * self.max_jitter
            eff_length = int(self.output_length) + 2 * self.jitter
            if eff_length > max_output_length:
                raise ValueError(
                    f"Effective length (out_len={self.output_length}) + 2 * ({self.jitter=}) = {eff_length} must be less"
                    f" than or equal to the maximum output length of the dataset ({max_output_length})."
                    f" The maximum output length is the minimum region length ({min_r_len}) + 2 * (max_jitter={self.max_jitter}).")
        elif self.output_length not in ("ragged", "variable"):
            raise ValueError(f"Output length must be 'ragged', 'variable' or a positive integer, got {self.output_length!r}")
        if self.unphased_union and self.sequence_type in ("haplotypes", "annotated"):
            raise ValueError("unphased_union is incompatible with 'haplotypes'/'annotated' output (a union of phased sequences is "
                             "ill-defined). Use 'variant-windows' or 'variants', or clear the flag with "
                             "with_settings(unphased_union=False).")  # _impl.py:769-779
        if (self.min_af is not None or self.max_af is not None) and self.sequence_type not in ("variants", "variant-windows"):
            raise NotImplementedError("Filtering by AF is not supported for haplotype output yet.")  # _haps.py:695-698
        if self.encoding != "bytes" and self.sequence_type not in ("haplotypes", "reference"):
            raise ValueError("one-hot encoding applies to 'haplotypes' / 'reference' sequences only")
        if self.encoding == "onehot_cf":
            if not isinstance(self.output_length, (int, np.integer)):
                raise ValueError("channels-first one-hot needs a fixed output length")
            if int(self.output_length) % 4:
                raise ValueError("channels-first one-hot needs an output length that is a multiple of 4")

    def _evolve(self, **kw) -> "Dataset":
        ds = replace(self, **kw, _cache={})
        ds._check_valid_state()
        return ds

    def with_settings(self, jitter=None, rng=None, deterministic=None, rc_neg=None, var_filter=None, realign_tracks=None,
                      min_af=None, max_af=None, splice_info=None, **unsupported) -> "Dataset":
        """Reference: `Dataset.with_settings`, _impl.py:228-499 (hot-path settings only)."""
        kw = {}
        if min_af is not None:
            kw["min_af"] = None if min_af is False else float(min_af)
        if max_af is not None:
            kw["max_af"] = None if max_af is False else float(max_af)
        ta, ut = unsupported.pop("token_alphabet", None), unsupported.pop("unknown_token", None)
        if ta is not None or ut is not None:  # _impl.py:432-441
            if ta is None or ut is None:
                raise ValueError("token_alphabet and unknown_token must be set together.")
            kw["token_alphabet"] = ta.encode("ascii") if isinstance(ta, str) else bytes(ta)
            kw["unknown_token"] = int(ut)
        fl = unsupported.pop("flank_length", None)
        if fl is not None:
            if fl is not False and kw.get("token_alphabet", self.token_alphabet) is None:
                raise ValueError("flank_length requires a token LUT; pass token_alphabet and unknown_token to with_settings(...)"
                                 " (in this or a prior call).")  # _impl.py:442-446
            kw["flank_length"] = 0 if fl is False else int(fl)
        for name in ("var_fields", "dummy_variant", "unphased_union"):
            if name in unsupported:
                v = unsupported.pop(name)
                if v is None:
                    continue
                if name == "var_fields":
                    avail = self.available_var_fields
                    missing = sorted(set(v) - set(avail))
                    if missing:
                        raise ValueError(f"Missing variant fields: {missing}")  # _impl.py:343-346
                    v = tuple(v)
                elif name == "dummy_variant":
                    v = None if v is False else v
                else:
                    v = bool(v)
                kw[name] = v
        if splice_info is False:
            kw["splice_rows"], kw["splice_names"] = None, None
        elif splice_info is not None:
            kw["splice_rows"], kw["splice_names"] = self._build_splice_rows(splice_info)
        if unsupported:
            raise NotImplementedError(f"settings outside the hot-path scope: {sorted(unsupported)}")
        if jitter is not None:
            kw["jitter"] = int(jitter)
        if rng is not None:
            kw["rng"] = np.random.default_rng(rng)
        if deterministic is not None:
            kw["deterministic"] = bool(deterministic)
        if rc_neg is not None:
            kw["rc_neg"] = bool(rc_neg)
        if var_filter not in (None, False) and self.engine.svar2 is not None:
            raise NotImplementedError("var_filter is not supported with the svar2 source")
        if var_filter is not None:
            if var_filter not in (False, "exonic"):
                raise ValueError(f"var_filter must be False or 'exonic', got {var_filter!r}")
            kw["var_filter"] = None if var_filter is False else "exonic"
        if realign_tracks is not None:
            kw["realign_tracks"] = bool(realign_tracks)
        return self._evolve(**kw)

    def with_len(self, output_length) -> "Dataset":
        """Reference: `Dataset.with_len`, _impl.py:570-647."""
        if isinstance(output_length, (int, np.integer)) and not isinstance(output_length, bool):
            if output_length < 1:
                raise ValueError(f"Output length ({output_length}) must be a positive integer.")
            output_length = int(output_length)
        return self._evolve(output_length=output_length)

    def with_seqs(self, kind, window_opt=None) -> "Dataset":
        """Reference: `Dataset.with_seqs`, _impl.py:649-783."""
        if kind in ("variants", "variant-windows") and getattr(self.engine, "svar2", None) is not None:
            raise NotImplementedError(f"with_seqs({kind!r}) needs the SVAR1 genotype CSR (the svar2 source decodes variants "
                                      "through its store)")
        if kind not in (None, "reference", "haplotypes", "annotated", "variants", "variant-windows"):
            raise ValueError(f"Unknown sequence type {kind!r}")
        kw = {}
        if kind == "variant-windows":
            if not isinstance(window_opt, VarWindowOpt):
                raise ValueError("with_seqs('variant-windows') requires window_opt=VarWindowOpt(...)")  # _impl.py:741-747
            kw["window_opt"] = window_opt
        elif window_opt is not None:
            raise ValueError("window_opt only applies to with_seqs('variant-windows')")
        enc = self.encoding if kind in ("haplotypes", "reference") else "bytes"
        return self._evolve(sequence_type=kind, encoding=enc, **kw)

    def with_encoding(self, encoding: str) -> "Dataset":
        """Fused one-hot output (this build's extension; the reference leaves one-hot to `seqpro.DNA.ohe`,
        docs/source/index.md:108-119).  "onehot": uint8 (..., L, 4); "onehot_cf": uint8 (..., 4, L); alphabet ACGT."""
        if encoding not in ("bytes", "onehot", "onehot_cf"):
            raise ValueError(f"Unknown encoding {encoding!r}")
        return self._evolve(encoding=encoding)

    def with_tracks(self, tracks=None, kind=None) -> "Dataset":
        """Reference: `Dataset.with_tracks`, _impl.py:785-839.  `False`/`[]` disables tracks."""
        if kind not in (None, "tracks"):
            raise NotImplementedError("only kind='tracks' (base-pair resolution) is in the hot-path scope")
        if tracks is None:
            names = tuple(self.track_kinds)
        elif tracks is False:
            names = ()
        else:
            names = (tracks,) if isinstance(tracks, str) else tuple(tracks)
        missing = [t for t in names if t not in self.track_kinds]
        if missing:
            raise ValueError(f"Track(s) {missing} not found. Available tracks: {self.available_tracks}")
        return self._evolve(active_tracks=names)

    def with_insertion_fill(self, strategy) -> "Dataset":
        """Reference: `Dataset.with_insertion_fill`, _impl.py:841-878 (one strategy for all tracks or a dict)."""
        if not self.track_kinds:
            raise ValueError("Dataset has no tracks; cannot configure insertion fill.")
        if self.sequence_type not in ("haplotypes", "annotated"):
            raise ValueError("with_insertion_fill is only meaningful for datasets with both haplotypes and tracks "
                             "(use with_seqs to activate haplotypes first).")
        if not self.active_tracks:
            raise ValueError("with_insertion_fill is only meaningful when tracks are active (use with_tracks to activate tracks first).")
        if not self.realign_tracks:
            raise ValueError("with_insertion_fill has no effect when realign_tracks=False (insertion fill only applies during "
                             "track re-alignment). Set with_settings(realign_tracks=True) first, or drop the call.")
        if isinstance(strategy, InsertionFill):
            fills = {name: strategy for name in self.track_kinds}
        else:
            fills = dict(self.insertion_fill)
            for name, s in dict(strategy).items():
                if name not in self.track_kinds:
                    raise ValueError(f"Track {name!r} not found. Available tracks: {self.available_tracks}")
                if not isinstance(s, InsertionFill):
                    raise TypeError("strategies must be InsertionFill instances")
                fills[name] = s
        return self._evolve(insertion_fill=fills)

    def with_output_format(self, fmt: str) -> "Dataset":
        """Reference: `Dataset.with_output_format`, _impl.py:880-952.  "flat" returns `Ragged` triples untouched."""
        if fmt not in ("ragged", "flat"):
            raise ValueError(f"Unknown output format {fmt!r}")
        return self._evolve(output_format=fmt)

    def subset_to(self, regions=None, samples=None) -> "Dataset":
        """Reference: `Dataset.subset_to`, _impl.py:1153-1221 (integer / slice / boolean / sample-name selectors)."""
        kw = {}
        if regions is not None:
            if self.splice_rows is not None:
                raise ValueError("subset_to(regions=...) on a spliced dataset would mix splice rows with regions: subset first, "
                                 "then with_settings(splice_info=...)")
            kw["region_subset"] = self._r_idx[_idx_to_array(regions, self.n_regions)]
        if samples is not None:
            if isinstance(samples, str) or (np.ndim(samples) > 0 and len(samples) and isinstance(samples[0], str)):
                names = [samples] if isinstance(samples, str) else list(samples)
                lut = {n: i for i, n in enumerate(self.sample_names)}
                try:
                    kw["sample_subset"] = np.array([lut[n] for n in names], np.int64)
                except KeyError as e:
                    raise KeyError(f"Sample {e.args[0]!r} not found") from None
            else:
                kw["sample_subset"] = self._s_idx[_idx_to_array(samples, self.n_samples)]
        return self._evolve(**kw)

    def to_full_dataset(self) -> "Dataset":
        return self._evolve(region_subset=None, sample_subset=None)

    # ------------------------------------------------------------------ indexing
    def _parse_idx(self, idx):
        """`DatasetIndexer.parse_idx`, _indexing.py:208-264: flat dataset indices (row-major over the FULL
        (regions, samples) grid), squeeze flag, optional outer reshape -- "basic" (ints / slices), "adv" (two
        arrays, paired) and "combo" (one array, one basic -> outer product) indexing."""
        if not isinstance(idx, tuple):
            r, s = idx, slice(None)
        elif len(idx) == 1:
            r, s = idx[0], slice(None)
        elif len(idx) == 2:
            r, s = idx
        else:
            raise IndexError("a Dataset is indexed by (regions, samples)")
        if (type(r) is np.ndarray and type(s) is np.ndarray and r.ndim == 1 and s.shape == r.shape
                and r.dtype.kind in "iu" and s.dtype.kind in "iu"):
            # the loader's case, two paired 1-D integer arrays: fancy indexing does the bounds check, the negative wrap
            # and the subset mapping in one step each
            flat = self._r_idx[r] * len(self.sample_names) + self._s_idx[s]
            return (flat if flat.dtype == np.int64 else flat.astype(np.int64)), False, None
        is_basic = lambda x: isinstance(x, (int, np.integer, slice))
        is_int = lambda x: isinstance(x, (int, np.integer))
        n_s = len(self.sample_names)
        r_raw = self._r_idx[_idx_to_array(r, self.n_regions)]
        s_raw = self._s_idx[_idx_to_array(s, self.n_samples)]
        squeeze, out_reshape = False, None
        if is_basic(r) and is_basic(s):
            squeeze = is_int(r) and is_int(s)
            ri, si = np.atleast_1d(r_raw), np.atleast_1d(s_raw)
            flat = (ri[:, None] * n_s + si[None, :]).squeeze()
            if isinstance(r, slice) and isinstance(s, slice):
                out_reshape = (len(ri), len(si))
            if flat.ndim > 1:
                out_reshape = flat.shape
        elif not is_basic(r) and not is_basic(s):
            ri, si = np.broadcast_arrays(r_raw, s_raw)
            flat = ri * n_s + si
            if flat.ndim > 1:
                out_reshape = flat.shape
        else:
            flat = r_raw.ravel()[:, None] * n_s + s_raw.ravel()[None, :]
            if r_raw.ndim > 1 or s_raw.ndim > 1:
                out_reshape = (*r_raw.shape, *s_raw.shape)
            else:
                out_reshape = flat.shape
        return np.asarray(flat, np.int64).ravel(), squeeze, out_reshape

    # ------------------------------------------------------------------ splicing (reference: _splice.py, _query.py:206-330)
    def _build_splice_rows(self, splice_info):
        """Splice rows = ordered lists of regions (e.g. the exons of a transcript) whose haplotypes are concatenated.

        `splice_info` follows the reference (`SpliceMap.from_bed`, _splice.py:163-230) for datasets opened from disk:
        a column name of the input BED (rows with the same value form a splice row, in BED order) or a
        `(id_column, order_column)` pair (elements sorted by the second column).  In-memory datasets -- and any
        dataset -- may pass the mapping directly: `{name: [region indices]}` or a sequence of index sequences
        (indices address regions like `ds[r, s]` does)."""
        r_map = self._r_idx
        names = None
        by_column = isinstance(splice_info, str) or (isinstance(splice_info, tuple) and len(splice_info) == 2
                                                     and all(isinstance(x, str) for x in splice_info))
        if by_column:
            cols = self.bed_columns or {}
            id_col, order_col = (splice_info, None) if isinstance(splice_info, str) else splice_info
            for c in (id_col, order_col):
                if c is not None and c not in cols:
                    raise ValueError(f"column {c!r} is not in the dataset's input BED (available: {sorted(cols)})")
            ids = np.asarray(cols[id_col])
            groups: dict = {}
            for i, v in enumerate(ids.tolist()):
                groups.setdefault(v, []).append(i)
            if order_col is not None:
                order = np.asarray(cols[order_col])
                groups = {k: sorted(v, key=lambda i: order[i]) for k, v in groups.items()}
            # BED columns are in INPUT order: rows map to storage through region_map (not through a subset); with a region
            # subset active, elements outside it drop out and rows left empty disappear
            full_map = self.region_map if self.region_map is not None else np.arange(len(self.full_regions), dtype=np.int64)
            allowed = None if self.region_subset is None else set(np.asarray(self.region_subset).tolist())
            names, rows = [], []
            for k, v in groups.items():
                st = [int(full_map[i]) for i in v]
                if allowed is not None:
                    st = [x for x in st if x in allowed]
                if st:
                    names.append(k)
                    rows.append(np.asarray(st, np.int64))
            if not rows:
                raise ValueError("no splice row is left after subsetting")
            return tuple(rows), tuple(names)
        elif isinstance(splice_info, dict):
            names, lists = list(splice_info), list(splice_info.values())
        else:
            lists = list(splice_info)
        rows = []
        for l in lists:
            a = _idx_to_array(np.asarray(l, np.int64), len(r_map))
            if a.ndim != 1 or a.size == 0:
                raise ValueError("every splice row needs at least one region index")
            rows.append(np.ascontiguousarray(r_map[a], np.int64))
        return tuple(rows), (tuple(names) if names is not None else None)

    @property
    def is_spliced(self) -> bool:
        return self.splice_rows is not None

    def _getitem_spliced(self, idx):
        """One ragged haplotype per (splice row, sample, ploid): the row's elements are reconstructed as separate
        kernel rows, laid out in (row, sample, ploid, element) order (the reference's permuted write,
        _splice.py:56-160 / _haps.py:872-919), so the concatenation is free and the group offsets are the kernel's
        row offsets sampled at cell boundaries.  Negative-strand elements are reverse-complemented one by one
        (_query.py:269-288); jitter and fixed lengths do not apply to spliced output."""
        if self.sequence_type not in ("haplotypes", "annotated"):
            raise NotImplementedError("spliced output is implemented for 'haplotypes' and 'annotated' sequences")
        if self.jitter or isinstance(self.output_length, (int, np.integer)):
            raise ValueError("spliced datasets return ragged haplotypes: use with_len('ragged') and jitter=0")
        rows_all = self.splice_rows
        n_rows_all, n_s_all = len(rows_all), len(self._s_idx)
        if not isinstance(idx, tuple):
            idx = (idx, slice(None))
        r_sel, s_sel = idx
        ri = np.atleast_1d(_idx_to_array(r_sel, n_rows_all))
        si = np.atleast_1d(self._s_idx[_idx_to_array(s_sel, n_s_all)])
        is_int = lambda x: isinstance(x, (int, np.integer))
        is_basic = lambda x: isinstance(x, (int, np.integer, slice))
        if not is_basic(r_sel) and not is_basic(s_sel):  # paired
            ri, si = np.broadcast_arrays(ri, si)
            pairs = list(zip(ri.ravel().tolist(), si.ravel().tolist()))
            outer = (len(pairs),)
        else:
            pairs = [(r, s_) for r in ri.ravel().tolist() for s_ in si.ravel().tolist()]
            outer = (ri.size, si.size)
        eng, dev, p, S = self.engine, self.engine.device, self.ploidy, len(self.sample_names)
        q_reg, q_goi, cell_len = [], [], []
        for r, s_ in pairs:
            elems = rows_all[r]
            for e in range(p):
                q_reg.append(elems)
                q_goi.append((elems * S + s_) * p + e)
                cell_len.append(len(elems))
        q_reg = np.concatenate(q_reg)
        goi = np.concatenate(q_goi)[:, None].astype(np.int64)
        regions = self.full_regions[q_reg]
        to_rc = (regions[:, 3] == -1) if self.rc_neg else None
        pk = _Packer(dev)
        i_reg, i_goi = pk.add(regions[:, :3], np.int32), pk.add(goi, np.int64)
        i_sh = pk.add(np.zeros((len(goi), 1), np.int32), np.int32)
        i_rc = pk.add(to_rc, np.uint8) if to_rc is not None else None
        cell_starts = np.concatenate([[0], np.cumsum(cell_len)]).astype(np.int64)
        i_cs = pk.add(cell_starts, np.int64)
        pk.upload()
        keep = keep_off = None
        if self.var_filter == "exonic":
            keep, keep_off = self._exonic_keep(pk.get(i_goi), pk.get(i_reg), goi)
        oo = eng.plan(pk.get(i_reg), pk.get(i_sh), pk.get(i_goi), -1, eng.max_records(goi), keep, keep_off,
                      pk.get(i_rc) if i_rc is not None else None)
        total = eng.total()
        group_offsets = oo[pk.get(i_cs)]
        shape = (*outer, p, None)
        if self.sequence_type == "annotated":
            h, av, ap = eng.execute("annotated")
            out = RaggedAnnotatedHaps(Ragged(h, group_offsets, shape), Ragged(av, group_offsets, shape), Ragged(ap, group_offsets, shape))
        elif self.encoding == "onehot":
            out = Ragged(eng.execute("onehot").view(total, 4), group_offsets, shape)
        elif self.encoding == "bytes":
            out = Ragged(eng.execute("haplotypes"), group_offsets, shape)
        else:
            raise ValueError("channels-first one-hot needs a fixed output length; spliced output is ragged")
        if self.output_length == "variable" and self.output_format != "flat":  # pad like the unspliced path (_query.py:94-127)
            out = self._shape_output(out, None, False)
        if is_int(r_sel) and is_int(s_sel):
            out = out.squeeze(0).squeeze(0) if hasattr(out, "squeeze") else out
        return out

    # ------------------------------------------------------------------ iteration (reference: to_dataloader, _impl.py:1963-2072)
    def to_dataloader(self, batch_size: int = 1, shuffle: bool = False, sampler=None, num_workers: int = 0, collate_fn=None,
                      pin_memory: bool = False, drop_last: bool = False, generator=None, *, return_indices: bool = False,
                      transform=None, mode=None, copy: bool = True, ring: int = 0, to_host: bool = False, **ignored):
        """Batches over the flat `(region, sample)` index like the reference's DataLoader (which wraps its sampler in a
        `BatchSampler` so that the dataset is indexed with lists of indices, _torch.py:160-228).  The batches are born on
        the GPU, so there are no workers, no pinning and no collation: `num_workers`, `pin_memory`, `collate_fn`,
        `buffer_bytes` ... are accepted and ignored.  `generator`: seed / numpy Generator / torch.Generator for `shuffle`.

        `mode=None`: one synchronous `ds[r, s]` per batch (any output kind).
        `mode="buffered"` / `"double_buffered"` (reference: `_impl.py:2029-2046`, prefetching producers): the GPU-side
        analogue for fixed-length output -- the loader reads `ring` batches ahead with ONE device call per ring (0: sized
        for ~256 MiB of output), two ring halves produced alternately while the other is consumed (`_pipeline.py`).  `copy=False` hands out zero-copy
        views into the ring that stay valid until `ring` more batches have been drawn (the reference's `copy=False`
        contract: "only valid until the next batch is yielded"); `copy=True` (default) clones every batch.
        `to_host=True` (pipelined modes): batches are delivered as NUMPY arrays in pinned host memory -- every ring is copied
        device-to-host on a copy stream while the next one is produced, so a CPU consumer sees the PCIe copy rate instead of
        copy + compute + launch latency per batch (`copy=False`: views valid until the half comes around again).
        Datasets the pipeline cannot serve (ragged / variable lengths, random shifts, splicing, var_filter) fall back to
        `mode=None`."""
        if sampler is not None and shuffle:
            raise ValueError("sampler option is mutually exclusive with shuffle")
        if mode not in (None, "buffered", "double_buffered"):
            raise ValueError(f"Unknown dataloader mode {mode!r}")
        if mode is not None:
            from . import _pipeline

            if _pipeline.supports(self) is None:
                return _pipeline.PipelinedLoader(self, int(batch_size), bool(shuffle), sampler, bool(drop_last), generator,
                                                 bool(return_indices), transform, bool(copy), ring, to_host=bool(to_host))
        return BatchLoader(self, int(batch_size), bool(shuffle), sampler, bool(drop_last), generator, bool(return_indices), transform)

    # ------------------------------------------------------------------ the hot path
    def __getitem__(self, idx):
        """Reference: `Dataset.__getitem__` _impl.py:2074-2121 -> `_query.getitem` _query.py:66-204."""
        if self.sequence_type is None and not self.active_tracks:
            raise ValueError("Dataset has neither sequences nor tracks active.")
        if self.splice_rows is not None:
            if self.sequence_type in ("variants", "variant-windows"):
                raise NotImplementedError("spliced 'variants' output is not built")
            return self._getitem_spliced(idx)
        ds_idx, squeeze, out_reshape = self._parse_idx(idx)
        pipe = self._eager_pipeline(len(ds_idx))
        if pipe is not None:
            # fixed-length fast path: the O(batch) prep runs on the device (gvl_dev_batch_prep), one small H2D copy per call
            jit = self.rng.integers(-self.jitter, self.jitter + 1, size=len(ds_idx), dtype=np.int32) if self.jitter else None
            out = pipe.run_eager(ds_idx, jit)
            if out_reshape is None and not squeeze and type(out) is torch.Tensor:
                return out  # (paired array indices, sequences only: nothing to reshape)
            out = out if isinstance(out, tuple) else (out,)
            out = tuple(self._shape_output(o, out_reshape, squeeze) for o in out)
            return out[0] if len(out) == 1 else out
        S = len(self.sample_names)
        r_idx, s_idx = ds_idx // S, ds_idx % S

        # ---- _getitem_unspliced, _query.py:161-175 ----
        regions = self.full_regions[r_idx].copy()
        lengths = regions[:, 2] - regions[:, 1]
        jitter_off = self.rng.integers(-self.jitter, self.jitter + 1, size=len(regions), dtype=np.int32)
        regions[:, 1] += jitter_off
        regions[:, 2] = regions[:, 1] + lengths

        out = self._reconstruct(ds_idx, r_idx, s_idx, regions)
        out = tuple(self._shape_output(o, out_reshape, squeeze) for o in out)
        return out[0] if len(out) == 1 else out

    def _eager_pipeline(self, n: int):
        """The fixed-length pipeline serving `__getitem__` (None when this dataset state needs the general path)."""
        c = self._cache
        if "fast" not in c:
            from . import _pipeline

            c["fast"] = _pipeline.supports(self) is None
        if not c["fast"] or n == 0:
            return None
        pipe = c.get("pipe")
        if pipe is None or pipe.b < n:
            from . import _pipeline

            cap = max(n, 2 * pipe.b if pipe is not None else 0)
            pipe = c["pipe"] = _pipeline.FixedPipeline(self, cap, ring=0)
        return pipe

    # ---- reconstructors: Haps / HapsTracks / Tracks (_haps.py:578-870, _reconstruct.py:132-307, _tracks.py:370-420)
    def _reconstruct(self, ds_idx, r_idx, s_idx, regions):
        eng, dev, p = self.engine, self.engine.device, self.ploidy
        b = len(ds_idx)
        fixed = isinstance(self.output_length, (int, np.integer))
        want_seqs = self.sequence_type is not None
        want_tracks = len(self.active_tracks) > 0
        is_ref = self.sequence_type == "reference"
        realign = (want_tracks and want_seqs and not is_ref and self.realign_tracks
                   and self.sequence_type not in ("variants", "variant-windows"))
        to_rc_q = (self.full_regions[r_idx, 3] == -1) if self.rc_neg else None
        lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)

        # geno_offset_idx = ravel (region, sample, ploid), _haps.py:757-768.  "reference" rows use the engine's
        # empty CSR slot: the zero-variant case of the same kernels (get_reference, src/reference/mod.rs:56-120).
        rows_p = 1 if is_ref else p
        if is_ref:
            goi = np.full((b, 1), eng.empty_slot, np.int64)
        else:
            goi = ds_idx[:, None] * p + np.arange(p, dtype=np.int64)[None, :]
        to_rc = None if to_rc_q is None else np.repeat(to_rc_q, rows_p)  # _haps.py:838-843

        pk = _Packer(dev)  # one pinned staging buffer, one H2D copy for all O(batch) arrays
        i_reg = pk.add(regions[:, :3], np.int32)
        i_goi = pk.add(goi, np.int64)
        i_rc = pk.add(to_rc, np.uint8) if to_rc is not None else None
        i_sh = pk.add(np.zeros((b, rows_p), np.int32), np.int32)
        i_tr = i_trc = None
        if want_tracks:
            oi = np.stack([ds_idx if self.track_kinds[n] == "sample" else r_idx for n in self.active_tracks])
            i_tr = pk.add(oi, np.int64)
            if to_rc_q is not None:
                i_trc = pk.add(to_rc_q, np.uint8)
        pk.upload()
        t_reg, t_goi, t_shifts = pk.get(i_reg), pk.get(i_goi), pk.get(i_sh)
        t_rc = pk.get(i_rc) if i_rc is not None else None

        max_rec = 0 if is_ref else eng.max_records(goi)
        keep = keep_off = None
        if self.var_filter == "exonic" and not is_ref:
            keep, keep_off = self._exonic_keep(t_goi, t_reg, goi)

        results = []
        oo = total = diffs = None
        if self.sequence_type in ("variants", "variant-windows"):
            # Haps._get_variants -> get_variants_flat (_haps.py:602-609, _flat_variants.py:869-1112), on the device
            results.append(self._get_variants(t_goi, t_rc, b, regions, max_rec))
            want_seqs = False
        if want_seqs:
            out_len = int(self.output_length) if fixed else -1
            if fixed and not self.deterministic and not is_ref:
                # random shifts need the diffs first (_haps.py:720-730): one extra device pass + sync; the
                # draw itself stays on the host generator so seeded runs follow the reference's RNG order
                dd = eng.get_diffs(t_goi, t_reg[:, 1].contiguous(), t_reg[:, 2].contiguous(), keep, keep_off, regions=t_reg,
                                   max_records=max_rec).cpu().numpy()
                max_shift = dd.clip(min=0) + (lengths - out_len).clip(min=0)[:, None]
                t_shifts = torch.from_numpy(self.rng.integers(0, max_shift + 1, dtype=np.int32)).to(dev)
            if realign:
                diffs = torch.empty((b, rows_p), dtype=torch.int32, device=dev)
            oo = eng.plan(t_reg, t_shifts, t_goi, out_len, max_rec, keep, keep_off, t_rc, diffs=diffs, use_svar2=not is_ref)
            total = eng.total()
            shape = (b, None) if is_ref else (b, p, None)  # Ref has no ploidy axis (_dataset/_ref.py)
            if self.sequence_type == "annotated":
                h, av, ap = eng.execute("annotated")
                results.append(RaggedAnnotatedHaps(Ragged(h, oo, shape), Ragged(av, oo, shape), Ragged(ap, oo, shape)))
            elif self.encoding == "bytes":
                results.append(Ragged(eng.execute("haplotypes"), oo, shape))
            elif self.encoding == "onehot":
                results.append(Ragged(eng.execute("onehot").view(total, 4), oo, shape))
            else:
                cf = eng.execute("onehot_cf")
                results.append(cf.view(b, 4, out_len) if is_ref else cf.view(b, p, 4, out_len))
        if want_tracks:
            names = list(self.active_tracks)
            t = len(names)
            if realign:
                # HapsTracks.__call__, _reconstruct.py:182-300
                lengths_d = torch.from_numpy(lengths).to(dev)
                track_lengths = (lengths_d - diffs.clamp(max=0).min(1).values.to(torch.int64)).to(torch.int32)
                ids, params = lower([self.insertion_fill.get(n, Repeat5p()) for n in names])
                if self.deterministic:
                    base_seed = int(np.bitwise_xor.reduce(ds_idx.astype(np.uint64)))  # :215-218
                else:
                    base_seed = int(self.rng.integers(0, np.iinfo(np.uint64).max, dtype=np.uint64))  # :219-222
                # the reference's flat buffer is track-major (_reconstruct.py:238) while its offsets describe
                # (b, t, p, ~l) (:292-300); the kernel writes the latter directly so data and offsets agree for t > 1
                out = eng.realign_tracks(names, t_reg, t_shifts, t_goi, pk.get(i_tr), track_lengths, oo, total, ids, params,
                                         base_seed, max_rec, keep, keep_off, t_rc, layout="btp")
                lens_bp = (oo[1:] - oo[:-1]).view(b, p)
                lens = lens_bp.view(b, 1, p).expand(b, t, p).reshape(-1)
                offsets = torch.zeros(b * t * p + 1, dtype=torch.int64, device=dev)
                torch.cumsum(lens, 0, out=offsets[1:])
                results.append(Ragged(out, offsets, (b, t, p, None)))
            else:
                # Tracks alone / un-realigned: paint the stored intervals (Tracks._call_float32, _tracks.py:370-420)
                out_len_q = np.full(b, int(self.output_length), np.int64) if fixed else lengths
                offs = np.concatenate([[0], np.cumsum(out_len_q)]).astype(np.int64)
                d_off = torch.from_numpy(offs).to(dev)
                starts = t_reg[:, 1].contiguous()
                per = int(offs[-1])
                # all tracks in one launch, written in (b, t, ~l) order, negative-strand rows reversed
                out = eng.paint_tracks(names, pk.get(i_tr), starts, d_off, per, pk.get(i_trc) if i_trc is not None else None)
                lens_q = d_off[1:] - d_off[:-1]
                lens = lens_q.view(b, 1).expand(b, t).reshape(-1)
                offsets = torch.zeros(b * t + 1, dtype=torch.int64, device=dev)
                torch.cumsum(lens, 0, out=offsets[1:])
                results.append(Ragged(out, offsets, (b, t, None)))
        return tuple(results)

    def _get_variants(self, t_goi, t_rc, b: int, regions, n_variants: int) -> RaggedVariants:
        p, eng = self.ploidy, self.engine
        fold = p if self.unphased_union else 1
        windows = self.sequence_type == "variant-windows"
        tokens = row_contigs = None
        if windows:
            o = self.window_opt
            lut, _ = build_token_lut(o.token_alphabet, o.unknown_token)
            tokens = dict(lut=lut, unk=o.unknown_token, L=o.flank_length, ref=1 if o.ref == "window" else 2,
                          alt=1 if o.alt == "window" else 2)
        elif self.flank_length and self.token_alphabet is not None:
            lut, _ = build_token_lut(self.token_alphabet, self.unknown_token)
            tokens = dict(lut=lut, unk=self.unknown_token, L=self.flank_length, flank=True)
        if tokens is not None:  # contig of every (b*p) row (_flat_variants.py:985-989)
            row_contigs = torch.from_numpy(np.repeat(regions[:, 0].astype(np.int32), p)).to(eng.device)
        # (n_variants = sum of the rows' genotype slice lengths, known from the host copy of the offsets: no sync for it)
        g = eng.gather_variants(t_goi, t_rc, self.var_fields, self.dummy_variant, self.min_af, self.max_af, fold, tokens, row_contigs,
                                n_hint=n_variants)
        shape = (b, 1 if self.unphased_union else p, None)
        off = g["row_offsets"]
        fields = {}
        names = [n for n in self.var_fields if not (windows and n in ("alt", "ref"))]
        names += [n for n in ("ref_window", "alt_window", "ref", "alt") if windows and n in g]  # token buffers of the windows tail
        for name in names:
            v = g[name]
            fields[name] = RaggedAlleles(v[0], v[1], off, shape) if isinstance(v, tuple) else Ragged(v, off, shape)
        if "flank_tokens" in g:  # (b, p, ~v, 2 L): the trailing token axis stays on `data`
            fields["flank_tokens"] = Ragged(g["flank_tokens"].view(-1, 2 * self.flank_length), off, shape)
        return RaggedVariants(fields, off, shape)

    def _exonic_keep(self, t_goi, t_reg, goi):
        """choose_exonic_variants (src/genotypes/mod.rs:132-176): variants fully inside the query, on the device."""
        eng = self.engine
        return eng.choose_exonic_variants(t_reg[:, 1].contiguous(), t_reg[:, 2].contiguous(), t_goi, eng.max_records(goi))

    # ---- output shaping, _query.py:94-127 ----
    def _shape_output(self, o, out_reshape, squeeze):
        if isinstance(o, (torch.Tensor, AnnotatedHaps, RaggedVariants)) or self.output_format == "flat" or self.output_length == "ragged":
            res = o  # already dense (fixed-length pipeline, channels-first one-hot); "flat"/"ragged" hand the flat triple back
        elif self.output_length == "variable":
            if isinstance(o, RaggedAnnotatedHaps):
                res = o.to_padded()
            elif o.data.dtype == torch.float32:
                res = o.to_padded(0.0)                               # tracks pad with 0 (_query.py:545-548)
            else:
                res = o.to_padded(0 if o.data.dim() > 1 else N_CHAR)  # bytes pad with N; one-hot with zeros
        else:
            res = o.to_fixed(int(self.output_length))
        if out_reshape is not None:
            if isinstance(res, torch.Tensor):
                res = res.reshape(*out_reshape, *res.shape[1:])
            elif isinstance(res, AnnotatedHaps):
                res = res.reshape((*out_reshape, *res.haps.shape[1:]))
            else:
                res = res.reshape((*out_reshape, *(d for d in res.shape[1:] if d is not None)))
        if squeeze:
            res = res.squeeze(0)  # (1 [p] l) -> ([p] l), _query.py:120-122
        return res


class _Packer:
    """All O(batch) host arrays of one call in ONE pinned buffer and ONE async H2D copy."""

    def __init__(self, device):
        self.device = device
        self.items = []
        self.total = 0
        self.dev = None

    def add(self, arr, dtype) -> int:
        a = np.ascontiguousarray(arr, dtype)
        self.items.append((a, self.total))
        self.total += (a.nbytes + 15) & ~15
        return len(self.items) - 1

    def upload(self) -> None:
        host = torch.empty(max(self.total, 16), dtype=torch.uint8, pin_memory=True)
        hv = host.numpy()
        for a, off in self.items:
            hv[off: off + a.nbytes] = a.reshape(-1).view(np.uint8)
        self.dev = host.to(self.device, non_blocking=True)
        self._host = host  # keep the pinned buffer alive until the copy has been consumed

    def get(self, i: int) -> torch.Tensor:
        a, off = self.items[i]
        tdt = torch.from_numpy(np.empty(0, a.dtype)).dtype
        return self.dev[off: off + a.nbytes].view(tdt).view(a.shape)


class BatchLoader:
    """Re-iterable batch iterator over a `Dataset` (see `Dataset.to_dataloader`)."""

    def __init__(self, ds: Dataset, batch_size: int, shuffle: bool, sampler, drop_last: bool, generator, return_indices: bool,
                 transform):
        if batch_size < 1:
            raise ValueError("batch_size must be a positive integer")
        self.ds, self.batch_size, self.shuffle, self.sampler = ds, batch_size, shuffle, sampler
        self.drop_last, self.return_indices, self.transform = drop_last, return_indices, transform
        if generator is None or isinstance(generator, (int, np.integer)):
            self._rng = np.random.default_rng(generator)
        elif isinstance(generator, np.random.Generator):
            self._rng = generator
        else:  # torch.Generator
            self._rng = np.random.default_rng(int(generator.initial_seed()))

    def __len__(self) -> int:
        n = len(self.sampler) if self.sampler is not None else len(self.ds)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        if self.sampler is not None:
            order = (np.ascontiguousarray(self.sampler, np.int64).ravel() if isinstance(self.sampler, np.ndarray)
                     else np.fromiter(iter(self.sampler), np.int64))
        elif self.shuffle:
            order = self._rng.permutation(len(self.ds))
        else:
            order = np.arange(len(self.ds), dtype=np.int64)
        n_s = self.ds.n_samples
        for lo in range(0, len(order), self.batch_size):
            idx = order[lo: lo + self.batch_size]
            if len(idx) < self.batch_size and self.drop_last:
                return
            r, s_ = idx // n_s, idx % n_s  # np.unravel_index(idx, dataset.shape), _torch.py:293
            batch = self.ds[r, s_]
            if not isinstance(batch, tuple):
                batch = (batch,)
            if self.return_indices:  # the (region, sample) indices the dataset was indexed with, _torch.py:299-300
                batch = (*batch, r, s_)
            if self.transform is not None:
                batch = self.transform(*batch)  # _torch.py:302-303
            elif len(batch) == 1:
                batch = batch[0]
            yield batch


class _MapDataset:
    """Map-style view of a Dataset (reference `TorchDataset`, _torch.py:272-307); usable with torch.utils.data.DataLoader
    (batch_size=None + a BatchSampler, as the reference's own loader does)."""

    def __init__(self, dataset: Dataset, include_indices: bool, transform):
        self.dataset, self.include_indices, self.transform = dataset, include_indices, transform

    def __len__(self) -> int:
        return len(self.dataset)

    def __getitem__(self, idx):
        r_idx, s_idx = np.unravel_index(idx, self.dataset.shape)
        batch = self.dataset[r_idx, s_idx]
        if not isinstance(batch, tuple):
            batch = (batch,)
        if self.include_indices:
            batch = (*batch, r_idx, s_idx)
        if self.transform is not None:
            return self.transform(*batch)
        return batch[0] if len(batch) == 1 else batch
