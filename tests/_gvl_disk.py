"""Writes a `genvarloader_b200.synth.SynthData` in the on-disk layout of a GenVarLoader dataset
(reference docs/source/format.md:8-49; writer side: python/genvarloader/_dataset/_write.py:186-330) with numpy +
pyarrow only -- the fixture for the `Dataset.open` tests.  Regions are stored SORTED, the input BED is a shuffled
copy of them with `r_idx_map` (input row -> storage row), exactly what `gvl.write` leaves behind."""
import json
from pathlib import Path

import numpy as np


def write_svar_store(svar_dir, d, with_ilen=False):
    """A `.svar` (SVAR1 / genoray SparseVar) directory as far as GenVarLoader reads it (_haps.py:428-446): the global
    sparse genotype array `variant_idxs.npy` (raw int32) and the variant table `index.arrow` (1-based POS, REF, ALT as a
    list column; ILEN optional -- without it the reader derives it from the allele lengths)."""
    import pyarrow as pa
    import pyarrow.ipc as ipc

    svar_dir = Path(svar_dir)
    svar_dir.mkdir(parents=True)
    np.asarray(d.geno_v_idxs, np.int32).tofile(svar_dir / "variant_idxs.npy")
    alts = [d.alt_alleles[d.alt_offsets[i]: d.alt_offsets[i + 1]].tobytes().decode() for i in range(d.v_starts.size)]
    refs = ["N" * (len(a) - int(il)) for a, il in zip(alts, d.ilens)]  # len(ALT) - len(REF) = ILEN
    cols = {"POS": d.v_starts.astype(np.int64) + 1, "REF": refs, "ALT": pa.array([[a] for a in alts], pa.list_(pa.utf8()))}
    if with_ilen:
        cols["ILEN"] = pa.array([[int(x)] for x in d.ilens], pa.list_(pa.int32()))
    vt = pa.table(cols)
    with pa.OSFile(str(svar_dir / "index.arrow"), "wb") as f, ipc.new_file(f, vt.schema) as w:
        w.write_table(vt)
    return dict(n_variants=int(d.v_starts.size), variant_idxs_bytes=int((svar_dir / "variant_idxs.npy").stat().st_size))


def write_gvl_dataset(path, d, contigs, samples, input_order, version="0.19.0", strand_as_str=True, pos_one_based=True,
                      annot_tracks=(), svar_dir=None, svar_fingerprint=None):
    import pyarrow as pa
    import pyarrow.ipc as ipc

    path = Path(path)
    (path / "genotypes").mkdir(parents=True)
    n_regions = d.n_regions
    meta = dict(samples=list(samples), contigs=list(contigs), n_regions=n_regions, ploidy=d.ploidy, max_jitter=d.max_jitter,
                version=version, format_version=None, svar_link=None, svar2_link=None, variants_fingerprint=None)
    (path / "metadata.json").write_text(json.dumps(meta))
    # input BED: storage row r_idx_map[i] holds input row i
    order = np.asarray(input_order)  # input row i shows storage region order[i]
    reg = d.regions[order]
    strand = ["+" if s == 1 else "-" for s in reg[:, 3]] if strand_as_str else reg[:, 3].astype(np.int32)
    bed = pa.table({"chrom": [contigs[c] for c in reg[:, 0]], "chromStart": reg[:, 1].astype(np.int64),
                    "chromEnd": reg[:, 2].astype(np.int64), "strand": strand, "name": [f"r{i}" for i in range(n_regions)],
                    "r_idx_map": order.astype(np.int64)})
    with pa.OSFile(str(path / "input_regions.arrow"), "wb") as f, ipc.new_file(f, bed.schema) as w:
        w.write_table(bed)
    go = np.asarray(d.geno_offsets)
    if svar_dir is not None:
        # genotypes stay in the linked .svar store (_write.py: svar-backed datasets): only the (2, r, s, p) starts/stops
        # into its variant_idxs.npy are stored here, plus the link record in metadata.json
        import os

        n_samples = len(samples)
        (path / "genotypes" / "svar_meta.json").write_text(json.dumps(
            {"shape": [2, n_regions, n_samples, d.ploidy], "dtype": "int64"}))
        np.ascontiguousarray(go, np.int64).tofile(path / "genotypes" / "offsets.npy")
        meta["svar_link"] = dict(relative_path=os.path.relpath(svar_dir, start=path).replace(os.sep, "/"),
                                 absolute_path=str(Path(svar_dir).resolve()), fingerprint=svar_fingerprint)
        (path / "metadata.json").write_text(json.dumps(meta))
    else:
        alts = [d.alt_alleles[d.alt_offsets[i]: d.alt_offsets[i + 1]].tobytes().decode() for i in range(d.v_starts.size)]
        vt = pa.table({"POS": (d.v_starts.astype(np.int64) + int(pos_one_based)), "ILEN": d.ilens.astype(np.int32), "ALT": alts,
                       "AF": np.linspace(0, 1, d.v_starts.size)})
        with pa.OSFile(str(path / "genotypes" / "variants.arrow"), "wb") as f, ipc.new_file(f, vt.schema) as w:
            w.write_table(vt)
        # sparse genotypes: contiguous CSR over (region, sample, ploid)
        starts, stops = go[0], go[1]
        lens = np.maximum(stops - starts, 0)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        v_idx = np.concatenate([d.geno_v_idxs[s:e] for s, e in zip(starts, stops)] + [np.zeros(0, np.int32)]).astype(np.int32)
        v_idx.tofile(path / "genotypes" / "variant_idxs.npy")
        offs.tofile(path / "genotypes" / "offsets.npy")
    for name, (s, e, v, o) in d.tracks.items():
        sub = "annot_intervals" if name in annot_tracks else "intervals"
        tdir = path / sub / name
        tdir.mkdir(parents=True)
        np.asarray(s, np.int32).tofile(tdir / "starts.npy")
        np.asarray(e, np.int32).tofile(tdir / "ends.npy")
        np.asarray(v, np.float32).tofile(tdir / "values.npy")
        np.asarray(o, np.int64).tofile(tdir / "offsets.npy")


def write_fasta(path, reference, ref_offsets, contigs, width=70, lower_every=0):
    with open(path, "wb") as f:
        for i, c in enumerate(contigs):
            seq = reference[ref_offsets[i]: ref_offsets[i + 1]].tobytes()
            if lower_every:
                seq = b"".join(seq[k:k + lower_every].lower() if (k // lower_every) % 2 else seq[k:k + lower_every]
                               for k in range(0, len(seq), lower_every))
            f.write(b">" + c.encode() + b" some description\n")
            for k in range(0, len(seq), width):
                f.write(seq[k:k + width] + b"\n")
