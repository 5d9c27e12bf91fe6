"""Golden vectors produced by EXECUTING the reference's own pure-Python kernels.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_pyref_golden.py

The reference keeps statement-for-statement Python twins of both Rust cores for its own
unit tests:
  * reconstruct_haplotype_from_sparse   python/genvarloader/_dataset/_genotypes.py:125-248
  * shift_and_realign_track_sparse      python/genvarloader/_dataset/_tracks.py:706-824
    (+ _apply_insertion_fill :639-703, _hash4/_xorshift64 :621-636)
Their modules import the compiled extension, so the function bodies are lifted with `ast`
and executed with numpy only -- nothing is re-implemented here.  The cases go beyond the
reference's frozen hypothesis goldens: annotations, negative starts / contig overshoot,
long indels, shifts, keep masks, duplicate positions, bigger windows, all five fills.

Outputs (pickle-free):  tests/golden/pyref_haps.npz, tests/golden/pyref_tracks.npz with the
same c{case}_a{arg} / c{case}_g{j} layout as make_golden.py.  Argument order = the batch
FFI entries (src/ffi/mod.rs:634-655 and :2439-2458) minus `out`.
"""
from __future__ import annotations

import ast
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/python/genvarloader/_dataset")
DST = Path(__file__).parent


def _lift(path: Path, names: set[str], consts: set[str] = frozenset()):
    tree = ast.parse(path.read_text())
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            body.append(node)
        elif isinstance(node, ast.Assign) and any(getattr(t, "id", None) in consts for t in node.targets):
            body.append(node)
    ns: dict = {"np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), str(path), "exec"), ns)
    return ns


def load_reference_python():
    g = _lift(REF / "_genotypes.py", {"reconstruct_haplotype_from_sparse"})
    t = _lift(
        REF / "_tracks.py",
        {"_xorshift64", "_hash4", "_apply_insertion_fill", "shift_and_realign_track_sparse"},
        {"_REPEAT_5P", "_REPEAT_5P_NORM", "_CONSTANT", "_FLANK_SAMPLE", "_INTERPOLATE"},
    )
    return g["reconstruct_haplotype_from_sparse"], t["shift_and_realign_track_sparse"]


# --------------------------------------------------------------------------------------
def _variant_table(rng, n_unique, max_pos, max_indel):
    v_starts = np.sort(rng.integers(0, max_pos, n_unique)).astype(np.int32)
    kind = rng.choice(3, n_unique, p=[0.5, 0.25, 0.25])
    ilens = np.where(kind == 0, 0, np.where(kind == 1, rng.integers(1, max_indel + 1, n_unique),
                                            -rng.integers(1, max_indel + 1, n_unique))).astype(np.int32)
    alt_lens = np.maximum(1, 1 + ilens).astype(np.int64)
    alt_offsets = np.concatenate([[0], np.cumsum(alt_lens)]).astype(np.int64)
    alt_alleles = rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), int(alt_offsets[-1])).astype(np.uint8)
    return v_starts, ilens, alt_alleles, alt_offsets


def _genos(rng, n_groups, n_unique, max_per):
    counts = rng.integers(0, max_per + 1, n_groups)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    v = [np.sort(rng.integers(0, n_unique, c)) for c in counts]  # sorted, duplicates allowed
    geno_v_idxs = (np.concatenate(v) if len(v) else np.empty(0)).astype(np.int32)
    return geno_v_idxs, np.stack([off[:-1], off[1:]]).astype(np.int64), off


def hap_case(rng, recon_py):
    n_contigs = int(rng.integers(1, 3))
    contig_lens = rng.integers(30, 400, n_contigs)
    ref_offsets = np.concatenate([[0], np.cumsum(contig_lens)]).astype(np.int64)
    reference = rng.choice(np.frombuffer(b"ACGTN", np.uint8), int(ref_offsets[-1]), p=[.24, .24, .24, .24, .04])
    n_unique = int(rng.integers(1, 40))
    v_starts, ilens, alt_alleles, alt_offsets = _variant_table(rng, n_unique, int(contig_lens.min()), 12)
    n_q, ploidy = int(rng.integers(1, 5)), int(rng.integers(1, 3))
    geno_v_idxs, geno_offsets, off1d = _genos(rng, n_q * ploidy, n_unique, 25)
    geno_offset_idx = rng.permutation(n_q * ploidy).astype(np.int64).reshape(n_q, ploidy)
    regions = np.empty((n_q, 3), np.int32)
    lengths = []
    for i in range(n_q):
        c = int(rng.integers(0, n_contigs))
        clen = int(contig_lens[c])
        start = int(rng.integers(-8, clen - 1))  # may be negative -> leading pad
        length = int(rng.integers(1, clen - max(start, 0) + 12))  # may overshoot the contig end
        if start < 0:
            length = max(length, -start + 1)  # in-contract: pad never exceeds the window
        regions[i] = (c, start, start + length)
        lengths.append(length)
    out_len = np.repeat(np.array(lengths, np.int64), ploidy)
    # some rows get a fixed length different from the region length (truncate / extend)
    if rng.random() < 0.4:
        out_len = np.maximum(1, out_len + rng.integers(-10, 10, out_len.shape))
        for k in range(len(out_len)):
            s = int(regions[k // ploidy, 1])
            if s < 0:
                out_len[k] = max(out_len[k], -s + 1)
    out_offsets = np.concatenate([[0], np.cumsum(out_len)]).astype(np.int64)
    shifts = np.zeros((n_q, ploidy), np.int32)
    if rng.random() < 0.6:
        shifts = rng.integers(0, np.maximum(1, np.array(lengths)[:, None] // 3 + 1), (n_q, ploidy)).astype(np.int32)
    if rng.random() < 0.4 and off1d[-1] > 0:
        # keep is laid out per flat work item k (keep_offsets[k]), sized by that row's geno slice
        sizes = np.array([geno_offsets[1, o] - geno_offsets[0, o] for o in geno_offset_idx.ravel()], np.int64)
        keep_offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        keep = rng.random(int(keep_offsets[-1])) < 0.7
    else:
        keep = keep_offsets = None
    pad_char = np.uint8(ord("N"))

    total = int(out_offsets[-1])
    out = np.zeros(total, np.uint8)
    av = np.zeros(total, np.int32)
    ap = np.zeros(total, np.int32)
    for k in range(n_q * ploidy):
        q = k // ploidy
        o = int(geno_offset_idx.ravel()[k])
        vi = geno_v_idxs[geno_offsets[0, o]:geno_offsets[1, o]]
        c = int(regions[q, 0])
        cref = reference[ref_offsets[c]:ref_offsets[c + 1]]
        kk = None if keep is None else keep[keep_offsets[k]:keep_offsets[k + 1]]
        s, e = int(out_offsets[k]), int(out_offsets[k + 1])
        recon_py(vi, v_starts, ilens, int(shifts.ravel()[k]), alt_alleles, alt_offsets, cref, int(regions[q, 1]),
                 out[s:e], int(pad_char), kk, av[s:e], ap[s:e])
    inputs = (out_offsets, regions, shifts, geno_offset_idx, geno_offsets, geno_v_idxs, v_starts, ilens, alt_alleles,
              alt_offsets, reference, ref_offsets, pad_char, keep, keep_offsets)
    return inputs, (out, av, ap)


def track_case(rng, realign_py, strategy_id):
    n_unique = int(rng.integers(1, 30))
    v_starts = np.sort(rng.integers(0, 300, n_unique)).astype(np.int32)
    kind = rng.choice(3, n_unique, p=[0.3, 0.4, 0.3])
    ilens = np.where(kind == 0, 0, np.where(kind == 1, rng.integers(1, 10, n_unique),
                                            -rng.integers(1, 10, n_unique))).astype(np.int32)
    n_q, ploidy = int(rng.integers(1, 5)), int(rng.integers(1, 3))
    q_starts = rng.integers(0, 200, n_q)
    region_lengths = rng.integers(4, 120, n_q)
    regions = np.stack([np.zeros(n_q, np.int64), q_starts, q_starts + region_lengths], 1).astype(np.int32)
    track_lengths = region_lengths + 40  # headroom for deletions (production contract: track >= region)
    track_offsets = np.concatenate([[0], np.cumsum(track_lengths)]).astype(np.int64)
    tracks = rng.normal(0, 50, int(track_offsets[-1])).astype(np.float32)
    tracks[rng.random(tracks.size) < 0.3] = 0.0
    geno_v_idxs, geno_offsets, off1d = _genos(rng, n_q * ploidy, n_unique, 20)
    geno_offset_idx = rng.permutation(n_q * ploidy).astype(np.int64).reshape(n_q, ploidy)
    out_len = np.repeat(region_lengths.astype(np.int64), ploidy)
    if rng.random() < 0.4:
        out_len = np.maximum(1, out_len + rng.integers(-10, 6, out_len.shape))
    out_offsets = np.concatenate([[0], np.cumsum(out_len)]).astype(np.int64)
    shifts = np.zeros((n_q, ploidy), np.int32)
    if rng.random() < 0.7:
        shifts = rng.integers(0, region_lengths[:, None] // 3 + 1, (n_q, ploidy)).astype(np.int32)
    if rng.random() < 0.4 and off1d[-1] > 0:
        sizes = np.array([geno_offsets[1, o] - geno_offsets[0, o] for o in geno_offset_idx.ravel()], np.int64)
        keep_offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        keep = rng.random(int(keep_offsets[-1])) < 0.7
    else:
        keep = keep_offsets = None
    param = {0: 0.0, 1: 0.0, 2: float(rng.normal()), 3: float(rng.integers(0, 6)), 4: float(rng.integers(1, 4))}[strategy_id]
    params = np.array([param], np.float64)
    base_seed = np.uint64(rng.integers(0, 2**63)) * np.uint64(2) + np.uint64(rng.integers(0, 2))
    out = np.zeros(int(out_offsets[-1]), np.float32)
    for k in range(n_q * ploidy):
        q, h = divmod(k, ploidy)
        kk = None if keep is None else keep[keep_offsets[k]:keep_offsets[k + 1]]
        realign_py(int(geno_offset_idx.ravel()[k]), geno_v_idxs, geno_offsets, v_starts, ilens, int(shifts.ravel()[k]),
                   tracks[track_offsets[q]:track_offsets[q + 1]], int(regions[q, 1]),
                   out[out_offsets[k]:out_offsets[k + 1]], params, kk, strategy_id, int(base_seed), q, h)
    inputs = (out_offsets, regions, shifts, geno_offset_idx, geno_v_idxs, geno_offsets, v_starts, ilens, tracks,
              track_offsets, params, keep, keep_offsets, np.int64(strategy_id), base_seed)
    return inputs, (out,)


def _save(name, cases):
    out = {"n_cases": np.array(len(cases), np.int64), "n_args": np.array(len(cases[0][0]), np.int64),
           "n_gold": np.array(len(cases[0][1]), np.int64)}
    for ci, (inputs, gold) in enumerate(cases):
        none = []
        for j, a in enumerate(inputs):
            if a is None:
                none.append(j)
            else:
                out[f"c{ci}_a{j}"] = np.asarray(a)
        out[f"c{ci}_none"] = np.array(none, np.int64)
        for j, g in enumerate(gold):
            out[f"c{ci}_g{j}"] = np.asarray(g)
    np.savez_compressed(DST / f"ref_{name}.npz", **out)
    print(f"{name}: {len(cases)} cases")


if __name__ == "__main__":
    if not REF.is_dir():
        sys.exit("needs /root/reference (build container only)")
    recon_py, realign_py = load_reference_python()
    rng = np.random.default_rng(20261017)
    _save("pyref_haps", [hap_case(rng, recon_py) for _ in range(150)])
    _save("pyref_tracks", [track_case(rng, realign_py, i % 5) for i in range(150)])
