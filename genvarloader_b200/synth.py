"""Synthetic in-memory datasets shaped like SURVEY.md section 8(d) -- used by bench.py and the tests.

Everything is generated with ``numpy.random.default_rng(seed)``; layouts are the reference's
on-disk / in-memory layouts:
  * reference: uint8 ASCII contigs concatenated + int64 offsets (``Reference``,
    python/genvarloader/_dataset/_reference.py:53-120), pad char ``N``
  * variant table: ``v_starts`` i32 sorted per contig, ``ilens`` i32, atomised left-aligned ALT
    alleles incl. the anchor base (SNP 1 B, INS 1+ilen B, DEL 1 B) as ragged u8 + i64 offsets
  * sparse genotypes: CSR over (region, sample, ploid) slots holding variant indices sorted by
    position (``genotypes/variant_idxs.npy`` + ``offsets.npy``, docs/source/format.md:8-49)
  * regions: int32 (R, 4) = contig_idx, start, end, strand
  * tracks: interval SoA (starts, ends, values) + CSR offsets per (region, sample) slot
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)


@dataclass
class SynthData:
    reference: np.ndarray          # u8
    ref_offsets: np.ndarray        # i64 (C+1)
    v_starts: np.ndarray           # i32 (V)
    ilens: np.ndarray              # i32 (V)
    alt_alleles: np.ndarray        # u8
    alt_offsets: np.ndarray        # i64 (V+1)
    geno_v_idxs: np.ndarray        # i32 (G)
    geno_offsets: np.ndarray       # i64 (2, R*S*P)
    regions: np.ndarray            # i32 (R, 4)
    n_samples: int
    ploidy: int
    max_jitter: int = 0
    tracks: dict = field(default_factory=dict)  # name -> (itv_starts, itv_ends, itv_values, itv_offsets)

    @property
    def n_regions(self) -> int:
        return int(self.regions.shape[0])


def make_reference(rng, contig_lens, n_frac=0.01) -> tuple[np.ndarray, np.ndarray]:
    """Uniform ACGT with ~n_frac of the bases inside short runs of N."""
    ref_offsets = np.concatenate([[0], np.cumsum(contig_lens)]).astype(np.int64)
    total = int(ref_offsets[-1])
    ref = ACGT[rng.integers(0, 4, total, dtype=np.uint8)]
    if n_frac > 0 and total > 1000:
        run = 50
        n_runs = max(1, int(total * n_frac / run))
        for s in rng.integers(0, total - run, n_runs):
            ref[s:s + run] = ord("N")
    return ref, ref_offsets


def make_variants(rng, contig_len: int, n_variants: int, snp_frac=0.8, max_indel=20):
    """Sorted unique positions; SNP/INS/DEL mix (snp_frac, rest split evenly)."""
    n_variants = min(n_variants, contig_len - 2)
    pos = np.sort(rng.choice(contig_len - 1, n_variants, replace=False)).astype(np.int32)
    u = rng.random(n_variants)
    ins = u >= snp_frac + (1 - snp_frac) / 2
    dele = (u >= snp_frac) & ~ins
    mag = rng.integers(1, max_indel + 1, n_variants)
    ilens = np.where(ins, mag, np.where(dele, -mag, 0)).astype(np.int32)
    alt_lens = np.where(ilens > 0, ilens + 1, 1).astype(np.int64)
    alt_offsets = np.concatenate([[0], np.cumsum(alt_lens)]).astype(np.int64)
    alt = ACGT[rng.integers(0, 4, int(alt_offsets[-1]), dtype=np.uint8)]
    return pos, ilens, alt, alt_offsets


def make_genotypes(rng, regions, v_starts, n_samples, ploidy, max_jitter=0, af_a=0.5, af_b=5.0, dense_af=None):
    """CSR of variant indices per (region, sample, ploid): every variant overlapping the jitter-
    expanded region is carried by each haplotype with probability AF ~ Beta(af_a, af_b)
    (or the constant ``dense_af``)."""
    V = len(v_starts)
    af = np.full(V, dense_af) if dense_af is not None else rng.beta(af_a, af_b, V)
    R = regions.shape[0]
    lo = np.searchsorted(v_starts, regions[:, 1] - max_jitter - 64, "left")
    hi = np.searchsorted(v_starts, regions[:, 2] + max_jitter + 64, "left")
    chunks, lengths = [], np.zeros(R * n_samples * ploidy, np.int64)
    slot = 0
    for r in range(R):
        idx = np.arange(lo[r], hi[r], dtype=np.int32)
        n = idx.size
        if n == 0:
            slot += n_samples * ploidy
            continue
        carry = rng.random((n_samples * ploidy, n)) < af[idx][None, :]
        cnt = carry.sum(1)
        lengths[slot:slot + n_samples * ploidy] = cnt
        chunks.append(np.broadcast_to(idx, carry.shape)[carry])
        slot += n_samples * ploidy
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    geno_v_idxs = np.concatenate(chunks).astype(np.int32) if chunks else np.empty(0, np.int32)
    return geno_v_idxs, np.ascontiguousarray(np.stack([off[:-1], off[1:]]))


def make_track(rng, regions, n_slots_per_region, max_jitter=0, mean_run=50, headroom=64):
    """Interval SoA per slot: contiguous runs (mean length ``mean_run``) covering the expanded
    region, ~25% of runs zero-valued and omitted (holes)."""
    starts, ends, values, counts = [], [], [], []
    for r in range(regions.shape[0]):
        s0 = int(regions[r, 1]) - max_jitter
        e0 = int(regions[r, 2]) + max_jitter + headroom
        for _ in range(n_slots_per_region):
            n_runs = max(1, (e0 - s0) // mean_run)
            cuts = np.sort(rng.choice(np.arange(s0 + 1, e0), min(n_runs, e0 - s0 - 1), replace=False))
            b = np.concatenate([[s0], cuts, [e0]])
            keep = rng.random(b.size - 1) > 0.25
            starts.append(b[:-1][keep])
            ends.append(b[1:][keep])
            values.append(rng.gamma(2.0, 2.0, int(keep.sum())).astype(np.float32))
            counts.append(int(keep.sum()))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return (np.concatenate(starts).astype(np.int32), np.concatenate(ends).astype(np.int32),
            np.concatenate(values).astype(np.float32), off)


def make_track_fast(rng, regions, n_slots_per_region, max_jitter=0, mean_run=50, headroom=64):
    """`make_track` for large tables (millions of intervals): the same kind of track -- contiguous runs of mean length
    `mean_run` over the expanded region, ~25% of them omitted -- drawn for all slots of a region at once."""
    starts, ends, values, counts = [], [], [], []
    for r in range(regions.shape[0]):
        s0 = int(regions[r, 1]) - max_jitter
        T = int(regions[r, 2]) + max_jitter + headroom - s0
        n_runs = max(1, T // mean_run)
        cuts = np.sort(rng.integers(1, T, size=(n_slots_per_region, n_runs - 1), dtype=np.int32), axis=1)
        b = np.concatenate([np.zeros((n_slots_per_region, 1), np.int32), cuts, np.full((n_slots_per_region, 1), T, np.int32)], 1)
        keep = (b[:, 1:] > b[:, :-1]) & (rng.random((n_slots_per_region, n_runs)) > 0.25)
        starts.append((b[:, :-1] + s0)[keep])
        ends.append((b[:, 1:] + s0)[keep])
        counts.append(keep.sum(1))
        values.append(rng.gamma(2.0, 2.0, int(keep.sum())).astype(np.float32))
    off = np.concatenate([[0], np.cumsum(np.concatenate(counts))]).astype(np.int64)
    return (np.concatenate(starts).astype(np.int32), np.concatenate(ends).astype(np.int32),
            np.concatenate(values).astype(np.float32), off)


def make_dataset(seed: int, contig_len: int, n_samples: int, n_regions: int, region_len: int,
                 variants_per_kb: float = 1.0, ploidy: int = 2, max_jitter: int = 0, snp_frac: float = 0.8,
                 neg_strand_frac: float = 0.0, straddle_ends: bool = True, n_tracks: int = 0,
                 sample_tracks: bool = True, dense_af: float | None = None, max_indel: int = 20,
                 fast_tracks: bool = False) -> SynthData:
    # `variants_per_kb` is the density PER HAPLOTYPE (what the kernels see); the table is denser by
    # 1/E[AF] so that Bernoulli(AF ~ Beta(0.5, 5)) carriers hit that density.
    rng = np.random.default_rng(seed)
    ref, ref_offsets = make_reference(rng, [contig_len])
    mean_af = dense_af if dense_af is not None else 0.5 / 5.5
    n_var = int(contig_len * variants_per_kb / 1000 / mean_af)
    v_starts, ilens, alt, alt_off = make_variants(rng, contig_len, max(n_var, 1), snp_frac, max_indel)
    lo, hi = max_jitter, max(contig_len - region_len - max_jitter, max_jitter + 1)
    starts = np.sort(rng.integers(lo, hi, n_regions)).astype(np.int64)
    if straddle_ends and n_regions >= 4 and max_jitter == 0:
        starts[0] = -min(37, region_len // 4)                  # leading pad
        starts[-1] = contig_len - region_len + min(53, region_len // 4)  # trailing pad
    strand = np.where(rng.random(n_regions) < neg_strand_frac, -1, 1)
    regions = np.stack([np.zeros(n_regions, np.int64), starts, starts + region_len, strand], 1).astype(np.int32)
    gv, go = make_genotypes(rng, regions, v_starts, n_samples, ploidy, max_jitter, dense_af=dense_af)
    d = SynthData(ref, ref_offsets, v_starts, ilens, alt, alt_off, gv, go, regions, n_samples, ploidy, max_jitter)
    for t in range(n_tracks):
        d.tracks[f"track{t}"] = (make_track_fast if fast_tracks else make_track)(rng, regions, n_samples if sample_tracks else 1,
                                                                              max_jitter)
    return d


# ---- the configurations of BASELINE.json / SURVEY.md section 8(d) -----------------------------
def cfg1(seed=1, n_regions=1000) -> SynthData:
    """1 Mb contig, 8 diploid samples, ~1 variant/kb, 1,000 regions x 16,384 bp."""
    return make_dataset(seed, 1_000_000, 8, n_regions, 16_384, 1.0)


def cfg2(seed=2, contig_len=50_000_000, n_samples=32, n_regions=64) -> SynthData:
    """chr22-scale contig, 131,072-bp windows.  ``n_samples`` defaults below 2,504 to bound host RAM
    of the synthetic generator; the kernel cost per row does not depend on the cohort size."""
    return make_dataset(seed, contig_len, n_samples, n_regions, 131_072, 1.0)


def cfg3(seed=3, contig_len=20_000_000, n_samples=4, n_regions=16, variants_per_kb=1.0) -> SynthData:
    """Borzoi-style 524,288-bp windows, >=10% indels, jitter, 50% negative strand, 2 tracks."""
    return make_dataset(seed, contig_len, n_samples, n_regions, 524_288 + 2 * 128, variants_per_kb, max_jitter=128,
                        snp_frac=0.8, neg_strand_frac=0.5, straddle_ends=False, n_tracks=2)


def cfg4(seed=4, contig_len=5_000_000, n_samples=8, n_regions=256) -> SynthData:
    """DNA-LM 6,144-bp haplotypes, thousands of rows per call, annotated."""
    return make_dataset(seed, contig_len, n_samples, n_regions, 6_144, 1.0)


def batch_args(d: SynthData, r_idx, s_idx, rng=None, jitter: int = 0):
    """O(batch) arguments of the fused FFI entries for (region, sample) pairs, built the way the
    reference's Python host does: regions + jitter (_dataset/_query.py:161-175), geno_offset_idx =
    ravel of (region, sample, ploid) (_dataset/_haps.py:757-768), per-row to_rc (_haps.py:838-843)."""
    r_idx = np.asarray(r_idx, np.int64)
    s_idx = np.asarray(s_idx, np.int64)
    regions = d.regions[r_idx].copy()
    if jitter:
        rng = rng or np.random.default_rng(0)
        lengths = regions[:, 2] - regions[:, 1]
        regions[:, 1] += rng.integers(-jitter, jitter + 1, size=len(regions), dtype=np.int32)
        regions[:, 2] = regions[:, 1] + lengths
    ploid = np.arange(d.ploidy, dtype=np.int64)
    goi = (r_idx[:, None] * d.n_samples + s_idx[:, None]) * d.ploidy + ploid[None, :]
    to_rc = np.repeat(d.regions[r_idx, 3] == -1, d.ploidy)
    ds_idx = r_idx * d.n_samples + s_idx
    return np.ascontiguousarray(regions[:, :3]), np.ascontiguousarray(goi), np.ascontiguousarray(to_rc), ds_idx


def to_svar2_channels(d: SynthData, regions, ds_idx, dense_frac: float = 0.5, seed: int = 0, pure_del: bool = True):
    """Re-express a batch of an SVAR1-style dataset in the svar2 two-channel flat layout
    (reference `FlatChannels`, src/svar2/mod.rs:150-160) at the decoded level:

      * a random `dense_frac` of the variant table is declared "dense" (shared): per query the dense window is
        every dense variant overlapping the (64-bp expanded) region; each haplotype gets LSB-first presence bits
      * everything else a haplotype carries goes to its private var_key list
      * keys index a decoded key table (ilen, ALT bytes); deletions become PURE deletions (empty ALT, the
        anchor comes from the reference) when `pure_del` and their stored ALT equals the reference base

    Returns a dict with the arrays of `gvl_svar2_channels` + the key table."""
    rng = np.random.default_rng(seed)
    V = d.v_starts.size
    is_dense = rng.random(V) < dense_frac
    b, p = len(ds_idx), d.ploidy
    key_ilen = d.ilens.copy()
    alt_off = d.alt_offsets.copy()
    key_alt = d.alt_alleles
    if pure_del:
        # drop the anchor byte of deletions whose ALT is exactly the reference base at that position
        contig0 = d.reference[d.ref_offsets[0]:d.ref_offsets[1]]
        lens = np.diff(d.alt_offsets)
        is_del = (d.ilens < 0) & (lens == 1) & (d.alt_alleles[d.alt_offsets[:-1]] == contig0[d.v_starts])
        new_lens = np.where(is_del, 0, lens)
        alt_off = np.concatenate([[0], np.cumsum(new_lens)]).astype(np.int64)
        keep_bytes = np.repeat(~is_del, lens)
        key_alt = d.alt_alleles[keep_bytes]
    vk_pos, vk_key, vk_off = [], [], [0]
    dense_pos, dense_key, dense_range = [], [], []
    bits, bit_off = [], [0]
    dense_idx = np.flatnonzero(is_dense)
    for q in range(b):
        lo = np.searchsorted(d.v_starts[dense_idx], regions[q, 1] - 64, "left")
        hi = np.searchsorted(d.v_starts[dense_idx], regions[q, 2] + 64, "left")
        win = dense_idx[lo:hi]
        dense_range.append((len(dense_pos), len(dense_pos) + len(win)))
        dense_pos.extend(d.v_starts[win].tolist())
        dense_key.extend(win.tolist())
        for h in range(p):
            slot = int(ds_idx[q]) * p + h
            vi = d.geno_v_idxs[d.geno_offsets[0, slot]:d.geno_offsets[1, slot]]
            mine_dense = vi[is_dense[vi]]
            mine_vk = vi[~is_dense[vi]]
            vk_pos.extend(d.v_starts[mine_vk].tolist())
            vk_key.extend(mine_vk.tolist())
            vk_off.append(len(vk_pos))
            bits.append(np.isin(win, mine_dense))
            bit_off.append(bit_off[-1] + len(win))
    allbits = np.concatenate(bits) if bits else np.zeros(0, bool)
    dense_present = np.packbits(allbits, bitorder="little")
    return dict(vk_pos=np.array(vk_pos, np.int32), vk_key=np.array(vk_key, np.int32), vk_off=np.array(vk_off, np.int64),
                dense_pos=np.array(dense_pos, np.int32), dense_key=np.array(dense_key, np.int32),
                dense_range=np.array(dense_range, np.int32).reshape(b, 2), dense_present=dense_present,
                dense_present_off=np.array(bit_off, np.int64), key_ilen=key_ilen.astype(np.int32),
                key_alt=np.ascontiguousarray(key_alt), key_alt_off=alt_off)


def to_svar2_dataset(d: SynthData, dense_frac: float = 0.5, seed: int = 0, pure_del: bool = True, pad: int = 64) -> dict:
    """The WHOLE dataset re-expressed as a resident svar2 two-channel source (the layout a dataset replica keeps in HBM;
    the reference keeps the same information as range tables into the .svar2 store, docs/source/format.md:88-96):

      * a random `dense_frac` of the variant table is "dense" (shared): per REGION the dense window is every dense
        variant overlapping the (max_jitter + `pad`)-expanded region, `dense_range` (R, 2) into `dense_pos/dense_key`;
        every (region, sample, ploid) slot owns LSB-first presence bits over its region's window, starting at bit
        `present_off[slot]` of `dense_present`
      * everything else a haplotype carries lives in its private var_key list `vk_pos/vk_key[vk_range[0, slot] : vk_range[1, slot]]`
      * keys index the decoded key table (key_ilen, key_alt, key_alt_off); deletions become PURE deletions (empty ALT)
        when `pure_del` and their stored ALT equals the reference base

    Vectorised (cohort-scale slot counts); equivalent to `to_svar2_channels` batch by batch."""
    rng = np.random.default_rng(seed)
    V = d.v_starts.size
    is_dense = rng.random(V) < dense_frac
    key_ilen, alt_off, key_alt = d.ilens.copy(), d.alt_offsets.copy(), d.alt_alleles
    if pure_del:
        contig0 = d.reference[d.ref_offsets[0]:d.ref_offsets[1]]
        lens = np.diff(d.alt_offsets)
        is_del = (d.ilens < 0) & (lens == 1) & (d.alt_alleles[d.alt_offsets[:-1]] == contig0[d.v_starts])
        alt_off = np.concatenate([[0], np.cumsum(np.where(is_del, 0, lens))]).astype(np.int64)
        key_alt = d.alt_alleles[np.repeat(~is_del, lens)]
    go = np.asarray(d.geno_offsets)
    n_slots = go.shape[1]
    lengths = go[1] - go[0]
    assert (go[0][1:] == go[1][:-1]).all() and go[0][0] == 0, "expects a gap-free genotype CSR"
    gv = np.asarray(d.geno_v_idxs)
    slot_of = np.repeat(np.arange(n_slots, dtype=np.int64), lengths)
    ent_dense = is_dense[gv]
    # var_key channel
    vk_key = gv[~ent_dense].astype(np.int32)
    vk_pos = d.v_starts[vk_key].astype(np.int32)
    vk_cnt = np.bincount(slot_of[~ent_dense], minlength=n_slots).astype(np.int64)
    vk_stop = np.cumsum(vk_cnt)
    vk_range = np.stack([vk_stop - vk_cnt, vk_stop]).astype(np.int64)
    # dense channel: one window per region
    dense_idx = np.flatnonzero(is_dense).astype(np.int32)
    dpos_all = d.v_starts[dense_idx]
    R = d.n_regions
    lo = np.searchsorted(dpos_all, d.regions[:, 1] - d.max_jitter - pad, "left")
    hi = np.searchsorted(dpos_all, d.regions[:, 2] + d.max_jitter + pad, "left")
    win = (hi - lo).astype(np.int64)
    w_stop = np.cumsum(win)
    dense_range = np.stack([w_stop - win, w_stop], 1).astype(np.int32)
    take = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]) if R else np.empty(0, np.int64)
    dense_key = dense_idx[take].astype(np.int32)
    dense_pos = dpos_all[take].astype(np.int32)
    # presence bits: slot -> window of its region
    spr = n_slots // R  # slots per region (samples x ploidy)
    slot_region = np.arange(n_slots) // spr
    bits_per_slot = win[slot_region]
    present_off = np.concatenate([[0], np.cumsum(bits_per_slot)]).astype(np.int64)
    total_bits = int(present_off[-1])
    bits = np.zeros(total_bits + 64, np.bool_)
    ds = slot_of[ent_dense]
    rank = np.searchsorted(dense_idx, gv[ent_dense])  # index of the variant among the dense ones
    inside = (rank >= lo[slot_region[ds]]) & (rank < hi[slot_region[ds]])
    assert inside.all(), "a carried dense variant lies outside its region's dense window"
    bits[present_off[ds] + (rank - lo[slot_region[ds]])] = True
    dense_present = np.packbits(bits, bitorder="little")
    return dict(vk_pos=vk_pos, vk_key=vk_key, vk_range=vk_range, dense_pos=dense_pos, dense_key=dense_key,
                dense_range=dense_range, dense_present=dense_present, present_off=present_off[:-1].copy(),
                slots_per_region=int(spr), key_ilen=key_ilen.astype(np.int32), key_alt=np.ascontiguousarray(key_alt),
                key_alt_off=alt_off, win=win)


def svar2_batch_channels(sv: dict, ds_idx, ploidy: int, n_samples: int) -> dict:
    """The per-call flat channel arrays (`gvl_svar2_channels` / the oracle's arguments) of a batch of a resident svar2
    dataset (`to_svar2_dataset`): what the reference gathers per call from its range cache."""
    ds_idx = np.asarray(ds_idx, np.int64)
    slots = (ds_idx[:, None] * ploidy + np.arange(ploidy)[None, :]).ravel()
    r = ds_idx // n_samples
    vk_lo, vk_hi = sv["vk_range"][0, slots], sv["vk_range"][1, slots]
    vk_off = np.concatenate([[0], np.cumsum(vk_hi - vk_lo)]).astype(np.int64)
    take = np.concatenate([np.arange(a, b) for a, b in zip(vk_lo, vk_hi)]) if len(slots) else np.empty(0, np.int64)
    dr = sv["dense_range"][r]
    d_off = np.concatenate([[0], np.cumsum(dr[:, 1] - dr[:, 0])]).astype(np.int64)
    d_take = np.concatenate([np.arange(a, b) for a, b in dr]) if len(r) else np.empty(0, np.int64)
    allbits = np.unpackbits(sv["dense_present"], bitorder="little")
    row_bits, bit_off = [], [0]
    for k, s in enumerate(slots):
        w = int(sv["win"][r[k // ploidy]])
        row_bits.append(allbits[sv["present_off"][s]: sv["present_off"][s] + w])
        bit_off.append(bit_off[-1] + w)
    bits = np.concatenate(row_bits) if row_bits else np.zeros(0, np.uint8)
    return dict(vk_pos=sv["vk_pos"][take], vk_key=sv["vk_key"][take], vk_off=vk_off, dense_pos=sv["dense_pos"][d_take],
                dense_key=sv["dense_key"][d_take], dense_range=np.stack([d_off[:-1], d_off[1:]], 1).astype(np.int32),
                dense_present=np.packbits(bits, bitorder="little"), dense_present_off=np.array(bit_off, np.int64),
                key_ilen=sv["key_ilen"], key_alt=sv["key_alt"], key_alt_off=sv["key_alt_off"])
