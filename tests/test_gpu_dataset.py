"""GPU Dataset vs the oracle driven by a numpy restatement of the reference's read-time prep
(jitter / strand / shifts / seeds, python/genvarloader/_dataset/_query.py:153-204, _haps.py:678-768,
_reconstruct.py:168-300).  Bit-exact; RNG draws follow the reference's order so seeded runs match."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = ord("N")


@pytest.fixture(scope="module")
def env(cuda_device):
    from genvarloader_b200 import synth
    from genvarloader_b200._dataset import Dataset
    from oracle import oracle as O

    d = synth.make_dataset(77, 400_000, 5, 14, 3000 + 2 * 16, 5.0, max_jitter=16, neg_strand_frac=0.5,
                           straddle_ends=False, n_tracks=2, max_indel=18, snp_frac=0.5)
    ds = Dataset.from_synth(cuda_device, d, rng=123)
    return d, ds, O, synth


def _prep(d, ds_idx, jitter, rng):
    S = d.n_samples
    r_idx, s_idx = ds_idx // S, ds_idx % S
    regions = d.regions[r_idx].copy()
    lengths = regions[:, 2] - regions[:, 1]
    regions[:, 1] += rng.integers(-jitter, jitter + 1, size=len(regions), dtype=np.int32)
    regions[:, 2] = regions[:, 1] + lengths
    goi = ds_idx[:, None] * d.ploidy + np.arange(d.ploidy)[None, :]
    to_rc = np.repeat(d.regions[r_idx, 3] == -1, d.ploidy)
    return r_idx, np.ascontiguousarray(regions[:, :3]), goi, to_rc


def _hap_args(d, regions, shifts, goi, out_len):
    return (regions, shifts, goi, d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
            d.reference, d.ref_offsets, N, out_len)


def test_fixed_length_haplotypes_shapes_and_values(env):
    d, ds, O, _ = env
    L = 2048
    dsl = ds.with_len(L).with_tracks(False)
    out = dsl[:4, :3]
    assert tuple(out.shape) == (4, 3, 2, L) and out.dtype.__str__() == "torch.uint8"
    ds_idx = (np.arange(4)[:, None] * d.n_samples + np.arange(3)[None, :]).ravel()
    _, regions, goi, to_rc = _prep(d, ds_idx, 0, np.random.default_rng(0))
    exp, _ = O.reconstruct_haplotypes_fused(*_hap_args(d, regions, np.zeros(goi.shape, np.int32), goi, L), None, None, to_rc)
    assert (out.cpu().numpy().ravel() == exp).all()
    one = dsl[2, 1]
    assert tuple(one.shape) == (2, L)
    assert (one.cpu().numpy() == out[2, 1].cpu().numpy()).all()
    col = dsl[[0, 3, 3], [2, 0, 1]]
    assert tuple(col.shape) == (3, 2, L)
    assert (col[1].cpu().numpy() == out[3, 0].cpu().numpy()).all()


def test_ragged_variable_annotated_and_onehot(env):
    d, ds, O, _ = env
    base = ds.with_tracks(False)
    ds_idx = np.array([1 * d.n_samples + 2, 6 * d.n_samples + 0, 9 * d.n_samples + 4])
    _, regions, goi, to_rc = _prep(d, ds_idx, 0, np.random.default_rng(0))
    a = _hap_args(d, regions, np.zeros(goi.shape, np.int32), goi, -1)
    e_out, e_av, e_ap, e_oo = O.reconstruct_annotated_haplotypes_fused(*a, None, None, to_rc)
    idx = ([1, 6, 9], [2, 0, 4])
    rag = base[idx]
    assert rag.shape == (3, 2, None)
    assert (rag.offsets.cpu().numpy() == e_oo).all() and (rag.data.cpu().numpy() == e_out).all()
    ann = base.with_seqs("annotated")[idx]
    assert (ann.haps.data.cpu().numpy() == e_out).all()
    assert (ann.var_idxs.data.cpu().numpy() == e_av).all() and (ann.ref_coords.data.cpu().numpy() == e_ap).all()
    var = base.with_len("variable")[idx]
    lens = np.diff(e_oo)
    assert tuple(var.shape) == (3, 2, lens.max())
    v = var.cpu().numpy().reshape(6, -1)
    for k in range(6):
        assert (v[k, :lens[k]] == e_out[e_oo[k]:e_oo[k + 1]]).all() and (v[k, lens[k]:] == N).all()
    annv = base.with_seqs("annotated").with_len("variable")[idx]
    assert (annv.ref_coords.cpu().numpy().reshape(6, -1)[0, lens[0]:] == np.iinfo(np.int32).max).all()
    oh = base.with_encoding("onehot")[idx]
    assert (oh.data.cpu().numpy() == O.onehot(e_out)).all()
    L = 1024
    cf = base.with_len(L).with_encoding("onehot_cf")[idx]
    e_fix, _ = O.reconstruct_haplotypes_fused(*_hap_args(d, regions, np.zeros(goi.shape, np.int32), goi, L), None, None, to_rc)
    assert (cf.cpu().numpy() == O.onehot(e_fix).reshape(3, 2, L, 4).transpose(0, 1, 3, 2)).all()


def test_jitter_and_random_shifts_follow_reference_rng_order(env):
    d, ds, O, _ = env
    L, J = 2500, 16
    dsj = ds.with_tracks(False).with_len(L).with_settings(jitter=J, deterministic=False, rng=2024)
    out = dsj[3:9, 1:3]
    ds_idx = (np.arange(3, 9)[:, None] * d.n_samples + np.arange(1, 3)[None, :]).ravel()
    rng = np.random.default_rng(2024)
    _, regions, goi, to_rc = _prep(d, ds_idx, J, rng)  # jitter draw first (_query.py:167-171)
    diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
    lengths = regions[:, 2] - regions[:, 1]
    max_shift = diffs.clip(min=0) + (lengths - L).clip(min=0)[:, None]
    shifts = rng.integers(0, max_shift + 1, dtype=np.int32)  # then the shifts (_haps.py:728-730)
    assert shifts.max() > 0
    exp, _ = O.reconstruct_haplotypes_fused(*_hap_args(d, regions, shifts, goi, L), None, None, to_rc)
    assert (out.cpu().numpy().ravel() == exp).all()


@pytest.mark.parametrize("fixed", [True, False])
def test_haplotypes_plus_realigned_tracks(env, fixed):
    from genvarloader_b200 import FlankSample, Interpolate

    d, ds, O, _ = env
    L = 2400
    dst = ds.with_insertion_fill({"track0": Interpolate(2), "track1": FlankSample(3)})
    if fixed:
        dst = dst.with_len(L)
    idx = ([2, 5, 5, 11], [0, 3, 4, 1])
    haps, tracks = dst[idx]
    ds_idx = np.array([2 * d.n_samples + 0, 5 * d.n_samples + 3, 5 * d.n_samples + 4, 11 * d.n_samples + 1])
    r_idx, regions, goi, to_rc = _prep(d, ds_idx, 0, np.random.default_rng(0))
    sh = np.zeros(goi.shape, np.int32)
    diffs = O.get_diffs_sparse(goi, d.geno_v_idxs, d.geno_offsets, d.ilens, None, None, regions[:, 1], regions[:, 2], d.v_starts)
    lengths = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    out_len = np.full(goi.shape, L, np.int64) if fixed else lengths[:, None] + diffs
    oo = np.concatenate([[0], np.cumsum(out_len.ravel())]).astype(np.int64)
    to = np.concatenate([[0], np.cumsum(lengths - diffs.clip(max=0).min(1))]).astype(np.int64)
    seed = int(np.bitwise_xor.reduce(ds_idx.astype(np.uint64)))
    exp_t = []
    for name, (sid, par) in zip(("track0", "track1"), ((4, 2.0), (3, 3.0))):
        s, e, v, io = d.tracks[name]
        buf = np.zeros(int(oo[-1]), np.float32)
        O.intervals_and_realign_track_fused(buf, oo, regions, sh, goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens,
                                            ds_idx, s, e, v, io, to, np.array([par]), sid, seed, None, None, to_rc)
        exp_t.append(buf)
    b, p = goi.shape
    if fixed:
        assert tuple(tracks.shape) == (b, 2, p, L)
        got = tracks.cpu().numpy()
        for ti in range(2):
            assert (got[:, ti].reshape(-1).view(np.uint32) == exp_t[ti].view(np.uint32)).all()
        e_h, _ = O.reconstruct_haplotypes_fused(*_hap_args(d, regions, sh, goi, L), None, None, to_rc)
        assert (haps.cpu().numpy().ravel() == e_h).all()
    else:
        assert tracks.shape == (b, 2, p, None)
        data, offs = tracks.data.cpu().numpy(), tracks.offsets.cpu().numpy()
        k = 0
        for q in range(b):
            for ti in range(2):
                for h in range(p):
                    row = q * p + h
                    seg = data[offs[k]:offs[k + 1]]
                    assert (seg.view(np.uint32) == exp_t[ti][oo[row]:oo[row + 1]].view(np.uint32)).all()
                    k += 1


def test_reference_sequences_and_unrealigned_tracks(env):
    d, ds, O, _ = env
    L = 1500
    dsr = ds.with_seqs("reference").with_len(L).with_tracks(["track1"])
    seq, trk = dsr[:5, 2]
    assert tuple(seq.shape) == (5, L) and tuple(trk.shape) == (5, 1, L)
    ds_idx = np.arange(5) * d.n_samples + 2
    r_idx, regions, goi, to_rc = _prep(d, ds_idx, 0, np.random.default_rng(0))
    oo = np.arange(6, dtype=np.int64) * L
    reg_fixed = regions.copy()
    reg_fixed[:, 2] = reg_fixed[:, 1] + L
    e_ref = O.get_reference(reg_fixed, oo, d.reference, d.ref_offsets, N, False, d.regions[r_idx, 3] == -1)
    assert (seq.cpu().numpy().ravel() == e_ref).all()
    def paint(name, offs):
        s, e, v, io = d.tracks[name]
        buf = np.zeros(int(offs[-1]), np.float32)
        O.intervals_to_tracks(ds_idx, regions[:, 1], s, e, v, io, buf, offs)
        O.reverse_flat_rows_inplace(buf, offs, d.regions[r_idx, 3] == -1)
        return buf

    assert (trk.cpu().numpy().ravel().view(np.uint32) == paint("track1", oo).view(np.uint32)).all()
    # two tracks in one launch, (b, t, l) order; then ragged rows (lengths = region lengths)
    _, trk2 = ds.with_seqs("reference").with_len(L).with_tracks(["track1", "track0"])[:5, 2]
    got2 = trk2.cpu().numpy()
    assert got2.shape == (5, 2, L)
    for ti, name in enumerate(("track1", "track0")):
        assert (got2[:, ti].ravel().view(np.uint32) == paint(name, oo).view(np.uint32)).all()
    _, trk3 = ds.with_seqs("reference").with_tracks(["track0", "track1"])[:5, 2]
    lens = (regions[:, 2] - regions[:, 1]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    data, offs = trk3.data.cpu().numpy(), trk3.offsets.cpu().numpy()
    exp3 = [paint("track0", ro), paint("track1", ro)]
    assert trk3.shape == (5, 2, None) and offs[-1] == 2 * ro[-1]
    for q in range(5):
        for ti in range(2):
            seg = data[offs[q * 2 + ti]:offs[q * 2 + ti + 1]]
            assert (seg.view(np.uint32) == exp3[ti][ro[q]:ro[q + 1]].view(np.uint32)).all()


def test_exonic_filter_and_subset(env):
    d, ds, O, _ = env
    sub = ds.with_tracks(False).with_settings(var_filter="exonic").subset_to(regions=[8, 3], samples=[4, 1, 0])
    out = sub[:, :]
    assert out.shape == (2, 3, 2, None)
    ds_idx = (np.array([8, 3])[:, None] * d.n_samples + np.array([4, 1, 0])[None, :]).ravel()
    _, regions, goi, to_rc = _prep(d, ds_idx, 0, np.random.default_rng(0))
    keep, ko = O.choose_exonic_variants(regions[:, 1], regions[:, 2], goi, d.geno_v_idxs, d.geno_offsets, d.v_starts, d.ilens)
    exp, eoo = O.reconstruct_haplotypes_fused(*_hap_args(d, regions, np.zeros(goi.shape, np.int32), goi, -1), keep, ko, to_rc)
    assert (out.offsets.cpu().numpy() == eoo).all() and (out.data.cpu().numpy() == exp).all()


def test_to_dataloader_batches_match_direct_indexing(env):
    """`to_dataloader` (reference _impl.py:1963-2072): batches over the flat (region, sample) index, born on the GPU."""
    d, ds, O, _ = env
    dsl = ds.with_len(1024).with_tracks(False).with_encoding("onehot")
    dl = dsl.to_dataloader(batch_size=7, shuffle=False, return_indices=True, num_workers=4, pin_memory=True)
    n = 0
    for k, (x, r, s) in enumerate(dl):
        assert x.is_cuda and x.shape[1:] == (d.ploidy, 1024, 4)
        idx = np.arange(k * 7, min((k + 1) * 7, len(dsl)))
        assert (r == idx // dsl.n_samples).all() and (s == idx % dsl.n_samples).all()
        assert (x == dsl[idx // dsl.n_samples, idx % dsl.n_samples]).all()
        n += x.shape[0]
    assert n == len(dsl) and len(dl) == -(-len(dsl) // 7)
    sh = dsl.to_dataloader(batch_size=16, shuffle=True, drop_last=True, generator=3)
    assert sum(b.shape[0] for b in sh) == (len(dsl) // 16) * 16


def test_read_side_helpers_like_the_reference(env):
    """n_variants / n_intervals / haplotype_lengths / to_torch_dataset / has_* (reference _impl.py:954-1110, 1234-1340,
    1848-1895, _torch.py:272-307) against the synthetic tables and the dataset's own ragged output."""
    d, ds, O, _ = env
    assert ds.has_reference and ds.has_genotypes and ds.has_intervals and "shape" in repr(ds)
    r, s = np.array([0, 3, 5, 3]), np.array([1, 0, 2, 0])
    goi = (r[:, None] * d.n_samples + s[:, None]) * d.ploidy + np.arange(d.ploidy)[None, :]
    go = np.asarray(d.geno_offsets).reshape(2, -1) if np.asarray(d.geno_offsets).ndim == 2 else np.stack([d.geno_offsets[:-1], d.geno_offsets[1:]])
    assert (ds.n_variants(r, s) == (go[1, goi] - go[0, goi])).all()
    assert ds.n_variants().shape == (ds.n_regions, ds.n_samples, d.ploidy) and ds.n_variants(2, 1).shape == (d.ploidy,)
    ni = ds.n_intervals(r, s)
    assert ni.shape == (4, len(ds.active_tracks))
    for t, name in enumerate(ds.active_tracks):
        off = np.asarray(d.tracks[name][3])
        slot = r * d.n_samples + s
        assert (ni[:, t] == off[slot + 1] - off[slot]).all()
    # haplotype lengths = row lengths of the ragged output (jitter 0), and grow by 2 * jitter with it
    rag = ds.with_tracks(False).with_len("ragged")[r, s]
    lens = (rag.offsets[1:] - rag.offsets[:-1]).cpu().numpy().reshape(4, d.ploidy)
    assert (ds.haplotype_lengths(r, s) == lens).all()
    assert ds.with_settings(deterministic=False).haplotype_lengths(r, s) is None
    tds = ds.with_tracks(False).with_len(512).to_torch_dataset(return_indices=True)
    x, ri, si = tds[[0, 7, 9]]
    assert len(tds) == len(ds) and x.shape[0] == 3 and (ri == np.array([0, 7, 9]) // ds.n_samples).all()
    assert (x == ds.with_tracks(False).with_len(512)[ri, si]).all()
