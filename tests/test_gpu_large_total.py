"""Maximum sizes: ONE call whose output has more than 2^31 positions (2.2 x 10^9 bp: 8.8 GB of one-hot, 8.8 GB of one
float track, ragged offsets beyond int32).  The oracle cannot hold this in seconds, so the giant call is compared ON THE
DEVICE with the same rows computed in chunks of 700 (region, sample) pairs -- the small-call path that every other
parity test pins to the oracle and the reference's goldens."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_more_than_2_31_positions_in_one_call(cuda_device):
    import torch

    from genvarloader_b200 import synth
    from genvarloader_b200._engine import Engine

    dev = cuda_device
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 60 << 30:
        pytest.skip("needs ~45 GB of device memory")
    L = 131_072
    d = synth.make_dataset(3, 6_000_000, 32, 32, L, 1.0, neg_strand_frac=0.5, straddle_ends=False, n_tracks=1)
    eng = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs,
                 d.geno_offsets)
    eng.add_track("track0", *d.tracks["track0"])
    b = 8_400  # 16,800 rows x 131,072 = 2.2 x 10^9 positions
    rng = np.random.default_rng(9)
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
    p = goi.shape[1]
    assert b * p * L > 2**31
    t_reg, t_goi = torch.from_numpy(regions).to(dev), torch.from_numpy(goi).to(dev)
    t_sh = torch.zeros(goi.shape, dtype=torch.int32, device=dev)
    t_rc = torch.from_numpy(to_rc).to(dev)
    chunks = [(lo, min(lo + 700, b)) for lo in range(0, b, 700)]
    small = eng.fork()

    def sub(lo, hi):
        return t_reg[lo:hi].contiguous(), t_sh[lo:hi].contiguous(), t_goi[lo:hi].contiguous(), t_rc[lo * p:hi * p].contiguous()

    # ---- fixed length, one-hot (packed kernel): 8.8 GB ----
    eng.plan(t_reg, t_sh, t_goi, L, eng.max_records(goi), to_rc=t_rc)
    big = eng.execute("onehot")
    assert big.numel() == b * p * L * 4
    for lo, hi in chunks:
        r, s, g, c = sub(lo, hi)
        small.plan(r, s, g, L, small.max_records(goi[lo:hi]), to_rc=c)
        assert torch.equal(small.execute("onehot"), big[lo * p * L * 4: hi * p * L * 4]), f"one-hot rows {lo}:{hi}"
    del big

    # ---- ragged, annotated: offsets beyond int32, bytes + 2 x int32 annotations ----
    oo = eng.plan(t_reg, t_sh, t_goi, -1, eng.max_records(goi), to_rc=t_rc)
    total = eng.total()
    assert total > 2**31 and int(oo[-1]) == total
    h, av, ap = eng.execute("annotated")
    for lo, hi in chunks:
        r, s, g, c = sub(lo, hi)
        so = small.plan(r, s, g, -1, small.max_records(goi[lo:hi]), to_rc=c)
        sh_, sv, sp = small.execute("annotated")
        a, e = int(oo[lo * p]), int(oo[hi * p])
        assert torch.equal(so + a, oo[lo * p: hi * p + 1])
        assert torch.equal(sh_, h[a:e]) and torch.equal(sv, av[a:e]) and torch.equal(sp, ap[a:e]), f"annotated rows {lo}:{hi}"
    del h, av, ap

    # ---- one realigned track, fixed length: 2.2 x 10^9 float32 values ----
    oo_f = torch.arange(b * p + 1, dtype=torch.int64, device=dev) * L
    oidx = torch.from_numpy(ds_idx[None, :].copy()).to(dev)
    tlen = torch.full((b,), L + 4096, dtype=torch.int32, device=dev)  # source window: query + room for net deletions
    big_t = eng.realign_tracks(["track0"], t_reg, t_sh, t_goi, oidx, tlen, oo_f, b * p * L, [4], [1.0], 11,
                               eng.max_records(goi), to_rc=t_rc, query_seed=torch.arange(b, dtype=torch.int64, device=dev))
    for lo, hi in chunks[::3]:
        r, s, g, c = sub(lo, hi)
        n = (hi - lo) * p * L
        got = small.realign_tracks(["track0"], r, s, g, oidx[:, lo:hi].contiguous(), tlen[lo:hi].contiguous(),
                                   oo_f[: (hi - lo) * p + 1].contiguous(), n, [4], [1.0], 11, small.max_records(goi[lo:hi]),
                                   to_rc=c, query_seed=torch.arange(lo, hi, dtype=torch.int64, device=dev))
        assert torch.equal(got.view(torch.int32), big_t[lo * p * L: hi * p * L].view(torch.int32)), f"track rows {lo}:{hi}"
    eng.check()
    small.check()
