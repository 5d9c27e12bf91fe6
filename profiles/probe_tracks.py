"""Track realignment (gvl_dev_realign_tracks) on the cfg3 shape: 32 haplotype rows x 524,288 values x 2 tracks
(Repeat5p + Interpolate), intervals with a mean run of 50 bp.  Back-to-back eager calls over a ring of 4
output buffers (4 x 134 MB > L2).  Not a bench.py number; context for DESIGN.md section 4.3."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from genvarloader_b200 import synth  # noqa: E402
from genvarloader_b200._engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
L = 524_288
import os
d = synth.cfg3(n_samples=8, variants_per_kb=float(os.environ.get('PROBE_VKB', 1.0)))
eng = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs, d.geno_offsets)
names = sorted(d.tracks)
n_itv = 0
for n in names:
    eng.add_track(n, *d.tracks[n])
    n_itv += int(d.tracks[n][0].size)
rng = np.random.default_rng(5)
import os
b = int(os.environ.get('PROBE_B', 16))
ring = []
for i in range(int(os.environ.get('PROBE_RING', 4))):
    r_idx, s_idx = rng.integers(0, d.n_regions, b), rng.integers(0, d.n_samples, b)
    regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx, rng, jitter=128)
    regions[:, 2] = regions[:, 1] + L  # fixed-length query inside the stored window
    nvar = int((d.geno_offsets[1, goi.ravel()] - d.geno_offsets[0, goi.ravel()]).sum())
    t = dict(regions=torch.from_numpy(regions).to(dev), shifts=torch.zeros(goi.shape, dtype=torch.int32, device=dev),
             goi=torch.from_numpy(goi).to(dev), to_rc=torch.from_numpy(to_rc).to(dev),
             oidx=torch.from_numpy(np.tile(ds_idx, (len(names), 1))).to(dev),
             tlen=torch.full((b,), L + 4096, dtype=torch.int32, device=dev),  # source window: query + room for net deletions
             oo=(torch.arange(goi.size + 1, dtype=torch.int64, device=dev) * L), nvar=nvar,
             out=torch.empty(len(names) * goi.size * L, dtype=torch.float32, device=dev), eng=eng if i == 0 else eng.fork())
    ring.append(t)


def call(t):
    t["eng"].realign_tracks(names, t["regions"], t["shifts"], t["goi"], t["oidx"], t["tlen"], t["oo"], t["goi"].numel() * L,
                            [0, 4], [0.0, 1.0], 7, t["nvar"], to_rc=t["to_rc"], out=t["out"])


for t in ring:
    call(t)
torch.cuda.synchronize()
# (eager launches: the per-track descriptors are copied from HOST arrays at call time, so this entry is not
#  meant to be captured in a CUDA graph with temporary host arguments)
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
n_it = 25
for _ in range(n_it):
    for t in ring:
        call(t)
e.record()
torch.cuda.synchronize()
us = a.elapsed_time(e) * 1e3 / (n_it * len(ring))
vals = len(names) * 2 * b * L
byts = 4.0 * vals  # + 12 B per interval actually read (a fraction of the stored ones)
print(f"tracks cfg3: {len(names)} tracks x {2 * b} rows x {L} values = {vals / 1e6:.1f} M values, {byts / 1e6:.0f} MB written per call")
print(f"realign_tracks (plan + execute, one stream, back to back): {us:.1f} us per call -> {byts / us / 1e3:.0f} GB/s written, "
      f"{vals / us / 1e3:.1f} G values/s")

if os.environ.get("GVL_LIB_NAME", "").startswith("libgvl_trace"):
    import ctypes as C
    from genvarloader_b200 import _ffi
    tr = torch.zeros(4096 * 64, dtype=torch.int64, device=dev)
    _ffi.check(_ffi.lib.gvl_debug_set_trk_trace(C.c_void_p(tr.data_ptr())))
    call(ring[0])
    torch.cuda.synchronize()
    a = tr.cpu().numpy().reshape(-1, 64)
    a = a[a[:, 0] > 0]
    t0 = a[:, 0].min()
    print(f"trace: {len(a)} CTAs; kernel span {(a[:, :60].max() - t0) / 1e3:.1f} us")
    full = tr.cpu().numpy().reshape(-1, 64)
    n_cta = full.shape[0]
    ends = full[:, :60].max(1)
    live = full[:, 0] > 0
    half = np.flatnonzero(live).max() // 2 + 1
    for nm, sel in (("track 0 (Repeat5p)", live & (np.arange(n_cta) < half)), ("track 1 (Interpolate)", live & (np.arange(n_cta) >= half))):
        e_ = (ends[sel] - t0) / 1e3
        print(f"  {nm}: CTA end p10 {np.percentile(e_, 10):.1f} p50 {np.percentile(e_, 50):.1f} p90 {np.percentile(e_, 90):.1f} max {e_.max():.1f} us")
    names_ = ["stage records", "window/interval search", "markers", "scan+fill", "group loop (output)"]
    for ps in range(8):
        st = a[:, ps * 6:ps * 6 + 6]
        ok = st[:, 5] > 0
        if not ok.any():
            break
        d_ = np.diff(st[ok], axis=1) / 1e3
        print(f"pass {ps}: start p50 {(np.median(st[ok, 0]) - t0) / 1e3:7.1f} us | " + " | ".join(f"{n} {np.median(d_[:, i]):.2f}" for i, n in enumerate(names_)) + f" | total {np.median(d_.sum(1)):.2f} us")
