"""`Dataset.with_seqs("variants")[r, s]` on the bench workloads' shapes: time per call (device drained) and variants per
second.  A call is gather_rows -> take(start, ilen) -> gather_alleles -> rc_alleles [-> dummy fill], with one host
synchronisation per ragged level (row offsets, allele offsets[, fill offsets]).  Not a bench.py number; context for
DESIGN.md section 4.7."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200 import Dataset, DummyVariant  # noqa: E402

for name, pairs in (("cfg2d", None), ("cfg4", None), ("cfg1", 8000)):
    w, d = bench.build_workload(name, 2)
    pairs = pairs or w["pairs"]
    if getattr(d, "geno_v_idxs", None) is None:
        continue
    try:
        ds = Dataset.from_synth(torch.device("cuda", 0), d, rng=0).with_tracks(False).with_seqs("variants")
    except NotImplementedError as e:  # (cfg2 is built on the svar2 source)
        print(f"{name}: {e}")
        continue
    rng = np.random.default_rng(0)
    idx = [(rng.integers(0, ds.n_regions, pairs), rng.integers(0, ds.n_samples, pairs)) for _ in range(8)]
    for label, dsv in (("alt,ilen,start", ds), ("+ dummy fill", ds.with_settings(dummy_variant=DummyVariant()))):
        for r, s in idx[:3]:
            out = dsv[r, s]
        torch.cuda.synchronize()
        n = 40
        t0 = time.perf_counter()
        for k in range(n):
            r, s = idx[k % len(idx)]
            out = dsv[r, s]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        nv = int(out.offsets[-1].item())
        nb = int(out["alt"].data.numel())
        print(f"{name} variants [{label}]: {pairs} pairs x {d.ploidy} rows, {nv} variants, {nb} ALT bytes per call: "
              f"{dt * 1e6:.0f} us per call = {nv / dt / 1e6:.1f} M variants/s")
