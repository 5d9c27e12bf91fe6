"""Host overhead of the public API: `Dataset.__getitem__` calls per second on the cfg2 shape (32 (region, sample) pairs =
64 haplotypes x 131,072 bp, fixed length, one-hot) -- numpy prep + one pinned upload + plan + execute per call, no sync
between calls.  Not a bench.py number; context for DESIGN.md section 5."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200 import Dataset  # noqa: E402

w, d = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2", 2)
L = w["window"]
ds = Dataset.from_synth(torch.device("cuda", 0), d, rng=0).with_tracks(False).with_len(L).with_encoding("onehot")
rng = np.random.default_rng(0)
idx = [(rng.integers(0, ds.n_regions, w["pairs"]), rng.integers(0, ds.n_samples, w["pairs"])) for _ in range(64)]
for r, s in idx[:8]:
    out = ds[r, s]
torch.cuda.synchronize()
n = 400
t0 = time.perf_counter()
for k in range(n):
    r, s = idx[k % len(idx)]
    out = ds[r, s]
t_host = (time.perf_counter() - t0) / n   # time until the calls have RETURNED (host side; the device runs behind)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"Dataset.__getitem__ ({w['pairs']} pairs x 2 haplotypes x {L} bp, one-hot): host {t_host * 1e6:.1f} us per call, "
      f"{dt * 1e6:.1f} us per call with the device drained -> {w['pairs'] * 2 * L / dt / 1e9:.1f} Gbp/s")
import cProfile
import pstats

pr = cProfile.Profile()
pr.enable()
for k in range(100):
    r, s = idx[k % len(idx)]
    out = ds[r, s]
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
