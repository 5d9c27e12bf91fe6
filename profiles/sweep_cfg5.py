"""BASELINE.json configs[4] on one GPU: window length x variant density sweep of the one-hot haplotype path
(plan + execute per step, SWEEP_SLOTS = 12 batches in flight, one CUDA graph per ring of SWEEP_RING = 64 batches,
>= 1 GiB written per measurement).
Prints a markdown table (profiles/r1_cfg5_sweep.md).  Multi-GPU: run under torchrun via bench.py --gpus N per cell."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200 import synth  # noqa: E402
from genvarloader_b200._engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak_gbs()
print("| window bp | variants/kb | rows/batch | Gbp/s | us/step | algorithmic GB/s | frac of %.0f GB/s |" % peak)
print("|---|---|---|---|---|---|---|")
for L in (16_384, 65_536, 131_072, 524_288):
    for vkb in (0.1, 1.0, 10.0):
        rows = max(32, (8 << 20) // L)  # ~8 Mbp per batch
        pairs = rows // 2
        n_regions = 16
        d = synth.make_dataset(5, max(4 * L, 2_000_000) * 4, 8, n_regions, L, vkb, neg_strand_frac=0.5, straddle_ends=False)
        w = dict(window=L, pairs=pairs)
        batches = bench.make_batches(d, w, int(os.environ.get("SWEEP_RING", 64)), 6)
        eng0 = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs, d.geno_offsets)
        n_streams = int(os.environ.get("SWEEP_SLOTS", 12))
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        slots = []
        for i, b in enumerate(batches):
            slots.append(dict(eng=eng0 if i == 0 else eng0.fork(), stream=streams[i % n_streams],
                              t={k: torch.from_numpy(b[k]).to(dev) for k in ("regions", "shifts", "goi", "to_rc")}, nvar=b["nvar"],
                              oo=torch.empty(rows + 1, dtype=torch.int64, device=dev),
                              out=torch.empty(rows * L * 4, dtype=torch.uint8, device=dev)))

        def step(s):
            t = s["t"]
            s["eng"].plan(t["regions"], t["shifts"], t["goi"], L, s["nvar"], to_rc=t["to_rc"], out_offsets=s["oo"])
            s["eng"].execute("onehot", out=s["out"])

        for s in slots:
            with torch.cuda.stream(s["stream"]):
                step(s)
        torch.cuda.synchronize()
        cap = torch.cuda.Stream(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            ev0 = torch.cuda.Event()
            ev0.record(cap)
            for st in streams:
                st.wait_event(ev0)
            for s in slots:
                with torch.cuda.stream(s["stream"]):
                    step(s)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                cap.wait_event(ev)
        reps = max(4, int((1 << 30) / (len(slots) * rows * L * 4)) + 1)
        with torch.cuda.stream(cap):
            for _ in range(2):
                g.replay()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(cap)
            for _ in range(reps):
                g.replay()
            e.record(cap)
        torch.cuda.synchronize()
        us = a.elapsed_time(e) * 1e3 / (reps * len(slots))
        nvar = float(np.mean([b["nvar"] for b in batches]))
        ab = rows * L * 5.0 + nvar * 29 + rows * 56
        print(f"| {L} | {vkb} | {rows} | {rows * L / us / 1e3:.0f} | {us:.2f} | {ab / us / 1e3:.0f} | {ab / us / 1e3 / peak:.2f} |", flush=True)
        del slots, eng0, g
        torch.cuda.empty_cache()
