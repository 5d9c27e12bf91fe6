"""Tiny driver for ncu: builds one workload, runs a few eager steps (plan + execute, one-hot).

    ncu --set full --clock-control none --import-source on -k regex:hap_ -s 6 -c 4 -o gpurun_out/prof \
        python profiles/prof_driver.py cfg2 6
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200._engine import Engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "onehot"
dev = torch.device("cuda", 0)
w, d = bench.build_workload(name, 2)
batches = bench.make_batches(d, w, 4, 3)
eng = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs, d.geno_offsets)
L, rows = w["window"], w["pairs"] * 2
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for i in range(steps):
    b = batches[i % len(batches)]
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("regions", "shifts", "goi", "to_rc")}
    flush.zero_()
    eng.plan(t["regions"], t["shifts"], t["goi"], L, b["nvar"], to_rc=t["to_rc"])
    out = eng.execute(mode)
    torch.cuda.synchronize()
eng.check()
print("done", name, steps, int(out[:16].sum()))
