"""Micro-probe: execute-kernel time per mode, with and without the L2 flush (not a bench number)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200._engine import Engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda", 0)
w, d = bench.build_workload(name, 2)
batches = bench.make_batches(d, w, 4, 3)
eng = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs, d.geno_offsets)
L, rows = w["window"], w["pairs"] * 2
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
b = batches[0]
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("regions", "shifts", "goi", "to_rc")}
eng.plan(t["regions"], t["shifts"], t["goi"], L, b["nvar"], to_rc=t["to_rc"])
outs = {m: eng.execute(m) for m in ("u8", "onehot", "annotated", "onehot_cf")}
torch.cuda.synchronize()
main = torch.cuda.current_stream()
for do_flush in (True, False):
    for m in ("u8", "onehot", "onehot_cf", "annotated"):
        durs = []
        for i in range(30):
            if do_flush:
                flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main)
            if m == "annotated":
                eng.execute(m, out=outs[m][0], annot_v=outs[m][1], annot_pos=outs[m][2])
            else:
                eng.execute(m, out=outs[m])
            e.record(main)
            torch.cuda.synchronize()
            if i >= 5:
                durs.append(a.elapsed_time(e) * 1e3)
        print(f"{name} flush={do_flush} mode={m:10s} exec us: mean {np.mean(durs):7.2f} min {np.min(durs):7.2f}")
# empty-kernel launch + event overhead reference
durs = []
x = torch.zeros(1, device=dev)
for i in range(30):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    x.add_(1)
    e.record(main)
    torch.cuda.synchronize()
    durs.append(a.elapsed_time(e) * 1e3)
print(f"tiny torch kernel between events us: mean {np.mean(durs[5:]):.2f} min {np.min(durs):.2f}")
# a plain device copy of the same byte volume as the one-hot output (for scale)
src = torch.empty(rows * L * 4, dtype=torch.uint8, device=dev)
dst = torch.empty_like(src)
durs = []
for i in range(30):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    dst.copy_(src)
    e.record(main)
    torch.cuda.synchronize()
    durs.append(a.elapsed_time(e) * 1e3)
print(f"torch copy of {src.numel()>>20} MiB (read+write) after flush us: mean {np.mean(durs[5:]):.2f} min {np.min(durs):.2f}")
durs = []
for i in range(30):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    dst.zero_()
    e.record(main)
    torch.cuda.synchronize()
    durs.append(a.elapsed_time(e) * 1e3)
print(f"torch fill of {src.numel()>>20} MiB (write only) after flush us: mean {np.mean(durs[5:]):.2f} min {np.min(durs):.2f}")
