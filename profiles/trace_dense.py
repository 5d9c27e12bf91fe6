"""Where a CTA of hap_exec_oh_kernel spends its cycles, phase by phase (trace build only, never a bench number).

    GVL_LIB_NAME=libgvl_trace2.so GVL_EXTRA_NVCC_FLAGS=-DGVL_TRACE=2 python -m genvarloader_b200._build
    GVL_LIB_NAME=libgvl_trace2.so python profiles/trace_dense.py cfg2d [n_batches]

Thread 0 of every CTA accumulates clock64() deltas per phase over the passes of its tile (gvl_hap_oh.cuh, GVL_TC)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200 import Dataset, _ffi  # noqa: E402
from genvarloader_b200._pipeline import FixedPipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2d"
ring = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
w, d = bench.build_workload(name, 2)
L, pairs = w["window"], w["pairs"]
ds = Dataset.from_synth(dev, d, rng=0).with_tracks(False).with_len(L).with_encoding("onehot")
pipe = FixedPipeline(ds, pairs, ring=ring, graph=False)
H = pipe.halves[0]
n_q = ring * pairs
idx = torch.from_numpy(bench.draw_indices(d, n_q, 7, 0, 1)).to(dev)
H.idx[:n_q].copy_(idx)
trace = torch.zeros(8 * 400_000, dtype=torch.int64, device=dev)
for it in range(4):
    pipe._stage_plan(H.eng, H.scr, H.idx, None, n_q, sub_batch=pairs)
    torch.cuda.synchronize()
    if it == 3:
        _ffi.check(_ffi.lib.gvl_debug_set_trace(C.c_void_p(trace.data_ptr())))
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pipe._stage_exec(H.eng, H.scr, H.out, n_q, sub_batch=pairs)
    e.record()
    torch.cuda.synchronize()
    print(f"iter {it}: execute launch {a.elapsed_time(e) * 1e3:.1f} us (trace build)")
tr = trace.cpu().numpy().reshape(-1, 8)
tr = tr[tr[:, 6] > 0]
names = ["prologue (directory, row header, LUT)", "record staging (loads + barrier)", "group table (+ barrier)", "edge slots part 1 (classify, cp.async)",
         "group loop (loads, encode, stores)", "edge slots part 2 (blend, stores)"]
tot = tr[:, :6].sum(1)
print(f"{name}: {len(tr)} CTAs, passes per tile mean {tr[:, 6].mean():.2f}; thread-0 cycles per CTA: mean {tot.mean():.0f} = {tot.mean() / 1.965e3:.2f} us")
for i, nm in enumerate(names):
    x = tr[:, i]
    print(f"  {nm:42s} mean {x.mean():8.0f} cyc  {100 * x.sum() / tot.sum():5.1f} %   p50 {np.percentile(x, 50):8.0f}  p90 {np.percentile(x, 90):8.0f}")
