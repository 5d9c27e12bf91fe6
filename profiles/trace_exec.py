"""Per-CTA timeline of hap_exec_oh_kernel (trace build only, never a bench number).

    GVL_LIB_NAME=libgvl_trace.so GVL_EXTRA_NVCC_FLAGS=-DGVL_TRACE=1 python -m genvarloader_b200._build
    GVL_LIB_NAME=libgvl_trace.so python profiles/trace_exec.py cfg2
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from genvarloader_b200 import _ffi  # noqa: E402
from genvarloader_b200._engine import Engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda", 0)
w, d = bench.build_workload(name, 2)
batches = bench.make_batches(d, w, 4, 3)
eng = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs, d.geno_offsets)
L, rows = w["window"], w["pairs"] * 2
trace = torch.zeros(65536 * 6, dtype=torch.int64, device=dev)
out = None
for it in range(6):
    b = batches[it % len(batches)]
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("regions", "shifts", "goi", "to_rc")}
    eng.plan(t["regions"], t["shifts"], t["goi"], L, b["nvar"], to_rc=t["to_rc"])
    torch.cuda.synchronize()
    if it == 5:
        _ffi.check(_ffi.lib.gvl_debug_set_trace(C.c_void_p(trace.data_ptr())))
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = eng.execute("onehot", out=out)
    e.record()
    torch.cuda.synchronize()
    print(f"iter {it}: event time around the execute launch {a.elapsed_time(e) * 1e3:.2f} us")
tr = trace.cpu().numpy().reshape(-1, 6)
tr = tr[tr[:, 1] > 0]
n = len(tr)
t0 = tr[:, 1].min()
st, t_hdr, t_loop, t_end = (tr[:, i] - t0 for i in (1, 2, 3, 4))
print(f"{name}: {n} CTAs on {len(np.unique(tr[:, 0]))} SMs; kernel span (first CTA start -> last CTA end) {t_end.max() / 1e3:.2f} us")
q = lambda x: "min %.2f  p10 %.2f  p50 %.2f  p90 %.2f  max %.2f us" % tuple(np.percentile(x, [0, 10, 50, 90, 100]) / 1e3)
print("CTA start (dispatch ramp)      ", q(st))
print("header + directory round trip  ", q(t_hdr - st))
print("records staged + group table   ", q(t_loop - t_hdr))
print("streaming (loads/encode/stores)", q(t_end - t_loop))
print("CTA lifetime                   ", q(t_end - st))
print("CTA end                        ", q(t_end))
