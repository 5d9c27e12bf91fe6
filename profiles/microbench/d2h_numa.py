"""D2H copy rate into pinned host memory, with the allocating thread bound to the GPU's local CPUs or not (context for the
host-buffer `e2e` path, which is one 67 MB device-to-host copy per step)."""
import os
import time

import torch

try:
    import pynvml

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n_words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
    local = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
except Exception as e:  # pragma: no cover
    local = []
    print("nvml:", e)
print("cpus:", os.cpu_count(), "allowed:", len(os.sched_getaffinity(0)), "gpu0-local:", len(local), local[:4], "...")
os.system("nvidia-smi topo -m 2>/dev/null | head -6; lscpu | grep -i 'numa\\|socket' | head -6")
dev = torch.device("cuda", 0)
src = torch.empty(64 << 20, dtype=torch.uint8, device=dev)


def rate(tag):
    dst = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
    dst.zero_()  # first touch
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    print(f"{tag}: {20 * (64 << 20) / (a.elapsed_time(e) * 1e-3) / 1e9:.1f} GB/s")


all_cpus = os.sched_getaffinity(0)
rate("default affinity")
if local:
    ok = sorted(set(local) & all_cpus)
    far = sorted(all_cpus - set(local))
    if ok:
        os.sched_setaffinity(0, ok)
        rate(f"bound to the GPU's local cpus ({len(ok)})")
    if far:
        os.sched_setaffinity(0, far)
        rate(f"bound to the OTHER cpus ({len(far)})")
    os.sched_setaffinity(0, all_cpus)
