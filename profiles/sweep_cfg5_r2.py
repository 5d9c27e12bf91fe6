"""BASELINE.json configs[4] on one GPU, round 2: window length x variant density sweep of the one-hot haplotype path THROUGH
THE PRODUCT PIPELINE (`bench.py`, one cell per run: rings of batches per device call, two halves, CUDA graphs; ~8 Mbp per batch,
--steps 640 so that the plan of one call runs under the execute of the previous one).  Prints the table of
profiles/r2_cfg5_sweep.md.

    python profiles/sweep_cfg5_r2.py [steps]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
steps = sys.argv[1] if len(sys.argv) > 1 else "640"
print("| window bp | variants/kb | rows/batch | Gbp/s | us/step | `whole_step_frac` | execute launch `roofline.frac` | prep + plan per call |")
print("|---|---|---|---|---|---|---|---|")
for L in (16_384, 65_536, 131_072, 524_288):
    for vkb in (0.1, 1.0, 10.0):
        rows = max(32, (8 << 20) // L)
        w = dict(desc=f"configs[4] cell: {L}-bp windows, {vkb} variants/kb, {rows} haplotypes/batch", contig_len=max(64 * L, 20_000_000),
                 n_samples=8, n_regions=max(64, min(512, 33_554_432 // L)), window=L, pairs=rows // 2, vkb=vkb, neg=0.5)
        env = dict(os.environ, GVL_BENCH_CUSTOM=json.dumps(w))
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "custom", "--steps", steps, "--warmup", "16",
                              "--cpu-seconds", "0.1"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            r = d["roofline"]
            print(f"| {L} | {vkb} | {rows} | {d['value'] / 1e9:.0f} | {d['ms_per_step'] * 1e3:.2f} | {d['whole_step_frac']:.2f} | {r['frac']:.2f} | "
                  f"{r['plan_kernel_ms'] * 1e3:.0f} us / {r['batches_per_launch']} batches |", flush=True)
        except Exception as e:  # pragma: no cover
            print(f"| {L} | {vkb} | {rows} | failed: {e} {out.stderr[-200:]!r} |", flush=True)
