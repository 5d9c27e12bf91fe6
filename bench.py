#!/usr/bin/env python
"""bench.py -- haplotype bp/s (one-hot) of the B200-native path, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg1|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic (region, sample) pairs: plan
(variant state machine -> segment table) + execute (fused copy / ALT scatter / pad / RC / one-hot),
with every input already resident in HBM.  Default workload = BASELINE.json configs[1]
(131,072-bp windows, 64 haplotypes per batch, 2,504-sample cohort on a 50 Mb contig).

Printed JSON line (rank 0): the driver's contract + `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches`.  See DESIGN.md "Measurement" for every definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_CHAR = ord("N")
ALG_BYTES_PER_BP = 5.0        # 1 reference byte read + 4 one-hot bytes written
ALG_BYTES_PER_VARIANT = 29.0  # 4 v_idx + 4 pos + 4 ilen + 16 alt_offsets + >=1 ALT
ALG_BYTES_PER_ROW = 56.0


# ----------------------------------------------------------------------------------------------
# workloads (SURVEY.md 8d / BASELINE.json configs)
# ----------------------------------------------------------------------------------------------
WORKLOADS = {
    # name: (description, builder kwargs, window, pairs per batch, mode)
    "cfg1": dict(desc="configs[0]: 1 Mb contig, 8 diploid samples, ~1 variant/kb, 16,384-bp windows, 64 haplotypes/batch",
                 contig_len=1_000_000, n_samples=8, n_regions=1000, window=16_384, pairs=32, vkb=1.0),
    "cfg2": dict(desc="configs[1]: 50 Mb contig, 2,504 diploid samples, ~1 variant/kb/haplotype (~0.55 M variant table), "
                      "131,072-bp windows, 64 haplotypes/batch",
                 contig_len=50_000_000, n_samples=2504, n_regions=16, window=131_072, pairs=32, vkb=1.0),
    "cfg3": dict(desc="configs[2] (haplotype part): 524,288-bp indel-bearing windows, 32 haplotypes/batch, 50% negative strand",
                 contig_len=20_000_000, n_samples=16, n_regions=16, window=524_288, pairs=16, vkb=1.0, neg=0.5),
    "cfg3t": dict(desc="configs[2]: 524,288-bp indel-bearing windows, 32 haplotypes/batch, 50% negative strand, PLUS 2 realigned "
                       "float tracks per haplotype (intervals with a mean run of 50 bp; fills Repeat5p and Interpolate(1))",
                  contig_len=20_000_000, n_samples=16, n_regions=16, window=524_288, pairs=16, vkb=1.0, neg=0.5, tracks=2),
    "cfg2d": dict(desc="configs[4] dense cell: 131,072-bp windows, 64 haplotypes/batch, ~10 variants/kb/haplotype",
                  contig_len=8_000_000, n_samples=16, n_regions=16, window=131_072, pairs=32, vkb=10.0, neg=0.5),
    "cfg4": dict(desc="configs[3] (one-hot instead of annotated): 6,144-bp windows, 4,096 haplotypes/batch",
                 contig_len=5_000_000, n_samples=64, n_regions=512, window=6_144, pairs=2048, vkb=1.0),
}


def build_workload(name: str, seed: int):
    from genvarloader_b200 import synth

    w = WORKLOADS[name]
    d = synth.make_dataset(seed, w["contig_len"], w["n_samples"], w["n_regions"], w["window"], w["vkb"],
                           neg_strand_frac=w.get("neg", 0.0), straddle_ends=False, n_tracks=w.get("tracks", 0))
    return w, d


def make_batches(d, w, n_batches: int, seed: int, rank: int = 0, world: int = 1):
    """Distinct (region, sample) batches, built like the reference's host prep (see synth.batch_args).
    With `world` ranks every GLOBAL batch holds world x pairs (region, sample) pairs drawn from one shared
    seed; rank r works on its contiguous block of it (genvarloader_b200._dist.shard_bounds) -- weak scaling."""
    from genvarloader_b200 import synth
    from genvarloader_b200._dist import shard_bounds

    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        lo, hi = shard_bounds(w["pairs"] * world, rank, world)
        r_idx = rng.integers(0, d.n_regions, w["pairs"] * world)[lo:hi]
        s_idx = rng.integers(0, d.n_samples, w["pairs"] * world)[lo:hi]
        regions, goi, to_rc, ds_idx = synth.batch_args(d, r_idx, s_idx)
        shifts = np.zeros(goi.shape, np.int32)
        nvar = int((d.geno_offsets[1, goi.ravel()] - d.geno_offsets[0, goi.ravel()]).sum())
        out.append(dict(regions=regions, goi=goi, to_rc=to_rc, shifts=shifts, nvar=nvar, ds_idx=ds_idx))
    return out


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes(w, nvar: int, mode: str = "onehot") -> float:
    rows = w["pairs"] * 2
    per_bp = {"onehot": ALG_BYTES_PER_BP, "u8": 2.0, "annotated": 10.0}[mode]
    return rows * w["window"] * per_bp + nvar * ALG_BYTES_PER_VARIANT + rows * ALG_BYTES_PER_ROW


def track_bytes(w) -> float:
    """4 B per realigned value written (the intervals read add 12 B each, not counted)."""
    return 4.0 * w.get("tracks", 0) * w["pairs"] * 2 * w["window"]


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's Rust/rayon path) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_arm(d, w, batches, budget_s: float, threads: int, min_reps: int = 2):
    """Times reconstruct_haplotypes_fused (+ the separate one-hot pass a seqpro user pays) exactly as the
    reference runs it: one crossing per batch, rayon-style (query, hap) tasks on `threads` threads."""
    from oracle import oracle as O

    O.set_threads(threads)
    timers = [O.FusedTimer(b["regions"], b["shifts"], b["goi"], d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens,
                           d.alt_alleles, d.alt_offsets, d.reference, d.ref_offsets, N_CHAR, w["window"], b["to_rc"],
                           onehot=True) for b in batches[:4]]
    timers[0](parallel=threads > 1)  # warm-up (page faults, thread start)
    t0 = time.perf_counter()
    n, bp = 0, 0
    while True:
        bp += timers[n % len(timers)](parallel=threads > 1)
        n += 1
        el = time.perf_counter() - t0
        if (el > budget_s and n >= min_reps) or n >= 4096:
            break
    return bp / el, n, el


# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    w, d = build_workload(args.workload, args.seed)
    batches = make_batches(d, w, 8, args.seed + 1)
    threads = O.default_threads()
    O.set_threads(threads)
    timers = [O.FusedTimer(b["regions"], b["shifts"], b["goi"], d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens,
                           d.alt_alleles, d.alt_offsets, d.reference, d.ref_offsets, N_CHAR, w["window"], b["to_rc"],
                           onehot=True) for b in batches]
    for i in range(max(args.warmup, 1)):
        timers[i % len(timers)](parallel=threads > 1)
    t0 = time.perf_counter()
    bp = 0
    for i in range(args.steps):
        bp += timers[i % len(timers)](parallel=threads > 1)
    el = time.perf_counter() - t0
    v = bp / el
    line = {
        "impl": "reference", "metric": "haplotype bp/s (one-hot)", "value": v, "unit": "bp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "window_bp": w["window"], "haplotypes_per_batch": w["pairs"] * 2,
                   "variants_per_batch": batches[0]["nvar"], "output": "uint8 one-hot (L,4)", "source": "SVAR1-style sparse CSR",
                   "parallelism": f"{threads} host threads, one task per (query, hap) row"},
        "cpu_baseline": {"value": v, "unit": "bp/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} batches of {w['pairs'] * 2} haplotypes x {w['window']} bp: C restatement of "
                                   "reconstruct_haplotypes_fused + separate one-hot pass, one task per (query, hap)"},
        "e2e": {"value": v, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from genvarloader_b200 import _ffi, _kernels
    from genvarloader_b200._engine import Engine

    w, d = build_workload(args.workload, args.seed)  # every rank holds a full replica
    L, rows = w["window"], w["pairs"] * 2
    n_slots = args.slots
    # ring length: as asked, but never longer than the timed run (a run shorter than the ring would fall back to one
    # host launch per batch, which the host cannot issue as fast as the device finishes them)
    ring = min(args.ring, max(n_slots, (args.steps // max(n_slots, 1)) * max(n_slots, 1)))
    n_batches = max(ring, n_slots)
    batches = make_batches(d, w, n_batches, args.seed + 1, rank, world)  # this rank's shard of every global batch
    eng0 = Engine(dev, d.reference, d.ref_offsets, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.geno_v_idxs,
                  d.geno_offsets)
    step_bytes_out = rows * L * {"onehot": 4, "u8": 1, "annotated": 1}[args.mode]
    track_names = sorted(d.tracks)
    for nm in track_names:
        eng0.add_track(nm, *d.tracks[nm])

    # ---- per-batch device inputs + output ring (ring > L2 so writes cannot stay cache-resident) ----
    streams = [torch.cuda.Stream(dev) for _ in range(max(1, n_slots))]
    slots = []
    for i, b in enumerate(batches):
        eng = eng0 if i == 0 else eng0.fork()
        s = dict(eng=eng, stream=streams[i % len(streams)],
                 regions=torch.from_numpy(b["regions"]).to(dev), shifts=torch.from_numpy(b["shifts"]).to(dev),
                 goi=torch.from_numpy(b["goi"]).to(dev), to_rc=torch.from_numpy(b["to_rc"]).to(dev),
                 out_offsets=torch.empty(rows + 1, dtype=torch.int64, device=dev),
                 out=torch.empty(step_bytes_out, dtype=torch.uint8, device=dev), nvar=b["nvar"], graph=None)
        if args.mode == "annotated":
            s["av"] = torch.empty(rows * L, dtype=torch.int32, device=dev)
            s["ap"] = torch.empty(rows * L, dtype=torch.int32, device=dev)
        if track_names:  # realigned tracks of the same batch: (n_tracks, rows, L) float32
            s["oidx"] = torch.from_numpy(np.tile(b["ds_idx"], (len(track_names), 1))).to(dev)
            s["tlen"] = torch.full((b["regions"].shape[0],), L + 4096, dtype=torch.int32, device=dev)  # window + room for deletions
            s["tout"] = torch.empty(len(track_names) * rows * L, dtype=torch.float32, device=dev)
        slots.append(s)

    MODE = args.mode

    def exec_(s_, out=None):
        o = s_["out"] if out is None else out
        if MODE == "annotated":
            s_["eng"].execute("annotated", out=o, annot_v=s_["av"], annot_pos=s_["ap"])
        else:
            s_["eng"].execute("onehot" if MODE == "onehot" else "haplotypes", out=o)

    def step(s):
        s["eng"].plan(s["regions"], s["shifts"], s["goi"], L, s["nvar"], to_rc=s["to_rc"], out_offsets=s["out_offsets"])
        exec_(s)
        if track_names:
            s["eng"].realign_tracks(track_names, s["regions"], s["shifts"], s["goi"], s["oidx"], s["tlen"], s["out_offsets"],
                                    rows * L, [0, 4][: len(track_names)], [0.0, 1.0][: len(track_names)], 7, s["nvar"],
                                    to_rc=s["to_rc"], out=s["tout"])

    # warm every slot once outside any capture (workspace growth may allocate)
    for s in slots:
        with torch.cuda.stream(s["stream"]):
            step(s)
    torch.cuda.synchronize()
    for s in slots:
        s["eng"].check()

    use_graph = not args.no_graph
    if use_graph:
        for s in slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s["stream"]):
                step(s)
            s["graph"] = g

    # One graph for the WHOLE ring: every slot stream is a parallel branch that runs its batches back to back
    # (memset + plan + execute each).  One host launch then feeds len(slots) steps, so the host's graph-launch
    # rate (~11 us per launch here) does not bound a step that takes less than that on the device.
    ring_graph = None
    if use_graph and len(slots) > 1:
        cap = torch.cuda.Stream(dev)
        ring_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ring_graph, stream=cap):
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

    def enqueue(i):
        s = slots[i % len(slots)]
        if use_graph:
            with torch.cuda.stream(s["stream"]):
                s["graph"].replay()
        else:
            with torch.cuda.stream(s["stream"]):
                step(s)

    main = torch.cuda.current_stream()

    def timed(n_steps: int, sync_ranks: bool = True) -> float:
        """n_steps steps, round-robin over the slot streams (several batches in flight); device time
        between one start event every stream waits on and one end event that waits on every stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if dist is not None and sync_ranks:
            dist.barrier()
            torch.cuda.synchronize()
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        n_ring = n_steps // len(slots) if ring_graph is not None else 0
        for _ in range(n_ring):
            ring_graph.replay()  # on `main`: len(slots) steps per launch
        if n_ring:
            ev = torch.cuda.Event()
            ev.record(main)
            for st in streams:
                st.wait_event(ev)
        for i in range(n_ring * len(slots), n_steps):  # remainder: one graph (or eager step) per batch
            enqueue(i)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        if dist is not None and sync_ranks:
            dist.barrier()
        return e0.elapsed_time(e1)  # ms

    timed(max(args.warmup, 3))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _ffi.launch_count(reset=True)
    with torch.cuda.stream(slots[0]["stream"]):
        step(slots[0])  # eager, to count the kernels one step launches (graph replays bypass the counter)
    torch.cuda.synchronize()
    launches_per_step = _ffi.launch_count()
    ms = timed(args.steps)
    launches = launches_per_step * args.steps
    # keep the GPU busy a little longer so the 100 ms clock sampler sees the loaded state
    t_end = time.perf_counter() + (0.6 if rank == 0 else 0.0)
    while time.perf_counter() < t_end:
        timed(min(args.steps, 200), sync_ranks=False)  # rank-local: no collective here
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bp_per_step = rows * L
    value = world * args.steps * bp_per_step / (ms * 1e-3)

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # ---- roofline of the dominant kernel (execute) ----
        # (a) steady state: the execute launches of all ring slots captured in ONE graph on ONE stream and
        #     replayed back to back (no host gaps, every launch writes a different 32-64 MiB buffer, the ring
        #     exceeds L2); launch duration = event time / launches.
        # (b) isolated: one launch after a 512 MiB L2 flush, events around the single launch.  A bare pair of
        #     events around an EMPTY stream already reads ~12 us on this box, so (b) overstates short kernels;
        #     it is reported for completeness only.
        for s_ in slots:
            with torch.cuda.stream(s_["stream"]):
                step(s_)  # every slot's context holds a valid plan for its batch
        torch.cuda.synchronize()
        g_exec = torch.cuda.CUDAGraph()
        cap_stream = torch.cuda.Stream(dev)
        with torch.cuda.graph(g_exec, stream=cap_stream):
            for s_ in slots:
                exec_(s_)
        reps = max(4, 256 // len(slots))
        with torch.cuda.stream(cap_stream):
            for _ in range(3):
                g_exec.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(cap_stream)
            for _ in range(reps):
                g_exec.replay()
            b.record(cap_stream)
        torch.cuda.synchronize()
        exec_ms_one_stream = a.elapsed_time(b) / (reps * len(slots))
        # (a') the same launches the way the product issues them: every slot's execute on its own slot stream
        #      (n_slots launches in flight), one graph, replayed back to back.  Launch duration = event time /
        #      launches: the time the GPU effectively spends per execute launch in steady state.
        g_exec_par = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_exec_par, stream=cap_stream):
            ev0 = torch.cuda.Event()
            ev0.record(cap_stream)
            for st in streams:
                st.wait_event(ev0)
            for s_ in slots:
                with torch.cuda.stream(s_["stream"]):
                    exec_(s_)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                cap_stream.wait_event(ev)
        with torch.cuda.stream(cap_stream):
            for _ in range(3):
                g_exec_par.replay()
            a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a1.record(cap_stream)
            for _ in range(reps):
                g_exec_par.replay()
            b1.record(cap_stream)
        torch.cuda.synchronize()
        exec_ms = a1.elapsed_time(b1) / (reps * len(slots))
        g_plan = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_plan, stream=cap_stream):
            for s_ in slots:
                s_["eng"].plan(s_["regions"], s_["shifts"], s_["goi"], L, s_["nvar"], to_rc=s_["to_rc"], out_offsets=s_["out_offsets"])
        with torch.cuda.stream(cap_stream):
            g_plan.replay()
            a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a2.record(cap_stream)
            for _ in range(reps):
                g_plan.replay()
            b2.record(cap_stream)
        torch.cuda.synchronize()
        plan_ms = a2.elapsed_time(b2) / (reps * len(slots))
        # (c) the plans of the whole ring issued like the product issues them (slot streams, --slots in flight): what the plan
        #     costs the GPU per batch when its latency is overlapped
        g_plan_par = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_plan_par, stream=cap_stream):
            ev0 = torch.cuda.Event()
            ev0.record(cap_stream)
            for st in streams:
                st.wait_event(ev0)
            for s_ in slots:
                with torch.cuda.stream(s_["stream"]):
                    s_["eng"].plan(s_["regions"], s_["shifts"], s_["goi"], L, s_["nvar"], to_rc=s_["to_rc"], out_offsets=s_["out_offsets"])
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                cap_stream.wait_event(ev)
        with torch.cuda.stream(cap_stream):
            g_plan_par.replay()
            a4, b4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a4.record(cap_stream)
            for _ in range(reps):
                g_plan_par.replay()
            b4.record(cap_stream)
        torch.cuda.synchronize()
        plan_ms_slots = a4.elapsed_time(b4) / (reps * len(slots))
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        s = slots[0]
        durs = []
        for i in range(15):
            flush.zero_()
            a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a3.record(main)
            exec_(s, out=slots[i % len(slots)]["out"])
            b3.record(main)
            torch.cuda.synchronize()
            if i >= 5:
                durs.append(a3.elapsed_time(b3))
        exec_ms_isolated = float(np.mean(durs))
        del flush
        ab = alg_bytes(w, s["nvar"], args.mode)
        achieved = ab / (exec_ms * 1e-3) / 1e9
        step_achieved = (ab + track_bytes(w)) * args.steps / (ms * 1e-3) / 1e9  # per GPU (every rank runs `steps` steps of its shard)
        packed_ran = int(_ffi.lib.gvl_debug_last_exec_kernel(eng0.ctx.handle)) == 1  # which execute kernel the step launched
        kernel_name = "hap_exec_oh_kernel (one-hot over the packed reference)" if packed_ran else f"hap_exec_kernel<{args.mode}> (byte reference)"
        traffic, traffic_note = None, None
        tp = ROOT / "profiles" / "ncu_exec_traffic.json"
        if tp.exists() and packed_ran:  # (the committed capture is of the packed kernel)
            try:
                tj = json.loads(tp.read_text())
                traffic = tj.get(args.workload)
                traffic_note = (tj.get(args.workload + "_detail") or {}).get("note")
            except Exception:
                traffic = None

        # ---- e2e: the reference-shaped host-buffer call (numpy in, pinned numpy out), copies included ----
        _kernels.pin_static(d.geno_offsets, d.geno_v_idxs, d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets,
                            d.reference, d.ref_offsets, ctx=eng0.ctx)
        e2e_bytes_out = rows * L * 4  # the e2e leg always returns the one-hot (the headline metric)
        pinned = _kernels.PinnedBuffer(e2e_bytes_out)
        def e2e_step(b):
            _kernels.reconstruct_haplotypes_fused(b["regions"], b["shifts"], b["goi"], d.geno_offsets, d.geno_v_idxs,
                                                  d.v_starts, d.ilens, d.alt_alleles, d.alt_offsets, d.reference,
                                                  d.ref_offsets, N_CHAR, L, None, None, b["to_rc"], mode="onehot",
                                                  out=pinned.array, ctx=eng0.ctx)
        for i in range(3):
            e2e_step(batches[i % len(batches)])
        n_e2e = max(3, min(args.steps, int(2.0 / max(e2e_bytes_out / 20e9, 1e-4))))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_step(batches[i % len(batches)])
        e2e_s = time.perf_counter() - t0
        h2d = int(sum(batches[0][k].nbytes for k in ("regions", "shifts", "goi", "to_rc")))
        e2e = {"value": n_e2e * bp_per_step / e2e_s, "unit": "bp/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(e2e_bytes_out + (rows + 1) * 8), "steps": n_e2e,
               "path": "genvarloader_b200._kernels.reconstruct_haplotypes_fused(mode='onehot') -> gvl_reconstruct_haplotypes_fused_begin/_finish, "
                       "host numpy in, pinned host numpy out"}

        # ---- CPU baseline (bounded sample of the same workload) ----
        from oracle import oracle as O

        threads = O.default_threads()
        cpu_v, cpu_n, cpu_s = cpu_arm(d, w, batches, budget_s=args.cpu_seconds, threads=threads)
        cpu1_v, cpu1_n, _ = cpu_arm(d, w, batches, budget_s=min(args.cpu_seconds, 4.0), threads=1)

        line = {
            "metric": "haplotype bp/s (%s)" % {"onehot": "one-hot", "u8": "uint8 bytes", "annotated": "annotated"}[args.mode],
            "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "window_bp": L, "haplotypes_per_batch": rows,
                       "variants_per_batch": s["nvar"],
                       "output": {"onehot": "uint8 one-hot (L,4)", "u8": "uint8 haplotype bytes", "annotated": "uint8 bytes + int32 variant index + int32 reference coordinate"}[args.mode] + (f" + {len(track_names)} float32 tracks (n_tracks, rows, L)" if track_names else ""), "source": "SVAR1-style sparse CSR",
                       "batches_in_flight": n_slots, "cuda_graph": use_graph,
                       "l2": f"ring of {len(slots)} distinct batches/outputs = {len(slots) * step_bytes_out >> 20} MiB written per cycle "
                             "(> 126 MB L2); roofline launches are preceded by a 512 MiB L2 flush",
                       "parallelism": f"dp{world} (replicated tables, (region,sample) shards, no collective)"},
            "output_GBps": value * {"onehot": 4, "u8": 1, "annotated": 9}[args.mode] / 1e9, "algorithmic_GBps": step_achieved * world,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                         "alg_bytes_per_launch": ab, "launch_ms": exec_ms, "plan_kernel_ms": plan_ms, "plan_kernel_ms_slot_streams": plan_ms_slots,
                         "launch_ms_one_stream": exec_ms_one_stream,
                         "frac_one_stream": ab / (exec_ms_one_stream * 1e-3) / 1e9 / peak,
                         "launch_ms_isolated_after_l2_flush": exec_ms_isolated,
                         "how": f"execute launches of the whole ring as ONE CUDA graph, issued like the product does (each slot's launch on its "
                                f"own stream, {n_slots} in flight), replayed back to back; event time / launches.  launch_ms_one_stream: same "
                                "launches serialised on a single stream (adds the ~3 us stream-order gap a 33 MB fill kernel also pays)",
                         "frac_of_nominal_8TBps": achieved / 8000.0,
                         "frac_layout_bytes": (ab - (0.5 * rows * L if args.mode == "onehot" else 0.0)) / (exec_ms * 1e-3) / 1e9 / peak,
                         "bytes_note": "algorithmic bytes = SURVEY.md 8d (5 B/bp: 1 reference byte + 4 one-hot bytes); the packed reference "
                                       "moves 0.5 B/bp, frac_layout_bytes counts 4.5 B/bp instead; the peak is a read+write copy figure, so a nearly "
                                       "write-only kernel can land a little above 1.0",
                         "whole_step_frac": step_achieved / peak},
            "cpu_baseline": {"value": cpu_v, "unit": "bp/s", "cores": threads, "kind": "port",
                             "sample": f"{cpu_n} batches ({cpu_s:.1f} s) of the same workload; C restatement of the reference's "
                                       "reconstruct_haplotypes_fused + separate one-hot pass (oracle/gvl_oracle.c)",
                             "single_thread_value": cpu1_v},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1280)
    ap.add_argument("--warmup", type=int, default=128)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--mode", default="onehot", choices=["onehot", "u8", "annotated"],
                    help="output of the execute kernel (the headline metric is one-hot)")
    ap.add_argument("--slots", type=int, default=12, help="batches in flight (streams)")
    ap.add_argument("--ring", type=int, default=128, help="distinct batches / output buffers cycled through (one graph launch)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
