#!/usr/bin/env python
"""bench.py -- the measurement contract for ggml-b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]

Headline workload (N=1): BASELINE.json configs[1] -- the K-quant dequantize sweep Q2_K..Q6_K of
4096x14336 synthetic GGUF weights to fp16.  One "step" = one pass of the sweep (5 launches, 166 MB of
blocks in, 587 MB of fp16 out, so every pass streams from HBM -- the working set is 6x the L2).
`value` = algorithmic GB/s (block bytes read + output bytes written) with inputs resident in HBM;
`e2e` = the same sweep through the public op `ggml.ggml_dequantize` with HOST (pinned) buffers, the
host->device copy of the blocks and the device->host copy of the result inside the timed region.
`roofline` is the Q4_K launch (the type BASELINE.json's metric names) against the measured HBM peak.
`extras` carries the other two families of the metric: MMVQ weight GB/s (Llama-3-8B Q4_K shapes,
batch 1..8) and MMQ int8 TOPS (M = 512..8192), each with its own roofline fraction.

`--impl reference` times the CPU side of the same path on the host cores (the oracle port of the
reference's ggml-cpu op; the reference's own op has no K-quant support) and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "ggml-libtorch_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

SWEEP = (10, 11, 12, 13, 14)                      # Q2_K .. Q6_K
ROWS, COLS = 4096, 14336                          # BASELINE configs[1]
METRIC = "dequant GB/s (K-quant sweep Q2_K-Q6_K 4096x14336 -> fp16, algorithmic bytes)"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0                         # /opt/skills/guides/B200_PROFILING.md


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": FALLBACK_HBM_GBS, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def traffic_table():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, exact PID is killed afterwards)
# ------------------------------------------------------------------------------------------------

class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.kill()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                inside = t0 - 0.05 <= ts <= t1 + 0.15
                if inside:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU side: oracle port / reference op (bench's cpu_baseline leg and the --impl reference arm)
# ------------------------------------------------------------------------------------------------

def cpu_sweep_once(weights, rows):
    """One pass of the sweep on the host: oracle port, all OpenMP threads, first `rows` rows -> fp32."""
    import oracle
    nbytes = 0
    t0 = time.perf_counter()
    for t, w in weights.items():
        qk, bs = oracle.block_elems(t), oracle.block_bytes(t)
        sub = w[:rows]
        oracle.dequantize(sub, t, rows, COLS, threads=True)
        nbytes += rows * (COLS // qk) * bs + rows * COLS * 4
    return time.perf_counter() - t0, nbytes


def cpu_baseline(weights, budget_s=12.0):
    import oracle
    rows = ROWS
    dt, nb = cpu_sweep_once(weights, rows)              # warm-up / calibration (page-faults the outputs)
    reps, tot_t, tot_b = 0, 0.0, 0
    while tot_t < budget_s and reps < 20:
        dt, nb = cpu_sweep_once(weights, rows)
        tot_t += dt
        tot_b += nb
        reps += 1
    out = {"value": tot_b / tot_t / 1e9, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
           "sample": f"{reps} passes of the full sweep (5 types x {rows}x{COLS}) -> fp32, oracle port of ggml-cpu "
                     f"(the reference CPU op has no K-quant case), OpenMP row split"}
    # the reference's own CPU op on its own case (configs[0]), single-threaded as shipped
    ref = oracle.ref_custom_ops()
    if ref is not None:
        import torch
        info = {}
        for t in (2, 8):
            w = torch.from_numpy(oracle.random_weights(t, 4096, 4096, seed=0))
            ref.ggml_dequantize(w, t, 4096, 4096)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                ref.ggml_dequantize(w, t, 4096, 4096)
                best = min(best, time.perf_counter() - t0)
            info[oracle.TYPE_NAMES[t]] = {"ms": best * 1e3, "GB/s": (w.numel() + 4096 * 4096 * 4) / best / 1e9}
        out["reference_op_config0"] = {"kind": "reference", "cores": 1, "what": "custom_ops.ggml_dequantize 4096x4096 -> fp32 (oracle/_ref)", **info}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    weights = {t: oracle.random_weights(t, ROWS, COLS, seed=t) for t in SWEEP}
    for _ in range(max(args.warmup, 1)):
        cpu_sweep_once(weights, ROWS)
    tot_t, tot_b = 0.0, 0
    for _ in range(args.steps):
        dt, nb = cpu_sweep_once(weights, ROWS)
        tot_t += dt
        tot_b += nb
    v = tot_b / tot_t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"K-quant dequantize sweep Q2_K-Q6_K, {ROWS}x{COLS} (BASELINE configs[1])",
                   "note": "CPU output is fp32 (the reference CPU op's dtype); bytes = blocks read + fp32 written"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                         "sample": f"{args.steps} passes of the full sweep; oracle port (ggml-cpu has no K-quant case), OpenMP"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------

def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import ggml
    from ggml import synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])

    # ---- synthetic weights (each rank owns its own shard: a 4096-row slice of a world*4096-row layer)
    w_host = {t: torch.from_numpy(synth.random_blocks(t, ROWS, COLS, seed=t + 100 * rank)).pin_memory() for t in SWEEP}
    w_dev = {t: w.to(dev) for t, w in w_host.items()}
    out_dev = {t: torch.empty((ROWS, COLS), dtype=torch.float16, device=dev) for t in SWEEP}
    in_bytes = {t: w_host[t].numel() for t in SWEEP}
    out_bytes = ROWS * COLS * 2
    step_bytes = sum(in_bytes.values()) + out_bytes * len(SWEEP)

    def step():
        for t in SWEEP:
            ggml.dequantize_out(w_dev[t], t, ROWS, COLS, out_dev[t])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)

    # ---- timed region: exactly K steps, CUDA events on the launching (current) stream.  Nothing but the
    # launches sits between the two events: the kernels are programmatic dependent launches, and an event
    # record between two of them would serialise what a real caller's stream overlaps.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = clocks.mark()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = clocks.mark()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = world * step_bytes * args.steps / (ms * 1e-3) / 1e9

    # ---- per-kernel durations (roofline): the same K steps again, one event pair per launch
    per = {t: [] for t in SWEEP}
    for _ in range(args.steps):
        for t in SWEEP:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ggml.dequantize_out(w_dev[t], t, ROWS, COLS, out_dev[t])
            b.record()
            per[t].append((a, b))
    barrier()
    t_wall1 = clocks.mark()
    kern_ms = {t: statistics.mean(a.elapsed_time(b) for a, b in per[t]) for t in SWEEP}
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- e2e: public op, host buffers, H2D + D2H inside the timed region
    out_host = {t: torch.empty((ROWS, COLS), dtype=torch.float16).pin_memory() for t in SWEEP}
    stage = {t: torch.empty_like(w_dev[t]) for t in SWEEP}

    # H2D of the next type's blocks runs on its own stream (its own copy engine) under the dequantize and the
    # D2H of the current one; every byte still crosses PCIe inside the timed region, every step.
    s_in = torch.cuda.Stream()
    staged = {t: torch.cuda.Event() for t in SWEEP}

    def e2e_step():
        main = torch.cuda.current_stream()
        s_in.wait_stream(main)                      # the previous step's kernels no longer read `stage`
        with torch.cuda.stream(s_in):
            for t in SWEEP:
                stage[t].copy_(w_host[t], non_blocking=True)
                staged[t].record(s_in)
        for t in SWEEP:
            main.wait_event(staged[t])
            y = ggml.ggml_dequantize(stage[t], t, ROWS, COLS)
            out_host[t].copy_(y, non_blocking=True)
        torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tm = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    e2e_val = world * step_bytes * e2e_steps / e2e_s / 1e9

    launches = 2 * args.steps * len(SWEEP)
    extras = {}
    if not args.no_extras:
        try:
            import bench_extras
            extras, n_extra = bench_extras.run(ggml, synth, dev, rank, world, peaks, peak_kind)
        except Exception as e:  # the headline must still print; the failure is visible in the line
            extras = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    q4k_bytes = in_bytes[12] + out_bytes
    q4k_gbs = q4k_bytes / (kern_ms[12] * 1e-3) / 1e9
    traffic = traffic_table().get("dequant_q4_k_f16")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"K-quant dequantize sweep Q2_K-Q6_K, {ROWS}x{COLS} -> fp16 (BASELINE configs[1])",
                   "l2": "inputs larger than L2: 166 MB blocks + 587 MB outputs per step (L2 is 126 MB)",
                   "sharding": "each rank dequantizes its own 4096-row shard; no collective on this path",
                   "arith": "fp32 mul/add in the CPU op's order, one RN cast to fp16"},
        "roofline": {"kernel": "dequant_tile_kernel<Q4_K, half>", "bound": "hbm", "achieved": q4k_gbs, "peak": hbm_peak,
                     "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "unit": "GB/s", "frac": q4k_gbs / hbm_peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": q4k_bytes, "avg_launch_ms": kern_ms[12],
                     "how": "event pair around every launch of a second pass of the K steps (serialised launches)"},
        "per_type_GBps": {synth.NAMES[t]: (in_bytes[t] + out_bytes) / (kern_ms[t] * 1e-3) / 1e9 for t in SWEEP},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": sum(in_bytes.values()),
                "d2h_bytes_per_step": out_bytes * len(SWEEP), "steps": e2e_steps,
                "api": "ggml.ggml_dequantize on pinned-host blocks (H2D on a second stream), result copied back to pinned host"},
        "gpu_launches": launches,
        "clocks": clk,
        "extras": extras,
    }
    if world == 1:
        import oracle  # cpu_baseline leg only
        weights = {t: w_host[t].numpy() for t in SWEEP}
        line["cpu_baseline"] = cpu_baseline(weights)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Route fd 1 to stderr for the whole run (NCCL / torchrun banners) and return a writer on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


_OUT = None


def emit(line: dict) -> None:
    """The one JSON line, on the real stdout."""
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
