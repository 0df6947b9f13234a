#!/usr/bin/env python
"""bench.py -- the measurement contract for ggml-b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras] [--layers L]

Headline workload (every N): BASELINE.json configs[4] -- the MLP layer stack of a Llama-3-70B-shape Q4_K_M model at
decode batch 1: 80 blocks of  down( silu(gate(x)) * up(x) )  with gate/up [28672, 8192] Q4_K and down [8192, 28672]
Q6_K or Q4_K (llama.cpp's Q4_K_M rule: half of the ffn_down tensors are Q6_K), every projection the repo's MMVQ kernel.
N = 1 runs the stack unsharded; N > 1 shards every projection along its output rows over the N ranks (column-parallel)
with the exchange of the output slices fused into the kernels' epilogues (peer stores over NVLink + one barrier per
projection): the total work is fixed, so this is STRONG scaling.  One "step" = one token through the 80 blocks
(36.6 GB of quantized weights, every byte streamed from HBM exactly once: the stack is 290x the L2).
`value` = weight bytes of the whole stack (all ranks) per second of device time, max over ranks.
`e2e`   = the same stack through the reference-named public ops (`ggml.ggml_mul_mat_vec_a8` inside
          ggml.parallel.QuantMLP, eager, allocating) with the token's activations copied from pinned host memory and
          the result copied back inside the timed region, every step.
`roofline` = the kernel and shape the north star sets its target on -- Q4_K MMVQ 14336x4096, batch 1 (BASELINE
          configs[2]) -- timed live in the same run; `roofline_step_kernel` = the dominant kernel of the step itself.
`extras` carries the other families of the metric (MMVQ all shapes/batches, MMQ TOPS, the dequantize sweep, the
          tensor-parallel prefill/decode table).

`--impl reference` times the CPU side of the same path on the host cores: the reference has no CPU MMVQ at all
(ggml-cpu/custom_ops.cpp is dequantize-only), so it is the oracle port of its CUDA algorithm, all host threads,
on a bounded sample of the same layer (stated in the line).
"""
from __future__ import annotations

import argparse
import json
import math
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

HIDDEN, INTER, LAYERS = 8192, 28672, 80           # Llama-3-70B MLP
Q4_K, Q6_K = 12, 14
METRIC = "MMVQ weight GB/s (Llama-3-70B-shape Q4_K_M MLP layer stack, decode batch 1)"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0                         # /opt/skills/guides/B200_PROFILING.md
CPU_SAMPLE_ROWS = 1024                            # rows of each projection per CPU step (bounded sample)


def down_type(layer: int, n_layers: int) -> int:
    """llama.cpp's Q4_K_M recipe for ffn_down (use_more_bits): Q6_K in the first and last eighth and every third layer
    between, Q4_K elsewhere."""
    more = layer < n_layers // 8 or layer >= 7 * n_layers // 8 or (layer - n_layers // 8) % 3 == 2
    return Q6_K if more else Q4_K


def block_bytes_of(t: int) -> int:
    return {Q4_K: 144, Q6_K: 210}[t]


def layer_weight_bytes(layer: int, n_layers: int) -> int:
    gu = 2 * INTER * (HIDDEN // 256) * block_bytes_of(Q4_K)
    dn = HIDDEN * (INTER // 256) * block_bytes_of(down_type(layer, n_layers))
    return gu + dn


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
# CPU side: oracle port (bench's cpu_baseline leg and the --impl reference arm)
# ------------------------------------------------------------------------------------------------

class CpuSample:
    """First CPU_SAMPLE_ROWS rows of gate, up (Q4_K) and down (Q6_K) of one block, q8_1 + MMVQ through the oracle port
    (oracle/ggml_oracle.c restates ggml_kernel.cu:13-66 + mmvq.cuh + vecdotq.cuh; the reference has no CPU MMVQ)."""

    def __init__(self):
        import oracle
        self.oracle = oracle
        self.threads = oracle.set_num_threads(os.cpu_count() or 1)       # explicit: torchrun exports OMP_NUM_THREADS=1
        r = CPU_SAMPLE_ROWS
        self.mats = [(Q4_K, r, HIDDEN, oracle.random_weights(Q4_K, r, HIDDEN, seed=1)),
                     (Q4_K, r, HIDDEN, oracle.random_weights(Q4_K, r, HIDDEN, seed=2)),
                     (Q6_K, r, INTER, oracle.random_weights(Q6_K, r, INTER, seed=3))]
        rng = np.random.default_rng(0)
        self.x = {HIDDEN: rng.standard_normal((1, HIDDEN)).astype(np.float32),
                  INTER: rng.standard_normal((1, INTER)).astype(np.float32)}
        self.bytes = sum(w.size for _, _, _, w in self.mats)
        self.what = (f"first {r} rows of gate, up (Q4_K, K={HIDDEN}) and down (Q6_K, K={INTER}) of one block = "
                     f"{self.bytes / 1e6:.1f} MB of weights per step, q8_1 quantization + exact-integer MMVQ, oracle port "
                     f"(the reference has no CPU MMVQ), OpenMP row split")

    def step(self) -> float:
        t0 = time.perf_counter()
        for t, n, k, w in self.mats:
            self.oracle.mmvq(w, t, self.x[k], n)
        return time.perf_counter() - t0


def cpu_baseline(budget_s: float = 12.0) -> dict:
    s = CpuSample()
    s.step()
    reps, tot = 0, 0.0
    while tot < budget_s and reps < 200:
        tot += s.step()
        reps += 1
    return {"value": s.bytes * reps / tot / 1e9, "unit": UNIT, "cores": s.threads, "kind": "port",
            "sample": f"{reps} steps of: {s.what}"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    s = CpuSample()
    for _ in range(max(args.warmup, 1)):
        s.step()
    tot = 0.0
    for _ in range(args.steps):
        tot += s.step()
    v = s.bytes * args.steps / tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int8 dot / f32 accumulate", "data": "synthetic",
        "config": {"workload": "Llama-3-70B-shape Q4_K_M MLP layer stack, decode batch 1 (BASELINE configs[4]) -- bounded sample",
                   "sample": s.what},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": s.threads, "kind": "port", "sample": f"{args.steps} steps of: {s.what}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------

def time_graph(torch, fn, iters: int, warm: int = 3) -> float:
    """Capture `fn` (a sequence of launches) into a CUDA graph, replay it, return ms per replay (median)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()                                  # first call outside capture: one-time function attributes
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        for _ in range(warm):
            g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            g.replay()
            b.record(s)
            b.synchronize()
            ts.append(a.elapsed_time(b))
    torch.cuda.current_stream().wait_stream(s)
    return statistics.median(ts)


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import ggml
    from ggml import synth
    from ggml.parallel import ColumnParallelQuantLinear, FusedQuantMLP, QuantMLP

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    if INTER % world or HIDDEN % world:
        raise SystemExit(f"the 70B MLP does not divide by --gpus {world}")
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    n_layers = args.layers
    ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)       # inference weights are immutable (include/ggml_b200.h)

    # ---- synthetic weights: this rank's row slices of every block (distinct buffers per layer: nothing is re-read)
    i_loc, h_loc = INTER // world, HIDDEN // world
    base_gu = torch.from_numpy(synth.random_blocks(Q4_K, 2 * i_loc, HIDDEN, seed=1 + 100 * rank)).to(dev)
    base_dn = {t: torch.from_numpy(synth.random_blocks(t, h_loc, INTER, seed=t + 100 * rank)).to(dev) for t in (Q4_K, Q6_K)}
    fused, plain = [], []
    for l in range(n_layers):
        w_gu = base_gu if l == 0 else base_gu.clone()
        td = down_type(l, n_layers)
        w_dn = base_dn[td].clone()
        gate = ColumnParallelQuantLinear(w_gu[:i_loc], Q4_K, INTER, rank, world)
        up = ColumnParallelQuantLinear(w_gu[i_loc:], Q4_K, INTER, rank, world)
        down = ColumnParallelQuantLinear(w_dn, td, HIDDEN, rank, world)
        fused.append(FusedQuantMLP(gate, up, down, W_gu=w_gu, share_buffers_with=fused[0] if fused else None))
        plain.append(QuantMLP(gate, up, down))
    stack_bytes = sum(layer_weight_bytes(l, n_layers) for l in range(n_layers))          # all ranks together
    x_host = torch.randn(1, HIDDEN, generator=torch.Generator().manual_seed(0)).to(torch.float16).pin_memory()
    x = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- one step = one token through the stack; captured once, replayed per step (decode runs from CUDA graphs: the
    # C ABI never allocates or synchronises).  Every block reads the same normalised activations (a real model has an
    # RMSNorm in between); the stream order still makes block l wait for block l-1.
    def one_step():
        y = None
        for m in fused:
            y = m(x)
        return y

    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    how = "cuda-graph replay, one graph = one step"
    with torch.cuda.stream(stream):
        one_step()                                # outside capture: symmetric buffers, function attributes
        barrier()
        graph = None
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                one_step()
        except Exception as e:                    # capture of the exchange not possible here: eager launches
            graph = None
            how = f"eager launches ({type(e).__name__} during capture)"
            torch.cuda.synchronize()

        def run_step():
            if graph is not None:
                graph.replay()
            else:
                one_step()

        for _ in range(max(args.warmup, 3)):
            run_step()
        barrier()

        clocks = Clocks(local_rank)
        if rank == 0:
            clocks.start()
            time.sleep(0.3)

        # ---- timed region: exactly K steps, CUDA events on the launching stream, barrier + synchronize on both sides
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = clocks.mark()
        ev0.record(stream)
        for _ in range(args.steps):
            run_step()
        ev1.record(stream)
        barrier()
        t_wall1 = clocks.mark()
        ms = ev0.elapsed_time(ev1)
    torch.cuda.current_stream().wait_stream(stream)
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = stack_bytes * args.steps / (ms * 1e-3) / 1e9
    launches = 2 * n_layers * (args.steps + max(args.warmup, 3) + 2)

    # ---- the step's dominant kernel: the stacked gate|up Q4_K MMVQ of this rank's shard, 80 launches back to back
    gu_rows = 2 * i_loc
    y_gu = torch.empty(1, gu_rows, dtype=torch.float16, device=dev)

    def gu_only():
        for m in fused:
            ggml.mul_mat_vec_a8_out(m.W_gu, x, Q4_K, gu_rows, y_gu)

    gu_ms = time_graph(torch, gu_only, iters=5) / n_layers
    launches += n_layers * 9
    gu_bytes = base_gu.numel()
    clk = clocks.stop(t_wall0, clocks.mark()) if rank == 0 else None

    # ---- e2e: the reference-named public ops (ggml.ggml_mul_mat_vec_a8 inside QuantMLP; every call allocates its output and
    # workspace), activations from pinned host memory and the result back to it inside the timed region, every step.
    # The step -- H2D copy, 80 x (gate, up, SwiGLU, down [+ NCCL all-gathers at N > 1]), D2H copy -- is captured ONCE with
    # torch.cuda.graph and replayed, the way decode serving drives these ops (the copies are graph nodes: every replay moves the
    # bytes); the eager number (Python + dispatcher bound: ~320 op calls per step) is reported beside it.  At N > 1 the NCCL
    # all-gathers are graph nodes too; all ranks take the graph or none does.
    y_host = torch.empty(1, HIDDEN, dtype=torch.float16).pin_memory()

    def e2e_eager_step():
        xd = x_host.to(dev, non_blocking=True)
        y = None
        for m in plain:
            y = m(xd)
        y_host.copy_(y, non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn, n):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        return time.perf_counter() - t0

    e2e_steps = max(2, min(args.steps, 5))
    e2e_eager_s = timed(e2e_eager_step, e2e_steps)
    y_eager = y_host.clone()
    launches += 3 * n_layers * (e2e_steps + 1)
    e2e_s, e2e_how = e2e_eager_s, "eager torch ops"

    def all_agree(ok: bool) -> bool:            # every rank takes the graph or none does (a rank that failed must not leave the others in a collective)
        if world == 1:
            return ok
        f = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        return bool(f.item())

    g2, cap_err = None, ""
    s2 = torch.cuda.Stream()
    s2.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(s2):
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=s2):
                xd = x_host.to(dev, non_blocking=True)
                y = None
                for m in plain:
                    y = m(xd)
                y_host.copy_(y, non_blocking=True)
    except Exception as e:                        # capture not possible here: the eager number stands
        g2, cap_err = None, type(e).__name__
        torch.cuda.synchronize()
    if all_agree(g2 is not None):
        def e2e_graph_step():
            with torch.cuda.stream(s2):
                g2.replay()
            torch.cuda.synchronize()

        y_host.zero_()
        e2e_graph_step()
        if all_agree(torch.equal(y_host, y_eager)):
            e2e_steps = max(e2e_steps, min(args.steps, 20))
            e2e_s = timed(e2e_graph_step, e2e_steps)
            e2e_how = "one torch.cuda.graph capture of the whole step (copies" + (", NCCL all-gathers" if world > 1 else "") + " included), replayed per step"
            launches += 3 * n_layers * (e2e_steps + 2)
        else:
            e2e_how = "eager torch ops (the graph replay differed from the eager result)"
    elif cap_err or world > 1:
        e2e_how = f"eager torch ops (graph capture failed on a rank: {cap_err or 'peer'})"
    torch.cuda.current_stream().wait_stream(s2)
    if world > 1:
        tm = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    e2e_val = stack_bytes * e2e_steps / e2e_s / 1e9
    e2e_eager_val = stack_bytes * max(2, min(args.steps, 5)) / e2e_eager_s / 1e9

    # ---- the north-star kernel and shape (BASELINE configs[2]): Q4_K MMVQ 14336x4096 and 4096x14336, batch 1 and 8, rank 0
    config3 = None
    if rank == 0:
        config3 = {"unit": "GB/s (weight bytes / time)", "peak": hbm_peak, "cases": [],
                   "method": "CUDA-graph replay of 4 walks over 8 distinct weight copies (264 MB > 2x L2), median of 10 replays, fp16 x/y"}
        for (N, K) in ((14336, 4096), (4096, 14336)):
            base = torch.from_numpy(synth.random_blocks(Q4_K, N, K, seed=N)).to(dev)
            copies = [base] + [base.clone() for _ in range(7)]
            for B in (1, 8):
                xb = torch.randn(B, K, dtype=torch.float16, device=dev)
                yb = torch.empty(B, N, dtype=torch.float16, device=dev)
                ws = torch.empty(max(ggml.mmvq_workspace_bytes(Q4_K, B, K, N), 16), dtype=torch.uint8, device=dev)

                def walk():
                    for _ in range(4):
                        for w in copies:
                            ggml.mul_mat_vec_a8_out(w, xb, Q4_K, N, yb, ws)

                us = time_graph(torch, walk, iters=10) / 32 * 1e3
                launches += 32 * 14
                gbs = base.numel() / us / 1e3
                config3["cases"].append({"type": "Q4_K", "N": N, "K": K, "B": B, "us_per_launch": us, "weight_GBps": gbs,
                                         "frac_of_hbm_peak": gbs / hbm_peak})
            del copies, base

    extras = {}
    if not args.no_extras:
        try:
            import bench_extras
            extras, n_extra = bench_extras.run(ggml, synth, dev, rank, world, peaks, peak_kind)
            launches += n_extra
        except Exception as e:  # the headline must still print; the failure is visible in the line
            extras = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            _leave_process_group(dist, g2)
        return

    c3 = config3["cases"][0]                                   # Q4_K 14336x4096, batch 1
    traffic = traffic_table().get("mmvq_q4_k_14336x4096_b1")
    gu_gbs = gu_bytes / (gu_ms * 1e-3) / 1e9
    peak_src = f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int8 dot / f32 accumulate", "data": "synthetic",
        "config": {"workload": f"Llama-3-70B-shape Q4_K_M MLP layer stack ({n_layers} blocks: gate/up [{INTER}, {HIDDEN}] Q4_K, down "
                               f"[{HIDDEN}, {INTER}] Q6_K/Q4_K), decode batch 1, fp16 activations (BASELINE configs[4])",
                   "l2": f"inputs larger than L2: {stack_bytes / world / 1e9:.1f} GB of weights per rank and step, each byte read once (L2 is 126 MB)",
                   "sharding": ("unsharded" if world == 1 else
                                f"column-parallel over {world} ranks (N/tp rows each), output slices exchanged by peer stores from the "
                                f"MMVQ epilogue + one barrier per projection (2 launches + 2 barriers per block)"),
                   "launch": how, "options": "GGB_OPT_STATIC_WEIGHTS=1", "weight_bytes_per_step": stack_bytes},
        "roofline": {"kernel": "mmvq_rg_kernel<Q4_K, half> (ggb_mmvq, batch 1)", "shape": "Q4_K 14336x4096 (BASELINE configs[2], the north-star target shape)",
                     "bound": "hbm", "achieved": c3["weight_GBps"], "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": c3["weight_GBps"] / hbm_peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": 14336 * 16 * 144, "avg_launch_ms": c3["us_per_launch"] * 1e-3,
                     "how": "CUDA events around graph replays of 32 dependent launches over 8 weight copies, in this run (not the step's "
                            "own dominant kernel: that one is roofline_step_kernel)"},
        "roofline_step_kernel": {"kernel": "mmvq_rg_kernel<Q4_K, half> (stacked gate|up projection of the step)",
                                 "shape": f"Q4_K {gu_rows}x{HIDDEN} per rank", "bound": "hbm", "achieved": gu_gbs, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": gu_gbs / hbm_peak, "algorithmic_bytes_per_launch": gu_bytes,
                                 "avg_launch_ms": gu_ms, "share_of_step_bytes": 2 * INTER * (HIDDEN // 256) * 144 * n_layers / stack_bytes},
        "config3_mmvq": config3,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 2, "d2h_bytes_per_step": y_host.numel() * 2,
                "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "api": "ggml.parallel.QuantMLP over ggml.ggml_mul_mat_vec_a8 (outputs and workspaces allocated per call"
                       + (", NCCL all-gather per projection" if world > 1 else "") + "); weights resident, activations from / result to pinned host memory",
                "how": e2e_how, "eager_value": e2e_eager_val},
        "gpu_launches": launches,
        "clocks": clk,
        "extras": extras,
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        _leave_process_group(dist, g2)


def _leave_process_group(dist, graph) -> None:
    """End of a multi-GPU run.  Destroying the process group while a captured graph with NCCL nodes is alive did not return (first
    2-GPU run of the graph-captured e2e: the line was printed, the processes never exited), so the graph is released first and a
    watchdog guarantees that the process leaves with status 0 once its work is done and printed."""
    import threading
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(45.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        if graph is not None:
            graph.reset()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


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
    ap.add_argument("--layers", type=int, default=LAYERS, help="blocks in the stack (default: the model's 80)")
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
