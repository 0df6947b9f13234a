"""The other two families of BASELINE.json's metric, reported under `extras` of bench.py's JSON line:

  * MMVQ weight GB/s (BASELINE configs[2]: Llama-3-8B Q4_K/Q6_K decode GEMV shapes, batch 1..8) against
    the measured HBM peak.  Each timed pass walks 8 distinct copies of the weight matrix (264-385 MB,
    > 2x the 126 MB L2), replayed from a CUDA graph so the 5-10 us kernels are not launch-bound.
  * MMQ int8 TOPS (BASELINE configs[3]: M = 512..8192 prefill tokens) against the int8 tensor peak.

Called only from bench.py; uses the product package only (no oracle).
"""
from __future__ import annotations

import statistics

import torch

NOMINAL_INT8_TOPS = 4500.0      # dense int8 tcgen05 peak of a B200 (2x the bf16 nominal of 2250)
# measured on this pool (scripts/probes/i8_probe.cu, profiles/r02_i8_probe.json): a bare tcgen05.mma.kind::i8 issue loop,
# M = 128 x N = 256 x K = 32 per instruction, 148 SMs, operands resident in shared memory
MEASURED_INT8_TOPS = 4416.6
MEASURED_F16_ISSUE_TFLOPS = 2100.3


def _time_graph(fn, iters: int, warm: int = 3) -> float:
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


def mmvq_section(ggml, synth, dev, peaks, peak_kind):
    hbm = float(peaks["hbm_gbs"])
    out = {"unit": "GB/s (weight bytes / time)", "peak": hbm, "peak_source": peak_kind,
           "method": "CUDA-graph replay of 4 walks over 8 distinct weight copies (> 2x L2), median of 10 replays, fp16 x/y",
           "cases": []}
    ncopies = 8
    walks = 4                                # 32 launches per graph replay: the replay's own launch cost is amortised
    launches = 0
    for tname, t in (("Q4_K", 12), ("Q6_K", 14)):
        for (N, K) in ((14336, 4096), (4096, 14336)):
            if t == 14 and N == 14336:
                continue                     # Q4_K_M stores ffn_down (4096 x 14336) in Q6_K; gate/up stay Q4_K
            base = torch.from_numpy(synth.random_blocks(t, N, K, seed=N + t)).to(dev)
            copies = [base] + [base.clone() for _ in range(ncopies - 1)]
            wbytes = base.numel()
            for B in (1, 2, 4, 8):
                x = torch.randn(B, K, dtype=torch.float16, device=dev)
                y = torch.empty(B, N, dtype=torch.float16, device=dev)

                ws = torch.empty(max(ggml.mmvq_workspace_bytes(t, B, K, N), 16), dtype=torch.uint8, device=dev)

                def run():
                    for _ in range(walks):
                        for w in copies:
                            ggml.mul_mat_vec_a8_out(w, x, t, N, y, ws)

                ms = _time_graph(run, iters=10) / walks
                launches += ncopies * walks * 14
                gbs = wbytes * ncopies / (ms * 1e-3) / 1e9
                out["cases"].append({"type": tname, "N": N, "K": K, "B": B, "us_per_launch": ms * 1e3 / ncopies,
                                     "weight_GBps": gbs, "frac_of_hbm_peak": gbs / hbm})
            del copies, base
    return out, launches


def mmq_section(ggml, synth, dev, peaks, peak_kind):
    out = {"unit": "TOPS (2*M*N*K / time)", "peak": NOMINAL_INT8_TOPS,
           "peak_source": "nominal dense int8; measured bare kind::i8 issue loop on this pool: 4416.6 TOPS (profiles/r02_i8_probe.json)",
           "int8_measured_issue_loop_tops": MEASURED_INT8_TOPS, "f16_measured_issue_loop_tflops": MEASURED_F16_ISSUE_TFLOPS,
           "exact_q8_1_int8_ceiling_tops": 417.6,
           "exact_q8_1_int8_ceiling_note": "kind::i8 128x256x32 tiles with the per-32 TMEM drain + fp32 rescale that exact q8_1 semantics "
                                           "need (8 drain warps, double-buffered accumulators), measured: 264-418 TOPS -- below this kernel",
           "bf16_measured_tflops": peaks.get("bf16_tflops"),
           "method": "CUDA events around 10 launches after 3 warm-ups, fp16 x/y, workspace preallocated", "cases": []}
    launches = 0
    for tname, t in (("Q4_K", 12), ("Q6_K", 14)):
        N, K = 14336, 4096
        w = torch.from_numpy(synth.random_blocks(t, N, K, seed=1)).to(dev)
        for M in (512, 1024, 2048, 4096, 8192):
            x = torch.randn(M, K, dtype=torch.float16, device=dev)
            y = torch.empty(M, N, dtype=torch.float16, device=dev)
            ws = torch.empty(max(ggml.mmq_workspace_bytes(t, M, K, N), 256), dtype=torch.uint8, device=dev)
            for _ in range(3):
                ggml.mul_mat_a8_out(w, x, t, N, y, ws)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            a.record()
            for _ in range(iters):
                ggml.mul_mat_a8_out(w, x, t, N, y, ws)
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b) / iters
            launches += iters + 3
            tops = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            out["cases"].append({"type": tname, "M": M, "N": N, "K": K, "ms": ms, "TOPS": tops,
                                 "frac_of_int8_peak": tops / NOMINAL_INT8_TOPS,
                                 "frac_of_measured_bf16_peak": tops / float(peaks.get("bf16_tflops", 1595.8))})
    return out, launches


def tp_mlp_section(ggml, synth, dev, rank, world, peaks, peak_kind):
    """BASELINE configs[4], one block, all exchange variants and batch sizes (the headline of bench.py is the 80-block stack
    at batch 1 on the fused path): Llama-3-70B MLP (hidden 8192, intermediate 28672), Q4_K gate/up and Q6_K down,
    column-parallel over `world` ranks.  Device time (CUDA events), max over ranks.  Weight GB/s counts every rank's shard
    bytes (the whole layer) per step.  NOTE: one block is 457 MB / tp per rank -- below the L2 size for tp >= 4, so the
    replayed block is partly L2-resident there; the headline stack is not."""
    import torch.distributed as dist
    from ggml.parallel import ColumnParallelQuantLinear, QuantMLP
    H, I = 8192, 28672
    if I % world or H % world:
        return {"skipped": f"70B MLP does not divide by tp={world}"}, 0

    def shard(t, n_total, k, seed, fused):
        n_local = n_total // world
        w = torch.from_numpy(synth.random_blocks(t, n_local, k, seed=seed + 1000 * rank)).to(dev)
        return ColumnParallelQuantLinear(w, t, n_total, rank, world, fused_gather=fused)

    from ggml.parallel import FusedQuantMLP
    variants = [("nccl all-gather" if world > 1 else "none (tp=1)", False)]
    if world > 1:
        variants.append(("fused peer stores + barrier (ggb_mmvq_gather)", True))
    mlps = {}
    for name, fused in variants:
        mlps[name] = QuantMLP(shard(12, I, H, 1, fused), shard(12, I, H, 2, fused), shard(14, H, I, 3, fused))
    # the whole block as two launches: stacked gate|up with split output columns, SwiGLU inside the down projection
    block_name = "fused block: 2 launches (stacked gate|up, SwiGLU in the down projection)" + (" + 2 barriers" if world > 1 else "")
    base = mlps[variants[0][0]]
    mlps[block_name] = FusedQuantMLP(base.gate, base.up, base.down)
    variants.append((block_name, True))
    mlp = mlps[variants[0][0]]
    layer_bytes = sum(l.W.numel() for l in (mlp.gate, mlp.up, mlp.down)) * world
    out = {"unit": "us per MLP block (device time, max over ranks)", "tp": world,
           "layer": "Llama-3-70B MLP: gate/up [28672, 8192] Q4_K, down [8192, 28672] Q6_K, N-sharded, one exchange per projection",
           "weight_bytes_total": layer_bytes, "cases": []}
    launches = 0
    for name, fused in variants:
      mlp = mlps[name]
      for M in ((1, 8) if fused else (1, 8, 512, 4096)):
        x = torch.randn(M, H, dtype=torch.float16, device=dev)
        reps = 8 if M <= 8 else 2
        how = "cuda-graph replay"
        try:
            # one graph = `reps` MLP blocks (3 matmuls + 3 exchanges each): host launch cost out of the picture
            def run():
                for _ in range(reps):
                    mlp(x)
            if world > 1:
                dist.barrier()
            ms = _time_graph(run, iters=5) / reps
        except Exception as e:                       # capture of the exchange not possible here: eager timing
            how = f"eager ({type(e).__name__})"
            torch.cuda.synchronize()
            for _ in range(3):
                mlp(x)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                mlp(x)
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b) / reps
        iters = 5 * reps
        if world > 1:
            tm = torch.tensor([ms], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        launches += (iters + 3) * 3
        case = {"M": M, "exchange": name, "us": ms * 1e3, "timing": how}
        if M <= 8:
            case["weight_GBps_aggregate"] = layer_bytes / (ms * 1e-3) / 1e9
            case["frac_of_hbm_peak_per_gpu"] = layer_bytes / world / (ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"])
        else:
            case["TOPS_aggregate"] = 2.0 * M * 3 * H * I / (ms * 1e-3) / 1e12
        out["cases"].append(case)
    return out, launches


def config0_section(ggml, synth, dev, peaks, peak_kind):
    """BASELINE configs[0] on the GPU: Q4_0 and Q8_0 dequantize of a 4096x4096 weight to fp32 (the reference CPU op's
    case and dtype; its own single-thread time is in cpu_baseline.reference_op_config0).  8 distinct weight/output
    pairs per pass (> L2), device time per launch."""
    out = {"unit": "GB/s (block bytes + fp32 output bytes)", "cases": []}
    launches = 0
    for tname, t in (("Q4_0", 2), ("Q8_0", 8)):
        ws = [torch.from_numpy(synth.random_blocks(t, 4096, 4096, seed=i)).to(dev) for i in range(8)]
        outs = [torch.empty((4096, 4096), dtype=torch.float32, device=dev) for _ in range(8)]
        for w, o in zip(ws, outs):
            ggml.dequantize_out(w, t, 4096, 4096, o)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            for w, o in zip(ws, outs):
                ggml.dequantize_out(w, t, 4096, 4096, o)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / (reps * 8)
        launches += 8 * (reps + 1)
        nbytes = ws[0].numel() + 4096 * 4096 * 4
        out["cases"].append({"type": tname, "us_per_launch": ms * 1e3, "GBps": nbytes / (ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"])})
    return out, launches


def dequant_sweep_section(ggml, synth, dev, peaks, peak_kind):
    """BASELINE configs[1] (round 1's headline): the K-quant dequantize sweep Q2_K..Q6_K of 4096x14336 -> fp16, back-to-back
    launches (5 per pass: 166 MB of blocks in, 587 MB out -- 6x the L2), algorithmic GB/s = blocks read + fp16 written."""
    hbm = float(peaks["hbm_gbs"])
    rows, cols = 4096, 14336
    sweep = (10, 11, 12, 13, 14)
    w = {t: torch.from_numpy(synth.random_blocks(t, rows, cols, seed=t)).to(dev) for t in sweep}
    o = {t: torch.empty((rows, cols), dtype=torch.float16, device=dev) for t in sweep}
    step_bytes = sum(x.numel() for x in w.values()) + rows * cols * 2 * len(sweep)

    def step():
        for t in sweep:
            ggml.dequantize_out(w[t], t, rows, cols, o[t])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps):
        step()
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / reps
    per = {}
    for t in sweep:
        a.record()
        for _ in range(10):
            ggml.dequantize_out(w[t], t, rows, cols, o[t])
        b.record()
        b.synchronize()
        per[synth.NAMES[t]] = (w[t].numel() + rows * cols * 2) / (a.elapsed_time(b) / 10 * 1e-3) / 1e9
    gbs = step_bytes / (ms * 1e-3) / 1e9
    return {"unit": "GB/s (block bytes + fp16 output bytes)", "sweep_GBps": gbs, "frac_of_hbm_peak": gbs / hbm, "ms_per_sweep": ms,
            "per_type_GBps_same_type_back_to_back": per}, 5 * (reps + 3) + 50


def all_types_section(ggml, synth, dev, peaks, peak_kind):
    """Every type the library serves, on one shape each: dequantize 4096x14336 -> fp16 (algorithmic bytes = blocks read +
    fp16 written) and MMVQ 14336x4096 at batch 1 and 8 (weight bytes / time).  The headline sweep and the MMVQ section
    above cover the K-quants of BASELINE's configs; this table shows where the legacy and IQ4 types stand."""
    hbm = float(peaks["hbm_gbs"])
    out = {"unit": "GB/s", "peak": hbm, "peak_source": peak_kind,
           "method": "CUDA-graph replay over distinct weight copies (> 2x L2), median of 5 replays, fp16", "cases": []}
    launches = 0
    for t in ggml.SUPPORTED_TYPES:
        case = {"type": synth.NAMES[t]}
        # dequantize: 4 weight / output pairs per replay (4 x 117 MB of output > L2)
        ws = [torch.from_numpy(synth.random_blocks(t, 4096, 14336, seed=i)).to(dev) for i in range(4)]
        outs = [torch.empty((4096, 14336), dtype=torch.float16, device=dev) for _ in range(4)]

        def run_dq():
            for w, o in zip(ws, outs):
                ggml.dequantize_out(w, t, 4096, 14336, o)

        ms = _time_graph(run_dq, iters=5) / 4
        launches += 4 * 9
        nbytes = ws[0].numel() + 4096 * 14336 * 2
        case["dequant_us"] = ms * 1e3
        case["dequant_GBps"] = nbytes / (ms * 1e-3) / 1e9
        del ws, outs
        # MMVQ: 8 copies of the 14336 x 4096 matrix
        N, K = 14336, 4096
        base = torch.from_numpy(synth.random_blocks(t, N, K, seed=t)).to(dev)
        copies = [base] + [base.clone() for _ in range(7)]
        for B in (1, 8):
            x = torch.randn(B, K, dtype=torch.float16, device=dev)
            y = torch.empty(B, N, dtype=torch.float16, device=dev)
            wsb = torch.empty(max(ggml.mmvq_workspace_bytes(t, B, K, N), 16), dtype=torch.uint8, device=dev)

            def run_mv():
                for w in copies:
                    ggml.mul_mat_vec_a8_out(w, x, t, N, y, wsb)

            ms = _time_graph(run_mv, iters=5) / 8
            launches += 8 * 9
            case[f"mmvq_b{B}_us"] = ms * 1e3
            case[f"mmvq_b{B}_weight_GBps"] = base.numel() / (ms * 1e-3) / 1e9
        del copies, base
        out["cases"].append(case)
    return out, launches


def run(ggml, synth, dev, rank, world, peaks, peak_kind):
    extras = {}
    n = 0
    # inference weights are immutable: weight streams may start under the previous kernel (include/ggml_b200.h).
    # Set for every section below (the headline sweep and e2e in bench.py ran with the default, 0).
    ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)
    extras["options"] = "GGB_OPT_STATIC_WEIGHTS=1 for all extras"
    try:
        tp, k = tp_mlp_section(ggml, synth, dev, rank, world, peaks, peak_kind)      # every rank takes part
        n += k
        if rank == 0:
            extras["tp_mlp"] = tp
    except Exception as e:
        if rank == 0:
            extras["tp_mlp"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        for name, fn in (("config0_gpu", config0_section), ("dequant_sweep", dequant_sweep_section), ("mmvq", mmvq_section),
                         ("mmq", mmq_section), ("all_types", all_types_section)):
            try:
                extras[name], k = fn(ggml, synth, dev, peaks, peak_kind)
                n += k
            except Exception as e:
                extras[name] = {"error": f"{type(e).__name__}: {e}"}
    return extras, n
