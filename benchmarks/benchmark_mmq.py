#!/usr/bin/env python
"""The reference's benchmark CLI (benchmarks/benchmark_mmq.py:148-171) on this library: same arguments, same two
measurements per quant type -- "dequantize + x @ W.T" against "ggml_mul_mat_a8" -- same console lines and the same CSV
(`benchmark_results_local_HD<hidden>xB<tokens>.csv`).  Differences: the sample GGUF files are generated locally
(benchmarks/fixtures.py; the reference downloads them), `--hidden-size` also takes 4096 / 8192, `--use-remote` is refused
(no hub), and every type the reference lists is timed with CUDA events as well as its host wall clock.

    python benchmarks/benchmark_mmq.py --num-tokens 4096 --hidden-size 1024 --dtype half
"""
from __future__ import annotations

import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(os.path.dirname(HERE), "ggml-libtorch_b200")]

import pandas as pd  # noqa: E402
import torch  # noqa: E402
from gguf import GGMLQuantizationType as QT  # noqa: E402

from fixtures import get_gguf_sample_tensors, seed_everything  # noqa: E402

QUANT_TYPES_MAP = {n: getattr(QT, n) for n in ("Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K", "Q4_0", "Q5_0", "Q8_0")}
DTYPES_MAP = {"half": torch.float16, "bfloat16": torch.bfloat16, "float": torch.float32}


def _timed(fn, iters: int, profile: bool) -> tuple[float, float]:
    """(host seconds, device seconds) per iteration"""
    torch.cuda.synchronize()
    if profile:
        torch.cuda.cudart().cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    host = (time.perf_counter() - t0) / iters
    if profile:
        torch.cuda.cudart().cudaProfilerStop()
    return host, e0.elapsed_time(e1) * 1e-3 / iters


@torch.inference_mode()
def bench_type(ops, quant_type, args, dtype):
    seed_everything(args.seed)
    torch.set_default_device("cuda")
    tensors = get_gguf_sample_tensors(hidden_size=args.hidden_size, quant_type=quant_type)
    weights = [torch.tensor(t.data, device="cuda") for t in tensors]
    shapes = [tuple(int(v) for v in t.name.split("_")[-1].split("x")) for t in tensors]
    xs = {k: torch.randn(args.num_tokens, k, dtype=dtype) for k in {s[1] for s in shapes}}

    def via_dequant():
        for w, (rows, cols) in zip(weights, shapes):
            xs[cols] @ ops.ggml_dequantize(w, quant_type, rows, cols).T.to(dtype)

    def via_mmq():
        for w, (rows, cols) in zip(weights, shapes):
            ops.ggml_mul_mat_a8(w, xs[cols], quant_type, w.size(0))

    out = []
    for label, fn in (("Dequant", via_dequant), ("MMQ", via_mmq)):
        _timed(fn, args.num_warmup_iters, False)
        host, dev = _timed(fn, 1 if args.profile else args.num_iters, args.profile)
        print(f"{quant_type.name} {label} Kernel running time: {host * 1e3 :.3f} ms   (device {dev * 1e3:.3f} ms)")
        out.append((host * 1e3, dev * 1e3))
    torch.set_default_device("cpu")
    return out


def main():
    ap = argparse.ArgumentParser(description="Benchmark the MMQ kernel for matrix multiplication with quantized weights")
    ap.add_argument("--num-tokens", type=int, default=4096)
    ap.add_argument("--hidden-size", type=int, choices=[256, 1024, 4096, 8192], default=1024)
    ap.add_argument("--quant-dtype", type=str, choices=QUANT_TYPES_MAP.keys(), default="Q4_0")
    ap.add_argument("--dtype", type=str, choices=DTYPES_MAP.keys(), default="half")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--use-remote", action="store_true", help="Use remote kernels (not available: no hub access)")
    ap.add_argument("--revision", type=str, default="main", help="The branch, tag or commit to use for remote kernel.")
    ap.add_argument("--num-warmup-iters", type=int, default=5)
    ap.add_argument("--num-iters", type=int, default=100, help="Number of benchmark iterations. If --profile is set, this number is ignored")
    args = ap.parse_args()
    print(args)
    if args.use_remote:
        raise SystemExit("--use-remote needs the Hugging Face hub (kernels.get_kernel); this environment has no network")
    print("Using local kernels.")
    import ggml as ops
    dtype = DTYPES_MAP[args.dtype]
    rows = {"Quantization": [], "Dequant Time (ms)": [], "MMQ Time (ms)": [], "Dequant device (ms)": [], "MMQ device (ms)": []}
    for name, qt in QUANT_TYPES_MAP.items():              # like the reference: every type, whatever --quant-dtype says
        (dq_h, dq_d), (mm_h, mm_d) = bench_type(ops, qt, args, dtype)
        rows["Quantization"].append(name)
        rows["Dequant Time (ms)"].append(dq_h)
        rows["MMQ Time (ms)"].append(mm_h)
        rows["Dequant device (ms)"].append(dq_d)
        rows["MMQ device (ms)"].append(mm_d)
    pd.DataFrame(rows).to_csv(f"benchmark_results_local_HD{args.hidden_size}xB{args.num_tokens}.csv", index=False)


if __name__ == "__main__":
    main()
