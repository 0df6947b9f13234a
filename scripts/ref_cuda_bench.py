"""Time the reference's own CUDA ops (oracle/_ref_cuda, built unmodified for sm_100a) on BASELINE's shapes: the same-box
GPU comparator.  Device time per op call (CUDA events around `iters` calls after warm-up; each call = the reference's
two launches + its two torch::empty allocations, exactly what a caller of the op pays).  Weights rotate over distinct
copies (> 2x L2) for the bandwidth-bound ops.

    python scripts/ref_cuda_bench.py [mmvq] [mmq] [dequant]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import torch

import oracle
from ggml import synth


def ev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    what = [a for a in sys.argv[1:]] or ["mmvq", "mmq", "dequant"]
    ref = oracle.ref_cuda_ops()
    assert ref is not None, "oracle/_ref_cuda is not built"
    dev = torch.device("cuda")
    out = {}
    if "mmvq" in what:
        rows = []
        for tname, t, N, K in (("Q4_K", 12, 14336, 4096), ("Q4_K", 12, 4096, 14336), ("Q6_K", 14, 4096, 14336),
                               ("Q4_K", 12, 28672, 8192), ("Q6_K", 14, 8192, 28672)):
            base = torch.from_numpy(synth.random_blocks(t, N, K, seed=N + t)).to(dev)
            nc = max(2, min(8, int(600e6 // base.numel())))
            copies = [base] + [base.clone() for _ in range(nc - 1)]
            x = torch.randn(1, K, dtype=torch.float16, device=dev)
            state = {"i": 0}

            def call():
                w = copies[state["i"] % nc]
                state["i"] += 1
                ref.ggml_mul_mat_vec_a8(w, x, t, N)

            ms = ev_time(call, 8 * nc, warm=nc)
            rows.append({"type": tname, "N": N, "K": K, "B": 1, "us": ms * 1e3, "weight_GBps": base.numel() / (ms * 1e-3) / 1e9})
            del copies, base
        out["mmvq"] = rows
    if "mmq" in what:
        rows = []
        for tname, t in (("Q4_K", 12), ("Q6_K", 14)):
            N, K = 14336, 4096
            w = torch.from_numpy(synth.random_blocks(t, N, K, seed=1)).to(dev)
            for M in (512, 2048, 8192):
                x = torch.randn(M, K, dtype=torch.float16, device=dev)
                ms = ev_time(lambda: ref.ggml_mul_mat_a8(w, x, t, N), 5, warm=2)
                rows.append({"type": tname, "M": M, "N": N, "K": K, "ms": ms, "TOPS": 2.0 * M * N * K / (ms * 1e-3) / 1e12})
        out["mmq"] = rows
    if "dequant" in what:
        rows = []
        for tname, t in (("Q2_K", 10), ("Q3_K", 11), ("Q4_K", 12), ("Q5_K", 13), ("Q6_K", 14)):
            ws = [torch.from_numpy(synth.random_blocks(t, 4096, 14336, seed=i)).to(dev) for i in range(4)]
            state = {"i": 0}

            def call():
                w = ws[state["i"] % 4]
                state["i"] += 1
                ref.ggml_dequantize(w, t, 4096, 14336)

            ms = ev_time(call, 12, warm=4)
            nbytes = ws[0].numel() + 4096 * 14336 * 2
            rows.append({"type": tname, "us": ms * 1e3, "GBps": nbytes / (ms * 1e-3) / 1e9})
            del ws
        out["dequant"] = rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()
