"""Debug probe: MMVQ error statistics against the oracle on a few shapes, then graph-replay timings
(8 distinct weight copies per replay, like bench_extras.mmvq_section).

    mmvq_probe.py [check] [time] [types=12,14] [static=0|1] [B=1,2,4,8]
"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.environ.get("GGML_PKG_DIR", os.path.join(ROOT, "ggml-libtorch_b200")), os.path.join(ROOT, "tests")]
import numpy as np, torch
import oracle, ggml
from ggml import synth, _capi

args = sys.argv[1:]
opt = dict(a.split("=") for a in args if "=" in a)
types = [int(v) for v in opt.get("types", "12").split(",")]
batches = [int(v) for v in opt.get("B", "1,2,4,8").split(",")]
static = int(opt.get("static", "0"))
_capi.lib.ggb_set_option(_capi.GGB_OPT_STATIC_WEIGHTS, static)
warps, cap = int(opt.get("warps", "0")), int(opt.get("cap", "0"))
_capi.lib.ggb_set_option(2, warps | (cap << 8))
if "mma" in opt:      # cs,s_log2+1,min_batch
    a, b, c = (int(v) for v in opt["mma"].split(","))
    _capi.lib.ggb_set_option(5, a | (b << 8) | (c << 16))
if "tiled" in opt:    # largest TMA box of the generic kernel's tiled stages (0 = per-row bulk copies)
    _capi.lib.ggb_set_option(6, int(opt["tiled"]))
l2 = int(opt.get("l2", "0"))
_capi.lib.ggb_set_option(4, l2)
_capi.lib.ggb_set_option(7, int(opt.get("rg", "1")))
shapes = ((14336, 4096), (4096, 14336), (28672, 8192), (8192, 28672))
if "shapes" in opt:
    shapes = tuple(tuple(int(v) for v in sh.split("x")) for sh in opt["shapes"].split(","))

if "check" in args or not [a for a in args if "=" not in a]:
    for t in types:
        for (N, K, B) in [(1, 256, 1), (5, 512, 3), (96, 1024, 2), (300, 4096, 1), (64, 6144, 3), (33, 2304, 2), (257, 2048, 8),
                          (257, 2048, 5), (1000, 14336, 1), (999, 8192, 4), (14336, 4096, 1), (14336, 4096, 8), (4096, 14336, 3), (1001, 2304, 7), (515, 28672, 8), (130, 14336, 8), (77, 16384, 6)]:
            w = oracle.random_weights(t, N, K, seed=1)
            x = torch.randn(B, K, generator=torch.Generator().manual_seed(0))
            want = oracle.mmvq(w, t, x.numpy(), N)
            got = ggml.ggml_mul_mat_vec_a8(torch.from_numpy(w).cuda(), x.cuda(), t, N).cpu().numpy()
            err = np.abs(got - want).max()
            scale = np.abs(want).max()
            print(f"type {t} N={N} K={K} B={B}: max err {err:.3e} scale {scale:.3e} rel {err/scale:.2e} "
                  f"{'OK' if err <= 2e-5 * scale else 'BAD'} nan {np.isnan(got).sum()}", flush=True)

if "time" in args:
    dev = torch.device("cuda")
    for t in types:
        for (N, K) in shapes:
            base = torch.from_numpy(synth.random_blocks(t, N, K, seed=N + t)).to(dev)
            ncopies = int(opt.get("copies", max(2, min(8, int(600e6 // base.numel())))))
            copies = [base] + [base.clone() for _ in range(ncopies - 1)]
            reps = int(opt.get("reps", "1"))             # launches per graph = ncopies * reps
            for B in batches:
                x = torch.randn(B, K, dtype=torch.float16, device=dev)
                y = torch.empty(B, N, dtype=torch.float16, device=dev)
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for w in copies:
                        ggml.mul_mat_vec_a8_out(w, x, t, N, y)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        for _ in range(reps):
                            for w in copies:
                                ggml.mul_mat_vec_a8_out(w, x, t, N, y)
                    for _ in range(3):
                        g.replay()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(10):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(s); g.replay(); b.record(s); b.synchronize()
                        ts.append(a.elapsed_time(b))
                ms = statistics.median(ts)
                us = ms * 1e3 / (ncopies * reps)
                print(f"type {t} N={N} K={K} B={B} static={static} l2={l2} warps={warps} cap={cap} reps={reps}: {us:.2f} us/launch  {base.numel()/us/1e3:.0f} GB/s "
                      f"({base.numel()/us/1e3/6552.3*100:.1f}% of 6552)", flush=True)
            del copies, base
