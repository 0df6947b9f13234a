"""Timing of the grouped MMVQ (ggml_moe_a8_vec) on a Mixtral-8x7B-shaped expert stack (8 experts, Q4_K gate/up
[14336, 4096] and down [4096, 14336], top-2), beside the obvious alternative: one plain ggml_mul_mat_vec_a8 per pair
(which needs the expert ids on the host).  CUDA-graph replays of 8 launches that cycle through all experts (264 MB per
cycle > 2x L2), CUDA events, one JSON line on stdout.

    python scripts/moe_bench.py [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ggml-libtorch_b200"))

import numpy as np
import torch

import ggml
from ggml import synth

HBM_PEAK = 6552.3
try:
    HBM_PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def graph_time(fns, reps):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))          # us per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    t, E = 12, 8
    ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)
    raw = np.stack([synth.random_blocks(t, 14336, 4096, seed=e) for e in range(E)])
    Wd = torch.from_numpy(raw).cuda()
    per_expert = raw[0].size
    out = {"what": "ggml_moe_a8_vec vs one ggml_mul_mat_vec_a8 per pair", "type": "Q4_K", "experts": E,
           "hbm_peak_gbs": HBM_PEAK, "cases": []}
    for name, N, K in (("gate/up", 14336, 4096), ("down", 4096, 14336)):
        W = Wd.view(E, N, -1)
        for tokens, top_k in ((1, 2), (4, 2), (8, 2)) if name == "gate/up" else ((2, 1), (8, 1), (16, 1)):
            pairs = tokens * top_k
            x = torch.randn(tokens, K, device="cuda", dtype=torch.float16)
            y = torch.empty(pairs, N, device="cuda", dtype=torch.float16)
            # 8 id sets: launch j uses experts j*pairs .. j*pairs + pairs - 1 (mod E), so every cycle touches all experts
            sets = [torch.tensor([(j * pairs + i) % E for i in range(pairs)], dtype=torch.int32) for j in range(8)]
            dev_sets = [s.cuda() for s in sets]
            grouped = [(lambda ids=ids: ggml.moe_a8_vec_out(x, W, ids, top_k, t, N, tokens, y)) for ids in dev_sets]

            def per_pair(ids_host):
                for i, e in enumerate(ids_host):
                    ggml.mul_mat_vec_a8_out(W[e], x[i // top_k:i // top_k + 1], t, N, y[i:i + 1])
            plain = [(lambda s=s.tolist(): per_pair(s)) for s in sets]
            distinct = min(pairs, E)
            rec = {"proj": name, "N": N, "K": K, "tokens": tokens, "top_k": top_k, "pairs": pairs}
            try:
                us = graph_time(grouped, a.reps)
                rec["grouped_us"] = round(us, 2)
                rec["grouped_weight_GBps_distinct"] = round(distinct * per_expert / us / 1e3, 1)
                rec["grouped_weight_GBps_streamed"] = round(pairs * per_expert / us / 1e3, 1)
                rec["frac_of_hbm_peak_distinct"] = round(distinct * per_expert / us / 1e3 / HBM_PEAK, 3)
                us2 = graph_time(plain, a.reps)
                rec["per_pair_launches_us"] = round(us2, 2)
            except Exception as ex:                      # keep what was measured
                rec["error"] = repr(ex)[:200]
            out["cases"].append(rec)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
