"""Quick MMQ timings: mmq_bench.py [types=12,14] [M=512,1024,2048,4096,8192] [nb=0|1] [shapes=14336x4096]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ggml-libtorch_b200")]
import torch, ggml
from ggml import synth, _capi
opt = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
types = [int(v) for v in opt.get("types", "12,14").split(",")]
Ms = [int(v) for v in opt.get("M", "512,1024,2048,4096,8192").split(",")]
shapes = [tuple(int(v) for v in sh.split("x")) for sh in opt.get("shapes", "14336x4096").split(",")]
_capi.lib.ggb_set_option(3, int(opt.get("nb", "0")))
for t in types:
    for (N, K) in shapes:
        w = torch.from_numpy(synth.random_blocks(t, N, K, seed=1)).cuda()
        for M in Ms:
            x = torch.randn(M, K, dtype=torch.float16, device="cuda")
            y = torch.empty(M, N, dtype=torch.float16, device="cuda")
            ws = torch.empty(max(ggml.mmq_workspace_bytes(t, M, K, N), 256), dtype=torch.uint8, device="cuda")
            f = lambda: ggml.mul_mat_a8_out(w, x, t, N, y, ws)
            for _ in range(3): f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): f()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            tops = 2.0 * M * N * K / ms / 1e9
            print(f"mmq type={t} N={N} K={K} M={M} nb={opt.get('nb','0')}: {ms*1e3:.1f} us  {tops:.0f} TOPS ({tops/1595.8*100:.1f}% of bf16 1595.8)", flush=True)
