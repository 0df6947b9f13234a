"""Run one op a few times (for ncu captures and quick timings):  op_run.py {mmq|mmvq|dequant} type N K M [iters] [opt:<id>=<value> ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ggml-libtorch_b200")]
import torch, ggml
from ggml import synth
from ggml import _capi
for a in [a for a in sys.argv if a.startswith("opt:")]:          # opt:<ggb_option id>=<value>
    k, v = a[4:].split("=")
    _capi.lib.ggb_set_option(int(k), int(v))
sys.argv = [a for a in sys.argv if not a.startswith("opt:")]
op, t, N, K, M = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
w = torch.from_numpy(synth.random_blocks(t, N, K, seed=1)).cuda()
x = torch.randn(M, K, dtype=torch.float16, device="cuda")
if op == "mmq":
    y = torch.empty(M, N, dtype=torch.float16, device="cuda")
    ws = torch.empty(max(ggml.mmq_workspace_bytes(t, M, K, N), 256), dtype=torch.uint8, device="cuda")
    f = lambda: ggml.mul_mat_a8_out(w, x, t, N, y, ws)
elif op == "mmvq":
    y = torch.empty(M, N, dtype=torch.float16, device="cuda")
    f = lambda: ggml.mul_mat_vec_a8_out(w, x, t, N, y)
else:
    y = torch.empty(N, K, dtype=torch.float16, device="cuda")
    f = lambda: ggml.dequantize_out(w, t, N, K, y)
for _ in range(3):
    f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    f()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
print(f"{op} type={t} N={N} K={K} M={M}: {ms*1e3:.1f} us/iter", f"{2.0*M*N*K/ms/1e9:.1f} TOPS" if op == "mmq" else "")
