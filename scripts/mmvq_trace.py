"""Debug: per-CTA timeline of consecutive MMVQ launches inside one CUDA graph (globaltimer stamps)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import numpy as np, torch
import ggml
from ggml import synth, _capi
opt = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
static, warps, cap = int(opt.get("static", "1")), int(opt.get("warps", "0")), int(opt.get("cap", "0"))
N, K, t, B = int(opt.get("N", 14336)), int(opt.get("K", 4096)), int(opt.get("t", 12)), int(opt.get("B", 1))
L = _capi.lib
L.ggb_set_option(1, static); L.ggb_set_option(2, warps | (cap << 8))
l2 = int(opt.get("l2", "0")); L.ggb_set_option(4, l2)
L.ggb_set_option(7, int(opt.get("rg", "1")))
L.ggb_debug_mmvq_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda")
base = torch.from_numpy(synth.random_blocks(t, N, K, seed=1)).to(dev)
copies = [base] + [base.clone() for _ in range(5)]
x = torch.randn(B, K, dtype=torch.float16, device=dev); y = torch.empty(B, N, dtype=torch.float16, device=dev)
nl = len(copies)
trace = torch.zeros(64 * 1024, dtype=torch.int64, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for w in copies: ggml.mul_mat_vec_a8_out(w, x, t, N, y)
    torch.cuda.synchronize()
    L.ggb_debug_mmvq_trace(trace.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for w in copies: ggml.mul_mat_vec_a8_out(w, x, t, N, y)
    L.ggb_debug_mmvq_trace(None)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
raw = trace.cpu().numpy().reshape(-1, 2048)[:nl]
rg = int(opt.get("rg", "1")) & 1
if rg and B == 1 and t == 12:
    # rg kernel: 8 stamps per CTA: clock64 at start, dep-wait done, first X loads landed, quantized, activations in registers,
    # first 4 rows done, end; [7] = globaltimer at start
    tr = raw.reshape(nl, 256, 8)[:, :148]
    for i in range(nl):
        cnt = np.bincount(tr[i, :, 7] & 0xff, minlength=148)[:148]
        print(f"launch {i}: {int((cnt == 0).sum())} SMs without a CTA, {int((cnt == 1).sum())} with one, {int((cnt >= 2).sum())} with two")
    tr[:, :, 7] &= ~0xff
    tr = tr.astype(np.float64)
    g0 = tr[0, :, 7].min()
    print(f"rg={opt.get('rg')} static={static} N={N} K={K} B={B}: per-CTA phases in SM cycles (min / median / max over CTAs); start = globaltimer us")
    names = ["start->dep-wait done", "->X loads landed", "->quantized", "->acts in regs", "->first 4 rows", "->end"]
    for i in range(nl):
        a = tr[i]
        row = f"launch {i}: start {(a[:,7].min()-g0)/1e3:6.2f}..{(a[:,7].max()-g0)/1e3:6.2f} us |"
        for k, nm in enumerate(names):
            d = a[:, k + 1] - a[:, k]
            row += f" {nm} {int(d.min())}/{int(np.median(d))}/{int(d.max())} |"
        tot = a[:, 6] - a[:, 1]
        row += f" dep-wait done->end {int(tot.min())}/{int(np.median(tot))}/{int(tot.max())}"
        print(row)
    sys.exit(0)
tr = raw[:, :1024].reshape(nl, 256, 4)[:, :148].astype(np.float64)
t0 = tr[0, :, 0].min()
tr = (tr - t0) / 1e3
print(f"l2={l2} static={static} warps={warps} cap={cap} N={N} K={K} B={B}   (us since first CTA start)")
for i in range(nl):
    a = tr[i]
    print(f"launch {i}: start {a[:,0].min():7.2f}..{a[:,0].max():7.2f}  dep-wait done {a[:,1].min():7.2f}..{a[:,1].max():7.2f}  "
          f"prologue done {a[:,2].min():7.2f}..{a[:,2].max():7.2f}  end {a[:,3].min():7.2f}..{a[:,3].max():7.2f}")
