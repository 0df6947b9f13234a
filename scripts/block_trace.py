"""Debug: per-CTA timeline of the headline block's launches (stacked gate|up rg kernel -> SwiGLU image launch -> down rg kernel),
globaltimer at CTA start + clock64 phase stamps converted at the SM clock."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import numpy as np, torch, ggml
from ggml import synth, _capi
L = _capi.lib
ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)
for a in sys.argv[1:]:
    if a.startswith("opt:"):
        k, v = a[4:].split("="); L.ggb_set_option(int(k), int(v))
tds = [int(v) for v in os.environ.get("TD", "12").split(",")]; plain = os.environ.get("PLAIN", "0") == "1"
H, I, B = 8192, 28672, 1
NB = max(4, len(tds))
tds = (tds * NB)[:NB]
dev = torch.device("cuda")
gu_w = [torch.from_numpy(synth.random_blocks(12, 2 * I, H, seed=1)).to(dev) for _ in range(1)]
gu_w += [gu_w[0].clone() for _ in range(NB - 1)]
dn_w = [torch.from_numpy(synth.random_blocks(t, H, I, seed=2)).to(dev) for t in tds]
x = torch.randn(B, H, dtype=torch.float16, device=dev)
gu = torch.randn(8, 2 * I, dtype=torch.float16, device=dev); dn = torch.empty(8, H, dtype=torch.float16, device=dev)
h = torch.randn(B, I, dtype=torch.float16, device=dev)
need = max([ggml.mmvq_workspace_bytes(12, B, H, 2 * I), 16] + [ggml.mmvq_workspace_bytes(t, B, I, H) for t in tds])
ws = torch.empty(need, dtype=torch.uint8, device=dev)
def blocks():
    for wg, wd, td in zip(gu_w, dn_w, tds):
        ggml.mul_mat_vec_a8_fused(wg, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)
        if plain: ggml.mul_mat_vec_a8_fused(wd, h, None, td, H, [dn.data_ptr()], H, 0, 0, 0, ws)
        else: ggml.mul_mat_vec_a8_fused(wd, gu[:B, :I], gu[:B, I:], td, H, [dn.data_ptr()], H, 0, 0, 0, ws)
L.ggb_debug_mmvq_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(64 * 2048, dtype=torch.int64, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    blocks(); torch.cuda.synchronize()
    L.ggb_debug_mmvq_trace(trace.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        blocks()
    L.ggb_debug_mmvq_trace(None)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
nl = 2 * NB
raw = trace.cpu().numpy().reshape(-1, 256, 8)[:nl, :148]
smid = raw[:, :, 7] & 0xff
raw[:, :, 7] &= ~0xff
tr = raw.astype(np.float64)
for i in range(nl):
    cnt = np.bincount(smid[i], minlength=148)[:148]
    print(f"launch {i} ({'gu' if i % 2 == 0 else 'down type %d' % tds[i // 2]}): CTAs per SM: {int((cnt == 0).sum())} SMs with none, {int((cnt == 1).sum())} with one, {int((cnt >= 2).sum())} with two or more")
clk = 1.965e3   # cycles per us
g0 = tr[0, :, 7].min()
names = ["dep", "Xin", "quant", "regs", "4rows", "end"]
for i in range(nl):
    a = tr[i]
    st = (a[:, 7] - g0) / 1e3
    rel = (a[:, 1:7] - a[:, 0:1]) / clk + st[:, None]       # absolute us of each stamp
    row = f"launch {i} ({'gu' if i % 2 == 0 else 'down'}): start {st.min():7.2f}..{st.max():7.2f} |"
    for k, nm in enumerate(names):
        row += f" {nm} {rel[:, k].min():7.2f}..{rel[:, k].max():7.2f} |"
    if os.environ.get("VERBOSE", "1") == "1": print(row)
    if i >= 2:
        dur = rel[:, 5] - rel[:, 0]
        print(f"  launch {i} ({'gu' if i % 2 == 0 else 'down'}): dep->end min {dur.min():6.2f} median {np.median(dur):6.2f} max {dur.max():6.2f} us; "
              f"first dep {rel[:, 0].min():7.2f} last end {rel[:, 5].max():7.2f}")
