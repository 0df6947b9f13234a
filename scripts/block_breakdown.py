"""Debug: where the time of the headline step goes (tp = 1): the stack's launches timed as sub-sequences, L layers of distinct weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import torch, ggml
from ggml import synth, _capi
from bench import time_graph, down_type
dev = torch.device("cuda")
ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)
for a in sys.argv[1:]:
    if a.startswith("opt:"):
        k, v = a[4:].split("="); _capi.lib.ggb_set_option(int(k), int(v))
H, I = 8192, 28672
L = int(os.environ.get("LAYERS", "16"))
B = 1
base_gu = torch.from_numpy(synth.random_blocks(12, 2 * I, H, seed=1)).to(dev)
base_dn = {t: torch.from_numpy(synth.random_blocks(t, H, I, seed=t)).to(dev) for t in (12, 14)}
gu_w = [base_gu.clone() for _ in range(L)]
tds = [down_type(l, 80) for l in range(L)] if os.environ.get("MIX", "1") == "1" else [int(os.environ["MIX"])] * L
dn_w = [base_dn[t].clone() for t in tds]
x = torch.randn(B, H, dtype=torch.float16, device=dev)
gu = torch.randn(8, 2 * I, dtype=torch.float16, device=dev)
dn = torch.empty(8, H, dtype=torch.float16, device=dev)
need = max(ggml.mmvq_workspace_bytes(12, B, H, 2 * I), max(ggml.mmvq_workspace_bytes(t, B, I, H) for t in (12, 14)), 16)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
h = torch.randn(B, I, dtype=torch.float16, device=dev)

def f_gu():
    for w in gu_w:
        ggml.mul_mat_vec_a8_fused(w, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)
def f_dn_swiglu():
    for w, t in zip(dn_w, tds):
        ggml.mul_mat_vec_a8_fused(w, gu[:B, :I], gu[:B, I:], t, H, [dn.data_ptr()], H, 0, 0, 0, ws)
def f_dn_plain():
    for w, t in zip(dn_w, tds):
        ggml.mul_mat_vec_a8_fused(w, h, None, t, H, [dn.data_ptr()], H, 0, 0, 0, ws)
def f_block():
    for wg, wd, t in zip(gu_w, dn_w, tds):
        ggml.mul_mat_vec_a8_fused(wg, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)
        ggml.mul_mat_vec_a8_fused(wd, gu[:B, :I], gu[:B, I:], t, H, [dn.data_ptr()], H, 0, 0, 0, ws)
def f_block_plain_dn():
    for wg, wd, t in zip(gu_w, dn_w, tds):
        ggml.mul_mat_vec_a8_fused(wg, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)
        ggml.mul_mat_vec_a8_fused(wd, h, None, t, H, [dn.data_ptr()], H, 0, 0, 0, ws)
r = {n: time_graph(torch, f, iters=5) / L * 1e3 for n, f in (("gate|up", f_gu), ("down(swiglu)", f_dn_swiglu), ("down(plain)", f_dn_plain),
                                                                ("block", f_block), ("block(plain down)", f_block_plain_dn))}
print(f"L={L} downs={sorted(set(tds))} args={sys.argv[1:]}: " + "  ".join(f"{k} {v:.1f} us" for k, v in r.items()), flush=True)
