"""Debug: time the two launches of the fused MLP block separately (tp = 1), with and without the pre-launched activation image."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import torch, ggml
from ggml import synth, _capi
sys.path.insert(0, ROOT)
from bench import time_graph
dev = torch.device("cuda")
ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)
H, I = 8192, 28672
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
gu_w = [torch.from_numpy(synth.random_blocks(12, 2 * I, H, seed=1)).to(dev)]
gu_w += [gu_w[0].clone() for _ in range(3)]
x = torch.randn(B, H, dtype=torch.float16, device=dev)
gu = torch.randn(8, 2 * I, dtype=torch.float16, device=dev)
dn = torch.empty(8, H, dtype=torch.float16, device=dev)
for td in (12, 14):
    dn_w = [torch.from_numpy(synth.random_blocks(td, H, I, seed=2)).to(dev)]
    dn_w += [dn_w[0].clone() for _ in range(3)]
    for pq in (12288, 0):
        _capi.lib.ggb_set_option(8, pq)
        need = max(ggml.mmvq_workspace_bytes(12, B, H, 2 * I), ggml.mmvq_workspace_bytes(td, B, I, H), 16)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)

        def f_gu():
            for w in gu_w:
                ggml.mul_mat_vec_a8_fused(w, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)

        def f_dn_swiglu():
            for w in dn_w:
                ggml.mul_mat_vec_a8_fused(w, gu[:B, :I], gu[:B, I:], td, H, [dn.data_ptr()], H, 0, 0, 0, ws)

        h = torch.randn(B, I, dtype=torch.float16, device=dev)

        def f_dn_plain():
            for w in dn_w:
                ggml.mul_mat_vec_a8_fused(w, h, None, td, H, [dn.data_ptr()], H, 0, 0, 0, ws)

        def f_block():
            for wg, wd in zip(gu_w, dn_w):
                ggml.mul_mat_vec_a8_fused(wg, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)
                ggml.mul_mat_vec_a8_fused(wd, gu[:B, :I], gu[:B, I:], td, H, [dn.data_ptr()], H, 0, 0, 0, ws)

        r = {n: time_graph(torch, f, iters=5) / 4 * 1e3 for n, f in (("gate|up", f_gu), ("down(swiglu)", f_dn_swiglu), ("down(plain)", f_dn_plain), ("block", f_block))}
        print(f"B={B} down type {td} prequant_min_k={pq} ws={need}: " + "  ".join(f"{k} {v:.1f} us" for k, v in r.items()), flush=True)
