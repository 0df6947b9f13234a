"""Small shapes of every kernel family for compute-sanitizer (memcheck / racecheck): dequantize, both q8_1 layouts, the four
MMVQ kernels (register-resident batch 1, dp4a ring, tensor-core batches, generic), the gather epilogue with two local
'peers', the fused SwiGLU block with its pre-launched activation image, and MMQ incl. the CTA-pair path and a tail launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import torch, ggml
from ggml import synth
dev = torch.device("cuda")
ggml.set_option(ggml.GGB_OPT_STATIC_WEIGHTS, 1)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def W(t, n, k, seed=0): return torch.from_numpy(synth.random_blocks(t, n, k, seed=seed)).to(dev)
if which in ("all", "dequant"):
    for t in (2, 8, 10, 11, 12, 13, 14, 20, 23):
        for dt in (torch.float16, torch.float32):
            ggml.ggml_dequantize(W(t, 40, 512), t, 40, 512, dtype=dt)
    for lay in (ggml.GGB_Q8_1, ggml.GGB_Q8_1_MMQ):
        ggml.ggml_quantize_q8_1(torch.randn(9, 1024, device=dev, dtype=torch.float16), lay, 12)
if which in ("all", "mmvq"):
    for t, n, k in ((12, 300, 4096), (12, 70, 14336), (12, 65, 2048), (14, 66, 2048), (13, 64, 1024), (23, 64, 1024), (2, 33, 1024), (8, 40, 256), (10, 64, 512), (11, 64, 512), (20, 40, 1024)):
        w = W(t, n, k)
        for B in (1, 2, 5, 8):
            ggml.ggml_mul_mat_vec_a8(w, torch.randn(B, k, device=dev, dtype=torch.float16), t, n)
    # gather epilogue, two local peers
    w = W(12, 264, 2048)
    bufs = [torch.zeros(8, 1000, dtype=torch.float16, device=dev) for _ in range(2)]
    for B in (1, 2, 6):
        ggml.mul_mat_vec_a8_gather(w, torch.randn(B, 2048, device=dev, dtype=torch.float16), 12, 264, [b.data_ptr() for b in bufs], 1000, 304)
    # fused block: stacked gate|up with split columns, SwiGLU down projection (pre-launched activation image for K >= 4096)
    H, I = 1024, 4096
    wgu, wd = W(12, 2 * I, H), W(14, H, I)
    gu = torch.zeros(8, 2 * I, dtype=torch.float16, device=dev); dn = torch.zeros(8, H, dtype=torch.float16, device=dev)
    for B in (1, 2, 4):
        x = torch.randn(B, H, device=dev, dtype=torch.float16)
        ws = torch.empty(max(ggml.mmvq_workspace_bytes(12, B, H, 2 * I), ggml.mmvq_workspace_bytes(14, B, I, H), 16), dtype=torch.uint8, device=dev)
        ggml.mul_mat_vec_a8_fused(wgu, x, None, 12, 2 * I, [gu.data_ptr()], 2 * I, 0, I, I, ws)
        ggml.mul_mat_vec_a8_fused(wd, gu[:B, :I], gu[:B, I:], 14, H, [dn.data_ptr()], H, 0, 0, 0, ws)
if which in ("all", "mmq"):
    for t, n, k, m in ((12, 256, 1024, 7), (14, 200, 512, 83), (8, 129, 2080, 20), (10, 128, 256, 300), (12, 148 * 128 + 64, 512, 520), (12, 14336, 512, 1100)):
        ggml.ggml_mul_mat_a8(W(t, n, k), torch.randn(m, k, device=dev, dtype=torch.float16), t, n)
torch.cuda.synchronize()
print("sanitize_run ok:", which)
