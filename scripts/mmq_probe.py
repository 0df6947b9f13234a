"""Debug probe: run MMQ on a few shapes and print error statistics against the operand-exact oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import oracle, ggml
from util import mmq_b_operand_fp16

cases = [(12, 128, 256, 16), (12, 128, 256, 128), (12, 256, 1024, 83), (12, 384, 2048, 300), (14, 256, 1024, 83), (8, 256, 1024, 40), (12, 1024, 4096, 512)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for t, N, K, M in cases:
    w = oracle.random_weights(t, N, K, seed=1)
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(0))
    got = ggml.ggml_mul_mat_a8(torch.from_numpy(w).cuda(), x.cuda(), t, N)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    a16 = oracle.mmq_a_operand(w, t, N, K).astype(np.float64)
    b16 = mmq_b_operand_fp16(x.numpy(), oracle.mmq_need_sum(t)).astype(np.float64)
    exact = b16 @ a16.T
    err = np.abs(got - exact)
    scale = np.abs(exact).max()
    print(f"type {t} N={N} K={K} M={M}: max err {err.max():.3e} scale {scale:.3e} rel {err.max()/scale:.2e} nan {np.isnan(got).sum()}", flush=True)
    if err.max() / scale > 1e-3:
        bad = np.argwhere(err > 1e-3 * scale)
        print("   first bad (tok,row):", bad[:5].tolist(), "count", len(bad), "of", err.size)
        print("   got", got[tuple(bad[0])], "want", exact[tuple(bad[0])])
        rows = np.unique(bad[:, 1]); toks = np.unique(bad[:, 0])
        print("   bad rows", rows[:10], "... n", len(rows), " bad toks", toks[:10], "... n", len(toks))
