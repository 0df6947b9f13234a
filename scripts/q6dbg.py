import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ggml-libtorch_b200")]
import numpy as np, torch
import oracle, ggml
np.set_printoptions(linewidth=200, precision=4, suppress=True)
for (N, K, B) in [(16, 256, 3), (16, 512, 3), (16, 2048, 3)]:
    t = 14
    w = oracle.random_weights(t, N, K, seed=1)
    x = torch.randn(B, K, generator=torch.Generator().manual_seed(0))
    want = oracle.mmvq(w, t, x.numpy(), N)
    got = ggml.ggml_mul_mat_vec_a8(torch.from_numpy(w).cuda(), x.cuda(), t, N).cpu().numpy()
    print(N, K, B, "max err", np.abs(got - want).max())
    print("want", want[:, :8]); print("got ", got[:, :8]); print("ratio", (got / want)[:, :8])
