"""GPU parity: ggml_mul_mat_vec_a8 (MMVQ, in-kernel q8_1 quantization) against the oracle's restatement
of quantize_q8_1 + mul_mat_vec_q (ggml_kernel.cu:13-66, mmvq.cuh:2-38, vecdotq.cuh) and, like the
reference's own test (test_cuda_kernels.py:61-81), against x @ dequant(W).T.

Stated tolerance: the integer dot products are exact and the q8_1 quantization is bit-identical to
the oracle's, so the only difference is fp32 summation order (the oracle accumulates in double):
|y - oracle| <= 2e-5 * max|oracle| for fp32 outputs, plus one output rounding for fp16 (2^-10
relative) / bf16 (2^-7 relative)."""
import numpy as np
import pytest
import torch

import oracle
from util import MMVQ_ATOL, MMVQ_RTOL

pytestmark = pytest.mark.gpu

TYPES = oracle.VEC_TYPES
IDS = [oracle.TYPE_NAMES[t] for t in TYPES]
DTYPES = [torch.float32, torch.float16, torch.bfloat16]
OUT_REL = {torch.float32: 0.0, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _check(ops, t, N, K, B, dtype, seed=0, dist="randn"):
    w = oracle.random_weights(t, N, K, seed=seed + t)
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, K, generator=g) if dist == "randn" else torch.rand(B, K, generator=g)).to(dtype)
    want = oracle.mmvq(w, t, x.float().numpy(), N)
    got = ops.ggml_mul_mat_vec_a8(torch.from_numpy(w).cuda(), x.cuda(), t, N)
    assert got.shape == (B, N) and got.dtype == dtype
    got = got.float().cpu().numpy()
    scale = float(np.abs(want).max()) + 1e-30
    tol = 2e-5 * scale + OUT_REL[dtype] * np.abs(want) * 1.01
    bad = np.abs(got - want) > tol
    assert not bad.any(), (f"{oracle.TYPE_NAMES[t]} N={N} K={K} B={B} {dtype}: {bad.sum()} of {bad.size} off; "
                           f"max err {np.abs(got - want).max():.3e} scale {scale:.3e}")
    return w, x, got


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_batches_fp32(ops, t, B):
    _check(ops, t, 257, 2048, B, torch.float32)


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
@pytest.mark.parametrize("shape", [(1, 256), (5, 512), (96, 1024), (300, 4096), (64, 6144), (33, 2304)])
def test_shapes(ops, t, shape, dtype):
    N, K = shape
    _check(ops, t, N, K, 1, dtype, seed=N)
    _check(ops, t, N, K, 3, dtype, seed=N + 1)


@pytest.mark.parametrize("t", oracle.BLOCK32_TYPES, ids=[oracle.TYPE_NAMES[t] for t in oracle.BLOCK32_TYPES])
def test_legacy_k_multiple_of_32(ops, t):
    """legacy types only need K % 32 == 0 (rows then are not 16-byte aligned: the non-TMA path)."""
    for N, K in [(17, 32), (40, 96), (9, 2080)]:
        _check(ops, t, N, K, 2, torch.float32, seed=K)


@pytest.mark.parametrize("hidden_size", [256, 1024])
@pytest.mark.parametrize("dtype", [torch.half, torch.bfloat16, torch.float32], ids=["half", "bf16", "f32"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_reference_style_mmvq(ops, t, dtype, hidden_size):
    """test_cuda_kernels.py:61-81: x = torch.rand(1, H), expected = x @ dequant(W).T, atol 1, rtol 1e-1."""
    torch.manual_seed(0)
    x = torch.rand((1, hidden_size), dtype=dtype, device="cuda")
    for rows in (hidden_size, 2 * hidden_size):
        w = oracle.random_weights(t, rows, hidden_size, seed=rows)
        weight = torch.tensor(oracle.dequantize(w, t, rows, hidden_size), device="cuda").to(dtype)
        ref_output = x @ weight.T
        qweight = torch.tensor(w, device="cuda")
        output = ops.ggml_mul_mat_vec_a8(qweight, x, t, qweight.shape[0]).to(dtype)
        torch.testing.assert_close(output, ref_output, atol=MMVQ_ATOL, rtol=MMVQ_RTOL)


def _check_sampled(ops, t, N, K, B, dtype, seed=0, nsample=192):
    """Full-size GEMV on the GPU, oracle on a sample of rows (rows are independent: a row's value does not depend on
    which rows share the launch -- test_linearity_property): first / last rows of the matrix and rows around every
    boundary a CTA's row range can have, plus random ones."""
    w = oracle.random_weights(t, N, K, seed=seed + t)
    x = torch.randn(B, K, generator=torch.Generator().manual_seed(seed)).to(dtype)
    got = ops.ggml_mul_mat_vec_a8(torch.from_numpy(w).cuda(), x.cuda(), t, N)
    assert got.shape == (B, N) and got.dtype == dtype and bool(torch.isfinite(got).all())
    rng = np.random.default_rng(seed)
    rows = np.unique(np.concatenate([np.arange(min(16, N)), np.arange(max(0, N - 16), N),
                                     (np.arange(1, 148)[:, None] * N // 148 + np.arange(-1, 1)[None, :]).ravel() % N,
                                     rng.integers(0, N, size=nsample)]))
    want = oracle.mmvq(w[rows], t, x.float().numpy(), len(rows))
    g = got.float().cpu().numpy()[:, rows]
    scale = float(np.abs(want).max()) + 1e-30
    tol = 2e-5 * scale + OUT_REL[dtype] * np.abs(want) * 1.01
    bad = np.abs(g - want) > tol
    assert not bad.any(), (f"{oracle.TYPE_NAMES[t]} N={N} K={K} B={B} {dtype}: {bad.sum()} of {bad.size} sampled values off; "
                           f"max err {np.abs(g - want).max():.3e} scale {scale:.3e}")


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K], ids=["Q4_K", "Q6_K"])
@pytest.mark.parametrize("shape", [(14336, 4096), (4096, 14336)], ids=["14336x4096", "4096x14336"])
def test_config3_llama8b_shapes(ops, t, shape, B):
    """BASELINE config 3: Llama-3-8B Q4_K_M decode GEMV shapes at full size, EVERY batch 1..8 (each batch can take a
    different kernel: register-resident batch 1, dp4a ring batch 2, tensor-core batches 3-8 / 2-8 for K > 8192), fp16."""
    N, K = shape
    _check_sampled(ops, t, N, K, B, torch.float16, seed=B)


@pytest.mark.parametrize("B", [1, 8])
@pytest.mark.parametrize("tp", [1, 2, 4, 8])
@pytest.mark.parametrize("proj", ["gate_up_Q4_K", "down_Q4_K", "down_Q6_K"])
def test_config5_llama70b_shard_shapes(ops, proj, tp, B):
    """BASELINE config 5: the per-rank shard shapes of the Llama-3-70B MLP for tp = 1, 2, 4, 8 (gate/up [28672/tp, 8192],
    down [8192/tp, 28672]; the stacked gate|up shard of the fused block has 2 * 28672/tp rows), batch 1 and 8."""
    t, N, K = {"gate_up_Q4_K": (oracle.Q4_K, 2 * 28672 // tp, 8192), "down_Q4_K": (oracle.Q4_K, 8192 // tp, 28672),
               "down_Q6_K": (oracle.Q6_K, 8192 // tp, 28672)}[proj]
    _check_sampled(ops, t, N, K, B, torch.float16, seed=tp + B)


@pytest.mark.parametrize("case", [(oracle.Q4_K, 130, 28672, 8), (oracle.Q6_K, 70, 14336, 8), (oracle.Q6_K, 40, 28672, 6),
                                  (oracle.Q5_K, 50, 28672, 7), (oracle.Q4_K, 33, 14336, 2)],
                         ids=["Q4_K-K28672-B8", "Q6_K-K14336-B8", "Q6_K-K28672-B6", "Q5_K-K28672-B7", "Q4_K-K14336-B2"])
def test_long_rows_k_phases(ops, case):
    """long rows x large batches: the q8_1 activation image does not fit in shared memory next to the weight ring, so
    the tensor-core kernel walks the row in K phases (and batch 2 takes that kernel only for K > 8192)."""
    t, N, K, B = case
    _check(ops, t, N, K, B, torch.float16, seed=B)
    _check(ops, t, N, K, B, torch.float32, seed=B + 1)


@pytest.mark.parametrize("case", [(oracle.Q4_K, 1), (oracle.Q4_K, 2), (oracle.Q4_K, 5), (oracle.Q6_K, 8), (oracle.Q5_K, 3),
                                  (oracle.IQ4_XS, 2), (oracle.IQ4_XS, 6)],
                         ids=["Q4_K-B1", "Q4_K-B2", "Q4_K-B5", "Q6_K-B8", "Q5_K-B3", "IQ4_XS-B2", "IQ4_XS-B6"])
def test_gather_epilogue_single_gpu(ops, case):
    """ggb_mmvq_gather with two 'peer' buffers on one GPU: every result lands at [tok, col_offset + row] of BOTH
    buffers, bit-identical to the plain op; nothing outside the slice is touched."""
    t, B = case
    N, K, ld, col0 = 264, 2048, 1000, 304
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=t)).cuda()
    x = torch.randn(B, K, generator=torch.Generator().manual_seed(B)).to(torch.float16).cuda()
    want = ops.ggml_mul_mat_vec_a8(w, x, t, N)
    bufs = [torch.full((8, ld), 7.0, dtype=torch.float16, device="cuda") for _ in range(2)]
    ops.mul_mat_vec_a8_gather(w, x, t, N, [b.data_ptr() for b in bufs], ld, col0)
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b[:B, col0:col0 + N], want)
        assert bool((b[:B, :col0] == 7.0).all()) and bool((b[:B, col0 + N:] == 7.0).all()) and bool((b[B:] == 7.0).all())
    with pytest.raises(ValueError):          # Q8_0 is not on the fused path
        w8 = torch.from_numpy(oracle.random_weights(oracle.Q8_0, 8, 256, seed=0)).cuda()
        ops.mul_mat_vec_a8_gather(w8, x[:, :256].contiguous(), oracle.Q8_0, 8, [bufs[0].data_ptr()], ld, 0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32], ids=["f16", "bf16", "f32"])
def test_fused_swiglu_mlp_block_single_gpu(ops, dtype):
    """the MLP block as two launches (stacked gate|up with split output columns, then a down
    projection that computes silu(g) * u inside its activation quantization) equals the three-matmul + eager
    SwiGLU path bit for bit, for every decode batch."""
    from ggml.parallel import ColumnParallelQuantLinear, FusedQuantMLP, QuantMLP
    hidden, inter = 2048, 4096
    dev = torch.device("cuda")
    mk = lambda t, n, k, seed: ColumnParallelQuantLinear(torch.from_numpy(oracle.random_weights(t, n, k, seed=seed)).to(dev), t, n, 0, 1)
    gate, up, down = mk(oracle.Q4_K, inter, hidden, 1), mk(oracle.Q4_K, inter, hidden, 2), mk(oracle.Q6_K, hidden, inter, 3)
    plain, fused = QuantMLP(gate, up, down), FusedQuantMLP(gate, up, down)
    for M in (1, 2, 3, 5, 8):
        x = (torch.randn(M, hidden, generator=torch.Generator().manual_seed(M)) * 0.25).to(dtype).to(dev)
        want = plain(x)
        for _ in range(2):                                  # both buffer sets
            got = fused(x)
            assert got.shape == want.shape and got.dtype == want.dtype
            assert torch.equal(got, want), f"M={M} {dtype}: max diff {(got.float() - want.float()).abs().max().item()}"
    assert fused.fused, "the fused path fell back"
    # the SwiGLU activation alone, against the oracle on silu(g) * u computed by torch
    g = torch.randn(3, inter, generator=torch.Generator().manual_seed(7)).to(dtype).to(dev)
    u = torch.randn(3, inter, generator=torch.Generator().manual_seed(8)).to(dtype).to(dev)
    h = torch.nn.functional.silu(g) * u
    y = torch.empty(3, hidden, dtype=dtype, device=dev)
    ops.mul_mat_vec_a8_fused(down.W, g, u, oracle.Q6_K, hidden, [y.data_ptr()], hidden, 0)
    assert torch.equal(y, ops.ggml_mul_mat_vec_a8(down.W, h, oracle.Q6_K, hidden))


@pytest.mark.parametrize("case", [(oracle.Q4_K, 28672, 8), (oracle.Q6_K, 14336, 8), (oracle.Q4_K, 28672, 6)], ids=["Q4_K-K28672-B8", "Q6_K-K14336-B8", "Q4_K-K28672-B6"])
def test_phased_walk_is_independent_of_n(ops, case):
    """round-1 advisor finding: in the K-phased walk of the tensor-core kernel the plan (phases, warps, chunk, team) -- and with it
    the fp32 summation order -- must not depend on N or the grid: a row shard gives the bits of the full matrix."""
    t, K, B = case
    N = 148 * 16 * 3 + 40
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=B)).cuda()
    x = torch.randn(B, K, generator=torch.Generator().manual_seed(1)).to(torch.float16).cuda()
    full = ops.ggml_mul_mat_vec_a8(w, x, t, N)
    for lo, hi in ((0, N // 8), (N // 2, N // 2 + 100), (N - 300, N)):
        part = ops.ggml_mul_mat_vec_a8(w[lo:hi].contiguous(), x, t, hi - lo)
        assert torch.equal(full[:, lo:hi], part), f"rows {lo}:{hi}"


K_QUANTS = (oracle.Q2_K, oracle.Q3_K, oracle.Q4_K, oracle.Q5_K, oracle.Q6_K)


@pytest.mark.parametrize("K", [2048, 6144, 8192, 12288, 14336, 16384, 24576, 28672])
@pytest.mark.parametrize("t", K_QUANTS, ids=[oracle.TYPE_NAMES[t] for t in K_QUANTS])
def test_register_kernel_row_lengths(ops, t, K):
    """batch 1 of every K-quant on the register-resident kernel (mmvq_rg.cu): one chunk, whole and ragged chunks (K = 6144: 16 + 8
    super-blocks; 14336: 16 + 16 + 16 + 8), two to seven chunks per row (12, 14 and 7 warps), the pre-launched activation image of
    K = 28672; N chosen so that CTAs own 0, 1 and several rows."""
    for N, dtype in ((77, torch.float16), (300, torch.float32)):
        _check(ops, t, N, K, 1, dtype, seed=K // 256 + N)


@pytest.mark.parametrize("t", (oracle.Q4_K, oracle.Q5_K, oracle.Q6_K), ids=["Q4_K", "Q5_K", "Q6_K"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_fused_swiglu_70b_down_projection(ops, t, dtype):
    """the 70B down projection's row length (K = 28672) with the fused SwiGLU activation: 7 warps behind the 128-thread activation
    pre-launch (mmvq_rg_prequant_kernel) -- bit-identical to the plain op on silu(g) * u computed by torch, batch 1 and 2."""
    inter, hidden = 28672, 264
    w = torch.from_numpy(oracle.random_weights(t, hidden, inter, seed=t)).cuda()
    for B in (1, 2):
        g = (torch.randn(B, inter, generator=torch.Generator().manual_seed(3 + B)) * 0.5).to(dtype).cuda()
        u = (torch.randn(B, inter, generator=torch.Generator().manual_seed(5 + B)) * 0.5).to(dtype).cuda()
        h = torch.nn.functional.silu(g) * u
        y = torch.full((B, hidden), 7.0, dtype=dtype, device="cuda")
        ops.mul_mat_vec_a8_fused(w, g, u, t, hidden, [y.data_ptr()], hidden, 0)
        want = ops.ggml_mul_mat_vec_a8(w, h, t, hidden)
        assert torch.equal(y, want), f"{oracle.TYPE_NAMES[t]} B={B} {dtype}: max diff {(y.float() - want.float()).abs().max().item()}"
        _ = oracle  # the plain op itself is checked against the oracle at this row length in test_register_kernel_row_lengths


@pytest.mark.parametrize("case", [(oracle.Q4_K, 300, 4096, 1), (oracle.Q4_K, 200, 28672, 1), (oracle.Q6_K, 200, 28672, 1), (oracle.Q5_K, 150, 14336, 1),
                                  (oracle.Q3_K, 300, 8192, 1), (oracle.Q2_K, 300, 4096, 1), (oracle.Q4_0, 300, 4096, 1), (oracle.IQ4_XS, 300, 4096, 2),
                                  (oracle.Q4_K, 300, 4096, 5), (oracle.Q4_0, 300, 4096, 8), (oracle.Q2_K, 300, 4096, 4), (oracle.Q6_K, 70, 14336, 8),
                                  (oracle.IQ2_XXS, 300, 4096, 1), (oracle.Q4_K, 300, 4096, 2)],
                         ids=lambda c: f"{oracle.TYPE_NAMES[c[0]]}-{c[1]}x{c[2]}-B{c[3]}")
def test_static_weights_option_gives_the_same_bits(ops, case):
    """GGB_OPT_STATIC_WEIGHTS = 1 (what bench.py runs with) lets every kernel start its weight stream before the previous kernel of
    the stream has finished.  It must not change a single bit: same kernels, same summation order.  Launches go back to back on
    one stream so that the programmatic-dependent overlap really happens (the output of one call is not an input of the next,
    exactly the situation the option is declared for)."""
    t, N, K, B = case
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=N + B)).cuda()
    xs = [torch.randn(B, K, generator=torch.Generator().manual_seed(i)).to(torch.float16).cuda() for i in range(4)]
    want = [ops.ggml_mul_mat_vec_a8(w, x, t, N) for x in xs]
    torch.cuda.synchronize()
    ops.set_option(ops.GGB_OPT_STATIC_WEIGHTS, 1)
    try:
        got = [ops.ggml_mul_mat_vec_a8(w, x, t, N) for x in xs]
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.GGB_OPT_STATIC_WEIGHTS, 0)
    for i, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a, b), f"launch {i}: max diff {(a.float() - b.float()).abs().max().item()}"


def test_linearity_property(ops):
    """size-independent property at full size: rows of Y are independent, so a row-slice of W gives the
    same values as the slice of the full result (bit-for-bit: same K order per row)."""
    t, N, K = oracle.Q4_K, 14336, 4096
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=5)).cuda()
    x = torch.randn(4, K, dtype=torch.float16, device="cuda")
    full = ops.ggml_mul_mat_vec_a8(w, x, t, N)
    part = ops.ggml_mul_mat_vec_a8(w[1000:3000].contiguous(), x, t, 2000)
    assert torch.equal(full[:, 1000:3000], part)
    # each batch row is independent of the rest of the batch: bit-for-bit inside one kernel family (batches
    # >= 3 take the tensor-core kernel, 1-2 the dp4a kernel), one fp16 ulp across the two (fp32 summation order)
    sub = ops.ggml_mul_mat_vec_a8(w, x[1:4].contiguous(), t, N)
    assert torch.equal(full[1:4], sub)
    # batch 1 (register-resident activations, mmvq_rg.cu), batch 2 (dp4a ring kernel) and batches >= 3 (tensor-core kernel)
    # are different kernels with different fp32 summation orders: equal within one fp16 ulp
    two = ops.ggml_mul_mat_vec_a8(w, x[2:4].contiguous(), t, N)
    one = ops.ggml_mul_mat_vec_a8(w, x[2:3].contiguous(), t, N)
    tol = 2e-5 * float(one.float().abs().max())
    torch.testing.assert_close(two[0:1].float(), one.float(), rtol=2e-3, atol=tol)
    torch.testing.assert_close(full[2:3].float(), one.float(), rtol=2e-3, atol=tol)
    # batch 1 is a pure function of (W row, x): the same row in two differently sized problems gives the same bits
    assert torch.equal(ops.ggml_mul_mat_vec_a8(w[1000:3000].contiguous(), x[2:3].contiguous(), t, 2000), one[:, 1000:3000])


def test_errors(ops):
    w = torch.zeros(144 * 4, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.ggml_mul_mat_vec_a8(w, torch.zeros(9, 256, device="cuda"), 12, 4)      # batch > 8
    with pytest.raises(ValueError):
        ops.ggml_mul_mat_vec_a8(w, torch.zeros(1, 512, device="cuda"), 12, 4)      # K mismatch
    with pytest.raises(ValueError):
        ops.ggml_mul_mat_vec_a8(w, torch.zeros(1, 256, device="cuda", dtype=torch.float64), 12, 4)
