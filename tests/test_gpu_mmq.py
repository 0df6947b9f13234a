"""GPU parity: ggml_mul_mat_a8 (MMQ on tcgen05) against the oracle.

Two checks per case:
  (a) reference semantics -- oracle.mmq restates quantize_mmq_q8_1 + mul_mat_q (mmq.cu:109-255,
      mmq.cuh vec_dot_*_q8_1_mma) with exact integer dots.  The kernel feeds the tensor cores fp16
      operands (A = fp16(dequant W), B = fp16(d8*q8)), i.e. one extra 2^-11 relative rounding per
      operand: stated tolerance |y - oracle| <= 2^-9 * max|oracle| (+ the output dtype's rounding).
  (b) operand-exact -- the same fp16 operands built on the CPU from the oracle's dequantize and
      q8_1_mmq blocks, multiplied in float64: only the fp32 accumulation order differs, so
      |y - exact| <= 1e-4 * max|exact|.  This pins the unpack path and the q8_1 quantization inside
      MMQ bit-for-bit (any wrong weight, scale or quant shows up far above 1e-4).
Plus the reference's own MMQ tests (test_cuda_kernels.py:84-183) with their tolerances."""
import numpy as np
import pytest
import torch

import oracle
from util import mmq_b_operand_fp16

pytestmark = pytest.mark.gpu

TYPES = oracle.ALL_TYPES
IDS = [oracle.TYPE_NAMES[t] for t in TYPES]
OUT_REL = {torch.float32: 0.0, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _check(ops, t, N, K, M, dtype, seed=0, dist="randn"):
    w = oracle.random_weights(t, N, K, seed=seed + t)
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(M, K, generator=g) if dist == "randn" else torch.rand(M, K, generator=g)).to(dtype)
    x32 = x.float().numpy()
    got = ops.ggml_mul_mat_a8(torch.from_numpy(w).cuda(), x.cuda(), t, N)
    assert got.shape == (M, N) and got.dtype == dtype
    got = got.float().cpu().numpy()
    name = f"{oracle.TYPE_NAMES[t]} N={N} K={K} M={M} {dtype}"
    # (b) operand-exact
    a16 = oracle.mmq_a_operand(w, t, N, K).astype(np.float64)
    b16 = mmq_b_operand_fp16(x32, oracle.mmq_need_sum(t)).astype(np.float64)
    exact = b16 @ a16.T
    scale = float(np.abs(exact).max()) + 1e-30
    tol = 1e-4 * scale + OUT_REL[dtype] * np.abs(exact) * 1.01
    bad = np.abs(got - exact) > tol
    assert not bad.any(), f"{name} vs operand-exact: {bad.sum()} of {bad.size} off; max err {np.abs(got - exact).max():.3e} scale {scale:.3e}"
    # (a) reference semantics
    want = oracle.mmq(w, t, x32, N)
    tol = 2.0 ** -9 * scale + OUT_REL[dtype] * np.abs(want) * 1.01
    bad = np.abs(got - want) > tol
    assert not bad.any(), f"{name} vs oracle.mmq: {bad.sum()} of {bad.size} off; max err {np.abs(got - want).max():.3e} scale {scale:.3e}"


@pytest.mark.parametrize("M", [1, 7, 16, 17, 33, 83, 128, 129, 300, 512])
@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K, oracle.Q8_0], ids=["Q4_K", "Q6_K", "Q8_0"])
def test_token_counts(ops, t, M):
    _check(ops, t, 256, 1024, M, torch.float32, seed=M)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
@pytest.mark.parametrize("shape", [(128, 256), (200, 512), (384, 2048), (130, 4096)])
def test_types_shapes(ops, t, shape, dtype):
    N, K = shape
    _check(ops, t, N, K, 83, dtype, seed=N)


@pytest.mark.parametrize("t", oracle.LEGACY_TYPES, ids=[oracle.TYPE_NAMES[t] for t in oracle.LEGACY_TYPES])
def test_legacy_ragged_k(ops, t):
    """K % 32 == 0 only: partial raw stages, rows that are not 16-byte aligned."""
    for N, K in [(40, 32), (64, 96), (129, 2080), (16, 288)]:
        _check(ops, t, N, K, 20, torch.float32, seed=K)


def test_many_tiles_persistent(ops):
    """more tiles than SMs: every role walks several tiles, both TMEM accumulator buffers recycle."""
    _check(ops, oracle.Q4_K, 148 * 128 + 64, 512, 520, torch.float16, seed=3)


def test_3d_input(ops):
    """X [B, T, K] is flattened to B*T rows (mmq.cu:201-204)."""
    t, N, K = oracle.Q4_K, 256, 1024
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=1)).cuda()
    x = torch.randn(4, 33, K, device="cuda", dtype=torch.float16)
    y3 = ops.ggml_mul_mat_a8(w, x, t, N)
    y2 = ops.ggml_mul_mat_a8(w, x.reshape(-1, K), t, N)
    assert y3.shape == (4, 33, N)
    assert torch.equal(y3.reshape(-1, N), y2)


NUM_TOKENS = [7, 83, 128, 2048]


@pytest.mark.parametrize("num_tokens", NUM_TOKENS)
@pytest.mark.parametrize("hidden_size", [256, 1024])
@pytest.mark.parametrize("dtype", [torch.half, torch.bfloat16, torch.float32], ids=["half", "bf16", "f32"])
@pytest.mark.parametrize("t", [10, 11, 12, 13, 14, 2, 6, 8], ids=["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K", "Q4_0", "Q5_0", "Q8_0"])
def test_reference_style_mmq(ops, t, dtype, hidden_size, num_tokens):
    """test_cuda_kernels.py:84-129."""
    torch.manual_seed(0)
    x = torch.rand((num_tokens, hidden_size), dtype=dtype, device="cuda")
    atols = {torch.half: 1, torch.bfloat16: 1.5, torch.float: 1.2}
    rtols = {torch.half: 1e-1, torch.bfloat16: 1e4, torch.float: 2e1}
    for rows in (hidden_size, 2 * hidden_size):
        w = oracle.random_weights(t, rows, hidden_size, seed=rows)
        weight = torch.tensor(oracle.dequantize(w, t, rows, hidden_size), device="cuda").to(dtype)
        ref_output = x @ weight.T
        qweight = torch.tensor(w, device="cuda")
        output = ops.ggml_mul_mat_a8(qweight, x, t, qweight.shape[0]).to(dtype)
        torch.testing.assert_close(output, ref_output, atol=atols[dtype], rtol=rtols[dtype])


@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K], ids=["Q4_K", "Q6_K"])
def test_config4_full_shape_properties(ops, t):
    """BASELINE config 4 at full size (14336 x 4096, M = 2048): too big for the CPU oracle in seconds, so
    size-independent properties -- token rows are independent (a slice of X gives the slice of Y, bit for
    bit), weight rows are independent, and a sampled block matches the oracle."""
    N, K, M = 14336, 4096, 2048
    w_np = oracle.random_weights(t, N, K, seed=2)
    w = torch.from_numpy(w_np).cuda()
    x = torch.randn(M, K, device="cuda", dtype=torch.float16)
    y = ops.ggml_mul_mat_a8(w, x, t, N)
    y_tok = ops.ggml_mul_mat_a8(w, x[512:768].contiguous(), t, N)
    assert torch.equal(y[512:768], y_tok)
    y_row = ops.ggml_mul_mat_a8(w[1024:1536].contiguous(), x, t, 512)
    assert torch.equal(y[:, 1024:1536], y_row)
    # sampled blocks against the oracle: one from the main launch, one from the rectangle the tail launch covers
    # (224 cluster tiles on 74 clusters: the last 2 = row blocks 108..111 x tokens 1536..2047 run as 128 x 64 tiles)
    for rows, toks in ((slice(3000, 3064), slice(0, 64)), (slice(14000, 14064), slice(1984, 2048)),
                       (slice(13824, 13888), slice(1536, 1600))):
        want = oracle.mmq(w_np[rows], t, x[toks].float().cpu().numpy(), 64)
        got = y[toks, rows].float().cpu().numpy()
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 2.0 ** -9 * scale + 2.0 ** -10 * np.abs(want).max()


def test_errors(ops):
    w = torch.zeros(144 * 4, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.ggml_mul_mat_a8(w, torch.zeros(3, 512, device="cuda"), 12, 4)
    with pytest.raises(ValueError):
        ops.ggml_mul_mat_a8(w, torch.zeros(256, device="cuda"), 12, 4)


@pytest.mark.parametrize("t", oracle.IQ_TYPES, ids=[oracle.TYPE_NAMES[t] for t in oracle.IQ_TYPES])
def test_iq4_has_no_mmq(ops, t):
    """IQ4_NL / IQ4_XS: dequantize + MMVQ only.  The reference's ggml_mul_mat_a8 has no case for them either and
    returns uninitialised memory (mmq.cu:222-252); here the call fails loudly."""
    w = torch.from_numpy(oracle.random_weights(t, 4, 512, seed=0)).cuda()
    with pytest.raises(ValueError):
        ops.ggml_mul_mat_a8(w, torch.zeros(16, 512, device="cuda"), t, 4)
