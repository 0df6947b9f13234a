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


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("shape", [(14336, 4096), (4096, 14336), (4096, 4096)], ids=["14336x4096", "4096x14336", "4096x4096"])
@pytest.mark.parametrize("M", [512, 1024, 4096, 8192])
@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K], ids=["Q4_K", "Q6_K"])
def test_config4_every_benchmarked_shape(ops, t, M, shape, dtype):
    """Every (type, M, shape, dtype) bench_extras times (BASELINE config 4): sampled 64 x 64 blocks against the oracle --
    first / middle / last row blocks x first / last token tiles, which covers every tile-picker branch (one or two
    256-token tiles per weight stage, clusters, CTA pairs) and the rectangle a tail launch covers."""
    N, K = shape
    w_np = oracle.random_weights(t, N, K, seed=M + N)
    w = torch.from_numpy(w_np).cuda()
    x = (torch.randn(M, K, generator=torch.Generator().manual_seed(M)) * 0.5).to(dtype).cuda()
    y = ops.ggml_mul_mat_a8(w, x, t, N)
    assert y.shape == (M, N) and bool(torch.isfinite(y).all())
    ulp = OUT_REL[dtype]
    for r0 in (0, (N // 2 // 64) * 64 + 64, N - 64, N - 320):
        for t0 in (0, M - 64, M // 2):
            rows, toks = slice(r0, r0 + 64), slice(t0, t0 + 64)
            want = oracle.mmq(w_np[rows], t, x[toks].float().cpu().numpy(), 64)
            got = y[toks, rows].float().cpu().numpy()
            scale = np.abs(want).max()
            err = np.abs(got - want)
            assert (err <= 2.0 ** -9 * scale + 1.01 * ulp * np.abs(want)).all(), \
                f"{oracle.TYPE_NAMES[t]} M={M} {N}x{K} {dtype} rows {r0} tokens {t0}: max err {err.max():.3e} scale {scale:.3e}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K, oracle.Q8_0, oracle.Q5_0, oracle.Q2_K], ids=["Q4_K", "Q6_K", "Q8_0", "Q5_0", "Q2_K"])
def test_activation_magnitude_range(ops, t, dtype):
    """The reference keeps int8 quants and fp32 scales, so it takes activations far outside fp16's range; the kernel's B
    operand is fp16(d8 * q8), stored per token divided by a power of two when the token's amax is outside [2^-8, 2^15] and
    multiplied back in the epilogue (exact).  Token rows of wildly different magnitudes in ONE call: every row within the
    usual tolerance of ITS OWN scale, nothing non-finite.  (Round-1 advisor finding: 1e5 used to give inf / NaN rows.)
    Q4_1 / Q5_1 are left out: the reference multiplies (d, m) by (d8, s8) in half precision (mmq.cuh:528), which is only
    good to ~0.5 % as soon as d * d8 is an fp16 subnormal -- there is no accurate reference value to compare with."""
    N, K = 384, 2048
    mags = [1e5, 3.0e38 ** 0.5, 1.0, 1e-6, 1e-12, 7e4, 2.0 ** -9, 0.0, 6.0e4, 1e-30]
    if oracle.mmq_need_sum(t):
        # these types carry half2(d8, sum x) in the reference's q8_1_mmq blocks (mmq.cu:149-153): its own range ends where
        # the fp16 sum of 32 inputs overflows (|x| ~ 3e3) and where d8 = amax / 127 becomes an fp16 subnormal (|x| ~ 1e-2),
        # so the comparison stays inside (the kernel rounds d8 to fp16 like the reference, its minima use d8 * sum(q8))
        mags = [3e3, 1.0, 0.05, 2.0 ** -4, 0.0, 1e3, 30.0, 0.5]
    M = 83
    w_np = oracle.random_weights(t, N, K, seed=5)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g)
    for i in range(M):
        x[i] *= mags[i % len(mags)]
    x = x.to(dtype)
    x32 = x.float().numpy()
    got = ops.ggml_mul_mat_a8(torch.from_numpy(w_np).cuda(), x.cuda(), t, N).float().cpu().numpy()
    want = oracle.mmq(w_np, t, x32, N)
    assert np.isfinite(want).all()
    assert np.isfinite(got).all(), f"non-finite outputs in token rows {sorted(set(np.argwhere(~np.isfinite(got))[:, 0].tolist()))}"
    for i in range(M):
        scale = np.abs(want[i]).max()
        err = np.abs(got[i] - want[i])
        tol = 2.0 ** -9 * scale + 1.01 * OUT_REL[dtype] * np.abs(want[i]) + 1e-44
        assert (err <= tol).all(), f"{oracle.TYPE_NAMES[t]} {dtype} token {i} (x ~ {mags[i % len(mags)]:.1e}): err {err.max():.3e} scale {scale:.3e}"


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


@pytest.mark.parametrize("case", [(oracle.Q4_K, 384, 2048, 83), (oracle.Q6_K, 256, 2048, 600), (oracle.Q4_0, 200, 1024, 128)],
                         ids=["Q4_K-M83", "Q6_K-M600", "Q4_0-M128"])
def test_static_weights_option_gives_the_same_bits(ops, case):
    """GGB_OPT_STATIC_WEIGHTS = 1: the weight tiles of a launch start loading before the previous kernel has finished; same bits."""
    t, N, K, M = case
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=M)).cuda()
    xs = [(torch.randn(M, K, generator=torch.Generator().manual_seed(i)) * 0.25).to(torch.float16).cuda() for i in range(3)]
    want = [ops.ggml_mul_mat_a8(w, x, t, N) for x in xs]
    torch.cuda.synchronize()
    ops.set_option(ops.GGB_OPT_STATIC_WEIGHTS, 1)
    try:
        got = [ops.ggml_mul_mat_a8(w, x, t, N) for x in xs]
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.GGB_OPT_STATIC_WEIGHTS, 0)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
