"""GPU parity against THE REFERENCE'S OWN CUDA KERNELS, compiled unmodified for sm_100a (oracle/build_ref_cuda.py ->
oracle/_ref_cuda/_ggml.so = torch.ops._ggml.*, the ops of torch-ext/torch_binding.cpp:6-21) and run on the same
W / X on the same device.  This is what pins the steps the reference has no CPU implementation and no test for:

  * quantize_q8_1      (ggml/ggml_kernel.cu:13-66)  -- ggb_quantize_q8_1 layout 0 is BYTE-IDENTICAL to it
  * quantize_mmq_q8_1  (ggml/mmq.cu:84-177)         -- ggb_quantize_q8_1 layout 1 is BYTE-IDENTICAL to it
    (both through oracle/ref_cuda_q8_shim.cu, which includes the reference's translation unit to reach its
    file-static launchers), and the CPU oracle's restatement equals both byte for byte as well;
  * ggml_mul_mat_vec_a8 (quantize_q8_1 + mul_mat_vec_q, ggml_kernel.cu:80-193) for every type the reference serves;
  * ggml_mul_mat_a8     (quantize_mmq_q8_1 + mul_mat_q, mmq.cu:180-255) at the reference's own token counts
    (tests/kernels/test_cuda_kernels.py:14: 7, 83, 128, 2048) and types;
  * ggml_dequantize     (the reference CUDA op computes in half and is not bit-exact with its own CPU op; ours is
    bit-exact with the CPU op, so this one is a tolerance check: the reference test's atol 1e-2 / rtol 4e-2).

Stated tolerances (observed maxima are written to gpurun_out/ref_cuda_parity.json and kept in profiles/):
  MMVQ: integer dots and q8_1 bytes are identical on both sides; what differs is the fp32 summation order and, for
        fp16/bf16, one output rounding -> |ours - ref| <= 4e-5 * max|ref| + 1.01 ulp(out dtype).
  MMQ : ours multiplies fp16-rounded operands (2^-11 per operand) -> 2^-9 * max|ref| + 1.01 ulp(out),
        plus, for the weight types whose reference MMQ kernel takes the block minimum times q8_1's `s` field (the fp16
        sum of the UNQUANTIZED inputs: Q4_1, Q5_1, Q4_K, Q5_K; mmq.cuh:529, 1351-1361) where ours -- like the reference's
        own MMVQ for the K-quants (vecdotq.cuh:280-283) -- multiplies the minimum by d8 * sum(q8): the exactly computed
        bound  sum_g |min_w[i,g]| * |s8[j,g] - d8[j,g] * sum(q8[j,g])|  (declared deviation; it is part of
        q8_1's own quantization noise).  Q4_1 / Q5_1 additionally round d*d8 and m*s8 to half in the reference itself.
The whole module skips when the library was not built (no /root/reference at build time).
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from util import DEQUANT_ATOL, DEQUANT_RTOL

pytestmark = pytest.mark.gpu

REF_TYPES = oracle.ALL_TYPES                      # the ten types north_star names
VEC_TYPES = oracle.VEC_TYPES                      # + IQ4_NL / IQ4_XS (dequantize + MMVQ in the reference)
NAME = oracle.TYPE_NAMES
DTYPES = [torch.float32, torch.float16, torch.bfloat16]
DT_IDS = ["f32", "f16", "bf16"]
ULP = {torch.float32: 0.0, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}
_OBSERVED = {}


@pytest.fixture(scope="module")
def ref():
    r = oracle.ref_cuda_ops()
    if r is None:
        pytest.skip("oracle/_ref_cuda is not built (python oracle/build_ref_cuda.py where /root/reference exists)")
    yield r
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "ref_cuda_parity.json"), "w") as f:
            json.dump(_OBSERVED, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _note(key, err, scale):
    rel = float(err) / (float(scale) + 1e-30)
    _OBSERVED[key] = max(_OBSERVED.get(key, 0.0), rel)


def _x(rows, K, dtype, seed, dist="randn", mul=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, K, generator=g) if dist == "randn" else torch.rand(rows, K, generator=g)
    return (x * mul).to(dtype).cuda()


# ------------------------------------------------------------------------------------------------ q8_1 bytes

@pytest.mark.parametrize("dtype", DTYPES, ids=DT_IDS)
@pytest.mark.parametrize("shape", [(1, 256), (1, 4096), (3, 4000), (8, 14336), (5, 800), (130, 1024)])
def test_q8_1_bytes_equal_reference_kernel(ref, ops, shape, dtype):
    rows, K = shape
    for seed, dist, mul in ((0, "randn", 1.0), (1, "rand", 1.0), (2, "randn", 300.0), (3, "randn", 1e-4)):
        x = _x(rows, K, dtype, seed, dist, mul)
        if seed == 2:
            x[0, : min(64, K)] = 0                              # an all-zero group (amax == 0 branch)
        want = ref.quantize_q8_1(x)
        got = ops.ggml_quantize_q8_1(x)
        assert got.shape == want.shape
        assert torch.equal(got, want), f"q8_1 bytes differ from the reference kernel ({rows}x{K} {dtype} {dist}*{mul})"
        cpu = oracle.quantize_q8_1(x.float().cpu().numpy())
        assert np.array_equal(cpu, want.cpu().numpy()), "oracle restatement differs from the reference kernel"


@pytest.mark.parametrize("dtype", DTYPES, ids=DT_IDS)
@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K, oracle.Q4_0, oracle.Q5_0, oracle.Q8_0],
                         ids=["Q4_K", "Q6_K", "Q4_0", "Q5_0", "Q8_0"])
@pytest.mark.parametrize("shape", [(7, 256), (83, 1024), (128, 4096), (33, 4000), (2048, 512)])
def test_q8_1_mmq_bytes_equal_reference_kernel(ref, ops, shape, t, dtype):
    rows, K = shape
    x = _x(rows, K, dtype, rows + t, "randn", 1.0 if t != oracle.Q8_0 else 40.0)
    want = ref.quantize_q8_1_mmq(x, t)
    got = ops.ggml_quantize_q8_1(x, ops.GGB_Q8_1_MMQ, t)
    assert got.shape == want.shape
    assert torch.equal(got, want), f"q8_1_mmq bytes differ from the reference kernel ({rows}x{K} {dtype} {NAME[t]})"
    cpu = oracle.quantize_q8_1_mmq(x.float().cpu().numpy(), oracle.mmq_need_sum(t))
    assert np.array_equal(cpu, want.cpu().numpy()), "oracle restatement differs from the reference kernel"


# ------------------------------------------------------------------------------------------------ MMVQ

def _mmvq_check(ref, ops, t, N, K, dtype, seed, dist="randn"):
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=seed + t)).cuda()
    x = _x(1, K, dtype, seed, dist)
    want = ref.ggml_mul_mat_vec_a8(w, x, t, N).float()
    got = ops.ggml_mul_mat_vec_a8(w, x, t, N)
    assert got.shape == (1, N) and got.dtype == dtype
    got = got.float()
    scale = float(want.abs().max())
    err = (got - want).abs()
    tol = 4e-5 * scale + 1.01 * ULP[dtype] * want.abs()
    _note(f"mmvq/{NAME[t]}/{DT_IDS[DTYPES.index(dtype)]}", float(err.max()), scale)
    assert bool((err <= tol).all()), (f"MMVQ {NAME[t]} N={N} K={K} {dtype}: max |ours - reference kernel| = {float(err.max()):.3e}, "
                                      f"scale {scale:.3e}")
    return w, x, want


@pytest.mark.parametrize("dtype", DTYPES, ids=DT_IDS)
@pytest.mark.parametrize("t", VEC_TYPES, ids=[NAME[t] for t in VEC_TYPES])
def test_mmvq_vs_reference_kernel(ref, ops, t, dtype):
    """the reference's own MMVQ test shape family (hidden 256 / 1024, test_cuda_kernels.py:61-81) plus Llama rows"""
    for (N, K) in ((256, 256), (1024, 1024), (2048, 1024), (300, 4096), (96, 14336)):
        _mmvq_check(ref, ops, t, N, K, dtype, seed=N, dist="rand" if K <= 1024 else "randn")


@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K], ids=["Q4_K", "Q6_K"])
@pytest.mark.parametrize("shape", [(14336, 4096), (4096, 14336), (28672, 8192), (8192, 28672)])
def test_mmvq_vs_reference_kernel_full_shapes(ref, ops, t, shape):
    """BASELINE configs 3 and 5 at full size (8B and 70B MLP shapes), fp16, against the reference kernel; batches 2..8 of
    OUR op are checked row by row against the reference's batch-1 op (it has no larger batch)."""
    N, K = shape
    w, x, want1 = _mmvq_check(ref, ops, t, N, K, torch.float16, seed=7)
    xb = _x(8, K, torch.float16, 11)
    wants = torch.cat([ref.ggml_mul_mat_vec_a8(w, xb[b:b + 1].contiguous(), t, N) for b in range(8)]).float()
    for B in (2, 3, 4, 5, 8):
        got = ops.ggml_mul_mat_vec_a8(w, xb[:B].contiguous(), t, N).float()
        want = wants[:B]
        scale = float(want.abs().max())
        err = (got - want).abs()
        _note(f"mmvq_batch/{NAME[t]}/B{B}", float(err.max()), scale)
        assert bool((err <= 4e-5 * scale + 1.01 * ULP[torch.float16] * want.abs()).all()), \
            f"MMVQ {NAME[t]} {N}x{K} B={B}: max err {float(err.max()):.3e} scale {scale:.3e}"


# ------------------------------------------------------------------------------------------------ MMQ

# weight types whose reference MMQ uses min * s8 (float sum of the inputs); payload byte ranges to zero to get the minima
_MIN_PAYLOAD = {oracle.Q4_1: [(4, 20)], oracle.Q5_1: [(4, 24)], oracle.Q4_K: [(16, 144)], oracle.Q5_K: [(16, 176)]}


def _min_term_bound(ops, w_np, t, N, K, x2d):
    """[M, N] fp32: sum_g |min_w[i, g]| * |s8[j, g] - d8[j, g] * sum(q8[j, g])| from the q8_1_mmq blocks of x."""
    if t not in _MIN_PAYLOAD:
        return 0.0
    bs = oracle.block_bytes(t)
    w0 = w_np.reshape(-1, bs).copy()
    for lo, hi in _MIN_PAYLOAD[t]:
        w0[:, lo:hi] = 0
    minw = torch.from_numpy(np.abs(oracle.dequantize(w0.reshape(N, -1), t, N, K))[:, ::32].copy()).cuda()     # [N, K/32]
    blk = ops.ggml_quantize_q8_1(x2d, ops.GGB_Q8_1_MMQ, t)                       # [Kp/128, M, 144], ds = half2(d, s)
    ds = blk[:, :, :16].contiguous().view(torch.float16).float().view(blk.shape[0], blk.shape[1], 4, 2)
    sq = blk[:, :, 16:].contiguous().view(torch.int8).float().view(blk.shape[0], blk.shape[1], 4, 32).sum(-1)
    delta = (ds[..., 1] - ds[..., 0] * sq).abs().permute(1, 0, 2).reshape(blk.shape[1], -1)[:, :K // 32]   # [M, K/32]
    return delta @ minw.T


def _mmq_check(ref, ops, t, N, K, M, dtype, seed, x3d=None, mul=0.25):
    if t in (oracle.Q4_1, oracle.Q5_1):
        # the reference's MMQ multiplies (d, m) by (d8, s8) in HALF precision (mmq.cuh:528): with these synthetic weights
        # (d ~ 2^-10) and |x| ~ 0.25 the product d * d8 ~ 6e-6 is an fp16 SUBNORMAL and the reference itself is only good
        # to ~0.5 %.  Scale the activations so its half products stay normal; ours is insensitive to this.
        mul = 16.0
    w_np = oracle.random_weights(t, N, K, seed=seed + t)
    w = torch.from_numpy(w_np).cuda()
    x = _x(M, K, dtype, seed, "randn", mul) if x3d is None else x3d
    want = ref.ggml_mul_mat_a8(w, x, t, N).float()
    got = ops.ggml_mul_mat_a8(w, x, t, N)
    assert got.shape == want.shape and got.dtype == dtype
    got = got.float()
    assert bool(torch.isfinite(got).all()), "non-finite MMQ output"
    scale = float(want.abs().max())
    err = (got - want).abs()
    bound = _min_term_bound(ops, w_np, t, N, K, x.reshape(-1, K))
    if not isinstance(bound, float):
        bound = bound.view(want.shape)
    half_ref = 2.0 ** -10 * scale if t in (oracle.Q4_1, oracle.Q5_1) else 0.0      # the reference's own __hmul2 roundings
    tol = 2.0 ** -9 * scale + half_ref + 1.01 * ULP[dtype] * torch.maximum(want.abs(), got.abs()) + 1.05 * bound
    _note(f"mmq/{NAME[t]}/{DT_IDS[DTYPES.index(dtype)]}", float(err.max()), scale)
    _note(f"mmq_excess_over_min_bound/{NAME[t]}/{DT_IDS[DTYPES.index(dtype)]}", float((err - 1.05 * bound).clamp(min=0).max()), scale)
    assert bool((err <= tol).all()), (f"MMQ {NAME[t]} M={M} N={N} K={K} {dtype}: max |ours - reference kernel| = "
                                      f"{float(err.max()):.3e}, scale {scale:.3e}")


@pytest.mark.parametrize("dtype", DTYPES, ids=DT_IDS)
@pytest.mark.parametrize("t", REF_TYPES, ids=[NAME[t] for t in REF_TYPES])
@pytest.mark.parametrize("M", [7, 83, 128, 2048])
def test_mmq_vs_reference_kernel(ref, ops, t, M, dtype):
    """test_cuda_kernels.py:84-129: NUM_TOKENS x types x dtypes on hidden 256 / 1024 weights"""
    for (N, K) in ((256, 256), (1024, 1024), (384, 2048)):
        _mmq_check(ref, ops, t, N, K, M, dtype, seed=M)


@pytest.mark.parametrize("t", [oracle.Q4_K, oracle.Q6_K, oracle.Q8_0], ids=["Q4_K", "Q6_K", "Q8_0"])
def test_mmq_3d_vs_reference_kernel(ref, ops, t):
    """test_cuda_kernels.py:132-183: X [batch, tokens, hidden]"""
    for batch, tokens in ((2, 7), (4, 83), (8, 128)):
        x3 = _x(batch * tokens, 1024, torch.float16, batch, "randn", 0.25).view(batch, tokens, 1024)
        _mmq_check(ref, ops, t, 512, 1024, batch * tokens, torch.float16, seed=batch, x3d=x3)


@pytest.mark.parametrize("case", [(oracle.Q4_K, 14336, 4096, 512), (oracle.Q4_K, 4096, 14336, 1024), (oracle.Q6_K, 4096, 14336, 512),
                                  (oracle.Q4_K, 4096, 4096, 4096), (oracle.Q6_K, 14336, 4096, 2048), (oracle.Q4_K, 14336, 4096, 8192)],
                         ids=["Q4_K-14336x4096-M512", "Q4_K-4096x14336-M1024", "Q6_K-4096x14336-M512", "Q4_K-4096x4096-M4096",
                              "Q6_K-14336x4096-M2048", "Q4_K-14336x4096-M8192"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_mmq_config4_vs_reference_kernel(ref, ops, case, dtype):
    """BASELINE config 4 shapes at full size: every element of our output against the reference kernel's."""
    t, N, K, M = case
    _mmq_check(ref, ops, t, N, K, M, dtype, seed=M)


# ------------------------------------------------------------------------------------------------ dequantize

@pytest.mark.parametrize("t", VEC_TYPES, ids=[NAME[t] for t in VEC_TYPES])
def test_dequantize_vs_reference_kernel(ref, ops, t):
    for (m, n) in ((256, 256), (1024, 1024), (512, 4096)):
        w = torch.from_numpy(oracle.random_weights(t, m, n, seed=m)).cuda()
        want = ref.ggml_dequantize(w, t, m, n)
        got = ops.ggml_dequantize(w, t, m, n)
        assert got.dtype == want.dtype == torch.float16 and got.shape == want.shape
        _note(f"dequant/{NAME[t]}", float((got.float() - want.float()).abs().max()), float(want.float().abs().max()))
        torch.testing.assert_close(got, want, atol=DEQUANT_ATOL, rtol=DEQUANT_RTOL)
