"""GPU parity: q8_1 / q8_1_mmq activation quantization, BIT-EXACT against the oracle's restatement of
quantize_q8_1 (ggml/ggml_kernel.cu:13-66) and quantize_mmq_q8_1 (ggml/mmq.cu:109-177)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.float16, torch.bfloat16]


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _inputs(rows, K, dtype, seed, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "randn":
        x = torch.randn(rows, K, generator=g) * 2.5
    elif kind == "rand":
        x = torch.rand(rows, K, generator=g)
    else:   # adversarial: zero groups, ties at .5 multiples, huge / tiny magnitudes
        x = torch.randn(rows, K, generator=g)
        x[0, :32] = 0
        if K >= 64:
            x[0, 32:64] = torch.arange(32) * 0.5 - 8.0
        x[-1, -32:] *= 1e4
        x[-1, :16] *= 1e-6
    return x.to(dtype)


@pytest.mark.parametrize("kind", ["randn", "rand", "adversarial"])
@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(1, 32), (1, 4096), (3, 544), (8, 14336), (83, 1024), (5, 40)])
def test_q8_1_bit_exact(ops, shape, dtype, kind):
    rows, K = shape
    x = _inputs(rows, K, dtype, rows * 7 + K, kind)
    want = oracle.quantize_q8_1(x.float().numpy())
    got = ops.ggml_quantize_q8_1(x.cuda(), ops.GGB_Q8_1).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} bytes differ"


@pytest.mark.parametrize("wtype", [oracle.Q4_K, oracle.Q6_K], ids=["need_sum", "fp32_d"])
@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(1, 512), (7, 4096), (83, 1024), (128, 640), (2048, 256)])
def test_q8_1_mmq_bit_exact(ops, shape, dtype, wtype):
    M, K = shape
    x = _inputs(M, K, dtype, M + K, "randn")
    want = oracle.quantize_q8_1_mmq(x.float().numpy(), oracle.mmq_need_sum(wtype))
    got = ops.ggml_quantize_q8_1(x.cuda(), ops.GGB_Q8_1_MMQ, wtype).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} bytes differ"


def test_q8_1_unaligned_rows(ops):
    """K not a multiple of 8 and a row pointer that is not 16-byte aligned (scalar load path)."""
    x = torch.randn(3, 100, dtype=torch.float16)
    want = oracle.quantize_q8_1(x.float().numpy())
    got = ops.ggml_quantize_q8_1(x.cuda(), ops.GGB_Q8_1).cpu().numpy()
    assert np.array_equal(got, want)
