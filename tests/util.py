"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dequant_golden.npz")

# reference tolerances (hf-kernels/ggml-kernels/tests/kernels/test_cuda_kernels.py)
DEQUANT_ATOL, DEQUANT_RTOL = 1e-2, 4e-2          # :58
MMVQ_ATOL, MMVQ_RTOL = 1.0, 1e-1                 # :81


def patch_gguf_iq3s_to_reference_table() -> None:
    """The reference vendors the FIRST IQ3_S codebook (hf-kernels/ggml-kernels/ggml/ggml-common.h:685 iq3xs_grid: levels
    4, 12, ..., 52, 62 used as d * (0.5 + ls) * 0.5 * grid; dequantize.cuh:333-354), today's gguf-py the later one (levels
    1, 3, ..., 13, 15 with d * (1 + 2 ls)).  The two give bit-identical values except on the top level: 62 / 4 = 15.5 vs 15.
    The product follows the reference, so gguf-py -- the expected-value source of the reference's tests -- gets its IQ3_S
    top level moved to 15.5 for this test session (idempotent)."""
    try:
        from gguf import quants
    except ImportError:
        return
    quants.IQ3_S.init_grid()
    g = quants.IQ3_S.grid
    if g is not None and (g == 15.0).any():
        g = g.copy()
        g[g == 15.0] = 15.5
        quants.IQ3_S.grid = g


def assert_bits_equal(got: np.ndarray, want: np.ndarray, what: str = "") -> None:
    """Bit-exact equality of two float arrays; NaNs match any NaN (payload/sign of a NaN produced by
    inf - inf differs between x86 and the GPU and is not part of the contract)."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    u = {2: np.uint16, 4: np.uint32}[got.dtype.itemsize]
    same = got.view(u) == want.view(u)
    if got.dtype.kind == "f":
        same |= np.isnan(got) & np.isnan(want)
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {got.size} values differ; first at {i}: got {got[i]!r} want {want[i]!r}")


def bf16_bits_from_f32(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns (NaN -> 0x7fff)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[np.isnan(x)] = 0x7FFF
    return r


def mmq_b_operand_fp16(x32: np.ndarray, need_sum: bool) -> np.ndarray:
    """fp16(d8 * q8) per element, from the ORACLE's q8_1_mmq blocks: the B operand the MMQ kernel feeds
    the tensor cores (d8 is fp16-rounded for need_sum weight types, fp32 otherwise; mmq.cu:149-153)."""
    import oracle
    M, K = x32.shape
    blocks = oracle.quantize_q8_1_mmq(x32, need_sum)            # [Kp/128, M, 144]
    nch = blocks.shape[0]
    q = blocks[:, :, 16:].view(np.int8).astype(np.float32).reshape(nch, M, 4, 32)
    if need_sum:
        d = blocks[:, :, :16].copy().view(np.float16).reshape(nch, M, 4, 2)[..., 0].astype(np.float32)
    else:
        d = blocks[:, :, :16].copy().view(np.float32).reshape(nch, M, 4)
    y = (d[..., None] * q).astype(np.float16)                    # one fp32 product, one RN to fp16
    return y.transpose(1, 0, 2, 3).reshape(M, nch * 128)[:, :K]
