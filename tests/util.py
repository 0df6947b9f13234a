"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dequant_golden.npz")

# reference tolerances (hf-kernels/ggml-kernels/tests/kernels/test_cuda_kernels.py)
DEQUANT_ATOL, DEQUANT_RTOL = 1e-2, 4e-2          # :58
MMVQ_ATOL, MMVQ_RTOL = 1.0, 1e-1                 # :81


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
