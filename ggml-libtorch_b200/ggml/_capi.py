"""ctypes binding of libggml_b200.so (the C ABI declared in include/ggml_b200.h).

There is NO fallback: if the shared library is missing the import raises, and every op raises when
its CUDA entry point reports an error.  Build it with `python ggml-libtorch_b200/build.py`.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GGML_B200_LIB", os.path.join(_HERE, "..", "lib", "libggml_b200.so"))

GGB_F32, GGB_F16, GGB_BF16 = 0, 1, 2
GGB_Q8_1, GGB_Q8_1_MMQ = 0, 1
GGB_OPT_STATIC_WEIGHTS = 1

# every symbol include/ggml_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "ggb_version", "ggb_last_error", "ggb_set_option", "ggb_block_elems", "ggb_block_bytes",
    "ggb_dequantize", "ggb_q8_1_bytes", "ggb_quantize_q8_1",
    "ggb_mmvq_workspace_bytes", "ggb_mmvq", "ggb_mmvq_gather", "ggb_mmvq_fused", "ggb_mmq_workspace_bytes", "ggb_mmq",
    "ggb_mmvq_fused_supported", "ggb_debug_mmvq_trace", "ggb_tp_barrier", "ggb_moe_mmvq",
)


class GgmlB200Error(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA library is not built (run `python ggml-libtorch_b200/build.py`). "
            "ggml-b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(path)
    i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    L.ggb_version.restype = ci
    L.ggb_last_error.restype = ctypes.c_char_p
    L.ggb_set_option.argtypes = [ci, ci]
    L.ggb_block_elems.argtypes = [ci]
    L.ggb_block_bytes.argtypes = [ci]
    L.ggb_dequantize.argtypes = [vp, ci, i64, i64, vp, ci, vp]
    L.ggb_q8_1_bytes.argtypes = [i64, i64, ci]
    L.ggb_q8_1_bytes.restype = i64
    L.ggb_quantize_q8_1.argtypes = [vp, ci, i64, i64, vp, ci, ci, vp]
    L.ggb_mmvq_workspace_bytes.argtypes = [ci, i64, i64, i64]
    L.ggb_mmvq_workspace_bytes.restype = i64
    L.ggb_mmvq.argtypes = [vp, vp, vp, ci, ci, i64, i64, i64, vp, i64, vp]
    L.ggb_mmvq_gather.argtypes = [vp, vp, ci, ci, i64, i64, i64, ctypes.POINTER(vp), ci, i64, i64, vp, i64, vp]
    L.ggb_mmvq_fused.argtypes = [vp, vp, vp, i64, ci, ci, i64, i64, i64, ctypes.POINTER(vp), ci, i64, i64, i64, i64, vp, i64, vp]
    L.ggb_mmvq_fused_supported.argtypes = [ci, i64, i64]
    L.ggb_debug_mmvq_trace.argtypes = [vp]
    L.ggb_debug_mmvq_trace.restype = None
    L.ggb_tp_barrier.argtypes = [ctypes.POINTER(vp), ci, ci, vp]
    L.ggb_moe_mmvq.argtypes = [vp, vp, vp, vp, ci, ci, i64, i64, i64, i64, i64, vp]
    L.ggb_mmq_workspace_bytes.argtypes = [ci, i64, i64, i64]
    L.ggb_mmq_workspace_bytes.restype = i64
    L.ggb_mmq.argtypes = [vp, vp, vp, ci, ci, i64, i64, i64, i64, vp, i64, vp]
    return L


lib = _load()


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.ggb_last_error().decode("utf-8", "replace")
        if rc in (-1, -2, -3, -4):
            raise ValueError(f"{what}: {msg} (status {rc})")
        raise GgmlB200Error(f"{what}: {msg} (status {rc})")
