"""numpy-facing ctypes wrapper of the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from the product package (ggml-libtorch_b200/), which
has no CPU fallback.  See oracle/ggml_oracle.c for what is restated and how it is pinned.
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libggml_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "custom_ops.so")

Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1 = 2, 3, 6, 7, 8, 9
Q2_K, Q3_K, Q4_K, Q5_K, Q6_K = 10, 11, 12, 13, 14
IQ4_NL, IQ4_XS = 20, 23
IQ2_XXS, IQ2_XS, IQ3_XXS, IQ1_S, IQ3_S, IQ2_S, IQ1_M = 16, 17, 18, 19, 21, 22, 29
LEGACY_TYPES = (Q4_0, Q4_1, Q5_0, Q5_1, Q8_0)
K_TYPES = (Q2_K, Q3_K, Q4_K, Q5_K, Q6_K)
ALL_TYPES = LEGACY_TYPES + K_TYPES           # dequantize + MMVQ + MMQ
IQ_GRID_TYPES = (IQ2_XXS, IQ2_XS, IQ2_S, IQ3_XXS, IQ3_S, IQ1_S, IQ1_M)    # codebook formats (tables: iq_tables.h)
IQ_TYPES = (IQ4_NL, IQ4_XS) + IQ_GRID_TYPES  # dequantize + MMVQ only (the reference has no MMQ for them either)
VEC_TYPES = ALL_TYPES + IQ_TYPES
BLOCK32_TYPES = LEGACY_TYPES + (IQ4_NL,)     # 32-value blocks: K only needs to be a multiple of 32
TYPE_NAMES = {2: "Q4_0", 3: "Q4_1", 6: "Q5_0", 7: "Q5_1", 8: "Q8_0", 9: "Q8_1",
              10: "Q2_K", 11: "Q3_K", 12: "Q4_K", 13: "Q5_K", 14: "Q6_K", 20: "IQ4_NL", 23: "IQ4_XS",
              16: "IQ2_XXS", 17: "IQ2_XS", 18: "IQ3_XXS", 19: "IQ1_S", 21: "IQ3_S", 22: "IQ2_S", 29: "IQ1_M"}


def build(force: bool = False) -> str:
    """Compile oracle/ggml_oracle.c (gcc) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "ggml_oracle.c")
    hdr = os.path.join(_HERE, "ggml_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        L.orc_block_elems.argtypes = [ci]
        L.orc_block_bytes.argtypes = [ci]
        L.orc_mmq_need_sum.argtypes = [ci]
        L.orc_dequantize.argtypes = [vp, ci, i64, vp]
        L.orc_dequantize_mt.argtypes = [vp, ci, i64, i64, vp]
        L.orc_q8_1_row_bytes.argtypes = [i64]
        L.orc_q8_1_row_bytes.restype = i64
        L.orc_quantize_q8_1.argtypes = [vp, i64, i64, vp]
        L.orc_q8_1_mmq_bytes.argtypes = [i64, i64]
        L.orc_q8_1_mmq_bytes.restype = i64
        L.orc_quantize_q8_1_mmq.argtypes = [vp, i64, i64, ci, vp]
        L.orc_mmq_a_operand.argtypes = [vp, ci, i64, vp]
        L.orc_mmvq.argtypes = [vp, ci, vp, i64, i64, i64, vp]
        L.orc_mmq.argtypes = [vp, ci, vp, i64, i64, i64, vp]
        L.orc_set_num_threads.argtypes = [ci]
        _lib = L
    return _lib


def block_elems(t: int) -> int:
    return lib().orc_block_elems(int(t))


def block_bytes(t: int) -> int:
    return lib().orc_block_bytes(int(t))


def mmq_need_sum(t: int) -> bool:
    return bool(lib().orc_mmq_need_sum(int(t)))


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


def dequantize(w, qtype: int, m: int, n: int, threads: bool = False) -> np.ndarray:
    """fp32 [m, n] from raw GGUF bytes; restates custom_ops.ggml_dequantize + gguf-py for K-quants."""
    wb = _u8(w)
    qtype = int(qtype)
    need = m * n // block_elems(qtype) * block_bytes(qtype)
    if wb.size != need:
        raise ValueError(f"expected {need} bytes for [{m},{n}] {TYPE_NAMES.get(qtype)}, got {wb.size}")
    out = np.empty((m, n), dtype=np.float32)
    if threads:
        rc = lib().orc_dequantize_mt(wb.ctypes.data, qtype, m, n, out.ctypes.data)
    else:
        rc = lib().orc_dequantize(wb.ctypes.data, qtype, m * n, out.ctypes.data)
    if rc:
        raise ValueError(f"orc_dequantize rc={rc}")
    return out


def quantize_q8_1(x: np.ndarray) -> np.ndarray:
    """x fp32 [rows, K] -> uint8 [rows, Kp/32*36] (block_q8_1 rows, ggml_kernel.cu:13-66)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, K = x.shape
    out = np.zeros((rows, lib().orc_q8_1_row_bytes(K)), dtype=np.uint8)
    lib().orc_quantize_q8_1(x.ctypes.data, rows, K, out.ctypes.data)
    return out


def quantize_q8_1_mmq(x: np.ndarray, need_sum: bool) -> np.ndarray:
    """x fp32 [M, K] -> uint8 [Kp/128, M, 144] (block_q8_1_mmq, chunk-major, mmq.cu:109-177)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    M, K = x.shape
    nbytes = lib().orc_q8_1_mmq_bytes(M, K)
    out = np.zeros(nbytes, dtype=np.uint8)
    lib().orc_quantize_q8_1_mmq(x.ctypes.data, M, K, int(bool(need_sum)), out.ctypes.data)
    return out.reshape(-1, M, 144)


def mmq_a_operand(w, qtype: int, m: int, n: int) -> np.ndarray:
    """fp16 [m, n]: the A operand the MMQ kernel feeds the tensor cores (half-FMA unpack for Q4_K/Q5_K/Q6_K,
    fp16(dequantize) for the other types); product-defined, used by the operand-exact parity test."""
    wb = _u8(w)
    out = np.empty((m, n), dtype=np.float16)
    rc = lib().orc_mmq_a_operand(wb.ctypes.data, int(qtype), m * n, out.ctypes.data)
    if rc:
        raise ValueError(f"orc_mmq_a_operand rc={rc}")
    return out


def mmvq(w, qtype: int, x: np.ndarray, N: int) -> np.ndarray:
    """Y[B,N] fp32 = MMVQ(W, q8_1(x)); x is the fp32-widened activation [B,K]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, K = x.shape
    q = quantize_q8_1(x)
    wb = _u8(w)
    y = np.empty((B, N), dtype=np.float32)
    rc = lib().orc_mmvq(wb.ctypes.data, int(qtype), q.ctypes.data, B, K, N, y.ctypes.data)
    if rc:
        raise ValueError(f"orc_mmvq rc={rc}")
    return y


def mmq(w, qtype: int, x: np.ndarray, N: int) -> np.ndarray:
    """Y[M,N] fp32 = MMQ(W, q8_1_mmq(x))."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    M, K = x.shape
    q = quantize_q8_1_mmq(x, mmq_need_sum(qtype))
    wb = _u8(w)
    y = np.empty((M, N), dtype=np.float32)
    rc = lib().orc_mmq(wb.ctypes.data, int(qtype), q.ctypes.data, M, K, N, y.ctypes.data)
    if rc:
        raise ValueError(f"orc_mmq rc={rc}")
    return y


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> int:
    """OpenMP threads of the multi-threaded oracle entry points (bench legs; torchrun exports OMP_NUM_THREADS=1)."""
    return lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------- the reference's own CPU op

_ref_mod = None


def ref_custom_ops():
    """The reference's `custom_ops` module compiled from /root/reference (oracle/_ref), or None."""
    global _ref_mod
    if _ref_mod is None:
        if not os.path.exists(_REF_PATH):
            try:
                from . import build_ref
                build_ref.build()
            except Exception:
                pass
        if not os.path.exists(_REF_PATH):
            return None
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location("custom_ops", _REF_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_mod = mod
    return _ref_mod


# ---------------------------------------------------------------- the reference's own CUDA ops (sm_100a build)

_REF_CUDA_DIR = os.path.join(_HERE, "_ref_cuda")
_ref_cuda = None


class _RefCudaOps:
    """torch.ops._ggml.* = the reference's three CUDA ops compiled unmodified for sm_100a (oracle/build_ref_cuda.py), plus
    torch.ops._ggml_q8shim.* = its two file-static q8_1 quantizers (oracle/ref_cuda_q8_shim.cu)."""

    def __init__(self, torch):
        self._t = torch
        self.ggml_dequantize = torch.ops._ggml.ggml_dequantize
        self.ggml_mul_mat_vec_a8 = torch.ops._ggml.ggml_mul_mat_vec_a8
        self.ggml_mul_mat_a8 = torch.ops._ggml.ggml_mul_mat_a8

    def quantize_q8_1(self, X):
        """uint8 [rows, Kp/32*36] exactly as quantize_row_q8_1_cuda leaves it (ggml_kernel.cu:52-66)."""
        out = self._t.ops._ggml_q8shim.quantize_q8_1(X)
        return out.view(self._t.uint8).reshape(X.shape[0], -1)

    def quantize_q8_1_mmq(self, X, qtype: int):
        """uint8 [Kp/128, rows, 144] exactly as quantize_mmq_q8_1_cuda leaves it (mmq.cu:157-177)."""
        out = self._t.ops._ggml_q8shim.quantize_q8_1_mmq(X, int(qtype))
        return out.view(self._t.uint8).reshape(-1, X.shape[0], 144)


def ref_cuda_ops():
    """The reference's CUDA ops built for sm_100a (oracle/_ref_cuda/*.so), or None when they are not built.
    Test / bench-comparator infrastructure only.  Never builds on the GPU box (the reference tree is absent there)."""
    global _ref_cuda
    if _ref_cuda is None:
        main, shim = os.path.join(_REF_CUDA_DIR, "_ggml.so"), os.path.join(_REF_CUDA_DIR, "_ggml_q8shim.so")
        if not (os.path.exists(main) and os.path.exists(shim)):
            return None
        import torch
        if not torch.cuda.is_available():
            return None
        torch.ops.load_library(main)
        torch.ops.load_library(shim)
        _ref_cuda = _RefCudaOps(torch)
    return _ref_cuda


# ---------------------------------------------------------------- synthetic GGUF weights

def random_weights(qtype: int, rows: int, cols: int, seed: int = 0) -> np.ndarray:
    """Random *valid* block bytes, uint8 [rows, cols/QK*bs] (SURVEY.md §8d 'C2' recipe):
    all payload bytes uniform, fp16 scale fields = fp16(U(0.5,1.5) * 2^-10) with random sign off."""
    qtype = int(qtype)
    qk, bs = block_elems(qtype), block_bytes(qtype)
    nb = rows * cols // qk
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, size=(nb, bs), dtype=np.uint8)
    half_offsets = {Q4_0: [0], Q4_1: [0, 2], Q5_0: [0], Q5_1: [0, 2], Q8_0: [0],
                    Q2_K: [80, 82], Q3_K: [108], Q4_K: [0, 2], Q5_K: [0, 2], Q6_K: [208],
                    IQ4_NL: [0], IQ4_XS: [0], IQ2_XXS: [0], IQ2_XS: [0], IQ2_S: [0], IQ3_XXS: [0], IQ3_S: [0], IQ1_S: [0],
                    IQ1_M: []}[qtype]
    for off in half_offsets:
        h = (rng.uniform(0.5, 1.5, size=nb) * 2.0 ** -10).astype(np.float16)
        blocks[:, off:off + 2] = h.view(np.uint8).reshape(nb, 2)
    if qtype == IQ1_M:      # its fp16 scale lives in the top nibbles of the four 16-bit scale words (bytes 49, 51, 53, 55)
        h = (rng.uniform(0.5, 1.5, size=nb) * 2.0 ** -10).astype(np.float16).view(np.uint16)
        for i in range(4):
            blocks[:, 49 + 2 * i] = (blocks[:, 49 + 2 * i] & 0x0f) | (((h >> (4 * i)) & 0x0f) << 4).astype(np.uint8)
    return blocks.reshape(rows, cols // qk * bs)
