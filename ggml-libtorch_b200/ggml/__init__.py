"""`ggml` -- drop-in for the reference's torch custom-op package (torch-ext/ggml/__init__.py:15-44).

Same three public functions, same positional signatures:

    ggml_dequantize(W, quant_type, m, n)            -> Tensor [m, n]     (fp16; `dtype=` extension)
    ggml_mul_mat_vec_a8(W, X, quant_type, row)      -> Tensor [B, row]   (MMVQ, 1 <= B <= 8)
    ggml_mul_mat_a8(W, X, quant_type, row)          -> Tensor [..., row] (MMQ, X 2-D or 3-D)

The ops are registered with torch.library (namespace `ggml_b200`, the reference's schemas, CUDA only, with
fake kernels so meta tracing works; `register_reference_namespace()` also defines them as `torch.ops._ggml.*`)
and call hand-written sm_100a CUDA through the C ABI of libggml_b200.so
(include/ggml_b200.h).  PyTorch supplies device memory and streams only.  There is no CPU path:
non-CUDA tensors raise.
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import (GGB_BF16, GGB_F16, GGB_F32, GGB_OPT_STATIC_WEIGHTS, GGB_Q8_1, GGB_Q8_1_MMQ,  # noqa: F401
                    GgmlB200Error)

__all__ = [
    "ggml_dequantize", "ggml_mul_mat_vec_a8", "ggml_mul_mat_a8", "ggml_quantize_q8_1",
    "dequantize_out", "mul_mat_vec_a8_out", "mul_mat_a8_out", "mmq_workspace_bytes", "mmvq_workspace_bytes",
    "set_option", "GGB_OPT_STATIC_WEIGHTS", "mul_mat_vec_a8_gather", "mul_mat_vec_a8_fused",
    "block_elems", "block_bytes", "SUPPORTED_TYPES", "MMQ_TYPES", "MMVQ_MAX_BATCH", "mmvq_fused_supported",
    "register_reference_namespace",
    "ggml_moe_a8_vec", "ggml_moe_a8", "ggml_moe_get_block_size", "moe_a8_vec_out", "moe_align_block_size",
]

_lib = _capi.lib
_DT = {torch.float32: GGB_F32, torch.float16: GGB_F16, torch.bfloat16: GGB_BF16}
_DT_INV = {v: k for k, v in _DT.items()}

# ggml_type ids, ggml/ggml-common.h:1128-1161
SUPPORTED_TYPES = (2, 3, 6, 7, 8, 10, 11, 12, 13, 14, 20, 23, 16, 17, 18, 19, 21, 22, 29)
MMQ_TYPES = (2, 3, 6, 7, 8, 10, 11, 12, 13, 14)      # the IQ types (IQ4_NL 20, IQ4_XS 23, IQ2_XXS 16, IQ2_XS 17, IQ3_XXS 18, IQ1_S 19,
                                                      # IQ3_S 21, IQ2_S 22, IQ1_M 29): dequantize + MMVQ only, as in the reference
MMVQ_MAX_BATCH = 8


def set_option(option: int, value: int) -> None:
    """Process-wide library options (include/ggml_b200.h `enum ggb_option`); e.g. GGB_OPT_STATIC_WEIGHTS = 1
    declares inference weights immutable so the MMVQ weight stream may start under the previous kernel."""
    _capi.check(_lib.ggb_set_option(int(option), int(value)), "ggb_set_option")


def block_elems(quant_type: int) -> int:
    return _lib.ggb_block_elems(int(quant_type))


def block_bytes(quant_type: int) -> int:
    return _lib.ggb_block_bytes(int(quant_type))


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_w(W: torch.Tensor, quant_type: int, rows: int, cols: int, what: str) -> None:
    if not W.is_cuda:
        raise ValueError(f"{what}: W must be a CUDA tensor (ggml-b200 has no CPU path)")
    if W.dtype != torch.uint8:
        raise ValueError(f"{what}: W must be uint8 raw GGUF bytes, got {W.dtype}")
    if not W.is_contiguous():
        raise ValueError(f"{what}: W must be contiguous")
    qk, bs = block_elems(quant_type), block_bytes(quant_type)
    if qk == 0:
        raise ValueError(f"{what}: unsupported ggml quant type {int(quant_type)}")
    if cols % qk != 0:
        raise ValueError(f"{what}: {cols} columns is not a multiple of the {qk}-value block")
    need = rows * (cols // qk) * bs
    if W.numel() != need:
        raise ValueError(f"{what}: W has {W.numel()} bytes, expected {need} for [{rows}, {cols}] of type {int(quant_type)}")


# ------------------------------------------------------------------------------------------------
# out-variants: no allocation, used by the custom ops below and by bench.py / CUDA-graph capture
# ------------------------------------------------------------------------------------------------

def dequantize_out(W: torch.Tensor, quant_type: int, m: int, n: int, out: torch.Tensor) -> torch.Tensor:
    _check_w(W, quant_type, m, n, "ggml_dequantize")
    if out.dtype not in _DT or not out.is_contiguous() or out.numel() != m * n or out.device != W.device:
        raise ValueError("ggml_dequantize: out must be a contiguous fp32/fp16/bf16 tensor of m*n values on W's device")
    with torch.cuda.device(W.device):
        rc = _lib.ggb_dequantize(W.data_ptr(), int(quant_type), m, n, out.data_ptr(), _DT[out.dtype], _stream(W))
    _capi.check(rc, "ggml_dequantize")
    return out


def q8_1_bytes(rows: int, K: int, layout: int) -> int:
    return _lib.ggb_q8_1_bytes(rows, K, layout)


def ggml_quantize_q8_1(X: torch.Tensor, layout: int = GGB_Q8_1, wtype: int = 12) -> torch.Tensor:
    """q8_1 activation quantization as a standalone op (parity tests / MMQ producer).
    layout GGB_Q8_1 -> uint8 [rows, Kp/32*36]; GGB_Q8_1_MMQ -> uint8 [Kp/128, rows, 144]."""
    if not X.is_cuda or X.dtype not in _DT or X.dim() != 2 or not X.is_contiguous():
        raise ValueError("ggml_quantize_q8_1: X must be a contiguous 2-D fp32/fp16/bf16 CUDA tensor")
    rows, K = X.shape
    nbytes = q8_1_bytes(rows, K, layout)
    out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=X.device)[:nbytes]
    with torch.cuda.device(X.device):
        rc = _lib.ggb_quantize_q8_1(X.data_ptr(), _DT[X.dtype], rows, K, out.data_ptr(), layout, int(wtype), _stream(X))
    _capi.check(rc, "ggml_quantize_q8_1")
    if layout == GGB_Q8_1:
        return out.view(rows, -1)
    return out.view(-1, rows, 144)


def _check_x(X: torch.Tensor, W: torch.Tensor, what: str) -> None:
    if not X.is_cuda or X.device != W.device:
        raise ValueError(f"{what}: X must be a CUDA tensor on W's device")
    if X.dtype not in _DT:
        raise ValueError(f"{what}: X must be fp32, fp16 or bf16, got {X.dtype}")
    if not X.is_contiguous():
        raise ValueError(f"{what}: X must be contiguous")


def mmvq_workspace_bytes(quant_type: int, B: int, K: int, N: int) -> int:
    return _lib.ggb_mmvq_workspace_bytes(int(quant_type), B, K, N)


def mul_mat_vec_a8_out(W: torch.Tensor, X: torch.Tensor, quant_type: int, row: int, out: torch.Tensor,
                       ws: torch.Tensor | None = None) -> torch.Tensor:
    what = "ggml_mul_mat_vec_a8"
    _check_x(X, W, what)
    if X.dim() != 2:
        raise ValueError(f"{what}: X must be [B, K]")
    B, K = X.shape
    if not 1 <= B <= MMVQ_MAX_BATCH:
        raise ValueError(f"{what}: batch {B} outside 1..{MMVQ_MAX_BATCH} (use ggml_mul_mat_a8)")
    _check_w(W, quant_type, row, K, what)
    if out.dtype != X.dtype or not out.is_contiguous() or out.numel() != B * row or out.device != W.device:
        raise ValueError(f"{what}: out must be a contiguous [B, row] tensor of X's dtype")
    need = mmvq_workspace_bytes(quant_type, B, K, row)
    if need and (ws is None or ws.numel() * ws.element_size() < need):
        ws = torch.empty(need, dtype=torch.uint8, device=W.device)
    with torch.cuda.device(W.device):
        rc = _lib.ggb_mmvq(W.data_ptr(), X.data_ptr(), out.data_ptr(), int(quant_type), _DT[X.dtype], B, K, row,
                           ws.data_ptr() if ws is not None else None,
                           ws.numel() * ws.element_size() if ws is not None else 0, _stream(W))
    _capi.check(rc, what)
    return out


def mul_mat_vec_a8_gather(W: torch.Tensor, X: torch.Tensor, quant_type: int, row: int, peer_ptrs, ldy: int,
                          col_offset: int, ws: torch.Tensor | None = None) -> None:
    """MMVQ of this rank's row slice with the column-parallel all-gather fused into the epilogue (ggb_mmvq_gather):
    results go to `peer_ptrs[i] + (tok * ldy + col_offset + r) * sizeof(X.dtype)` for every i.  `peer_ptrs` are the
    device pointers of one symmetric [B, ldy] buffer as mapped on every rank (torch symmetric memory `buffer_ptrs`).
    The caller runs a cross-rank barrier on the stream before reading the buffer.  Raises ValueError for shapes / types
    that are not on the fused path."""
    _check_x(X, W, "ggml_mul_mat_vec_a8")
    B, K = X.shape
    _check_w(W, quant_type, row, K, "ggml_mul_mat_vec_a8")
    import ctypes
    need = mmvq_workspace_bytes(quant_type, B, K, row)
    if need and (ws is None or ws.numel() * ws.element_size() < need):
        ws = torch.empty(need, dtype=torch.uint8, device=W.device)
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(W.device):
        rc = _lib.ggb_mmvq_gather(W.data_ptr(), X.data_ptr(), int(quant_type), _DT[X.dtype], B, K, row, arr, len(peer_ptrs),
                                  int(ldy), int(col_offset), ws.data_ptr() if ws is not None else None,
                                  ws.numel() * ws.element_size() if ws is not None else 0, _stream(W))
    _capi.check(rc, "ggml_mul_mat_vec_a8 (gather)")


def mul_mat_vec_a8_fused(W: torch.Tensor, X: torch.Tensor, X2: torch.Tensor | None, quant_type: int, row: int, peer_ptrs,
                         ldy: int, col_offset: int, split_row: int = 0, col_offset2: int = 0,
                         ws: torch.Tensor | None = None) -> None:
    """The general fused MMVQ (ggb_mmvq_fused).  X (and X2) are [B, K] views with unit column stride and a common row
    stride; with X2 the activation is silu(X) * X2 computed on the fly (torch eager's roundings).  Results go to
    `peer_ptrs[i] + (tok * ldy + col) * itemsize` with col = col_offset + r for r < split_row (or no split), else
    col_offset2 + r - split_row."""
    what = "ggml_mul_mat_vec_a8 (fused)"
    if not X.is_cuda or X.device != W.device or X.dtype not in _DT or X.dim() != 2 or X.stride(1) != 1:
        raise ValueError(f"{what}: X must be a CUDA [B, K] view with unit column stride on W's device (fp32/fp16/bf16)")
    if X2 is not None and (X2.shape != X.shape or X2.dtype != X.dtype or X2.stride() != X.stride() or X2.device != X.device):
        raise ValueError(f"{what}: X2 must match X in shape, dtype, strides and device")
    B, K = X.shape
    if not 1 <= B <= MMVQ_MAX_BATCH:
        raise ValueError(f"{what}: batch {B} outside 1..{MMVQ_MAX_BATCH}")
    _check_w(W, quant_type, row, K, what)
    import ctypes
    need = mmvq_workspace_bytes(quant_type, B, K, row)
    if need and (ws is None or ws.numel() * ws.element_size() < need):
        ws = torch.empty(need, dtype=torch.uint8, device=W.device)
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(W.device):
        rc = _lib.ggb_mmvq_fused(W.data_ptr(), X.data_ptr(), X2.data_ptr() if X2 is not None else None, X.stride(0),
                                 int(quant_type), _DT[X.dtype], B, K, row, arr, len(peer_ptrs), int(ldy), int(col_offset),
                                 int(split_row), int(col_offset2), ws.data_ptr() if ws is not None else None,
                                 ws.numel() * ws.element_size() if ws is not None else 0, _stream(W))
    _capi.check(rc, what)


def mmq_workspace_bytes(quant_type: int, M: int, K: int, N: int) -> int:
    return _lib.ggb_mmq_workspace_bytes(int(quant_type), M, K, N)


def mul_mat_a8_out(W: torch.Tensor, X2d: torch.Tensor, quant_type: int, row: int, out: torch.Tensor,
                   ws: torch.Tensor | None = None, ldy: int | None = None) -> torch.Tensor:
    what = "ggml_mul_mat_a8"
    _check_x(X2d, W, what)
    if X2d.dim() != 2:
        raise ValueError(f"{what}: X must be flattened to [M, K]")
    M, K = X2d.shape
    _check_w(W, quant_type, row, K, what)
    ldy = row if ldy is None else ldy
    if out.dtype != X2d.dtype or out.device != W.device:
        raise ValueError(f"{what}: out must have X's dtype and device")
    need = mmq_workspace_bytes(quant_type, M, K, row)
    if ws is None or ws.numel() * ws.element_size() < need:
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=W.device)
    with torch.cuda.device(W.device):
        rc = _lib.ggb_mmq(W.data_ptr(), X2d.data_ptr(), out.data_ptr(), int(quant_type), _DT[X2d.dtype], M, K, row,
                          ldy, ws.data_ptr(), ws.numel() * ws.element_size(), _stream(W))
    _capi.check(rc, what)
    return out


# ------------------------------------------------------------------------------------------------
# torch.library registration (schemas follow torch-ext/torch_binding.cpp:6-21)
# ------------------------------------------------------------------------------------------------

# Explicit schema strings (torch.library.custom_op would infer `SymInt type` from the Python annotation; the reference
# declares `int type`): identical to torch-ext/torch_binding.cpp:8-19.
_LIB = torch.library.Library("ggml_b200", "DEF")
_LIB.define("ggml_dequantize(Tensor W, int type, SymInt m, SymInt n) -> Tensor")
_LIB.define("ggml_mul_mat_vec_a8(Tensor W, Tensor X, int type, SymInt row) -> Tensor")
_LIB.define("ggml_mul_mat_a8(Tensor W, Tensor X, int type, SymInt row) -> Tensor")
# extension (no reference counterpart): the cast the reference's callers do afterwards, fused; out_dtype = GGB_F32/F16/BF16
_LIB.define("ggml_dequantize_dtype(Tensor W, int type, SymInt m, SymInt n, int out_dtype) -> Tensor")


def _op_dequantize(W: torch.Tensor, type: int, m: int, n: int) -> torch.Tensor:
    out = torch.empty((m, n), dtype=torch.float16, device=W.device)      # fp16 like the reference CUDA op
    return dequantize_out(W, type, m, n, out)


def _op_dequantize_dtype(W: torch.Tensor, type: int, m: int, n: int, out_dtype: int) -> torch.Tensor:
    out = torch.empty((m, n), dtype=_DT_INV[out_dtype], device=W.device)
    return dequantize_out(W, type, m, n, out)


def _op_mmvq(W: torch.Tensor, X: torch.Tensor, type: int, row: int) -> torch.Tensor:
    out = torch.empty((X.shape[0], row), dtype=X.dtype, device=W.device)
    return mul_mat_vec_a8_out(W, X, type, row, out)


def _op_mmq(W: torch.Tensor, X: torch.Tensor, type: int, row: int) -> torch.Tensor:
    # X 2-D [M, K] or 3-D [B, T, K] flattened to B*T rows, as ggml/mmq.cu:183-204
    if X.dim() not in (2, 3):
        raise ValueError("ggml_mul_mat_a8: X must be 2-D or 3-D")
    lead = X.shape[:-1]
    X2 = X.reshape(-1, X.shape[-1])
    out = torch.empty((X2.shape[0], row), dtype=X.dtype, device=W.device)
    mul_mat_a8_out(W, X2, type, row, out)
    return out.view(*lead, row)


# CUDA dispatch key only, like the reference (torch_binding.cpp:10,15,20); fake kernels so meta tracing / torch.compile see shapes
_LIB.impl("ggml_dequantize", _op_dequantize, "CUDA")
_LIB.impl("ggml_dequantize_dtype", _op_dequantize_dtype, "CUDA")
_LIB.impl("ggml_mul_mat_vec_a8", _op_mmvq, "CUDA")
_LIB.impl("ggml_mul_mat_a8", _op_mmq, "CUDA")


@torch.library.register_fake("ggml_b200::ggml_dequantize")
def _(W, type, m, n):
    return torch.empty((m, n), dtype=torch.float16, device=W.device)


@torch.library.register_fake("ggml_b200::ggml_dequantize_dtype")
def _(W, type, m, n, out_dtype):
    return torch.empty((m, n), dtype=_DT_INV[out_dtype], device=W.device)


@torch.library.register_fake("ggml_b200::ggml_mul_mat_vec_a8")
def _(W, X, type, row):
    return torch.empty((X.shape[0], row), dtype=X.dtype, device=W.device)


@torch.library.register_fake("ggml_b200::ggml_mul_mat_a8")
def _(W, X, type, row):
    return torch.empty((*X.shape[:-1], row), dtype=X.dtype, device=W.device)


# ------------------------------------------------------------------------------------------------
# public functions (reference names and positional signatures)
# ------------------------------------------------------------------------------------------------

def ggml_dequantize(W: torch.Tensor, quant_type: int, m: int, n: int, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Dequantize the GGML tensor.  fp16 output like the reference CUDA op (ggml/ggml_kernel.cu:71-73);
    `dtype=` (fp32/fp16/bf16) fuses the cast the reference's callers do afterwards."""
    dtype = torch.float16 if dtype is None else dtype
    if dtype not in _DT:
        raise ValueError(f"ggml_dequantize: unsupported output dtype {dtype}")
    _check_w(W, quant_type, int(m), int(n), "ggml_dequantize")
    if dtype == torch.float16:
        return torch.ops.ggml_b200.ggml_dequantize(W, int(quant_type), int(m), int(n))
    return torch.ops.ggml_b200.ggml_dequantize_dtype(W, int(quant_type), int(m), int(n), _DT[dtype])


def ggml_mul_mat_vec_a8(W: torch.Tensor, X: torch.Tensor, quant_type: int, row: int) -> torch.Tensor:
    """Mulmat with the MMVQ kernel.  The reference requires batch_size == 1; here 1 <= batch <= 8."""
    batch = X.size(0)
    if not 1 <= batch <= MMVQ_MAX_BATCH:
        raise ValueError(f"Batch size must be 1..{MMVQ_MAX_BATCH} for the MMVQ kernel, got {batch}")
    return torch.ops.ggml_b200.ggml_mul_mat_vec_a8(W, X, int(quant_type), int(row))


def ggml_mul_mat_a8(W: torch.Tensor, X: torch.Tensor, quant_type: int, row: int) -> torch.Tensor:
    """Mulmat through the MMQ kernel for arbitrary batch size."""
    return torch.ops.ggml_b200.ggml_mul_mat_a8(W, X, int(quant_type), int(row))


# ------------------------------------------------------------------------------------------------
# mixture-of-experts variants.  Not in this reference; names, argument order and output
# layout are those of its upstream's ops (vLLM `_custom_ops.ggml_moe_a8_vec / ggml_moe_a8 / ggml_moe_get_block_size`).
# ------------------------------------------------------------------------------------------------

def _check_experts(W: torch.Tensor, quant_type: int, row: int, K: int, what: str) -> int:
    if W.dim() != 3:
        raise ValueError(f"{what}: W must be the stacked expert weights [E, row, bytes_per_row]")
    E = W.shape[0]
    _check_w(W, quant_type, E * row, K, what)
    return E


def moe_a8_vec_out(X: torch.Tensor, W: torch.Tensor, topk_ids: torch.Tensor, top_k: int, quant_type: int, row: int,
                   tokens: int, out: torch.Tensor) -> torch.Tensor:
    """Grouped MMVQ (ggb_moe_mmvq): out[i] = q8_1(X[i // top_k]) . W[topk_ids.flatten()[i]]^T for i < tokens * top_k, one
    launch, no allocation (`topk_ids` must already be int32), CUDA-graph capturable."""
    what = "ggml_moe_a8_vec"
    _check_x(X, W, what)
    if X.dim() != 2 or X.shape[0] != tokens:
        raise ValueError(f"{what}: X must be [tokens, K]")
    K = X.shape[1]
    E = _check_experts(W, quant_type, row, K, what)
    pairs = tokens * top_k
    if topk_ids.dtype != torch.int32 or not topk_ids.is_contiguous() or topk_ids.numel() != pairs or topk_ids.device != W.device:
        raise ValueError(f"{what}: topk_ids must be a contiguous int32 tensor of tokens * top_k = {pairs} ids on W's device")
    if out.dtype != X.dtype or not out.is_contiguous() or out.numel() != pairs * row or out.device != W.device:
        raise ValueError(f"{what}: out must be a contiguous [tokens * top_k, row] tensor of X's dtype")
    with torch.cuda.device(W.device):
        rc = _lib.ggb_moe_mmvq(W.data_ptr(), X.data_ptr(), topk_ids.data_ptr(), out.data_ptr(), int(quant_type),
                               _DT[X.dtype], tokens, top_k, E, K, row, _stream(W))
    _capi.check(rc, what)
    return out


def ggml_moe_a8_vec(X: torch.Tensor, W: torch.Tensor, topk_ids: torch.Tensor, top_k: int, quant_type: int, row: int,
                    tokens: int) -> torch.Tensor:
    """MoE matmul through the grouped MMVQ kernel: X [tokens, K], W [E, row, bytes] raw GGUF expert stack, topk_ids
    [tokens, top_k] -> [tokens * top_k, row] (row i = token i // top_k through expert topk_ids[i // top_k, i % top_k]).
    The second projection of an expert MLP is the same call with top_k = 1 and tokens = the number of pairs."""
    ids = topk_ids.reshape(-1)
    if ids.dtype != torch.int32:
        ids = ids.to(torch.int32)
    out = torch.empty((int(tokens) * int(top_k), int(row)), dtype=X.dtype, device=W.device)
    return moe_a8_vec_out(X, W, ids.contiguous(), int(top_k), int(quant_type), int(row), int(tokens), out)


_MOE_BLOCK = 32


def ggml_moe_get_block_size(quant_type: int) -> int:
    """Padding granularity of the per-expert token lists `ggml_moe_a8` expects (upstream: the MMQ tile width per type).
    Here the tensor-core tile is picked per expert from its token count, so one granularity serves every type."""
    if int(quant_type) not in MMQ_TYPES:
        raise ValueError(f"ggml_moe_get_block_size: ggml quant type {int(quant_type)} has no MMQ kernel")
    return _MOE_BLOCK


def moe_align_block_size(topk_ids: torch.Tensor, block_size: int, num_experts: int):
    """The token-list layout `ggml_moe_a8` consumes (upstream: vllm `moe_align_block_size`, a separate kernel there; plain
    torch here -- host-side scheduling, not part of the hot path).  Returns (sorted_token_ids, expert_ids,
    num_tokens_post_padded): pair indices sorted by expert, every expert's list padded to a multiple of `block_size`
    with the value `topk_ids.numel()`, and the expert of every block."""
    flat = topk_ids.reshape(-1).to(torch.int64)
    numel = flat.numel()
    counts = torch.bincount(flat, minlength=num_experts)[:num_experts]
    padded = (counts + block_size - 1) // block_size * block_size
    ends = torch.cumsum(padded, 0)
    starts = ends - padded
    total = numel + num_experts * (block_size - 1)
    total = (total + block_size - 1) // block_size * block_size
    sorted_ids = torch.full((total,), numel, dtype=torch.int32, device=flat.device)
    order = torch.argsort(flat, stable=True)
    ranks = torch.arange(numel, device=flat.device) - (torch.cumsum(counts, 0) - counts)[flat[order]]
    sorted_ids[starts[flat[order]] + ranks] = order.to(torch.int32)
    nblocks = total // block_size
    expert_ids = torch.searchsorted(ends, torch.arange(nblocks, device=flat.device) * block_size, right=True)
    expert_ids = expert_ids.clamp_(max=num_experts - 1).to(torch.int32)
    return sorted_ids, expert_ids, ends[-1:].to(torch.int32)


def ggml_moe_a8(X: torch.Tensor, W: torch.Tensor, sorted_token_ids: torch.Tensor, expert_ids: torch.Tensor,
                num_tokens_post_padded: torch.Tensor, quant_type: int, row: int, top_k: int, tokens: int) -> torch.Tensor:
    """MoE matmul for prefill: every expert's token list goes through the tensor-core MMQ kernel (`ggml_mul_mat_a8`).
    X [tokens, K]; W [E, row, bytes]; (sorted_token_ids, expert_ids, num_tokens_post_padded) as produced by
    `moe_align_block_size(topk_ids, ggml_moe_get_block_size(type), E)`.  Returns [tokens * top_k, row]; pair i = token
    i // top_k.  Host-scheduled: ONE device->host read of the block -> expert map, then gather + MMQ + scatter per
    non-empty expert (a grouped tile scheduler inside the MMQ kernel is the next step)."""
    what = "ggml_moe_a8"
    _check_x(X, W, what)
    if X.dim() != 2 or X.shape[0] != tokens:
        raise ValueError(f"{what}: X must be [tokens, K]")
    if int(quant_type) not in MMQ_TYPES:
        raise ValueError(f"{what}: ggml quant type {int(quant_type)} has no MMQ kernel (use ggml_moe_a8_vec)")
    K = X.shape[1]
    E = _check_experts(W, quant_type, row, K, what)
    pairs = int(tokens) * int(top_k)
    block = _MOE_BLOCK
    # the one host read: the padded length and the block -> expert map together
    head = torch.cat([num_tokens_post_padded.reshape(-1)[:1].to(torch.int64), expert_ids.reshape(-1).to(torch.int64)]).tolist()
    n_pad = head[0]
    if n_pad % block or n_pad > sorted_token_ids.numel() or n_pad // block > len(head) - 1:
        raise ValueError(f"{what}: num_tokens_post_padded = {n_pad} is not a multiple of the block size {block} "
                         "inside sorted_token_ids / expert_ids")
    blocks = head[1:1 + n_pad // block]
    # one spare row catches the padding entries (id == pairs) of every scatter
    out = torch.zeros((pairs + 1, int(row)), dtype=X.dtype, device=W.device)
    slots_all = sorted_token_ids[:n_pad].to(torch.int64).clamp_(max=pairs)
    tok_all = (slots_all // int(top_k)).clamp_(max=int(tokens) - 1)
    b0 = 0
    while b0 < len(blocks):
        e = blocks[b0]
        b1 = b0
        while b1 < len(blocks) and blocks[b1] == e:
            b1 += 1
        if 0 <= e < E:                  # an id outside [0, E) (an expert held by another rank): its rows stay zero, as in the vec op
            sl = slice(b0 * block, b1 * block)
            ye = ggml_mul_mat_a8(W[e], X.index_select(0, tok_all[sl]), quant_type, row)
            out.index_copy_(0, slots_all[sl], ye)
        b0 = b1
    return out[:pairs]


def mmvq_fused_supported(quant_type: int, B: int, K: int) -> bool:
    """Pure query: is (type, batch, K) on the fused-epilogue path of mul_mat_vec_a8_gather / _fused?"""
    return bool(_lib.ggb_mmvq_fused_supported(int(quant_type), int(B), int(K)))


def register_reference_namespace(namespace: str = "_ggml") -> None:
    """Also expose the three ops as torch.ops.<namespace>.* with the reference's exact schemas
    (torch-ext/torch_binding.cpp:6-21; its TORCH_EXTENSION_NAME is `_ggml` in development builds), so that callers which
    go through torch.ops instead of the Python functions are drop-in too.  Opt-in: defining the namespace while the
    reference's own extension is loaded would be a duplicate registration."""
    lib = torch.library.Library(namespace, "FRAGMENT")
    lib.define("ggml_dequantize(Tensor W, int type, SymInt m, SymInt n) -> Tensor")
    lib.define("ggml_mul_mat_vec_a8(Tensor W, Tensor X, int type, SymInt row) -> Tensor")
    lib.define("ggml_mul_mat_a8(Tensor W, Tensor X, int type, SymInt row) -> Tensor")
    lib.impl("ggml_dequantize", lambda W, type, m, n: torch.ops.ggml_b200.ggml_dequantize(W, type, m, n), "CUDA")
    lib.impl("ggml_mul_mat_vec_a8", lambda W, X, type, row: torch.ops.ggml_b200.ggml_mul_mat_vec_a8(W, X, type, row), "CUDA")
    lib.impl("ggml_mul_mat_a8", lambda W, X, type, row: torch.ops.ggml_b200.ggml_mul_mat_a8(W, X, type, row), "CUDA")
    _REF_NS_LIBS.append(lib)            # keep the registration alive


_REF_NS_LIBS: list = []
