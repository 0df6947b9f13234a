"""CPU tests of the C-ABI boundary: libggml_b200.so loads without a GPU, exports every symbol that
include/ggml_b200.h declares, answers the geometry queries, and rejects bad arguments with a status
code + message BEFORE touching CUDA.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ggml_b200.h")


@pytest.fixture(scope="module")
def capi():
    import ggml._capi as c
    return c


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ggb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(capi):
    names = _declared_functions()
    assert len(names) >= 11
    raw = ctypes.CDLL(os.path.abspath(capi.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/ggml_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == names


def test_version_and_geometry(capi):
    L = capi.lib
    assert L.ggb_version() == 100
    want = {2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34),
            10: (256, 84), 11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210),
            20: (32, 18), 23: (256, 136),
            16: (256, 66), 17: (256, 74), 18: (256, 98), 19: (256, 50), 21: (256, 110), 22: (256, 82), 29: (256, 56)}
    for t, (qk, bs) in want.items():
        assert L.ggb_block_elems(t) == qk and L.ggb_block_bytes(t) == bs
    for t in (0, 1, 9, 15, 24, 28, 30, 99, -1):
        assert L.ggb_block_elems(t) == 0 and L.ggb_block_bytes(t) == 0


def test_q8_1_sizes(capi):
    L = capi.lib
    # ggml_kernel.cu:55,90: Kp = ceil(K/512)*512 ; mmq.cu:165-166,190-191: Kp = K - K%512 + 512
    assert L.ggb_q8_1_bytes(1, 4096, capi.GGB_Q8_1) == 4096 // 32 * 36
    assert L.ggb_q8_1_bytes(1, 4100, capi.GGB_Q8_1) == 4608 // 32 * 36
    assert L.ggb_q8_1_bytes(7, 4096, capi.GGB_Q8_1_MMQ) == 7 * 4608 // 32 * 36
    assert L.ggb_q8_1_bytes(7, 4100, capi.GGB_Q8_1_MMQ) == 7 * 4608 // 32 * 36
    assert L.ggb_q8_1_bytes(0, 4096, capi.GGB_Q8_1) == 0


def test_argument_errors_without_gpu(capi):
    L = capi.lib
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.addressof(buf)
    p16 = (p + 15) & ~15
    assert L.ggb_dequantize(p16, 99, 1, 256, p16, capi.GGB_F16, None) == -1       # GGB_E_TYPE
    assert b"unsupported" in L.ggb_last_error()
    assert L.ggb_dequantize(p16, 12, 1, 100, p16, capi.GGB_F16, None) == -2       # GGB_E_SHAPE
    assert L.ggb_dequantize(p16, 12, 1, 256, p16, 7, None) == -3                  # GGB_E_DTYPE
    assert L.ggb_dequantize(None, 12, 1, 256, p16, capi.GGB_F16, None) == -4      # GGB_E_ARG
    assert L.ggb_dequantize(p16, 12, 1, 256, p16 + 2, capi.GGB_F16, None) == -4   # misaligned out
    assert L.ggb_dequantize(p16, 12, 0, 256, p16, capi.GGB_F16, None) == 0        # empty is a no-op
    assert L.ggb_last_error() == b""
    assert L.ggb_mmvq(p16, p16, p16, 12, capi.GGB_F16, 9, 256, 4, None, 0, None) == -2
    assert L.ggb_mmvq(p16, p16, p16, 12, capi.GGB_F16, 1, 100, 4, None, 0, None) == -2
    assert L.ggb_mmq(p16, p16, p16, 5, capi.GGB_F16, 1, 256, 4, 4, None, 0, None) == -1
    assert L.ggb_mmq(p16, p16, p16, 12, capi.GGB_F16, 4, 256, 8, 4, p16, 1 << 20, None) == -2  # ldy < N
    assert L.ggb_quantize_q8_1(p16, capi.GGB_F32, 1, 32, p16, 5, 12, None) == -4


def test_python_wrappers_reject_cpu_tensors():
    import torch
    import ggml
    w = torch.zeros(144, dtype=torch.uint8)
    with pytest.raises(ValueError):
        ggml.ggml_dequantize(w, 12, 1, 256)
    with pytest.raises(ValueError):
        ggml.ggml_mul_mat_vec_a8(w, torch.zeros(9, 256), 12, 1)


def test_fake_tensor_shapes():
    """register_fake lets meta tracing see the ops (the reference registers CUDA impls only)."""
    import torch
    import ggml  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        W = torch.empty(8 * 144, dtype=torch.uint8, device="cuda")
        X = torch.empty(3, 5, 256, dtype=torch.bfloat16, device="cuda")
        y = torch.ops.ggml_b200.ggml_mul_mat_a8(W, X, 12, 8)
        assert y.shape == (3, 5, 8) and y.dtype == torch.bfloat16
        d = torch.ops.ggml_b200.ggml_dequantize(W, 12, 8, 256)           # the reference's 4-argument schema, fp16 out
        assert d.shape == (8, 256) and d.dtype == torch.float16
        d = torch.ops.ggml_b200.ggml_dequantize_dtype(W, 12, 8, 256, 2)  # extension: fused cast (GGB_BF16)
        assert d.shape == (8, 256) and d.dtype == torch.bfloat16
        v = torch.ops.ggml_b200.ggml_mul_mat_vec_a8(W, X[0, :2], 12, 8)
        assert v.shape == (2, 8)


def test_reference_schemas_and_namespace():
    """The three ops carry the reference's schemas (torch-ext/torch_binding.cpp:6-21) and can be exposed under its
    development namespace `_ggml` as well."""
    import torch
    import ggml
    want = {"ggml_dequantize": "(Tensor W, int type, SymInt m, SymInt n) -> Tensor",
            "ggml_mul_mat_vec_a8": "(Tensor W, Tensor X, int type, SymInt row) -> Tensor",
            "ggml_mul_mat_a8": "(Tensor W, Tensor X, int type, SymInt row) -> Tensor"}
    if not hasattr(torch.ops._ggml, "ggml_dequantize"):
        ggml.register_reference_namespace("_ggml")
    for name, sig in want.items():
        ours = str(getattr(torch.ops.ggml_b200, name).default._schema)
        assert ours == f"ggml_b200::{name}{sig}", ours
        ref = str(getattr(torch.ops._ggml, name).default._schema)
        assert ref.split("::")[1] == f"{name}{sig}", ref


def test_fused_supported_query(capi):
    L = capi.lib
    assert L.ggb_mmvq_fused_supported(12, 1, 8192) == 1 and L.ggb_mmvq_fused_supported(14, 8, 28672) == 1
    assert L.ggb_mmvq_fused_supported(14, 1, 4352) == 0          # Q6_K rows not 16-byte aligned
    assert L.ggb_mmvq_fused_supported(2, 1, 4096) == 0 and L.ggb_mmvq_fused_supported(12, 9, 4096) == 0


def test_iq4_types_and_options_without_gpu(capi):
    """The IQ types: geometry known, MMQ refused (like the reference, which has no MMQ case for them); every tuning
    option is accepted and restored, an unknown one is an argument error."""
    L = capi.lib
    assert (L.ggb_block_elems(20), L.ggb_block_bytes(20)) == (32, 18)
    assert (L.ggb_block_elems(23), L.ggb_block_bytes(23)) == (256, 136)
    buf = ctypes.create_string_buffer(8192)
    p16 = (ctypes.addressof(buf) + 255) & ~255
    for t in (20, 23, 16, 17, 18, 19, 21, 22, 29):
        assert L.ggb_mmq(p16, p16, p16, t, capi.GGB_F16, 4, 256, 8, 8, p16, 1 << 20, None) == -1   # GGB_E_TYPE
        assert b"MMVQ only" in L.ggb_last_error()
        assert L.ggb_mmq_workspace_bytes(t, 4, 256, 8) == 0
    for opt, val, restore in ((6, 0, -1), (6, 1024, -1), (4, 8, 0), (1, 1, 0)):
        assert L.ggb_set_option(opt, val) == 0
        assert L.ggb_set_option(opt, restore) == 0
    assert L.ggb_set_option(99, 1) < 0
