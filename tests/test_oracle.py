"""CPU tests pinning the oracle (oracle/ggml_oracle.c) to the reference.

  * dequantize: bit-exact vs the committed golden vectors (made by gguf-py + the reference's own
    ggml-cpu op, tests/golden/make_golden.py), vs gguf-py live, and vs oracle/_ref (the reference
    CPU op compiled from /root/reference) where it is present;
  * q8_1 / q8_1_mmq: the reference has no CPU implementation and no test of the bytes ("parity
    unpinned"); checked here against an independent pure-Python restatement of
    ggml_kernel.cu:13-50 / mmq.cu:109-154 on small cases and by reconstruction error bounds;
  * MMVQ / MMQ evaluators: within the reference's own test tolerances of x @ dequant(W).T.
"""
import math
import struct

import numpy as np
import pytest

import oracle
from util import DEQUANT_ATOL, DEQUANT_RTOL, GOLDEN, MMVQ_ATOL, MMVQ_RTOL, assert_bits_equal

gguf = pytest.importorskip("gguf")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("t", oracle.VEC_TYPES, ids=lambda t: oracle.TYPE_NAMES[t])
def test_dequant_golden(golden, t):
    name = oracle.TYPE_NAMES[t]
    for tag in ("", "_edge"):
        w, y = golden[f"{name}{tag}_w"], golden[f"{name}{tag}_y"]
        got = oracle.dequantize(w, t, y.shape[0], y.shape[1])
        assert_bits_equal(got, y, f"{name}{tag}")
    if t in oracle.LEGACY_TYPES:
        assert_bits_equal(oracle.dequantize(golden[f"{name}_w"], t, *golden[f"{name}_refcpu_y"].shape),
                          golden[f"{name}_refcpu_y"], f"{name} vs ggml-cpu golden")


@pytest.mark.parametrize("t", oracle.VEC_TYPES, ids=lambda t: oracle.TYPE_NAMES[t])
@pytest.mark.parametrize("shape", [(1, 256), (3, 1024), (256, 1024), (7, 4096)])
def test_dequant_vs_gguf_py(t, shape):
    rows, cols = shape
    w = oracle.random_weights(t, rows, cols, seed=rows * 31 + cols)
    want = gguf.quants.dequantize(w, gguf.GGMLQuantizationType(t)).astype(np.float32)
    assert_bits_equal(oracle.dequantize(w, t, rows, cols), want, oracle.TYPE_NAMES[t])
    assert_bits_equal(oracle.dequantize(w, t, rows, cols, threads=True), want, oracle.TYPE_NAMES[t] + " mt")


@pytest.mark.parametrize("t", oracle.LEGACY_TYPES, ids=lambda t: oracle.TYPE_NAMES[t])
def test_dequant_vs_reference_cpu_op(t):
    ref = oracle.ref_custom_ops()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import torch
    rows, cols = 256, 1024          # the reference's own fixture sizes (tests/test_dequantize.py:29)
    rng = np.random.default_rng(t)
    w = gguf.quants.quantize((rng.standard_normal((rows, cols)) * 0.02).astype(np.float32), gguf.GGMLQuantizationType(t))
    want = ref.ggml_dequantize(torch.from_numpy(w.copy()), t, rows, cols).numpy()
    got = oracle.dequantize(w, t, rows, cols)
    assert_bits_equal(got, want, oracle.TYPE_NAMES[t])
    # and the reference's own tolerance check against gguf-py (tests/test_dequantize.py:71)
    np.testing.assert_allclose(got, gguf.quants.dequantize(w, gguf.GGMLQuantizationType(t)), atol=DEQUANT_ATOL, rtol=DEQUANT_RTOL)


def test_dequant_empty_and_errors():
    assert oracle.dequantize(np.zeros((0,), np.uint8), oracle.Q4_K, 0, 256).shape == (0, 256)
    with pytest.raises(ValueError):
        oracle.dequantize(np.zeros(10, np.uint8), oracle.Q4_K, 1, 256)


# ------------------------------------------------------------------ q8_1: independent pure-Python restatement

def _f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def _py_q8_1_group(vals):
    """ggml_kernel.cu:26-49 with one 'lane' per value, butterfly masks 16,8,4,2,1, all in fp32."""
    v = [_f32(x) for x in vals]
    a = [abs(x) for x in v]
    s = list(v)
    for mask in (16, 8, 4, 2, 1):
        a = [max(a[i], a[i ^ mask]) for i in range(32)]
        s = [_f32(s[i] + s[i ^ mask]) for i in range(32)]
    amax = a[0]
    d = _f32(amax / 127.0) if amax != 0 else 0.0
    # float division then roundf (half away from zero)
    q = []
    for x in v:
        if amax == 0:
            q.append(0)
        else:
            r = float(np.float32(x) / np.float32(d))
            q.append(int(math.copysign(math.floor(abs(r) + 0.5), r)))
    return d, s[0], q


@pytest.mark.parametrize("K", [32, 96, 512, 544])
def test_q8_1_vs_python_restatement(K):
    rng = np.random.default_rng(K)
    x = (rng.standard_normal((3, K)) * 3).astype(np.float32)
    x[1, :32] = 0.0                                     # amax == 0 group
    x[2, 5] = 1e-30
    out = oracle.quantize_q8_1(x)
    Kp = (K + 511) // 512 * 512
    assert out.shape == (3, Kp // 32 * 36)
    for r in range(3):
        for g in range(Kp // 32):
            blk = out[r, g * 36:(g + 1) * 36]
            vals = [float(x[r, g * 32 + i]) if g * 32 + i < K else 0.0 for i in range(32)]
            d, s, q = _py_q8_1_group(vals)
            assert blk[:2].view(np.float16)[0] == np.float16(d) or (np.isnan(np.float16(d)))
            assert blk[2:4].view(np.float16)[0] == np.float16(np.float32(s))
            assert list(blk[4:].view(np.int8)) == q


@pytest.mark.parametrize("need_sum", [True, False])
def test_q8_1_mmq_layout(need_sum):
    rng = np.random.default_rng(7)
    M, K = 5, 640
    x = rng.standard_normal((M, K)).astype(np.float32)
    out = oracle.quantize_q8_1_mmq(x, need_sum)
    Kp = K - K % 512 + 512
    assert out.shape == (Kp // 128, M, 144)
    plain = oracle.quantize_q8_1(x)
    for r in range(M):
        for c in range(Kp // 128):
            for g in range(4):
                gi = 4 * c + g
                qs = out[c, r, 16 + 32 * g:16 + 32 * g + 32]
                if gi * 32 < (K + 511) // 512 * 512:
                    pb = plain[r, gi * 36:(gi + 1) * 36]
                    assert np.array_equal(qs, pb[4:])
                    if need_sum:
                        assert np.array_equal(out[c, r, 4 * g:4 * g + 4], pb[:4])
                    else:
                        d32 = out[c, r, 4 * g:4 * g + 4].view(np.float32)[0]
                        assert np.float16(d32) == pb[:2].view(np.float16)[0]
                else:
                    assert not qs.any()


def test_q8_1_reconstruction_bound():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 4096)).astype(np.float32)
    out = oracle.quantize_q8_1(x).reshape(4, -1, 36)
    d = out[:, :, :2].copy().view(np.float16).astype(np.float32)[..., 0]
    q = out[:, :, 4:].view(np.int8).astype(np.float32)
    rec = (q * d[..., None]).reshape(4, -1)[:, :4096]
    amax = np.abs(x.reshape(4, -1, 32)).max(-1, keepdims=True)
    err = np.abs(rec - x).reshape(4, -1, 32)
    assert (err <= amax / 127 * 0.51 + amax * 1e-3).all()


# ------------------------------------------------------------------ MMVQ / MMQ evaluators

@pytest.mark.parametrize("t", oracle.VEC_TYPES, ids=lambda t: oracle.TYPE_NAMES[t])
def test_mmvq_oracle_vs_dense(t):
    rng = np.random.default_rng(0)
    N, K = 96, 1024
    w = oracle.random_weights(t, N, K, seed=t)
    x = rng.random((2, K), dtype=np.float32)             # torch.rand-like, as test_cuda_kernels.py:69
    y = oracle.mmvq(w, t, x, N)
    want = x.astype(np.float64) @ oracle.dequantize(w, t, N, K).astype(np.float64).T
    np.testing.assert_allclose(y, want, atol=MMVQ_ATOL, rtol=MMVQ_RTOL)
    # q8_1 rounding noise only: relative to the output scale it is small
    scale = np.abs(want).max() + 1e-9
    assert np.abs(y - want).max() / scale < 2e-2


@pytest.mark.parametrize("t", oracle.ALL_TYPES, ids=lambda t: oracle.TYPE_NAMES[t])
def test_mmq_oracle_vs_dense(t):
    rng = np.random.default_rng(1)
    N, K, M = 64, 512, 7
    w = oracle.random_weights(t, N, K, seed=50 + t)
    x = rng.random((M, K), dtype=np.float32)
    y = oracle.mmq(w, t, x, N)
    want = x.astype(np.float64) @ oracle.dequantize(w, t, N, K).astype(np.float64).T
    np.testing.assert_allclose(y, want, atol=1.0, rtol=1e-1)   # test_cuda_kernels.py:122-129 (half)
    scale = np.abs(want).max() + 1e-9
    assert np.abs(y - want).max() / scale < 2e-2


def test_mmq_oracle_has_no_iq4():
    """the reference's ggml_mul_mat_a8 has no IQ case (mmq.cu:222-252); neither has the oracle"""
    w = oracle.random_weights(oracle.IQ4_XS, 8, 512, seed=0)
    with pytest.raises(ValueError):
        oracle.mmq(w, oracle.IQ4_XS, np.zeros((2, 512), np.float32), 8)


def test_iq_codebooks_are_pinned_to_gguf_py():
    """The codebook tables of the product (csrc/iq_tables.cuh) and of the oracle (oracle/iq_tables.h) are generated files: both must
    hold exactly what scripts/gen_iq_tables.py derives from gguf-py today (IQ3_S: with the reference's top level, see the
    generator), entry for entry."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_iq_tables", os.path.join(root, "scripts", "gen_iq_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    for path in (os.path.join(root, "ggml-libtorch_b200", "csrc", "iq_tables.cuh"), os.path.join(root, "oracle", "iq_tables.h")):
        txt = open(path).read()
        for name, cls, remap in gen.TABLES:
            want, _ = gen.grid_of(cls, remap)
            m = re.search(r"%s\[(\d+)\]\s*=\s*\{(.*?)\};" % name, txt, re.S)
            assert m, f"{name} missing in {path}"
            got = np.array([int(v) for v in re.findall(r"(\d+)[uU]", m.group(2))], dtype=np.uint64)
            assert int(m.group(1)) == len(want) and np.array_equal(got, want), f"{name} in {os.path.basename(path)}"
