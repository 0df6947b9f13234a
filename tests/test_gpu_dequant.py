"""GPU parity: ggml_dequantize through the torch op -> C ABI -> sm_100a kernels, BIT-EXACT against
the CPU oracle (fp32) and against one RN cast of it (fp16/bf16), for every supported type.

Also runs the reference's own dequantize tests (tests/test_dequantize.py:54-71,
hf-kernels/ggml-kernels/tests/kernels/test_cuda_kernels.py:39-58) on synthetic GGUF files named the
reference's way."""
import numpy as np
import pytest
import torch

import oracle
from util import DEQUANT_ATOL, DEQUANT_RTOL, GOLDEN, assert_bits_equal, bf16_bits_from_f32

pytestmark = pytest.mark.gpu

TYPES = oracle.VEC_TYPES
IDS = [oracle.TYPE_NAMES[t] for t in TYPES]
DTYPES = [torch.float32, torch.float16, torch.bfloat16]


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _expect(y32: np.ndarray, dtype):
    if dtype == torch.float32:
        return y32
    if dtype == torch.float16:
        with np.errstate(over="ignore"):
            return y32.astype(np.float16)
    return bf16_bits_from_f32(y32)


def _check(ops, w: np.ndarray, t: int, m: int, n: int, dtype, what: str):
    want32 = oracle.dequantize(w, t, m, n)
    got = ops.ggml_dequantize(torch.from_numpy(np.ascontiguousarray(w)).cuda(), t, m, n, dtype=dtype)
    assert got.shape == (m, n) and got.dtype == dtype
    want = _expect(want32, dtype)
    if dtype == torch.bfloat16:
        g = got.view(torch.int16).cpu().numpy().view(np.uint16)
        nan = np.isnan(want32)
        ok = (g == want) | (nan & ((g & 0x7FFF) > 0x7F80))
        assert ok.all(), f"{what}: {np.count_nonzero(~ok)} bf16 values differ"
    else:
        assert_bits_equal(got.cpu().numpy(), want, what)


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_golden(ops, t, dtype):
    g = np.load(GOLDEN)
    name = oracle.TYPE_NAMES[t]
    for tag in ("", "_edge"):
        w, y = g[f"{name}{tag}_w"], g[f"{name}{tag}_y"]
        got = ops.ggml_dequantize(torch.from_numpy(w.copy()).cuda(), t, y.shape[0], y.shape[1], dtype=torch.float32)
        assert_bits_equal(got.cpu().numpy(), y, f"{name}{tag} golden")
        _check(ops, w, t, y.shape[0], y.shape[1], dtype, f"{name}{tag}")


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
@pytest.mark.parametrize("shape", [(1, 256), (3, 768), (256, 1024), (33, 8192), (5, 14336)])
def test_random_shapes(ops, t, shape, dtype):
    m, n = shape
    w = oracle.random_weights(t, m, n, seed=m + n + t)
    _check(ops, w, t, m, n, dtype, f"{oracle.TYPE_NAMES[t]} {shape}")


@pytest.mark.parametrize("t", oracle.BLOCK32_TYPES, ids=[oracle.TYPE_NAMES[t] for t in oracle.BLOCK32_TYPES])
def test_legacy_block_granularity(ops, t):
    """legacy types only need n % 32 == 0: rows that are not a multiple of the 8192-value tile."""
    for m, n in [(1, 32), (7, 96), (129, 160)]:
        w = oracle.random_weights(t, m, n, seed=n)
        _check(ops, w, t, m, n, torch.float32, f"{oracle.TYPE_NAMES[t]} [{m},{n}]")


def test_unaligned_weight_pointer(ops):
    """W as a byte-offset view (2-byte aligned only) takes the non-TMA path and stays exact."""
    t, m, n = oracle.Q4_K, 40, 1024
    w = oracle.random_weights(t, m, n, seed=9)
    buf = torch.zeros(w.size + 64, dtype=torch.uint8, device="cuda")
    buf[2:2 + w.size] = torch.from_numpy(w.reshape(-1)).cuda()
    got = ops.ggml_dequantize(buf[2:2 + w.size], t, m, n, dtype=torch.float32)
    assert_bits_equal(got.cpu().numpy(), oracle.dequantize(w, t, m, n), "unaligned W")


def test_empty_and_errors(ops):
    e = ops.ggml_dequantize(torch.empty(0, dtype=torch.uint8, device="cuda"), 12, 0, 256)
    assert e.shape == (0, 256) and e.dtype == torch.float16
    w = torch.zeros(144, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.ggml_dequantize(w, 12, 1, 512)          # wrong byte count
    with pytest.raises(ValueError):
        ops.ggml_dequantize(w, 16, 1, 256)          # IQ2_XXS: codebook IQ types are not built (remainder)
    with pytest.raises(ValueError):
        ops.ggml_dequantize(w.view(torch.int8), 12, 1, 256)


@pytest.mark.parametrize("t", [oracle.Q4_0, oracle.Q8_0], ids=["Q4_0", "Q8_0"])
def test_config1_full_size(ops, t):
    """BASELINE config 1: 4096x4096 Q4_0/Q8_0, quantized from randn*0.02 by gguf-py, vs the oracle."""
    gguf = pytest.importorskip("gguf")
    rng = np.random.default_rng(0)
    w = gguf.quants.quantize((rng.standard_normal((4096, 4096)) * 0.02).astype(np.float32), gguf.GGMLQuantizationType(t))
    _check(ops, w, t, 4096, 4096, torch.float32, "config1")


@pytest.mark.parametrize("t", oracle.K_TYPES + (oracle.IQ4_XS,), ids=[oracle.TYPE_NAMES[t] for t in oracle.K_TYPES + (oracle.IQ4_XS,)])
def test_config2_full_size(ops, t):
    """BASELINE config 2: 4096x14336 K-quants to fp16/bf16, bit-exact vs the (threaded) oracle."""
    m, n = 4096, 14336
    w = oracle.random_weights(t, m, n, seed=0)
    want32 = oracle.dequantize(w, t, m, n, threads=True)
    wd = torch.from_numpy(w).cuda()
    got16 = ops.ggml_dequantize(wd, t, m, n).cpu().numpy()
    assert_bits_equal(got16, want32.astype(np.float16), "fp16")
    gotb = ops.ggml_dequantize(wd, t, m, n, dtype=torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(gotb, bf16_bits_from_f32(want32))
    # size-independent property: fp32 output cast on the GPU equals the fused-cast output
    got32 = ops.ggml_dequantize(wd, t, m, n, dtype=torch.float32)
    assert torch.equal(got32.to(torch.float16), torch.from_numpy(got16).cuda())


# ---------------------------------------------------------------- the reference's own test shape

def _write_gguf(path, t, hidden):
    import gguf
    qt = gguf.GGMLQuantizationType(t)
    wr = gguf.GGUFWriter(path, "llama")
    for i, rows in enumerate((hidden, 2 * hidden)):
        if t in oracle.LEGACY_TYPES:
            rng = np.random.default_rng(i)
            data = gguf.quants.quantize((rng.standard_normal((rows, hidden)) * 0.02).astype(np.float32), qt)
        else:
            data = oracle.random_weights(t, rows, hidden, seed=i)
        wr.add_tensor(f"tensor_{qt.name}_{rows}x{hidden}", data, raw_dtype=qt)
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()


@pytest.mark.parametrize("hidden_size", [256, 1024])
@pytest.mark.parametrize("dtype", [torch.half, torch.bfloat16, torch.float32], ids=["half", "bf16", "f32"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_reference_style_dequantize(ops, tmp_path, t, dtype, hidden_size):
    gguf = pytest.importorskip("gguf")
    qt = gguf.GGMLQuantizationType(t)
    path = str(tmp_path / f"Quant_{qt.name}_{hidden_size}.gguf")
    _write_gguf(path, t, hidden_size)
    for tensor in gguf.GGUFReader(path).tensors:
        shape = list(map(int, tensor.name.split("_")[-1].split("x")))
        ref_output = torch.tensor(gguf.dequantize(tensor.data, qt), device="cuda").to(dtype)
        output = ops.ggml_dequantize(torch.tensor(tensor.data, device="cuda"), qt, *shape).to(dtype)
        torch.testing.assert_close(output, ref_output, atol=DEQUANT_ATOL, rtol=DEQUANT_RTOL)
        # and the stronger contract: the fused-dtype output is bit-identical to the cast of gguf-py's fp32
        fused = ops.ggml_dequantize(torch.tensor(tensor.data, device="cuda"), qt, *shape, dtype=dtype)
        assert torch.equal(fused, ref_output)


@pytest.mark.parametrize("t", (oracle.Q4_0, oracle.Q4_K, oracle.Q6_K, oracle.IQ4_XS, oracle.IQ2_XS, oracle.IQ1_M),
                         ids=lambda t: oracle.TYPE_NAMES[t])
def test_static_weights_option_gives_the_same_bits(ops, t):
    """GGB_OPT_STATIC_WEIGHTS = 1: the tile loads start before the previous kernel of the stream has finished; the values must
    not change (back-to-back launches into different outputs)."""
    m, n = 96, 2048
    w = torch.from_numpy(oracle.random_weights(t, m, n, seed=t)).cuda()
    want = ops.ggml_dequantize(w, t, m, n)
    torch.cuda.synchronize()
    ops.set_option(ops.GGB_OPT_STATIC_WEIGHTS, 1)
    try:
        got = [ops.ggml_dequantize(w, t, m, n) for _ in range(4)]
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.GGB_OPT_STATIC_WEIGHTS, 0)
    for g in got:
        assert torch.equal(g, want)
