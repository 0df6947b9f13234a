"""GGUF file -> tensors with per-rank row slicing (ggml/loader.py), on CPU: a synthetic GGUF written with the
reference's fixture naming (tests/utils.py:25-31: Quant_<TYPE>_<hidden>.gguf, tensors named ..._<rows>x<cols>) is read
back whole and in two shards; the shards are byte slices of the original rows and dequantize to the same values."""
import numpy as np
import pytest

import oracle

gguf = pytest.importorskip("gguf")


def _write(path, specs):
    wr = gguf.GGUFWriter(path, "llama")
    for name, t, rows, cols, seed in specs:
        wr.add_tensor(name, oracle.random_weights(t, rows, cols, seed=seed), raw_dtype=gguf.GGMLQuantizationType(t))
    wr.add_tensor("norm.weight", np.ones(256, dtype=np.float32))       # a non-quantized tensor: skipped by the loader
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()


def test_load_whole_and_sharded(tmp_path):
    torch = pytest.importorskip("torch")
    pytest.importorskip("ggml", reason="the package needs libggml_b200.so (block geometry comes from the C ABI)")
    from ggml.loader import load_gguf
    specs = [("blk.0.ffn_gate_512x256", oracle.Q4_K, 512, 256, 1), ("blk.0.ffn_down_256x512", oracle.Q6_K, 256, 512, 2),
             ("blk.0.attn_q_64x256", oracle.Q8_0, 64, 256, 3)]
    path = str(tmp_path / "Quant_mix_256.gguf")
    _write(path, specs)
    whole = load_gguf(path, device="cpu")
    assert set(whole) == {s[0] for s in specs}
    for name, t, rows, cols, seed in specs:
        w = oracle.random_weights(t, rows, cols, seed=seed)
        q = whole[name]
        assert (q.quant_type, q.rows, q.cols, q.row_begin) == (t, rows, cols, 0)
        assert np.array_equal(q.data.numpy(), w)
    parts = [load_gguf(path, device="cpu", rank=r, world=2, shard=lambda n: "ffn" in n) for r in (0, 1)]
    for name, t, rows, cols, seed in specs:
        w = oracle.random_weights(t, rows, cols, seed=seed)
        if "ffn" in name:
            got = np.concatenate([p[name].data.numpy() for p in parts])
            assert [p[name].row_begin for p in parts] == [0, rows // 2]
            assert np.array_equal(got, w)
            # a row slice is a valid GGUF tensor: it dequantizes to the slice of the full matrix
            full = oracle.dequantize(w, t, rows, cols)
            half = oracle.dequantize(parts[1][name].data.numpy(), t, rows // 2, cols)
            assert np.array_equal(full[rows // 2:], half)
        else:
            for p in parts:                                   # replicated
                assert np.array_equal(p[name].data.numpy(), w)
    with pytest.raises(ValueError):
        load_gguf(path, device="cpu", rank=0, world=3)


def test_unsupported_quant_types_are_not_dropped_silently(tmp_path):
    """Advisor finding (round 1): a model with tensors of an unserved quantized type (then IQ2 / IQ3; now e.g. the ternary TQ2_0)
    used to load with layers missing and no error."""
    pytest.importorskip("torch")
    pytest.importorskip("ggml", reason="the package needs libggml_b200.so")
    from ggml.loader import UnsupportedQuantTypes, load_gguf
    path = str(tmp_path / "mixed.gguf")
    wr = gguf.GGUFWriter(path, "llama")
    wr.add_tensor("blk.0.ffn_gate_64x256", oracle.random_weights(oracle.Q4_K, 64, 256, seed=1), raw_dtype=gguf.GGMLQuantizationType(12))
    tq2 = np.zeros((4, 256 // 256 * 66), dtype=np.uint8)                 # TQ2_0: 66 bytes per 256 values
    wr.add_tensor("blk.0.ffn_up_4x256", tq2, raw_dtype=gguf.GGMLQuantizationType.TQ2_0)
    wr.add_tensor("norm.weight", np.ones(256, dtype=np.float32))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    with pytest.raises(UnsupportedQuantTypes, match="ffn_up"):
        load_gguf(path, device="cpu")
    skipped = []
    got = load_gguf(path, device="cpu", strict=False, skipped=skipped)
    assert set(got) == {"blk.0.ffn_gate_64x256"} and [n for n, _ in skipped] == ["blk.0.ffn_up_4x256"]


def test_expert_stack_is_flattened(tmp_path):
    """A 3-D tensor (MoE expert stack) loads as experts * rows independent weight rows instead of aborting the file."""
    pytest.importorskip("torch")
    pytest.importorskip("ggml", reason="the package needs libggml_b200.so")
    from ggml.loader import load_gguf
    path = str(tmp_path / "moe.gguf")
    w = oracle.random_weights(oracle.Q4_K, 4 * 32, 256, seed=3)           # 4 experts x 32 rows x 256 columns
    wr = gguf.GGUFWriter(path, "llama")
    wr.add_tensor("blk.0.ffn_gate_exps", w.reshape(4, 32, -1), raw_dtype=gguf.GGMLQuantizationType(12))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    q = load_gguf(path, device="cpu")["blk.0.ffn_gate_exps"]
    assert (q.rows, q.cols) == (128, 256) and np.array_equal(q.data.numpy(), w)
