"""GGUF file -> device loader on the GPU (ggml/loader.py): pinned staging + asynchronous H2D, per-rank row slices; the
loaded tensors feed the ops directly (dequantize of a shard == the rows of the full matrix, bit for bit)."""
import numpy as np
import pytest
import torch

import oracle

gguf = pytest.importorskip("gguf")
pytestmark = pytest.mark.gpu


def test_load_gguf_cuda_whole_and_sharded(tmp_path):
    import ggml
    from ggml.loader import load_gguf
    specs = [("blk.0.ffn_gate_512x1024", oracle.Q4_K, 512, 1024, 1), ("blk.0.ffn_down_1024x512", oracle.Q6_K, 1024, 2048, 2),
             ("blk.0.attn_q_64x256", oracle.Q8_0, 64, 256, 3)]
    path = str(tmp_path / "Quant_mix.gguf")
    wr = gguf.GGUFWriter(path, "llama")
    for name, t, rows, cols, seed in specs:
        wr.add_tensor(name, oracle.random_weights(t, rows, cols, seed=seed), raw_dtype=gguf.GGMLQuantizationType(t))
    wr.add_tensor("norm.weight", np.ones(256, dtype=np.float32))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    whole = load_gguf(path, device="cuda")
    for name, t, rows, cols, seed in specs:
        q = whole[name]
        assert q.data.is_cuda and q.data.dtype == torch.uint8 and q.data.is_contiguous()
        assert np.array_equal(q.data.cpu().numpy(), oracle.random_weights(t, rows, cols, seed=seed))
    world = 4
    x = torch.randn(3, 1024, device="cuda", dtype=torch.float16)
    for name, t, rows, cols, seed in specs[:2]:
        full = ggml.ggml_dequantize(whole[name].data, t, rows, cols, dtype=torch.float32)
        y_full = ggml.ggml_mul_mat_vec_a8(whole[name].data, x[:, :cols].contiguous() if cols <= 1024 else torch.randn(3, cols, device="cuda", dtype=torch.float16), t, rows) if cols <= 1024 else None
        for r in range(world):
            part = load_gguf(path, device="cuda", rank=r, world=world, names=[name])[name]
            assert part.row_begin == r * rows // world and part.rows_local == rows // world
            got = ggml.ggml_dequantize(part.data, t, part.rows_local, cols, dtype=torch.float32)
            assert torch.equal(got, full[part.row_begin:part.row_begin + part.rows_local])
            if y_full is not None:
                y = ggml.ggml_mul_mat_vec_a8(part.data, x[:, :cols].contiguous(), t, part.rows_local)
                assert torch.equal(y, y_full[:, part.row_begin:part.row_begin + part.rows_local])
