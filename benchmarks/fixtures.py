"""Synthetic stand-ins for the reference's GGUF sample files (benchmarks/utils.py:21-31 downloads
`Isotr0py/test-gguf-sample` from the hub; there is no network here): `Quant_<TYPE>_<hidden>.gguf`, tensors named
`..._<rows>x<cols>` exactly as the reference's callers parse them (benchmark_mmq.py:68-73), written once into a cache
directory with random *valid* blocks (ggml.synth)."""
from __future__ import annotations

import os
import random
from pathlib import Path

import numpy as np
import torch
from gguf import GGMLQuantizationType, GGUFReader, GGUFWriter, ReaderTensor

ROOT = Path(__file__).resolve().parent.parent
CACHE = Path(os.environ.get("GGML_B200_FIXTURES", ROOT / "benchmarks" / "_fixtures"))


def seed_everything(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def sample_file(hidden_size: int, quant_type: GGMLQuantizationType) -> Path:
    """Path of Quant_<TYPE>_<hidden>.gguf, creating it on first use: an attention-like [h, h] and two MLP-like
    [4h, h] / [h, 4h] matrices of the type."""
    from ggml import synth
    CACHE.mkdir(parents=True, exist_ok=True)
    path = CACHE / f"Quant_{quant_type.name}_{hidden_size}.gguf"
    if path.exists():
        return path
    tmp = path.with_suffix(".tmp")
    wr = GGUFWriter(str(tmp), "llama")
    for i, (rows, cols) in enumerate(((hidden_size, hidden_size), (4 * hidden_size, hidden_size), (hidden_size, 4 * hidden_size))):
        wr.add_tensor(f"tensor_{quant_type.name}_{rows}x{cols}", synth.random_blocks(int(quant_type), rows, cols, seed=i),
                      raw_dtype=quant_type)
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()
    tmp.rename(path)
    return path


def get_gguf_sample_tensors(hidden_size: int, quant_type: GGMLQuantizationType) -> list[ReaderTensor]:
    """Same name and return type as the reference's helper (benchmarks/utils.py:25-31)."""
    return GGUFReader(sample_file(hidden_size, quant_type)).tensors
