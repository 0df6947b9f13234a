"""GGUF file -> device tensors, with per-rank row slicing (the step just before the hot
path; the reference does it inline in its tests, tests/utils.py:25-31 and benchmarks/benchmark_mmq.py:61-66).

`gguf.GGUFReader` memory-maps the file and yields every quantized tensor as raw block bytes `uint8 [rows, row_bytes]`
-- exactly the `W` the ops take.  Because quant blocks never straddle rows, rank r of a column-parallel layer needs
only the contiguous byte range of rows `[r * N/tp, (r+1) * N/tp)`: the loader copies just that range from the mapping
into a pinned staging buffer and issues one asynchronous H2D copy per tensor on the current stream.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional

import numpy as np
import torch

from . import SUPPORTED_TYPES, block_bytes, block_elems


@dataclass
class QuantTensor:
    name: str
    data: torch.Tensor          # uint8 [rows_local, row_bytes]: this rank's rows, raw GGUF blocks
    quant_type: int             # ggml_type id
    rows: int                   # rows of the FULL tensor (N)
    cols: int                   # K
    row_begin: int              # first row held by this rank

    @property
    def rows_local(self) -> int:
        return int(self.data.shape[0])


def _shape_of(t) -> tuple[int, int]:
    """(rows, cols) of a GGUF weight tensor: GGUF stores dimensions innermost first.  Leading dimensions of a 3-D tensor
    (the expert stack of an MoE layer, [experts, rows, cols]) are flattened into rows: every (expert, row) is an
    independent weight row of K = cols values."""
    dims = [int(d) for d in t.shape]
    if len(dims) < 2:
        raise ValueError(f"{t.name}: not a matrix (dims {dims})")
    rows = 1
    for d in dims[1:]:
        rows *= d
    return rows, dims[0]


# GGUF tensor types that are not quantized blocks (norms, biases, embeddings kept in float): never an error to skip them
_FLOAT_TYPES = {0, 1, 30, 24, 25, 26, 27, 28}          # F32, F16, BF16, I8, I16, I32, I64, F64


class UnsupportedQuantTypes(ValueError):
    """Raised by load_gguf(strict=True) when the file holds quantized tensors of types this library does not serve."""


def load_gguf(path: str, device: torch.device | str = "cuda", rank: int = 0, world: int = 1,
              shard: Optional[Callable[[str], bool]] = None, names: Optional[Iterable[str]] = None,
              strict: bool = True, skipped: Optional[list] = None) -> Dict[str, QuantTensor]:
    """Load the quantized weight tensors of a GGUF file.

    `shard(name)` says which tensors are column-parallel (row-sliced over `world` ranks; default: all of them when
    world > 1); the others are loaded whole on every rank.  `names` restricts the load.  Float tensors (F32 / F16 / BF16
    norms and embeddings) are skipped silently.  QUANTIZED tensors of a type outside SUPPORTED_TYPES (the reference also
    serves IQ1 / IQ2 / IQ3, ggml ids 16-19, 21, 22, 29) are never dropped silently: with `strict` (default) the load
    raises UnsupportedQuantTypes listing them; with strict=False their (name, type id) pairs are appended to `skipped`.
    """
    import gguf
    device = torch.device(device)
    wanted = set(names) if names is not None else None
    out: Dict[str, QuantTensor] = {}
    missing = []
    pending = []
    for t in gguf.GGUFReader(path).tensors:
        qt = int(t.tensor_type)
        if wanted is not None and t.name not in wanted:
            continue
        if qt not in SUPPORTED_TYPES:
            if qt not in _FLOAT_TYPES:
                missing.append((t.name, qt))
            continue
        rows, cols = _shape_of(t)
        qk, bs = block_elems(qt), block_bytes(qt)
        if cols % qk:
            raise ValueError(f"{t.name}: K={cols} is not a multiple of the {qk}-value block")
        row_bytes = cols // qk * bs
        raw = np.asarray(t.data).reshape(rows, row_bytes)          # a view of the memory map
        lo, hi = 0, rows
        if world > 1 and (shard is None or shard(t.name)):
            if rows % world:
                raise ValueError(f"{t.name}: N={rows} is not divisible by the tensor-parallel size {world}")
            per = rows // world
            lo, hi = rank * per, (rank + 1) * per
        host = torch.from_numpy(np.array(raw[lo:hi], copy=True))    # reads only this rank's byte range from the file
        if device.type == "cuda":
            host = host.pin_memory()
            data = host.to(device, non_blocking=True)              # the copy of tensor i overlaps the file read of tensor i + 1
            pending.append(host)
        else:
            data = host
        out[t.name] = QuantTensor(t.name, data, qt, rows, cols, lo)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()            # the pinned staging buffers may now be released
        pending.clear()
    if missing:
        if strict:
            raise UnsupportedQuantTypes("quantized tensors of unsupported ggml types (name, type id): " + ", ".join(f"{n}:{q}" for n, q in missing))
        if skipped is not None:
            skipped.extend(missing)
    return out
