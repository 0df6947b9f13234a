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
    """(rows, cols) of a 2-D GGUF tensor: GGUF stores dimensions innermost first."""
    dims = [int(d) for d in t.shape]
    if len(dims) != 2:
        raise ValueError(f"{t.name}: only 2-D weight tensors are quantized matrices (got {dims})")
    return dims[1], dims[0]


def load_gguf(path: str, device: torch.device | str = "cuda", rank: int = 0, world: int = 1,
              shard: Optional[Callable[[str], bool]] = None, names: Optional[Iterable[str]] = None) -> Dict[str, QuantTensor]:
    """Load the quantized 2-D tensors of a GGUF file.

    `shard(name)` says which tensors are column-parallel (row-sliced over `world` ranks; default: all of them when
    world > 1); the others are loaded whole on every rank.  `names` restricts the load.  Tensors of types outside
    the supported set (F32/F16 norms, IQ types) are skipped.
    """
    import gguf
    device = torch.device(device)
    wanted = set(names) if names is not None else None
    out: Dict[str, QuantTensor] = {}
    for t in gguf.GGUFReader(path).tensors:
        qt = int(t.tensor_type)
        if qt not in SUPPORTED_TYPES or (wanted is not None and t.name not in wanted):
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
            data = host.to(device, non_blocking=True)
        else:
            data = host
        out[t.name] = QuantTensor(t.name, data, qt, rows, cols, lo)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()            # the pinned staging buffers may now be released
    return out
