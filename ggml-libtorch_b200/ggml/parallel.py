"""Column-parallel quantized linear layers over one 8xB200 box (new -- the reference is
single-GPU).

GGUF blocks run along K inside a weight row, so a contiguous slice of rows `W[r*N/tp : (r+1)*N/tp]`
is itself a valid GGUF tensor: rank r keeps only those bytes, computes `Y_r[M, N/tp]` from the
replicated activations with the same MMVQ / MMQ kernels, and ONE exchange -- an all-gather of the
output slices over NVLink (NCCL via torch.distributed) -- rebuilds `Y[M, N]`.  Every output element
is produced by exactly one rank with the same K order as on one GPU, so the sharded result equals the
single-GPU result bit for bit.

One process per GPU; `group` is a torch.distributed process group (NCCL on GPUs; the gloo CPU tests
inject a CPU matmul through `matmul_fn`).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    if n_rows % world != 0:
        raise ValueError(f"N={n_rows} is not divisible by the tensor-parallel size {world}")
    per = n_rows // world
    return rank * per, (rank + 1) * per


def shard_rows(W: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank `rank`'s contiguous row slice of raw GGUF bytes W [N, row_bytes]."""
    if W.dim() != 2:
        raise ValueError("W must be [rows, row_bytes]")
    lo, hi = shard_bounds(W.shape[0], rank, world)
    return W[lo:hi].contiguous()


def _default_matmul(W: torch.Tensor, X: torch.Tensor, quant_type: int, rows: int) -> torch.Tensor:
    from . import MMVQ_MAX_BATCH, ggml_mul_mat_a8, ggml_mul_mat_vec_a8
    if X.shape[0] <= MMVQ_MAX_BATCH:
        return ggml_mul_mat_vec_a8(W, X, quant_type, rows)
    return ggml_mul_mat_a8(W, X, quant_type, rows)


class ColumnParallelQuantLinear:
    """y = x @ dequant(W).T with W [N, K] row-sharded over the ranks of `group`.

    `W_shard` is this rank's slice (shard_rows); forward() returns the full [M, N] on every rank.
    """

    def __init__(self, W_shard: torch.Tensor, quant_type: int, n_total: int, rank: int, world: int,
                 group: Optional[dist.ProcessGroup] = None,
                 matmul_fn: Callable[[torch.Tensor, torch.Tensor, int, int], torch.Tensor] = _default_matmul):
        lo, hi = shard_bounds(n_total, rank, world)
        if W_shard.shape[0] != hi - lo:
            raise ValueError(f"shard has {W_shard.shape[0]} rows, expected {hi - lo}")
        self.W = W_shard
        self.quant_type = int(quant_type)
        self.n_total, self.n_local = n_total, hi - lo
        self.rank, self.world, self.group = rank, world, group
        self.matmul_fn = matmul_fn

    def local(self, X: torch.Tensor) -> torch.Tensor:
        """This rank's output slice Y[:, lo:hi] (no communication)."""
        return self.matmul_fn(self.W, X, self.quant_type, self.n_local)

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        y_local = self.local(X)
        if self.world == 1:
            return y_local
        M = y_local.shape[0]
        if M == 1:
            # decode: the rank-major gather IS Y[1, N]
            out = torch.empty((1, self.n_total), dtype=y_local.dtype, device=y_local.device)
            dist.all_gather_into_tensor(out.view(self.world, self.n_local), y_local.reshape(1, -1), group=self.group)
            return out
        # rank-major gather [tp * M, N/tp] (dim-0 concatenation), then one interleaving copy to [M, N]
        gathered = torch.empty((self.world * M, self.n_local), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(gathered, y_local.contiguous(), group=self.group)
        return gathered.view(self.world, M, self.n_local).permute(1, 0, 2).reshape(M, self.n_total)

    __call__ = forward


class QuantMLP:
    """Llama-style MLP block on column-parallel quantized weights (BASELINE configs[4]):
    down( silu(gate(x)) * up(x) ), each projection sharded along its N with an all-gather of its output."""

    def __init__(self, gate: ColumnParallelQuantLinear, up: ColumnParallelQuantLinear, down: ColumnParallelQuantLinear):
        self.gate, self.up, self.down = gate, up, down

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        g = self.gate(X)
        u = self.up(X)
        h = torch.nn.functional.silu(g) * u
        return self.down(h)

    __call__ = forward
