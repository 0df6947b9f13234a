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

`fused_gather=True` (decode batches, K-quants): the exchange moves INTO the MMVQ kernel.  The full output lives in
a symmetric buffer (torch symmetric memory: the same allocation mapped into every rank over NVLink); each rank's
kernel stores its slice straight into every rank's copy (`ggb_mmvq_gather`), and one device-side barrier on the
stream replaces the NCCL all-gather (whose launch + protocol latency is several times the 114 KB .. 14 KB it moves
at decode).  Two buffers alternate, so a result stays valid until the second-next call of the layer.
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


class TpBarrier:
    """The barrier of the fused exchange as ONE programmatic-dependent launch of the library (ggb_tp_barrier) instead of
    a separate barrier kernel of the framework: a symmetric block of 16 uint32 per rank (slot r = the last epoch rank r
    reached, word 8 = this rank's epoch counter on the device, so CUDA-graph replays need nothing from the host).
    Collective construction (every rank of `group`)."""

    def __init__(self, rank: int, world: int, group, device: torch.device):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = rank, world
        self.flags = symm_mem.empty(16, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.hdl = symm_mem.rendezvous(self.flags, group if group is not None else dist.group.WORLD)
        torch.cuda.current_stream(device).synchronize()
        self.hdl.barrier(channel=0)                     # every rank's block is zero before anybody signals
        torch.cuda.current_stream(device).synchronize()
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._arr = (ctypes.c_void_p * world)(*ptrs)
        self.device = device

    def __call__(self) -> None:
        from . import _capi, _stream
        with torch.cuda.device(self.device):
            rc = _capi.lib.ggb_tp_barrier(self._arr, self.world, self.rank, _stream(self.flags))
        _capi.check(rc, "ggb_tp_barrier")

    def lost_peer(self) -> int:
        """0, or 1 + the rank that never arrived at some barrier (checked by tests; reading it synchronises)."""
        return int(self.flags[9].item())


class ColumnParallelQuantLinear:
    """y = x @ dequant(W).T with W [N, K] row-sharded over the ranks of `group`.

    `W_shard` is this rank's slice (shard_rows); forward() returns the full [M, N] on every rank.
    """

    def __init__(self, W_shard: torch.Tensor, quant_type: int, n_total: int, rank: int, world: int,
                 group: Optional[dist.ProcessGroup] = None,
                 matmul_fn: Callable[[torch.Tensor, torch.Tensor, int, int], torch.Tensor] = _default_matmul,
                 fused_gather: bool = False):
        lo, hi = shard_bounds(n_total, rank, world)
        if W_shard.shape[0] != hi - lo:
            raise ValueError(f"shard has {W_shard.shape[0]} rows, expected {hi - lo}")
        self.W = W_shard
        self.quant_type = int(quant_type)
        self.n_total, self.n_local = n_total, hi - lo
        self.rank, self.world, self.group = rank, world, group
        self.matmul_fn = matmul_fn
        self.fused_gather = bool(fused_gather) and world > 1
        self._symm = {}          # dtype -> [(buffer, handle, peer pointers)] x 2
        self._calls = 0
        self._ws = None
        self._barrier = None

    def _symm_buffers(self, dtype: torch.dtype, device: torch.device):
        """Two symmetric [MMVQ_MAX_BATCH, N] buffers per dtype (collective on first use: every rank calls it)."""
        if dtype not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem
            from . import MMVQ_MAX_BATCH
            group = self.group if self.group is not None else dist.group.WORLD
            bufs = []
            for _ in range(2):
                buf = symm_mem.empty((MMVQ_MAX_BATCH, self.n_total), dtype=dtype, device=device)
                hdl = symm_mem.rendezvous(buf, group)
                bufs.append((buf, hdl, [int(p) for p in hdl.buffer_ptrs]))
            self._symm[dtype] = bufs
        return self._symm[dtype]

    def _forward_fused(self, X: torch.Tensor) -> Optional[torch.Tensor]:
        from . import mmvq_workspace_bytes, mul_mat_vec_a8_gather
        M, K = X.shape
        buf, hdl, ptrs = self._symm_buffers(X.dtype, X.device)[self._calls & 1]
        need = mmvq_workspace_bytes(self.quant_type, M, K, self.n_local)
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = torch.empty(need, dtype=torch.uint8, device=X.device)
        # (eligibility was decided from (type, batch, K) alone in forward(): the same verdict on every rank.  Any error here
        # is a real one -- misaligned pointer, workspace -- and must surface instead of silently changing the exchange.)
        mul_mat_vec_a8_gather(self.W, X, self.quant_type, self.n_local, ptrs, self.n_total, self.rank * self.n_local, self._ws)
        self._calls += 1
        if self._barrier is None:
            import os
            self._barrier = (lambda: hdl.barrier(channel=0)) if os.environ.get("GGML_B200_TP_BARRIER", "") == "torch" else \
                TpBarrier(self.rank, self.world, self.group, X.device)
        self._barrier()                      # every rank's slice has landed in this rank's copy
        return buf[:M]

    def local(self, X: torch.Tensor) -> torch.Tensor:
        """This rank's output slice Y[:, lo:hi] (no communication)."""
        return self.matmul_fn(self.W, X, self.quant_type, self.n_local)

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if self.fused_gather and X.dim() == 2 and X.shape[0] <= 8:
            from . import mmvq_fused_supported
            if mmvq_fused_supported(self.quant_type, X.shape[0], X.shape[1]):
                return self._forward_fused(X)
        return self.gather(self.local(X))

    def gather(self, y_local: torch.Tensor) -> torch.Tensor:
        """All-gather of the ranks' column slices [M, N/tp] into the full [M, N] (NCCL over NVLink)."""
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
    down( silu(gate(x)) * up(x) ), each projection sharded along its N with an exchange of its output."""

    def __init__(self, gate: ColumnParallelQuantLinear, up: ColumnParallelQuantLinear, down: ColumnParallelQuantLinear):
        self.gate, self.up, self.down = gate, up, down

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if self.gate.world > 1 and not (self.gate.fused_gather and X.dim() == 2 and X.shape[0] <= 8):
            # SwiGLU is elementwise: each rank applies it to ITS column slices of g and u, and only h is exchanged --
            # one all-gather of [M, I] instead of two, and 1/tp of the elementwise work (bit-identical: the same ops
            # on the same values, whichever rank holds them)
            h_local = torch.nn.functional.silu(self.gate.local(X)) * self.up.local(X)
            return self.down(self.gate.gather(h_local))
        g = self.gate(X)
        u = self.up(X)
        h = torch.nn.functional.silu(g) * u
        return self.down(h)

    __call__ = forward


class FusedQuantMLP:
    """The same block for decode batches (M <= 8, K-quants) as TWO launches and two barriers:

      1. one MMVQ over the stacked shard [gate_shard; up_shard]: its epilogue writes the gate slice and the up slice of
         a symmetric [8, 2 * I] buffer on every rank (ggb_mmvq_fused with split output columns);
      2. the down projection reads g = buf[:, :I], u = buf[:, I:] and computes silu(g) * u inside its activation
         quantization, then gathers its own output the same way.

    The result equals QuantMLP's bit for bit: rows are independent of how many share a launch, and the SwiGLU is
    computed with torch eager's roundings.  Larger batches / other types fall back to QuantMLP's path."""

    def __init__(self, gate: ColumnParallelQuantLinear, up: ColumnParallelQuantLinear, down: ColumnParallelQuantLinear,
                 W_gu: Optional[torch.Tensor] = None, share_buffers_with: Optional["FusedQuantMLP"] = None):
        """`W_gu`: the stacked shard [gate.W; up.W] when the caller already holds it (gate.W / up.W are then views of it and
        nothing is copied).  `share_buffers_with`: another block of the same shapes (the previous layer of a stack) whose
        two alternating output buffer sets this block uses too -- consecutive blocks alternate between the sets, so a
        result stays valid until the second-next block runs."""
        if gate.quant_type != up.quant_type or gate.n_total != up.n_total or gate.W.shape[1] != up.W.shape[1]:
            raise ValueError("gate and up must share type and shape")
        self.plain = QuantMLP(gate, up, down)
        self.gate, self.up, self.down = gate, up, down
        self.rank, self.world, self.group = gate.rank, gate.world, gate.group
        if W_gu is not None:
            if W_gu.shape != (2 * gate.n_local, gate.W.shape[1]) or not W_gu.is_contiguous():
                raise ValueError("W_gu must be the contiguous stack [gate.W; up.W]")
            self.W_gu = W_gu
        else:
            self.W_gu = torch.cat([gate.W, up.W], dim=0).contiguous()    # [2 * I / tp, row_bytes]
        self.inter, self.hidden = gate.n_total, down.n_total
        if share_buffers_with is not None:
            o = share_buffers_with
            if (o.inter, o.hidden, o.world) != (self.inter, self.hidden, self.world):
                raise ValueError("shared buffers need blocks of the same shape")
            self._shared = o._shared
        else:
            self._shared = {"bufs": {}, "calls": 0, "ws": None, "barrier": None}
        self.fused = True

    # the buffer sets, the call counter that alternates them and the workspace live in `_shared` (one per stack of blocks)
    @property
    def _bufs(self):
        return self._shared["bufs"]

    @property
    def _calls(self):
        return self._shared["calls"]

    @_calls.setter
    def _calls(self, v):
        self._shared["calls"] = v

    @property
    def _ws(self):
        return self._shared["ws"]

    @_ws.setter
    def _ws(self, v):
        self._shared["ws"] = v

    def _barrier(self, device):
        """the stack's barrier: ggb_tp_barrier (one PDL launch) unless GGML_B200_TP_BARRIER=torch asks for the framework's"""
        b = self._shared["barrier"]
        if b is None:
            import os
            if os.environ.get("GGML_B200_TP_BARRIER", "") == "torch":
                hdl = self._buffers(torch.float16, device)[0][0][1]
                b = lambda: hdl.barrier(channel=0)
            else:
                b = TpBarrier(self.rank, self.world, self.group, device)
            self._shared["barrier"] = b
        return b

    def _buffers(self, dtype, device):
        """per dtype: two (gate|up buffer, down buffer) pairs, each with its peer pointers and barrier handle"""
        if dtype not in self._bufs:
            from . import MMVQ_MAX_BATCH
            sets = []
            for _ in range(2):
                pair = []
                for cols in (2 * self.inter, self.hidden):
                    if self.world > 1:
                        import torch.distributed._symmetric_memory as symm_mem
                        buf = symm_mem.empty((MMVQ_MAX_BATCH, cols), dtype=dtype, device=device)
                        hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
                        pair.append((buf, hdl, [int(p) for p in hdl.buffer_ptrs]))
                    else:
                        buf = torch.empty((MMVQ_MAX_BATCH, cols), dtype=dtype, device=device)
                        pair.append((buf, None, [buf.data_ptr()]))
                sets.append(pair)
            self._bufs[dtype] = sets
        return self._bufs[dtype]

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if not (self.fused and X.dim() == 2 and X.shape[0] <= 8):
            return self.plain(X)
        from . import mmvq_fused_supported, mmvq_workspace_bytes, mul_mat_vec_a8_fused
        M, K = X.shape
        # eligibility from (type, batch, K) alone: the same verdict on every rank, decided before anything is launched
        if not (mmvq_fused_supported(self.gate.quant_type, M, K) and mmvq_fused_supported(self.down.quant_type, M, self.inter)):
            return self.plain(X)
        (gu, gu_h, gu_p), (dn, dn_h, dn_p) = self._buffers(X.dtype, X.device)[self._calls & 1]
        n_gu, n_dn, I = self.gate.n_local, self.down.n_local, self.inter
        need = max(mmvq_workspace_bytes(self.gate.quant_type, M, K, 2 * n_gu), mmvq_workspace_bytes(self.down.quant_type, M, I, n_dn))
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = torch.empty(need, dtype=torch.uint8, device=X.device)
        mul_mat_vec_a8_fused(self.W_gu, X, None, self.gate.quant_type, 2 * n_gu, gu_p, 2 * I, self.rank * n_gu,
                             n_gu, I + self.rank * n_gu, self._ws)
        barrier = self._barrier(X.device) if gu_h is not None else None
        if barrier is not None:
            barrier()                            # every rank's gate | up slices have landed in this rank's copy
        mul_mat_vec_a8_fused(self.down.W, gu[:M, :I], gu[:M, I:], self.down.quant_type, n_dn, dn_p, self.hidden,
                             self.rank * n_dn, 0, 0, self._ws)
        if barrier is not None:
            barrier()
        self._calls += 1
        return dn[:M]

    __call__ = forward
