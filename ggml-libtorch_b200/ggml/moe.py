"""Mixture-of-experts MLP on GGUF expert stacks: the caller of `ggml_moe_a8_vec` / `ggml_moe_a8` (SURVEY.md section 8f
rank 4).  Not in this reference; the block is the one its upstream builds around these ops (vLLM
`model_executor/layers/quantization/gguf.py::_fused_moe_gguf`): gate|up of every (token, expert) pair, SiLU-and-mul, the
down projection of every pair, routing weights, sum over the top_k rows of a token.

Expert parallelism over one box (one process per GPU): rank r keeps a contiguous range of experts (`expert_offset`); ids of
experts held elsewhere fall outside `[0, E_local)` after the offset and yield ZERO rows in both ops, so every rank runs the
same two launches on the full routing table and ONE all-reduce (sum) of the `[tokens, hidden]` result is the whole exchange.
A token's row is the sum of at most top_k non-zero terms in rank order, so the result equals the single-GPU one up to that
summation order (bit for bit when top_k <= 2: fp addition is commutative).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

PREFILL_MIN_TOKENS = 64          # upstream's switch: below it expert reuse is too low for the tile kernel to pay


def route_topk(router_logits: torch.Tensor, top_k: int, renormalize: bool = True):
    """softmax over experts -> top_k (weights fp32, ids int32), renormalised like upstream's `fused_topk`."""
    probs = torch.softmax(router_logits.float(), dim=-1)
    w, ids = torch.topk(probs, top_k, dim=-1)
    if renormalize:
        w = w / w.sum(dim=-1, keepdim=True)
    return w, ids.to(torch.int32)


class QuantMoE:
    """`y = sum_j w[t, j] * down_e(silu(gate_e(x_t)) * up_e(x_t))`, e = topk_ids[t, j].

    w13: uint8 [E_local, 2 * inter, bytes(K = hidden)] -- per expert the gate rows stacked on the up rows (upstream's w1 layout);
    w2:  uint8 [E_local, hidden, bytes(K = inter)].
    `ops` is the module providing the two ops (default: the `ggml` package = the CUDA library; CPU tests inject stand-ins)."""

    def __init__(self, w13: torch.Tensor, w2: torch.Tensor, qtype13: int, qtype2: int, top_k: int, *, expert_offset: int = 0,
                 group: Optional[dist.ProcessGroup] = None, world: int = 1, ops=None):
        if w13.dim() != 3 or w2.dim() != 3 or w13.shape[0] != w2.shape[0] or w13.shape[1] % 2:
            raise ValueError("w13 must be [E, 2*inter, bytes] and w2 [E, hidden, bytes] over the same experts")
        if ops is None:
            import ggml as ops
        self.ops = ops
        self.w13, self.w2, self.t13, self.t2 = w13, w2, int(qtype13), int(qtype2)
        self.n_local, self.inter, self.hidden = w13.shape[0], w13.shape[1] // 2, w2.shape[1]
        self.top_k, self.offset, self.group, self.world = int(top_k), int(expert_offset), group, int(world)

    def _pairs_vec(self, x, ids):
        tokens = x.shape[0]
        gu = self.ops.ggml_moe_a8_vec(x, self.w13, ids, self.top_k, self.t13, 2 * self.inter, tokens)
        h = F.silu(gu[:, :self.inter]) * gu[:, self.inter:]
        return self.ops.ggml_moe_a8_vec(h.contiguous(), self.w2, ids, 1, self.t2, self.hidden, tokens * self.top_k)

    def _pairs_tiles(self, x, ids):
        tokens = x.shape[0]
        block = self.ops.ggml_moe_get_block_size(self.t13)
        # experts held by other ranks are collected under one extra id (n_local): both ops leave those rows zero
        local = torch.where((ids >= 0) & (ids < self.n_local), ids, torch.full_like(ids, self.n_local))
        lists = self.ops.moe_align_block_size(local, block, self.n_local + 1)
        gu = self.ops.ggml_moe_a8(x, self.w13, *lists, self.t13, 2 * self.inter, self.top_k, tokens)
        h = F.silu(gu[:, :self.inter]) * gu[:, self.inter:]
        return self.ops.ggml_moe_a8(h.contiguous(), self.w2, *lists, self.t2, self.hidden, 1, tokens * self.top_k)

    def __call__(self, x: torch.Tensor, topk_weights: torch.Tensor, topk_ids: torch.Tensor) -> torch.Tensor:
        tokens = x.shape[0]
        if topk_ids.shape != (tokens, self.top_k) or topk_weights.shape != (tokens, self.top_k):
            raise ValueError("topk_weights / topk_ids must be [tokens, top_k]")
        ids = (topk_ids.to(torch.int32) - self.offset).contiguous()
        mmq = self.ops.MMQ_TYPES
        if tokens > PREFILL_MIN_TOKENS and self.t13 in mmq and self.t2 in mmq:
            out = self._pairs_tiles(x, ids)
        else:
            out = self._pairs_vec(x, ids)
        y = (out.view(tokens, self.top_k, self.hidden) * topk_weights.view(tokens, self.top_k, 1).to(out.dtype)).sum(dim=1)
        if self.world > 1:
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y
