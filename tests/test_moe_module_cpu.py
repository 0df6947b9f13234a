"""Host logic of `ggml.moe.QuantMoE` (the mixture-of-experts MLP around ggml_moe_a8_vec / ggml_moe_a8) on CPU: the ops are
replaced by oracle-backed stand-ins with the ops' semantics (pair i = token i // top_k through expert ids[i]; ids outside
[0, E) give zero rows), so what is under test is the wiring -- gate|up split, SiLU-and-mul, the second projection per pair,
routing weights, the sum over top_k, the decode / prefill switch -- and the expert-parallel exchange (gloo, world 2):
sharded == unsharded bit for bit."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

import oracle

T13, T2 = oracle.Q4_K, oracle.Q8_0
E, HIDDEN, INTER, TOP_K = 4, 256, 512, 2


def _standin_ops():
    import ggml

    def moe_vec(X, W, ids, top_k, t, N, tokens):
        w, x = W.numpy(), X.float().numpy()
        out = torch.zeros(tokens * top_k, N, dtype=X.dtype)
        for i, e in enumerate(ids.reshape(-1).tolist()):
            if 0 <= e < w.shape[0]:
                out[i] = torch.from_numpy(oracle.mmvq(w[e], t, x[i // top_k:i // top_k + 1], N)[0]).to(X.dtype)
        return out

    # the real host scheduling of ggml_moe_a8, its tensor-core matmul replaced by the oracle's MMQ
    ggml._check_x = lambda *a, **k: None
    ggml._check_w = lambda *a, **k: None
    ggml.ggml_mul_mat_a8 = lambda We, X, t, N: torch.from_numpy(oracle.mmq(We.numpy(), t, X.float().numpy(), N)).to(X.dtype)
    return types.SimpleNamespace(ggml_moe_a8_vec=moe_vec, ggml_moe_a8=ggml.ggml_moe_a8, MMQ_TYPES=ggml.MMQ_TYPES,
                                 moe_align_block_size=ggml.moe_align_block_size,
                                 ggml_moe_get_block_size=ggml.ggml_moe_get_block_size)


def _weights():
    w13 = np.stack([oracle.random_weights(T13, 2 * INTER, HIDDEN, seed=10 + e) for e in range(E)])
    w2 = np.stack([oracle.random_weights(T2, HIDDEN, INTER, seed=20 + e) for e in range(E)])
    return torch.from_numpy(w13), torch.from_numpy(w2)


def _inputs(tokens):
    from ggml.moe import route_topk
    g = torch.Generator().manual_seed(tokens)
    x = torch.randn(tokens, HIDDEN, generator=g)
    w, ids = route_topk(torch.randn(tokens, E, generator=g), TOP_K)
    return x, w, ids


def _expected(x, w, ids, w13, w2, matmul):
    y = torch.zeros(x.shape[0], HIDDEN)
    for t in range(x.shape[0]):
        terms = []
        for j in range(TOP_K):
            e = int(ids[t, j])
            gu = torch.from_numpy(matmul(w13[e].numpy(), T13, x[t:t + 1].numpy(), 2 * INTER))
            h = F.silu(gu[:, :INTER]) * gu[:, INTER:]
            terms.append(torch.from_numpy(matmul(w2[e].numpy(), T2, h.numpy(), HIDDEN))[0] * w[t, j])
        y[t] = torch.stack(terms).sum(dim=0)
    return y


def test_route_topk():
    from ggml.moe import route_topk
    logits = torch.tensor([[0.0, 2.0, 1.0, -1.0], [3.0, 3.0, 0.0, 0.0]])
    w, ids = route_topk(logits, 2)
    assert ids.dtype == torch.int32 and ids[0].tolist() == [1, 2] and sorted(ids[1].tolist()) == [0, 1]
    assert torch.allclose(w.sum(-1), torch.ones(2)) and w[0, 0] > w[0, 1] and torch.allclose(w[1], torch.tensor([0.5, 0.5]))
    w2, _ = route_topk(logits, 2, renormalize=False)
    assert (w2.sum(-1) < 1).all()


@pytest.mark.parametrize("tokens", [1, 5, 70])          # 70 > 64: the per-expert tile path
def test_quant_moe_wiring(tokens, monkeypatch):
    import ggml
    for name in ("_check_x", "_check_w", "ggml_mul_mat_a8"):
        monkeypatch.setattr(ggml, name, getattr(ggml, name))          # restored after the stand-ins patch the module
    from ggml.moe import QuantMoE
    w13, w2 = _weights()
    x, w, ids = _inputs(tokens)
    moe = QuantMoE(w13, w2, T13, T2, TOP_K, ops=_standin_ops())
    y = moe(x, w, ids)
    want = _expected(x, w, ids, w13, w2, oracle.mmq if tokens > 64 else oracle.mmvq)
    assert y.shape == (tokens, HIDDEN)
    assert torch.allclose(y, want, rtol=1e-6, atol=1e-6 * float(want.abs().max())), float((y - want).abs().max())
    with pytest.raises(ValueError):
        moe(x, w[:, :1], ids)
    with pytest.raises(ValueError):
        QuantMoE(w13[:, :-1], w2, T13, T2, TOP_K, ops=moe.ops)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ggml.moe import QuantMoE
        w13, w2 = _weights()
        per = E // world
        lo = rank * per
        moe = QuantMoE(w13[lo:lo + per].contiguous(), w2[lo:lo + per].contiguous(), T13, T2, TOP_K, expert_offset=lo,
                       world=world, ops=_standin_ops())
        res = {}
        for tokens in (3, 70):
            x, w, ids = _inputs(tokens)
            res[f"y{tokens}"] = moe(x, w, ids).numpy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_expert_parallel_world2_gloo(tmp_path, monkeypatch):
    import ggml
    for name in ("_check_x", "_check_w", "ggml_mul_mat_a8"):
        monkeypatch.setattr(ggml, name, getattr(ggml, name))
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    from ggml.moe import QuantMoE
    w13, w2 = _weights()
    single = QuantMoE(w13, w2, T13, T2, TOP_K, ops=_standin_ops())
    for tokens in (3, 70):
        x, w, ids = _inputs(tokens)
        y = single(x, w, ids).numpy()
        assert np.array_equal(r0[f"y{tokens}"].view(np.uint32), r1[f"y{tokens}"].view(np.uint32)), "ranks disagree"
        assert np.array_equal(r0[f"y{tokens}"].view(np.uint32), y.view(np.uint32)), f"tokens={tokens}: sharded != single"
