"""CPU tests of the host side of the mixture-of-experts variants: the token-list layout of `moe_align_block_size`, the
per-expert scheduling of `ggml_moe_a8` (the tensor-core matmul replaced by the oracle's MMQ, so only the gather / scatter
logic is under test here -- the CUDA path is tests/test_gpu_moe.py), and `ggb_moe_mmvq`'s argument validation."""
import ctypes

import numpy as np
import pytest
import torch

import oracle


@pytest.fixture()
def ops():
    import ggml
    return ggml


@pytest.mark.parametrize("tokens,top_k,E,block", [(1, 2, 8, 32), (70, 2, 5, 32), (33, 4, 3, 8), (5, 1, 4, 1)])
def test_moe_align_block_size_layout(ops, tokens, top_k, E, block):
    g = torch.Generator().manual_seed(tokens)
    ids = torch.randint(0, E, (tokens, top_k), generator=g)
    sorted_ids, expert_ids, n_pad = ops.moe_align_block_size(ids, block, E)
    numel = tokens * top_k
    n = int(n_pad.item())
    flat = ids.reshape(-1).tolist()
    counts = [flat.count(e) for e in range(E)]
    assert n == sum((c + block - 1) // block * block for c in counts)
    assert n % block == 0 and n <= sorted_ids.numel() and sorted_ids.dtype == torch.int32 and expert_ids.dtype == torch.int32
    s = sorted_ids.tolist()
    seen = []
    for b in range(n // block):
        e = int(expert_ids[b])
        for slot in s[b * block:(b + 1) * block]:
            if slot == numel:
                continue                       # padding
            assert 0 <= slot < numel and flat[slot] == e, "a pair sits in a block of another expert"
            seen.append(slot)
    assert sorted(seen) == list(range(numel)), "every pair exactly once"
    assert all(v == numel for v in s[n:]), "the tail is padding"
    # blocks of one expert are contiguous and experts ascend
    eb = expert_ids[: n // block].tolist()
    assert eb == sorted(eb)


def test_moe_a8_host_scheduling(ops, monkeypatch):
    t, E, N, K, tokens, top_k = oracle.Q4_K, 5, 24, 256, 37, 2
    w = np.stack([oracle.random_weights(t, N, K, seed=e) for e in range(E)])
    dense = [torch.from_numpy(oracle.dequantize(w[e], t, N, K)) for e in range(E)]
    W = torch.from_numpy(w)
    calls = []

    def fake_mmq(We, X, qt, row):                # dense stand-in: identifies the expert by its bytes
        e = next(i for i in range(E) if We.data_ptr() == W[i].data_ptr())
        calls.append((e, X.shape[0]))
        return X @ dense[e].T

    monkeypatch.setattr(ops, "ggml_mul_mat_a8", fake_mmq)
    monkeypatch.setattr(ops, "_check_x", lambda *a, **k: None)
    monkeypatch.setattr(ops, "_check_w", lambda *a, **k: None)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(tokens, K, generator=g)
    ids = torch.randint(0, E - 1, (tokens, top_k), generator=g)         # the last expert stays empty
    block = ops.ggml_moe_get_block_size(t)
    out = ops.ggml_moe_a8(x, W, *ops.moe_align_block_size(ids, block, E), t, N, top_k, tokens)
    assert out.shape == (tokens * top_k, N)
    for i, e in enumerate(ids.reshape(-1).tolist()):
        want = x[i // top_k] @ dense[e].T
        assert torch.allclose(out[i], want, rtol=1e-5, atol=1e-5), f"pair {i}"
    assert sorted(e for e, _ in calls) == sorted(set(ids.reshape(-1).tolist())), "one matmul per non-empty expert"
    assert all(m % block == 0 for _, m in calls)
    with pytest.raises(ValueError):
        ops.ggml_moe_a8(x, W, *ops.moe_align_block_size(ids, block, E), 23, N, top_k, tokens)     # IQ4_XS: no MMQ


def test_moe_block_size_query(ops):
    for t in ops.MMQ_TYPES:
        assert ops.ggml_moe_get_block_size(t) == 32
    for t in (20, 23, 16, 29, 0):
        with pytest.raises(ValueError):
            ops.ggml_moe_get_block_size(t)


def test_moe_mmvq_argument_errors_without_gpu(ops):
    L = ops._capi.lib
    buf = ctypes.create_string_buffer(4096)
    p16 = (ctypes.addressof(buf) + 15) & ~15
    f16 = ops.GGB_F16
    #                 W    X    ids  Y    type dt  tokens top_k E  K    N
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 99, f16, 1, 2, 4, 256, 8, None) == -1        # type
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 12, 7, 1, 2, 4, 256, 8, None) == -3          # dtype
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 12, f16, 1, 2, 4, 100, 8, None) == -2        # K not whole blocks
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 12, f16, 1, 0, 4, 256, 8, None) == -2        # top_k < 1
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 12, f16, 1, 2, 0, 256, 8, None) == -2        # no experts
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 12, f16, -1, 2, 4, 256, 8, None) == -2       # negative tokens
    assert L.ggb_moe_mmvq(p16, p16, None, p16, 12, f16, 1, 2, 4, 256, 8, None) == -4       # null ids
    assert L.ggb_moe_mmvq(p16, p16, p16 + 2, p16, 12, f16, 1, 2, 4, 256, 8, None) == -4    # misaligned ids
    assert L.ggb_moe_mmvq(None, p16, p16, p16, 12, f16, 1, 2, 4, 256, 8, None) == -4       # null W
    assert b"ggb_moe_mmvq" in L.ggb_last_error()
    assert L.ggb_moe_mmvq(p16, p16, p16, p16, 12, f16, 0, 2, 4, 256, 8, None) == 0         # no tokens: a no-op
    assert L.ggb_last_error() == b""


def test_moe_wrappers_reject_cpu_tensors(ops):
    w = torch.zeros(2, 4, 144, dtype=torch.uint8)
    x = torch.zeros(1, 256, dtype=torch.float16)
    ids = torch.zeros(1, 2, dtype=torch.int32)
    with pytest.raises(ValueError):
        ops.ggml_moe_a8_vec(x, w, ids, 2, 12, 4, 1)
