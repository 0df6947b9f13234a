"""Host logic of the column-parallel layers on CPU: world_size 2, gloo backend, the CPU oracle injected
as the matmul (the product's CUDA kernels are not involved).  Checks the byte-sharding of GGUF rows, the
all-gather reassembly for decode (M = 1) and batched (M > 1) outputs, and the sharded-equals-unsharded
parity property (bit for bit)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cpu_matmul(W, X, quant_type, rows):
    y = oracle.mmvq(W.numpy(), quant_type, X.float().numpy(), rows)
    return torch.from_numpy(y).to(X.dtype)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ggml.parallel import ColumnParallelQuantLinear, QuantMLP, shard_rows
        t, hidden, inter = oracle.Q4_K, 256, 512
        wg = torch.from_numpy(oracle.random_weights(t, inter, hidden, seed=1))
        wu = torch.from_numpy(oracle.random_weights(t, inter, hidden, seed=2))
        wd = torch.from_numpy(oracle.random_weights(oracle.Q6_K, hidden, inter, seed=3))
        mk = lambda w, tt, n: ColumnParallelQuantLinear(shard_rows(w, rank, world), tt, n, rank, world, matmul_fn=_cpu_matmul)
        gate, up, down = mk(wg, t, inter), mk(wu, t, inter), mk(wd, oracle.Q6_K, hidden)
        res = {}
        for M in (1, 3):
            x = torch.randn(M, hidden, generator=torch.Generator().manual_seed(M))
            res[f"gate{M}"] = gate(x).numpy()
            res[f"mlp{M}"] = QuantMLP(gate, up, down)(x).numpy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


def test_shard_rows_are_byte_slices():
    from ggml.parallel import shard_bounds, shard_rows
    w = torch.from_numpy(oracle.random_weights(oracle.Q4_K, 8, 512, seed=0))
    parts = [shard_rows(w, r, 4) for r in range(4)]
    assert torch.equal(torch.cat(parts), w)
    assert shard_bounds(8, 3, 4) == (6, 8)
    with pytest.raises(ValueError):
        shard_rows(w, 0, 3)
    # a row slice is a valid GGUF tensor: dequantizing the slice == slicing the dequantized matrix
    full = oracle.dequantize(w.numpy(), oracle.Q4_K, 8, 512)
    part = oracle.dequantize(parts[2].numpy(), oracle.Q4_K, 2, 512)
    assert np.array_equal(full[4:6], part)


@pytest.mark.timeout(300)
def test_column_parallel_world2_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # unsharded result computed the same way on one "device"
    t, hidden, inter = oracle.Q4_K, 256, 512
    wg = oracle.random_weights(t, inter, hidden, seed=1)
    wu = oracle.random_weights(t, inter, hidden, seed=2)
    wd = oracle.random_weights(oracle.Q6_K, hidden, inter, seed=3)
    for M in (1, 3):
        x = torch.randn(M, hidden, generator=torch.Generator().manual_seed(M))
        g = oracle.mmvq(wg, t, x.numpy(), inter)
        u = oracle.mmvq(wu, t, x.numpy(), inter)
        h = torch.nn.functional.silu(torch.from_numpy(g)) * torch.from_numpy(u)
        y = oracle.mmvq(wd, oracle.Q6_K, h.numpy(), hidden)
        for r in (r0, r1):
            assert np.array_equal(r[f"gate{M}"], g), "sharded gate != unsharded (must be bit-identical)"
            assert np.array_equal(r[f"mlp{M}"], y), "sharded MLP != unsharded"
