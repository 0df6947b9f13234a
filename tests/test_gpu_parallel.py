"""GPU parity of the column-parallel layers: world_size 2, 4 and 8 over NCCL, the product's CUDA
kernels on every rank.  The sharded MLP output must equal the single-GPU output bit for bit for decode
(M = 1, 8: MMVQ) and prefill (M = 160: MMQ), because every output element is produced by exactly one rank
with the same K order as on one GPU.  Each world size is skipped on boxes with fewer GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _weights(hidden, inter):
    wg = torch.from_numpy(oracle.random_weights(oracle.Q4_K, inter, hidden, seed=1))
    wu = torch.from_numpy(oracle.random_weights(oracle.Q4_K, inter, hidden, seed=2))
    wd = torch.from_numpy(oracle.random_weights(oracle.Q6_K, hidden, inter, seed=3))
    return wg, wu, wd


HIDDEN, INTER, BATCHES = 2048, 4096, (1, 8, 160)


def _inputs(M):
    return (torch.randn(M, HIDDEN, generator=torch.Generator().manual_seed(M)) * 0.25).to(torch.float16)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from ggml.parallel import ColumnParallelQuantLinear, QuantMLP, shard_rows
        wg, wu, wd = _weights(HIDDEN, INTER)
        mk = lambda w, tt, n: ColumnParallelQuantLinear(shard_rows(w, rank, world).to(dev), tt, n, rank, world)
        mlp = QuantMLP(mk(wg, oracle.Q4_K, INTER), mk(wu, oracle.Q4_K, INTER), mk(wd, oracle.Q6_K, HIDDEN))
        res = {}
        for M in BATCHES:
            x = _inputs(M).to(dev)
            res[f"gate{M}"] = mlp.gate(x).float().cpu().numpy()
            res[f"mlp{M}"] = mlp(x).float().cpu().numpy()
        # the same layers with the all-gather fused into the MMVQ epilogue (peer stores into a symmetric buffer)
        mkf = lambda w, tt, n: ColumnParallelQuantLinear(shard_rows(w, rank, world).to(dev), tt, n, rank, world, fused_gather=True)
        fmlp = QuantMLP(mkf(wg, oracle.Q4_K, INTER), mkf(wu, oracle.Q4_K, INTER), mkf(wd, oracle.Q6_K, HIDDEN))
        for M in BATCHES:
            x = _inputs(M).to(dev)
            for rep in range(3):                     # alternating buffers, back-to-back calls
                y = fmlp(x)
            res[f"fused{M}"] = y.float().cpu().numpy()
        res["fused_active"] = np.array([int(fmlp.gate.fused_gather), int(fmlp.down.fused_gather)])
        # ... and the whole block as two launches (stacked gate|up, SwiGLU inside the down projection)
        from ggml.parallel import FusedQuantMLP
        mkp = lambda w, tt, n: ColumnParallelQuantLinear(shard_rows(w, rank, world).to(dev), tt, n, rank, world)
        bmlp = FusedQuantMLP(mkp(wg, oracle.Q4_K, INTER), mkp(wu, oracle.Q4_K, INTER), mkp(wd, oracle.Q6_K, HIDDEN))
        for M in BATCHES:
            x = _inputs(M).to(dev)
            for rep in range(3):
                y = bmlp(x)
            res[f"block{M}"] = y.float().cpu().numpy()
        res["block_active"] = np.array([int(bmlp.fused)])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_column_parallel_nccl(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    # single-GPU result of the same layer stack
    import ggml
    from ggml.parallel import ColumnParallelQuantLinear, QuantMLP
    dev = torch.device("cuda", 0)
    wg, wu, wd = _weights(HIDDEN, INTER)
    mk = lambda w, tt, n: ColumnParallelQuantLinear(w.to(dev), tt, n, 0, 1)
    mlp = QuantMLP(mk(wg, oracle.Q4_K, INTER), mk(wu, oracle.Q4_K, INTER), mk(wd, oracle.Q6_K, HIDDEN))
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    r0 = ranks[0]
    for M in BATCHES:
        x = _inputs(M).to(dev)
        want_gate = mlp.gate(x).float().cpu().numpy()
        want_mlp = mlp(x).float().cpu().numpy()
        for r in ranks:
            assert np.array_equal(r[f"gate{M}"], want_gate), f"gate M={M}: sharded != single-GPU"
            assert np.array_equal(r[f"mlp{M}"], want_mlp), f"mlp M={M}: sharded != single-GPU"
            assert np.array_equal(r[f"fused{M}"], want_mlp), f"mlp M={M}: fused gather != single-GPU"
            assert np.array_equal(r[f"block{M}"], want_mlp), f"mlp M={M}: fused block != single-GPU"
    assert r0["fused_active"].tolist() == [1, 1], "the fused path fell back to NCCL"
    assert r0["block_active"].tolist() == [1], "the fused block fell back"
