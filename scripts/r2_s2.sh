#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_parallel.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -4
python scripts/mmvq_probe.py time types=12 static=1 B=1 reps=2 shapes=8192x28672,28672x8192,14336x4096,4096x14336,8192x24576 2>&1
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py
for t in 13 14 23; do
  for rg in 3 1; do
    python scripts/mmvq_probe.py time types=$t static=1 B=1 rg=$rg reps=4 shapes=14336x4096,4096x14336,28672x8192 2>&1 | sed "s/^/rg=$rg /"
  done
done
for mma in 0,0,0 2,1,0 2,2,0 2,3,0 4,2,0 4,3,0 4,4,0; do
  python scripts/mmvq_probe.py time types=12 static=1 B=4,8 mma=$mma reps=2 shapes=14336x4096,4096x14336,28672x8192,8192x28672 2>&1 | sed "s/^/mma=$mma /"
done
