#!/bin/bash
# cluster placement policy of the rg kernel in the 70B block: spread (new default) / driver default / load balancing / no cluster
for m in 1 16777217 33554433 1048577; do
  echo "== rg mode $m"
  LAYERS=24 MIX=12 python scripts/block_breakdown.py opt:7=$m
  LAYERS=24 MIX=14 python scripts/block_breakdown.py opt:7=$m
  python scripts/mmvq_probe.py time types=12 static=1 B=1 rg=$m reps=4 shapes=14336x4096,4096x14336,28672x8192,8192x28672 2>&1
done
TD=12 python scripts/block_trace.py
python scripts/mmvq_probe.py time types=12 static=1 B=4,8 reps=2 shapes=14336x4096,4096x14336,28672x8192,8192x28672
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_parallel.py -x -q -m gpu 2>&1 | tail -3
