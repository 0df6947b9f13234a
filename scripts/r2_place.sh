#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_canaries.py tests/test_gpu_parallel.py -x -q -m gpu 2>&1 | tail -12
echo "== placement: cluster launches (default)"
VERBOSE=0 TD=12 python scripts/block_trace.py
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py
python scripts/mmvq_trace.py N=14336 K=4096 B=1 static=1 | head -8
python scripts/mmvq_probe.py time types=12 static=1 B=1 reps=4 shapes=14336x4096,4096x14336,28672x8192,8192x28672
echo "== placement: plain launches (bit 26)"
VERBOSE=0 TD=12 python scripts/block_trace.py opt:7=67108865 | head -8
python scripts/mmvq_trace.py N=14336 K=4096 B=1 static=1 rg=67108865 | head -8
python scripts/mmvq_probe.py time types=12 static=1 B=1 rg=67108865 reps=4 shapes=14336x4096,4096x14336
echo "== tensor-core path: in-kernel pair quantization (default) vs image launch (mma=0,0,128)"
python scripts/mmvq_probe.py time types=12,14 static=1 B=3,4,8 reps=2 shapes=14336x4096,4096x14336,28672x8192
python scripts/mmvq_probe.py time types=12,14 static=1 B=3,4,8 reps=2 mma=0,0,128 shapes=14336x4096,4096x14336,28672x8192
