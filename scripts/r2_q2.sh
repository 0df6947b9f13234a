#!/bin/bash
python scripts/mmvq_probe.py check types=10 2>&1 | grep -v " OK " | head
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -4
python scripts/mmvq_probe.py time types=10 static=1 B=1 reps=4 shapes=14336x4096,4096x14336,28672x8192
python scripts/mmvq_probe.py time types=10 static=1 B=1 reps=4 rg=-2147483647 shapes=14336x4096,4096x14336,28672x8192
