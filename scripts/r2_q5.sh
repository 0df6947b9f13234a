#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_parallel.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -4
python scripts/mmvq_probe.py time types=13 static=1 B=1 reps=4 shapes=14336x4096,4096x14336,28672x8192,8192x28672
python scripts/mmvq_probe.py time types=13 static=1 B=1 reps=4 rg=536870913 shapes=14336x4096,4096x14336,28672x8192,8192x28672
