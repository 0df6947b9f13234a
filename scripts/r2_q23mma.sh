#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -6
python scripts/mmvq_probe.py time types=10,11 static=1 B=3,4,8 reps=2 shapes=14336x4096,4096x14336,28672x8192
