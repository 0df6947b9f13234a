#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -6
VERBOSE=0 TD=12 python scripts/block_trace.py
python scripts/mmvq_probe.py time types=2,3,6,7,8 static=1 B=1,4,8 reps=2 shapes=14336x4096,28672x8192
python scripts/mmvq_probe.py time types=16,17,22,18,21,19,29 static=1 B=1,8 reps=2 shapes=14336x4096
