#!/bin/bash
python -m pytest tests/test_gpu_dequant.py tests/test_gpu_mmvq.py tests/test_gpu_canaries.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_mmq.py -x -q -m gpu 2>&1 | tail -15
VERBOSE=0 TD=12 python scripts/block_trace.py
