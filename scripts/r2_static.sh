#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_dequant.py tests/test_gpu_mmq.py -x -q -m gpu -k "static_weights or phased_walk or row_lengths or swiglu_70b" 2>&1 | tail -8
