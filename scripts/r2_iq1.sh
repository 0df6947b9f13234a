#!/bin/bash
python -m pytest tests/test_gpu_dequant.py tests/test_gpu_mmvq.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -15
for v in "" "opt:1=0" "opt:7=262145" "opt:7=1048577" "opt:7=1310721"; do
  echo "== block trace $v"; VERBOSE=0 TD=12 python scripts/block_trace.py $v
done
