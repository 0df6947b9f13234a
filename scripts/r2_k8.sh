#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py -x -q -m gpu 2>&1 | tail -3
python scripts/mmvq_probe.py time types=12,14 static=1 B=1 reps=4 shapes=28672x8192,8192x8192,4096x16384,8192x24576
python scripts/mmvq_probe.py time types=12,14 static=1 B=1 reps=4 rg=513 shapes=28672x8192,8192x8192
LAYERS=40 MIX=1 python scripts/block_breakdown.py
LAYERS=40 MIX=1 python scripts/block_breakdown.py opt:7=513
