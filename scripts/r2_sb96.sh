#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_parallel.py -x -q -m gpu 2>&1 | tail -3
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py opt:8=0
VERBOSE=0 TD=14 python scripts/block_trace.py
python scripts/mmvq_probe.py time types=13,14,23 static=1 B=1 reps=4 shapes=14336x4096,4096x14336,28672x8192,8192x28672
python bench.py --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
