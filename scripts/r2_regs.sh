#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py -x -q -m gpu -k "shard or 70b or tp" 2>&1 | tail -3
VERBOSE=1 TD=12 python scripts/block_trace.py
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py
python scripts/mmvq_probe.py time types=12 static=1 B=1 reps=4 shapes=8192x28672
python bench.py --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
