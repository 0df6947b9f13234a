#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_canaries.py tests/test_gpu_parallel.py -x -q -m gpu 2>&1 | tail -5
echo "== block with the max-shared carve-out on every kernel"
VERBOSE=0 TD=12 python scripts/block_trace.py
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py
python scripts/mmvq_probe.py time types=12 static=1 B=1 reps=4 shapes=14336x4096,4096x14336,28672x8192,8192x28672
python scripts/mmvq_probe.py time types=12,14 static=1 B=2,4,8 reps=2 shapes=14336x4096,4096x14336
python scripts/mmvq_probe.py time types=13,14 static=1 B=1 reps=4 shapes=14336x4096,4096x14336
python bench.py --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
