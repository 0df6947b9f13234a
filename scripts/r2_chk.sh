#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_parallel.py tests/test_gpu_canaries.py tests/test_gpu_vs_ref_cuda.py -x -q -m gpu 2>&1 | tail -3
LAYERS=40 MIX=1 python scripts/block_breakdown.py
python scripts/mmvq_probe.py time types=12,14 static=1 B=1 reps=4 shapes=8192x28672,1024x28672
python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['eager_value'])"
