#!/bin/bash
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_parallel.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -4
LAYERS=24 MIX=14 python scripts/block_breakdown.py
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=40 MIX=1 python scripts/block_breakdown.py
VERBOSE=0 TD=14 python scripts/block_trace.py
python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'])"
VERBOSE=0 TD=12,12,14,12,12,14 python scripts/block_trace.py
echo "== Q4_K down on 14 warps, SwiGLU in-kernel (bit 28)"
LAYERS=40 MIX=1 python scripts/block_breakdown.py opt:7=268435457
