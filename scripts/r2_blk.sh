#!/bin/bash
LAYERS=4 MIX=12 python scripts/block_breakdown.py
LAYERS=4 MIX=14 python scripts/block_breakdown.py
LAYERS=24 MIX=12 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py
LAYERS=40 MIX=1 python scripts/block_breakdown.py
LAYERS=24 MIX=14 python scripts/block_breakdown.py opt:8=0
python bench.py --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline_step_kernel']['avg_launch_ms'])"
