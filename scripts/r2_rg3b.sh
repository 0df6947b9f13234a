#!/bin/bash
SH=14336x4096,4096x14336
for e in 1 3; do for l in 0 2 4 8; do
  m=$((1 + (e<<16) + (l<<20)))
  python scripts/mmvq_probe.py time types=12 static=1 B=1 rg=$m reps=4 shapes=$SH 2>&1 | sed "s/^/early=$e l2rows=$l /" >> gpurun_out/r2_rg3b_time.log
done; done
python scripts/mmvq_trace.py rg=$((1 + (3<<16) + (8<<20))) >> gpurun_out/r2_rg3b_trace.log 2>&1
python scripts/mmvq_trace.py rg=$((1 + (1<<16) + (8<<20))) >> gpurun_out/r2_rg3b_trace.log 2>&1
