#!/bin/bash
SH=14336x4096,4096x14336,28672x8192,8192x8192
for cl in 1 2 4; do
  m=$((1+(cl<<20)))
  python scripts/mmvq_probe.py check types=12 rg=$m B=1 2>&1 | grep -c OK | sed "s/^/cluster=$cl check OK count: /" >> gpurun_out/r2_rg3_time.log
  python scripts/mmvq_probe.py time types=12 static=1 B=1 rg=$m reps=4 shapes=$SH 2>&1 | sed "s/^/cluster=$cl /" >> gpurun_out/r2_rg3_time.log
done
python scripts/mmvq_probe.py time types=12 static=1 B=1 reps=4 shapes=$SH 2>&1 | sed "s/^/default /" >> gpurun_out/r2_rg3_time.log
python scripts/mmvq_trace.py rg=$((1+(2<<20))) N=4096 K=14336 > gpurun_out/r2_rg3_trace.log 2>&1
