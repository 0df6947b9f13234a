#!/bin/bash
# round-2 probe of the batch-1 register-resident MMVQ kernel (mmvq_rg.cu): parity vs the oracle, timings, per-CTA phases
python scripts/mmvq_probe.py check types=12 rg=1 B=1 > gpurun_out/r2_rg3_check.log 2>&1
SH=14336x4096,4096x14336,4096x4096,28672x8192,8192x8192
for m in 0 1 $((1+(3<<12))) $((1+(2<<16))) $((1+(4<<16))) 257 ; do
  python scripts/mmvq_probe.py time types=12 static=1 B=1 rg=$m reps=4 shapes=$SH 2>&1 | sed "s/^/rg=$m /" >> gpurun_out/r2_rg3_time.log
done
python scripts/mmvq_probe.py time types=12 static=0 B=1 rg=1 reps=4 shapes=$SH 2>&1 | sed "s/^/rg=1 /" >> gpurun_out/r2_rg3_time.log
python scripts/mmvq_trace.py rg=1 > gpurun_out/r2_rg3_trace.log 2>&1
python scripts/mmvq_trace.py rg=1 N=4096 K=14336 >> gpurun_out/r2_rg3_trace.log 2>&1
