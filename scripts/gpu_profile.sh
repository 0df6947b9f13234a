#!/bin/bash
# ncu evidence for profiles/: the launch list of bench.py and one full capture of each hot kernel.
# The .ncu-rep files stay on the box (tens of MB each); what comes back in gpurun_out/prof/ is the raw-metrics page of
# every capture as CSV plus the digest scripts/ncu_summary.py prints (key metrics, stall totals, hottest SASS lines).
# Usage (on the GPU box, after the plain commands have exited 0):  bash scripts/gpu_profile.sh
set -u
OUT=gpurun_out/prof; TMP=/tmp/ggb_prof
mkdir -p $OUT $TMP
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > $TMP/ncu_bench.log 2>&1
cap() {   # cap <name> <kernel regex> <launches to skip> <op_run args...>
  local name=$1 rx=$2 skip=$3; shift 3
  python scripts/op_run.py "$@" > $OUT/plain_$name.log 2>&1 && \
  $NCU -k regex:$rx -s $skip -c 1 -f -o $TMP/$name python scripts/op_run.py "$@" > $TMP/ncu_$name.log 2>&1
  ncu -i $TMP/$name.ncu-rep --page raw --csv > $OUT/$name.raw.csv 2>/dev/null
  python scripts/ncu_summary.py $TMP/$name.ncu-rep 30 > $OUT/$name.summary.txt 2>&1
  head -22 $OUT/$name.summary.txt
}
cap dequant_q4k_f16 dequant_tile 3 dequant 12 4096 14336 1 3
cap mmvq_sb_q4k_b1 mmvq_sb_kernel 3 mmvq 12 14336 4096 1 3
cap mmvq_mma_q4k_b8 mmvq_mma_kernel 3 mmvq 12 14336 4096 8 3
cap mmq_q4k_m4096 mmq_kernel 4 mmq 12 14336 4096 4096 3
cat $OUT/plain_*.log
du -sh gpurun_out
