#!/bin/bash
# round-2 final single-GPU evidence, part B: one ncu --set full capture per hot kernel (each only after the same command ran without
# ncu); the summary and the raw metric page are written on the box, the .ncu-rep files are too large to travel
cap() {   # cap <name> <kernel regex> <op_run args...>
  name=$1; rx=$2; shift 2
  python scripts/op_run.py "$@" > gpurun_out/${name}_plain.log 2>&1 || { echo "$name: plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o /tmp/$name -f python scripts/op_run.py "$@" > gpurun_out/${name}_ncu.log 2>&1
  python scripts/ncu_summary.py /tmp/$name.ncu-rep 14 > gpurun_out/$name.summary.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep; rm -f /tmp/$name.ncu-rep
}
cap r02_ncu_mmvq_rg_q4k_b1 mmvq_rg_kernel mmvq 12 14336 4096 1 5 opt:1=1
cap r02_ncu_mmvq_rg_q6k_b1 mmvq_rg_kernel mmvq 14 14336 4096 1 5 opt:1=1
cap r02_ncu_mmvq_rgx_q4_0_b1 mmvq_rgx_kernel mmvq 2 14336 4096 1 5 opt:1=1
cap r02_ncu_mmvq_mma_q4k_b8 mmvq_mma_kernel mmvq 12 14336 4096 8 5 opt:1=1
cap r02_ncu_mmq_q4k_m4096 mmq_kernel mmq 12 14336 4096 4096 3
ls -la gpurun_out
