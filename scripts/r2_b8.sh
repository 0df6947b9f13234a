#!/bin/bash
# batch 3-8 tensor-core MMVQ: per-CTA timeline + chunk / team sweeps
for sh in "N=14336 K=4096" "N=4096 K=14336"; do
  for B in 8 3; do python scripts/mmvq_trace.py $sh B=$B static=1 2>&1 | tail -4; done
done
for mma in 0,0,0 1,0,0 2,0,0 4,0,0 2,3,0 2,2,0 4,3,0 1,4,0; do
  python scripts/mmvq_probe.py time types=12 static=1 B=8 mma=$mma reps=4 shapes=14336x4096,4096x14336 2>&1 | sed "s/^/mma=$mma /"
done
