#!/bin/bash
# generic register-resident batch-1 kernel (mmvq_rgx.cu): parity for every type, then timings against the round-1 generic kernel
python -m pytest tests/test_gpu_mmvq.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_canaries.py -x -q -m gpu 2>&1 | tail -4
for t in 2 3 6 7 8 10 11 20; do
  for rg in 1 0; do
    python scripts/mmvq_probe.py time types=$t static=1 B=1 rg=$rg reps=4 shapes=14336x4096,4096x14336,28672x8192 2>&1 | sed "s/^/rg=$rg /"
  done
done
for t in 13 14 23; do
  for rg in 3 1; do
    python scripts/mmvq_probe.py time types=$t static=1 B=1 rg=$rg reps=4 shapes=14336x4096,4096x14336 2>&1 | sed "s/^/rg=$rg /"
  done
done
