#!/bin/bash
# scripts/grun_retry.sh <timeout> '<command>' <logfile>: like grun.sh, retried while the pod answers busy (exit 3)
cd "$(dirname "$0")/.."
python ggml-libtorch_b200/build.py > /dev/null || exit 1
make -s -C oracle || exit 1
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2" > "$3" 2>&1
  rc=$?
  if grep -q "status=transient" "$3"; then sleep 100; continue; fi
  exit $rc
done
