#!/bin/bash
# Build everything locally (so the .so files that travel are current), then run a command on the GPU box.
# Usage: scripts/grun.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
python ggml-libtorch_b200/build.py > /dev/null
make -s -C oracle
mkdir -p gpurun_out
exec /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2"
