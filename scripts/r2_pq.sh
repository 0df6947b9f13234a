#!/bin/bash
echo "== no pre-launch: SwiGLU inside the down kernel's pair-shared quantization"
LAYERS=24 MIX=12 python scripts/block_breakdown.py opt:8=0
VERBOSE=0 TD=12 python scripts/block_trace.py opt:8=0
echo "== same, clusters of 4"
LAYERS=24 MIX=12 python scripts/block_breakdown.py opt:8=0 opt:7=4194305
