#!/bin/bash
# last GPU call of round 2: MoE parity, MoE timing, then the full GPU suite on the final library (as far as the lease goes)
export PYTHONUNBUFFERED=1
timeout 55 python -m pytest tests/test_gpu_moe.py -q > gpurun_out/r2_moe_pytest.log 2>&1; echo "moe pytest rc=$?"; tail -3 gpurun_out/r2_moe_pytest.log
timeout 30 python scripts/moe_bench.py --reps 5 > gpurun_out/r2_moe_bench.json 2> gpurun_out/r2_moe_bench.err; echo "moe bench rc=$?"
timeout 100 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_moe.py > gpurun_out/r2_moe_full_pytest.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/r2_moe_full_pytest.log
