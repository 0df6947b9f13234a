#!/bin/bash
# round-2 final single-GPU evidence, part A: full GPU suite, smoke, bench (both arms), reference CUDA comparator, CLI, ncu launch list
python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; tail -2 gpurun_out/r2_final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r2_bench_final_reference.json 2>/dev/null
python scripts/ref_cuda_bench.py > gpurun_out/r2_ref_cuda_bench_final.json 2>/dev/null
python benchmarks/benchmark_mmq.py --num-tokens 4096 --hidden-size 4096 --num-iters 20 > gpurun_out/r2_benchmark_mmq_cli.log 2>&1; tail -4 gpurun_out/r2_benchmark_mmq_cli.log
python bench.py --steps 2 --warmup 3 --no-extras --layers 8 > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --layers 8 > gpurun_out/r2_ncu_launches.log 2>&1
ls -la gpurun_out | head -30
