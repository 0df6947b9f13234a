#!/bin/bash
# round-2 final single-GPU evidence: full GPU suite, smoke, bench (both arms), reference CUDA comparator, ncu launch list + captures
python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; tail -2 gpurun_out/r2_final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r2_bench_final_reference.json 2>/dev/null
python scripts/ref_cuda_bench.py > gpurun_out/r2_ref_cuda_bench_final.json 2>/dev/null
python benchmarks/benchmark_mmq.py --num-tokens 4096 --hidden-size 4096 --num-iters 20 > gpurun_out/r2_benchmark_mmq_cli.log 2>&1; tail -4 gpurun_out/r2_benchmark_mmq_cli.log
python bench.py --steps 2 --warmup 3 --no-extras --layers 8 > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --layers 8 > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/op_run.py mmvq 12 14336 4096 1 5 opt:1=1 > gpurun_out/r2_ncu_rg_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mmvq_rg_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_mmvq_rg_q4k_b1 -f python scripts/op_run.py mmvq 12 14336 4096 1 5 opt:1=1 > gpurun_out/r2_ncu_rg.log 2>&1
python scripts/op_run.py mmvq 14 14336 4096 1 5 opt:1=1 > gpurun_out/r2_ncu_rg6_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mmvq_rg_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_mmvq_rg_q6k_b1 -f python scripts/op_run.py mmvq 14 14336 4096 1 5 opt:1=1 > gpurun_out/r2_ncu_rg6.log 2>&1
python scripts/op_run.py mmvq 2 14336 4096 1 5 opt:1=1 > gpurun_out/r2_ncu_rgx_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mmvq_rgx_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_mmvq_rgx_q4_0_b1 -f python scripts/op_run.py mmvq 2 14336 4096 1 5 opt:1=1 > gpurun_out/r2_ncu_rgx.log 2>&1
python scripts/op_run.py mmvq 12 14336 4096 8 5 opt:1=1 > gpurun_out/r2_ncu_mma_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mmvq_mma_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_mmvq_mma_q4k_b8 -f python scripts/op_run.py mmvq 12 14336 4096 8 5 opt:1=1 > gpurun_out/r2_ncu_mma.log 2>&1
python scripts/op_run.py mmq 12 14336 4096 4096 3 > gpurun_out/r2_ncu_mmq_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mmq_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_mmq_q4k_m4096 -f python scripts/op_run.py mmq 12 14336 4096 4096 3 > gpurun_out/r2_ncu_mmq.log 2>&1
ls -la gpurun_out/*.ncu-rep
