#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_final_n8b.json 2> gpurun_out/r2_bench_final_n8b.err
echo "N=8 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final_n8b.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['e2e'])"
