#!/bin/bash
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "N=2 rc=$?"; tail -3 gpurun_out/r2_bench_n2b.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n2b.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['e2e'])"
