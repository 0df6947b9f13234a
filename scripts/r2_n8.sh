#!/bin/bash
# round-2 multi-GPU run: NCCL parity at world 2/4/8, then the bench at N = 8 (with extras), 4, 2
python -m pytest tests/test_gpu_parallel.py -x -q -m gpu > gpurun_out/r2_final_n8_pytest.log 2>&1
tail -3 gpurun_out/r2_final_n8_pytest.log
for n in 8 4; do
  extra="--no-extras"; [ $n = 8 ] && extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 $extra > gpurun_out/r2_bench_final_n$n.json 2> gpurun_out/r2_bench_final_n$n.err
  echo "N=$n rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final_n$n.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['config']['launch'])"
done
timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_final_n1_same_box_as_n8.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_final_n1_same_box_as_n8.json')); print(d['n_gpus'], round(d['value']), d['ms_per_step'])"
