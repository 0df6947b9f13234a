#!/bin/bash
# round-2 final 2-GPU run: NCCL / fused-exchange parity at world 2, then the bench at N = 2 (with extras) and N = 1 on the same box
python -m pytest tests/test_gpu_parallel.py -x -q -m gpu > gpurun_out/r2_n2_pytest.log 2>&1
tail -3 gpurun_out/r2_n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n2.json 2> gpurun_out/r2_bench_final_n2.err
echo "N=2 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['config']['launch'])
for c in d['extras']['tp_mlp']['cases']: print(c['M'], round(c['us'],1), c['exchange'][:70])"
timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_final_n1_samebox.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final_n1_samebox.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), d['ms_per_step'])"
