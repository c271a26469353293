#!/bin/bash
# 2-GPU pass: replica consistency + scaling point
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r5_n2.json 2> gpurun_out/bench_r5_n2.err
echo "bench n2 exit $?"; tail -3 gpurun_out/bench_r5_n2.err | cut -c1-300
python -c "
import json; r=json.load(open('gpurun_out/bench_r5_n2.json')); print(r['n_gpus'], r['value'], r['ms_per_step'], r['e2e']['value'], r['config']['cuda_graph'], r['step_breakdown_ms'])"
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r5_n1.json 2> gpurun_out/bench_r5_n1.err
python -c "
import json; r=json.load(open('gpurun_out/bench_r5_n1.json')); print(r['n_gpus'], r['value'], r['ms_per_step'], r['e2e']['value'], r['step_breakdown_ms'])"
timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -q 2>&1 | tail -3
