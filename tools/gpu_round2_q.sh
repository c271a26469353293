#!/bin/bash
# round-2 call q (2 GPUs): last data-parallel regression of the round — mgpu_check and the default bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_check.py > gpurun_out/r2q_mgpu_check.log 2>&1; echo "mgpu_check rc $?"; grep -E "mgpu_check ok|Error|assert" gpurun_out/r2q_mgpu_check.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "bench n2 rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_n2.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29548 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2q_ref_n2.json 2> gpurun_out/r2q_ref_n2.err; echo "reference arm n2 rc $?"; cut -c1-200 gpurun_out/r2q_ref_n2.json
