#!/bin/bash
# round-2 call d (2 GPUs): GPU suite (rank-local) then the data-parallel checks through pf_allreduce_flat
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_gputests.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/r2d_gputests.log | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_check.py > gpurun_out/r2d_mgpu_check.log 2>&1; echo "mgpu_check rc $?"; tail -12 gpurun_out/r2d_mgpu_check.log | cut -c1-250
for b in 2 1; do
PF_AR_BUCKETS=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$b bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2_buckets$b.json 2> gpurun_out/r2d_bench_n2_buckets$b.err; echo "bench n2 buckets=$b rc $?"; cut -c1-260 gpurun_out/r2d_bench_n2_buckets$b.json
done
