#!/bin/bash
# round-2 call h (4 GPUs): configs[3] (MobileNet-v1 channel-pruned step) at N = 4 — and, on GPU 0 meanwhile-idle boxes
# are charged anyway, the GPU tests changed since call f
mkdir -p gpurun_out
NGPU=4 WORKLOADS=mobilenet_cpg50_b256 bash tools/gpu_round2_g.sh
cp gpurun_out/r2g_bench_mobilenet_cpg50_b256_n4.json gpurun_out/r2h_bench_mobilenet_cpg50_b256_n4.json 2>/dev/null
timeout 900 python -m pytest tests/test_learners_gpu.py tests/test_layers_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2h_gputests.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/r2h_gputests.log | cut -c1-220
