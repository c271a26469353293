#!/bin/bash
# round-2 call m (4 GPUs): configs[3] at N = 4 with the final kernels; the eval-mode restore test incl. chn-pruned-gpu
mkdir -p gpurun_out
NGPU=4 WORKLOADS=mobilenet_cpg50_b256 bash tools/gpu_round2_g.sh
cp gpurun_out/r2g_bench_mobilenet_cpg50_b256_n4.json gpurun_out/r2m_bench_mobilenet_cpg50_b256_n4.json 2>/dev/null
timeout 600 python -m pytest tests/test_learners_gpu.py -m gpu -q --timeout 600 -k "exec_mode or restore" > gpurun_out/r2m_gputests.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2m_gputests.log | cut -c1-220
