#!/bin/bash
# round-2 call e (2 GPUs): re-run of the tests fixed after call d, the data-parallel check, per-layer conv table and
# the default bench line with the new AFF=2 epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learners_gpu.py tests/test_bench_configs_gpu.py tests/test_tc_gpu.py tests/test_levels_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2e_gputests.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r2e_gputests.log | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_check.py > gpurun_out/r2e_mgpu_check.log 2>&1; echo "mgpu_check rc $?"; grep -E "mgpu_check ok|Error|assert" gpurun_out/r2e_mgpu_check.log | head -5
rm -f gpurun_out/bench_conv.log; VARIANTS="levels" PASSES=fwd bash tools/gpu_bench_conv.sh > /dev/null 2>&1; tail -19 gpurun_out/bench_conv.log
timeout 600 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_n1.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_step'], d['step_breakdown_ms'])"
