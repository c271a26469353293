#!/bin/bash
# round-2 call s: interior fast path of the row-blocked depthwise kernels — tests, per-layer timings, MobileNet bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -q -k depthwise > gpurun_out/r2s_gputests.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r2s_gputests.log | cut -c1-200
timeout 200 python tools/prof_dw.py > gpurun_out/r2s_dw_layers.log 2>&1; cat gpurun_out/r2s_dw_layers.log
PF_BENCH_WORKLOAD=mobilenet_cpg50_b256 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_mobilenet.json 2> gpurun_out/r2s_bench_mobilenet.err; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_mobilenet.json')); print(d['ms_per_step'], d['e2e']['value'], d['step_breakdown_ms'])"
