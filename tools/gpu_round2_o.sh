#!/bin/bash
# round-2 call o: bn_apply_levels with four loads in flight — tests that cover it, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_levels_gpu.py tests/test_layers_gpu.py tests/test_bench_configs_gpu.py tests/test_step_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2o_gputests.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2o_gputests.log | cut -c1-220
timeout 600 python bench.py > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_n1.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['step_breakdown_ms'])"
