#!/bin/bash
# round-2 call j: the paired-pixel stem wgrad (MobileNet), NUQ RL test, MobileNet bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learners_gpu.py tests/test_search_and_preproc_e2e_gpu.py tests/test_bench_configs_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2j_gputests.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/r2j_gputests.log | cut -c1-220
PF_BENCH_WORKLOAD=mobilenet_cpg50_b256 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_mobilenet.json 2> gpurun_out/r2j_bench_mobilenet.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_mobilenet.json')); print(d['ms_per_step'], d['e2e']['value'], d['step_breakdown_ms'])"
