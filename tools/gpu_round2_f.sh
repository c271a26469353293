#!/bin/bash
# round-2 call f: full GPU suite (row-blocked depthwise, new AFF epilogue, CPG / NUQ work), MobileNet config A/B
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2f_gputests.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/r2f_gputests.log | cut -c1-220
for rows in 1 0; do
PF_DW_ROWS=$rows PF_BENCH_WORKLOAD=mobilenet_cpg50_b256 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_mobilenet_rows$rows.json 2> gpurun_out/r2f_bench_mobilenet_rows$rows.err; echo "bench mobilenet rows=$rows rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_mobilenet_rows$rows.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d.get('step_breakdown_ms'))"
done
