#!/bin/bash
# round-2 call k: GPU suite; A/B of the tile width for the conv3 + residual layers; MobileNet + default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2k_gputests.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2k_gputests.log | cut -c1-220
echo "== conv3 + residual, split planes (teacher) and levels (student): BN 256 (default) vs 128"
for bn in 0 128; do for v in tma levels; do
  echo "-- PF_TC_BN=$bn $v"
  PF_TC_BN=$bn RESIDUAL=1 PASSES=fwd VARIANT=$v ONLY="64->256" timeout 300 python tools/bench_conv_tc.py ab 2>&1 | grep "64->256"
  PF_TC_BN=$bn RESIDUAL=1 PASSES=fwd VARIANT=$v ONLY="128->512" timeout 300 python tools/bench_conv_tc.py ab 2>&1 | grep "128->512"
  PF_TC_BN=$bn RESIDUAL=1 PASSES=fwd VARIANT=$v ONLY="256->1024" timeout 300 python tools/bench_conv_tc.py ab 2>&1 | grep "256->1024"
done; done
PF_BENCH_WORKLOAD=mobilenet_cpg50_b256 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_mobilenet.json 2> gpurun_out/r2k_bench_mobilenet.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_mobilenet.json')); print(d['ms_per_step'], d['e2e']['value'], d['step_breakdown_ms'])"
timeout 600 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_n1.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_step'], d['step_breakdown_ms'])"
