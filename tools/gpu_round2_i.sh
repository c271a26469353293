#!/bin/bash
# round-2 call i: tests of the kernels changed since call f, per-layer conv tables (incl. conv3 + residual),
# default bench line, MobileNet launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2i_gputests.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2i_gputests.log | cut -c1-220
rm -f gpurun_out/bench_conv.log
VARIANTS="tma levels" bash tools/gpu_bench_conv.sh > /dev/null 2>&1; tail -40 gpurun_out/bench_conv.log
echo "== conv3 + residual layers"
for v in tma levels; do RESIDUAL=1 PASSES=fwd VARIANT=$v ONLY="->256" timeout 300 python tools/bench_conv_tc.py res_$v 2>&1 | grep -E "64->256|->" | head -6; RESIDUAL=1 PASSES=fwd VARIANT=$v ONLY="128->512" timeout 300 python tools/bench_conv_tc.py res2_$v 2>&1 | grep "128->512"; done
timeout 600 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_n1.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_step'], d['step_breakdown_ms'])"
PF_BENCH_WORKLOAD=mobilenet_cpg50_b256 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_mobilenet.json 2> gpurun_out/r2i_bench_mobilenet.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_mobilenet.json')); print(d['ms_per_step'], d['e2e']['value'], d['step_breakdown_ms'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_mobilenet.csv python tools/one_step.py mobilenet_cpg50_b256 2 > gpurun_out/launchlist_mobilenet.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_mobilenet.csv > gpurun_out/r2i_launchlist_mobilenet.txt 2>&1; head -40 gpurun_out/r2i_launchlist_mobilenet.txt
