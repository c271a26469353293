#!/bin/bash
# round-2 call l (final evidence of the round): GPU suite, default bench, ncu launch list of ONE step of the same
# binary + the conv stack's DRAM bytes stamped with the kernel sources' hash, then the bench line that reads the stamp
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=12 > gpurun_out/r2l_gputests.log 2>&1; echo "pytest rc $?"; tail -20 gpurun_out/r2l_gputests.log | cut -c1-200
bash tools/gpu_launchlist.sh
python tools/conv_traffic_from_launchlist.py gpurun_out/launches_resnet50.csv resnet50_uq8_dst_b256 256 && cp profiles/r2_ncu_conv_traffic.json gpurun_out/
python tools/summarize_launches.py gpurun_out/launches_resnet50.csv > gpurun_out/r2l_launchlist_summary.txt 2>&1; head -32 gpurun_out/r2l_launchlist_summary.txt
timeout 600 python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_n1.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_step'], d['roofline']['traffic'], d['step_breakdown_ms'])"
