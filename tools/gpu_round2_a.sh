#!/bin/bash
# round-2 checkpoint call: GPU suite, default bench line, launch list of one step + conv traffic stamp
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2a_gputests.log
timeout 600 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc $?"; cut -c1-1500 gpurun_out/r2a_bench_n1.json
bash tools/gpu_launchlist.sh
python tools/conv_traffic_from_launchlist.py gpurun_out/launches_resnet50.csv resnet50_uq8_dst_b256 256 && cp profiles/r2_ncu_conv_traffic.json gpurun_out/
python tools/summarize_launches.py gpurun_out/launches_resnet50.csv > gpurun_out/r2a_launchlist_summary.txt 2>&1; head -50 gpurun_out/r2a_launchlist_summary.txt
