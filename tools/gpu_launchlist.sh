#!/bin/bash
# ncu launch list of the default bench workload's training step (eager, no CUDA graph): per-launch device time and
# DRAM bytes (single replay pass; cold-cache serialised -> compare SHARES).  tools/summarize_launches.py cuts ONE
# step out of it (between two softmax-CE launches).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/launches_resnet50.csv python tools/one_step.py resnet50_uq8_dst_b256 2 > gpurun_out/launchlist_bench.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/launchlist_bench.log | cut -c1-300
wc -l gpurun_out/launches_resnet50.csv
