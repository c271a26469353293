#!/bin/bash
# ncu launch list of training steps of the default bench workload (eager, no CUDA graph):
# per-launch device time and DRAM bytes (single replay pass; cold-cache serialised -> compare SHARES).
# tools/summarize_launches.py cuts ONE step out of it (between two softmax-CE launches).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -s 2400 -c 2200 --csv --log-file gpurun_out/launches_resnet50.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launchlist_bench.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/launchlist_bench.log | cut -c1-300
wc -l gpurun_out/launches_resnet50.csv
