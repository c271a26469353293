#!/bin/bash
# GPU probe of the TMA-fed conv kernels: one process per group under `timeout` (a device trap ends one group only).
# usage: gpurun --timeout 900 -- 'bash tools/gpu_tma_probe.sh [groups...]'
mkdir -p gpurun_out
groups="${@:-onehot fwd dgrad wgrad levels}"
for g in $groups; do
  echo "===== $g" | tee -a gpurun_out/tma_probe.log
  timeout 240 python tools/tma_probe.py $g 2>&1 | tail -150 | tee -a gpurun_out/tma_probe.log
done
grep -E "^PROBE" gpurun_out/tma_probe.log
