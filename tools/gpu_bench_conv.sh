#!/bin/bash
# per-layer conv timings of the three operand feeds at ResNet-50 / batch-256 shapes
mkdir -p gpurun_out
for v in ${VARIANTS:-lsu tma levels}; do
  echo "===== $v" | tee -a gpurun_out/bench_conv.log
  VARIANT=$v timeout 600 python tools/bench_conv_tc.py r2_$v 2>&1 | tail -25 | tee -a gpurun_out/bench_conv.log
done
