#!/bin/bash
# one ncu --set full capture of a single conv layer of tools/bench_conv_tc.py
# usage: ONLY='s1 1x1 64->256' PASSES=fwd VARIANT=levels KERNEL=conv_tma_kernel TAG=x bash tools/gpu_ncu_conv.sh
mkdir -p gpurun_out
FLUSH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-conv_tma} -s 3 -c 2 \
  -o gpurun_out/ncu_${TAG:-conv} -f python tools/bench_conv_tc.py ncu > gpurun_out/ncu_${TAG:-conv}.log 2>&1
tail -3 gpurun_out/ncu_${TAG:-conv}.log
ncu -i gpurun_out/ncu_${TAG:-conv}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/ncu_${TAG:-conv}.txt
cat gpurun_out/ncu_${TAG:-conv}.txt | head -80
