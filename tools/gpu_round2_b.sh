#!/bin/bash
# round-2 call b: per-layer table of the current conv kernels (tma / levels feeds) + ncu --set full captures of the
# TMA-fed kernels on a stage-3 3x3 (tensor-bound candidate) and a stage-1 1x1 (HBM / epilogue-bound) layer
mkdir -p gpurun_out
rm -f gpurun_out/bench_conv.log
VARIANTS="tma levels" bash tools/gpu_bench_conv.sh > /dev/null 2>&1
cat gpurun_out/bench_conv.log | tail -40
cap() {  # tag layer passes variant kernel
  ONLY="$2" PASSES=$3 VARIANT=$4 KERNEL=$5 TAG=$1 bash tools/gpu_ncu_conv.sh > gpurun_out/ncu_$1.stdout 2>&1
  grep -E "gpu__time_duration.sum|pipe_tensor_cycles|dram_throughput|lts__throughput" gpurun_out/ncu_$1.txt | cut -c1-170
}
cap r2_s3_3x3_fwd_levels 's3 3x3 256->256' fwd levels conv_tma_kernel
cap r2_s3_3x3_fwd_split 's3 3x3 256->256' fwd tma conv_tma_kernel
cap r2_s3_3x3_wgrad_levels 's3 3x3 256->256' wgrad levels conv_tma_wgrad
cap r2_s1_1x1_64_256_fwd_levels 's1 1x1 64->256' fwd levels conv_tma_kernel
cap r2_s2_3x3_fwd_levels 's2 3x3 128->128' fwd levels conv_tma_kernel
ls -la gpurun_out | head -40
