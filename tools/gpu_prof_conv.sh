#!/bin/bash
# ncu --set full of one launch of the fwd (MODE 0, planes) and the wgrad persistent kernels at the ResNet-50
# stage-3 3x3 shape (N=256, 14x14, 256 -> 256); reports land in gpurun_out/, summaries go to profiles/.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist_kernel -s 2 -c 1 \
  -f -o gpurun_out/prof_conv_fwd_v5 python tools/prof_conv.py s3 > gpurun_out/ncu_conv_fwd.log 2>&1
tail -2 gpurun_out/ncu_conv_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad_persist_kernel -s 2 -c 1 \
  -f -o gpurun_out/prof_conv_wgrad_v5 python tools/prof_conv.py s3 > gpurun_out/ncu_conv_wgrad.log 2>&1
tail -2 gpurun_out/ncu_conv_wgrad.log
