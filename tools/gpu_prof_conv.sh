#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/prof_conv_tc2 python tools/prof_conv.py s3 > gpurun_out/ncu_conv2.log 2>&1
tail -2 gpurun_out/ncu_conv2.log
