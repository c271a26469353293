#!/bin/bash
# round-2 call r: depthwise kernels per layer + one ncu --set full capture of the row-blocked stride-1 kernels
mkdir -p gpurun_out
timeout 300 python tools/prof_dw.py > gpurun_out/r2r_dw_layers.log 2>&1; cat gpurun_out/r2r_dw_layers.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dw3x3s1 -c 3 -o gpurun_out/ncu_r2_dw_rows -f python tools/prof_dw.py ncu > gpurun_out/ncu_r2_dw_rows.log 2>&1
ncu -i gpurun_out/ncu_r2_dw_rows.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/ncu_r2_dw_rows.txt
grep -E "^==|gpu__time_duration.sum|dram_throughput|lts__throughput.avg|l1tex__throughput|issue_active|registers_per_thread|warps_active|long_scoreboard_per|lg_throttle_per|mio_throttle_per|wait_per|not_selected_per" gpurun_out/ncu_r2_dw_rows.txt | cut -c1-170
