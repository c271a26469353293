#!/bin/bash
# First GPU pass: parity tests, HBM micro-benchmarks, ncu evidence for the fake-quant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r1_gpu.csv 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_tests.log
tail -25 gpurun_out/r1_tests.log
python tools/microbench.py --out gpurun_out/microbench_r1.json > gpurun_out/microbench_r1.log 2>&1
echo "microbench exit $?"
tail -40 gpurun_out/microbench_r1.log | cut -c1-260
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uq_act_quant_kernel -s 6 -c 1 \
  -o gpurun_out/prof_r1_act_quant python tools/microbench.py --only 'uq_act_quant 256x112' --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:momentum_kernel -s 6 -c 1 \
  -o gpurun_out/prof_r1_momentum python tools/microbench.py --only 'masked_momentum' --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uq_weight -s 10 -c 2 \
  -o gpurun_out/prof_r1_uq_weight python tools/microbench.py --only 'uq_weight_fwd[channel]' --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
