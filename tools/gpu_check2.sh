#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_kernels_gpu.py > gpurun_out/r2_tests.log 2>&1
echo "pytest(layers+step) exit $?" >> gpurun_out/r2_tests.log
tail -40 gpurun_out/r2_tests.log
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r2_tests_k.log 2>&1
echo "pytest(kernels) exit $?" >> gpurun_out/r2_tests_k.log
tail -15 gpurun_out/r2_tests_k.log
python tools/microbench.py --out gpurun_out/microbench_r2.json --only uq_ > gpurun_out/microbench_r2.log 2>&1
cut -c1-200 gpurun_out/microbench_r2.log
