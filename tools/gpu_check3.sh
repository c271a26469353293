#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_step_gpu.py -m gpu -q > gpurun_out/r3_tests.log 2>&1
echo "pytest(step) exit $?" >> gpurun_out/r3_tests.log
tail -30 gpurun_out/r3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r3_smoke.log
python bench.py --workload resnet20_uq8_dst_b256 --steps 20 --warmup 3 > gpurun_out/bench_r3_resnet20.json 2> gpurun_out/bench_r3_resnet20.err
echo "bench resnet20 exit $?"; tail -3 gpurun_out/bench_r3_resnet20.err; cut -c1-1500 gpurun_out/bench_r3_resnet20.json
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r3_resnet50.json 2> gpurun_out/bench_r3_resnet50.err
echo "bench resnet50 exit $?"; tail -3 gpurun_out/bench_r3_resnet50.err; cut -c1-2500 gpurun_out/bench_r3_resnet50.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
