#!/bin/bash
# First GPU call of the next round: everything that was written after the previous round's GPU budget ran out.
#   1. pf_preprocess_images against its numpy statement (tests/test_preproc_gpu.py, opt-in)
#      (then: run a real-data ILSVRC-12 step with --enbl_device_preprocess and compare the image placeholder with the
#      host pipeline's batch before turning the flag on by default)
#   2. the RL bit search through the real UniformQuantLearner (tools/rl_smoke.py)
#   3. the regular GPU suite + a bench line, to confirm nothing else moved
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_validate_unverified.sh'
mkdir -p gpurun_out
PF_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_preproc_gpu.py -q 2>&1 | tail -5 | tee gpurun_out/unverified_preproc.log
timeout 300 python tools/preproc_e2e_check.py 2>&1 | tail -4 | tee gpurun_out/unverified_preproc_e2e.log
timeout 600 python tools/rl_smoke.py 2>&1 | tail -8 | tee gpurun_out/unverified_rl.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/gpu_suite.log
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json
