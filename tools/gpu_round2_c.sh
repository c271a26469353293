#!/bin/bash
# round-2 call c: GPU suite after the eval-mode / CPG selection / NUQ cluster-mode work
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2c_gputests.log 2>&1; echo "pytest rc $?"; tail -40 gpurun_out/r2c_gputests.log
