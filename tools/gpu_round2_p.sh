#!/bin/bash
# round-2 call p: the GPU suite as the driver runs it (final state)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2p_gputests.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2p_gputests.log | cut -c1-220
