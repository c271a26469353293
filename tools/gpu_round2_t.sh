#!/bin/bash
# round-2 call t: HEAD sanity — smoke() and a short default bench
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2t_smoke.log
timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['traffic'] is not None)"
