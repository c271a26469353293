#!/bin/bash
# round-2 call n: smoke() as the driver runs it, then the default bench line (final state of the round)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2n_smoke.log
timeout 600 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_n1.json')); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_hbm'])"
