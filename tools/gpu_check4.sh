#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_step_gpu.py -m gpu -q 2>&1 | tail -25
timeout 200 python tools/bench_conv.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-22s fwd tc %.3f ms %5.1f TF | fp32 %.3f ms | dgrad tc %.3f ms %5.1f TF | fp32 %.3f ms | wgrad tc %.3f ms %5.1f TF | fp32 %.3f ms' % (r['layer'], r['fwd_tc_ms'], r['fwd_tc_tflops'], r['fwd_fp32_ms'], r['dgrad_tc_ms'], r['dgrad_tc_tflops'], r['dgrad_fp32_ms'], r['wgrad_tc_ms'], r['wgrad_tc_tflops'], r['wgrad_fp32_ms']))"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r4_resnet50.json 2> gpurun_out/bench_r4_resnet50.err
echo "bench resnet50 exit $?"; tail -3 gpurun_out/bench_r4_resnet50.err; python -c "
import json; r=json.load(open('gpurun_out/bench_r4_resnet50.json')); print(r['value'], r['ms_per_step'], r['e2e'], r['roofline']['achieved'], r['step_breakdown_ms'], r['losses_last_step'])"
