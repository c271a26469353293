#!/bin/bash
# round-2 call g (8 GPUs): bench lines of the 8-GPU configs (BASELINE.json configs[2], configs[4]) and of the default
# workload, data-parallel through pf_allreduce_flat with the bucketed overlap
mkdir -p gpurun_out
N=${NGPU:-8}
port=29600
for wl in ${WORKLOADS:-resnet50_uq8_dst_b256 resnet50_ws50_dst_b256 resnet50_nuq4_dst_b256}; do
  port=$((port+1))
  PF_BENCH_WORKLOAD=$wl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_${wl}_n$N.json 2> gpurun_out/r2g_bench_${wl}_n$N.err
  echo "bench $wl n=$N rc $?"
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2g_bench_${wl}_n$N.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('no line', e); print(open('gpurun_out/r2g_bench_${wl}_n$N.err').read()[-1500:])"
done
