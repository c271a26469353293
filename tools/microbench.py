"""Per-kernel HBM roofline micro-benchmarks at the ResNet-50 / batch-256 shapes of SURVEY.md §8(d).

    python tools/microbench.py [--out gpurun_out/microbench.json] [--only NAME]

Timing: CUDA events on the launching stream, 5 warm-ups, 20 timed iterations, L2 flushed (a 256 MB
write) between iterations for tensors smaller than L2.  achieved = ALGORITHMIC bytes / time.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pocketflow_b200 import ops  # noqa: E402


def resnet50_kernel_shapes():
    """HWIO kernels of ResNet-50 v2 (utils/external/resnet_model.py), in creation order."""
    shapes = [(7, 7, 3, 64)]
    cin = 64
    for filters, blocks in zip([64, 128, 256, 512], [3, 4, 6, 3]):
        for b in range(blocks):
            if b == 0:
                shapes.append((1, 1, cin, filters * 4))          # projection shortcut
            shapes += [(1, 1, cin, filters), (3, 3, filters, filters), (1, 1, filters, filters * 4)]
            cin = filters * 4
    shapes.append((2048, 1001))
    return shapes


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured'
    return 6650.0, 'fallback'


class Timer:
    def __init__(self, flush=True):
        self.flush_buf = torch.empty(256 * 1024 * 1024 // 4, device='cuda') if flush else None

    def run(self, fn, iters=20, warm=5):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            if self.flush_buf is not None:
                self.flush_buf.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return float(np.median(ts)), float(ts[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/microbench.json')
    ap.add_argument('--only', default=None)
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    peak, peak_kind = peaks()
    T = Timer()
    res = []

    def report(name, algo_bytes, fn, note=''):
        if args.only and args.only not in name:
            return
        med, best = T.run(fn, args.iters)
        gbs = algo_bytes / (med * 1e-3) / 1e9
        r = dict(kernel=name, ms_median=med, ms_best=best, algorithmic_bytes=algo_bytes, achieved_gbs=gbs,
                 frac_of_peak=gbs / peak, peak_gbs=peak, peak_kind=peak_kind, note=note)
        res.append(r)
        print(json.dumps(r), flush=True)

    shapes = resnet50_kernel_shapes()
    torch.manual_seed(0)
    ws = [torch.randn(s, device='cuda') * (2.0 / np.prod(s[:-1])) ** 0.5 for s in shapes]
    # a1: the 52 quantized tensors (first conv and final dense excluded, utils.py:122-123)
    qsrc = ws[1:-1]
    qdst = [torch.empty_like(w) for w in qsrc]
    nq = sum(w.numel() for w in qsrc)
    for mode, kw in (('layer', {}), ('channel', dict(use_buckets=True, bucket_type='channel')),
                     ('split256', dict(use_buckets=True, bucket_type='split', bucket_size=256))):
        q = ops.UniformWeightQuantizer(qsrc, qdst, 8, **kw)
        report('uq_weight_fwd[%s] resnet50 52 tensors' % mode, 8 * nq, q.forward, 'minmax+quant, 4 launches')
        report('uq_weight_minmax[%s]' % mode, 4 * nq, q.minmax)
        report('uq_weight_quant[%s]' % mode, 8 * nq, q.quantize)
    # a2: activations: the largest ResNet-50 ReLU output (784 MiB) and a mid one
    for label, n in (('256x112x112x64 (784 MiB)', 256 * 112 * 112 * 64), ('256x28x28x512', 256 * 28 * 28 * 512),
                     ('256x32x32x16 (resnet20)', 256 * 32 * 32 * 16)):
        x = torch.relu(torch.randn(n, device='cuda'))
        y = torch.empty_like(x)
        mm = torch.empty(2, dtype=torch.int32, device='cuda')
        ops.act_range_reset(mm)
        report('uq_act_minmax ' + label, 4 * n, lambda: ops.act_minmax(x, mm))
        report('uq_act_quant ' + label, 8 * n, lambda: ops.act_quant(x, y, mm, 8))
        report('uq_act_fwd(minmax+quant) ' + label, 8 * n, lambda: ops.act_fake_quant(x, 8, out=y, minmax=mm),
               'two-pass: 12 B/elem of traffic caps this at 0.67')
        del x, y
    # a6/a9: flat optimizer steps over all 25.5M parameters
    n = sum(w.numel() for w in ws)
    w, acc, g, v = (torch.randn(n, device='cuda') for _ in range(4))
    v.abs_()
    mask = (torch.rand(n, device='cuda') > 0.5).float()
    hp = torch.tensor([0.1, 0.9, 0.999, 0.0], device='cuda')
    report('masked_momentum resnet50 25.5M', 24 * n, lambda: ops.momentum_step(w, acc, g, mask, hp, 0.9, 1e-4, 0.125))
    report('momentum(no mask) resnet50 25.5M', 20 * n, lambda: ops.momentum_step(w, acc, g, None, hp, 0.9, 1e-4, 0.125))
    report('adam resnet50 25.5M', 28 * n, lambda: ops.adam_step(w, acc, v, g, hp, wd=1e-4))
    # a5: mask build over the 54 maskable tensors
    bk = [x.clone() for x in ws]
    mk = [torch.ones_like(x) for x in ws]
    mb = ops.MaskBuilder(ws, bk, mk)
    ratios = [0.5] * len(ws)
    report('ws_mask_build resnet50 54 tensors', 24 * n, lambda: mb.build(ratios),
           '4 radix-select passes + apply; includes one small H2D of the ranks')
    # a7
    s, t = torch.randn(256, 1001, device='cuda'), torch.randn(256, 1001, device='cuda')
    lab = torch.eye(1001, device='cuda')[torch.randint(0, 1001, (256,), device='cuda')].contiguous()
    dl, out, rw = torch.empty_like(s), torch.empty(4, device='cuda'), torch.empty(1024, device='cuda')
    report('softmax_ce 256x1001 hard+dst', 16 * 256 * 1001, lambda: ops.softmax_ce(s, lab, t, 4.0, 4.0, dl, out, rw),
           'latency-bound (1 MB)')
    l2o, l2p = torch.zeros(4, device='cuda'), torch.empty(ops.L2_PARTIALS, device='cuda')
    report('l2_loss 25.5M', 4 * n, lambda: ops.l2_loss(w, 1e-4, l2o, l2p))
    # a11
    cq = ops.CodebookWeightQuantizer(qsrc, qdst, 4)
    cq.quantile_init()
    report('nuq_weight_fwd 4-bit resnet50 52 tensors', 8 * nq, cq.forward, 'minmax + 16-centroid search')
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
