"""Depthwise 3x3 kernels at MobileNet-v1 / batch-256 shapes: CUDA-event timings (default) or a short run for ncu.
usage: python tools/prof_dw.py [ncu]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pocketflow_b200 import ops  # noqa: E402

SHAPES = [(256, 112, 112, 32, 1), (256, 112, 112, 64, 2), (256, 56, 56, 128, 1), (256, 56, 56, 128, 2), (256, 28, 28, 256, 1),
          (256, 28, 28, 256, 2), (256, 14, 14, 512, 1), (256, 14, 14, 512, 2), (256, 7, 7, 1024, 1)]


def main():
    short = len(sys.argv) > 1
    dev = 'cuda:0'
    flush = torch.empty(64 << 20, device=dev)
    for n, h, w, c, st in SHAPES[:1] if short else SHAPES:
        p = (h + st - 1) // st
        pt = max((p - 1) * st + 3 - h, 0) // 2
        d = ops.conv_desc(n, h, w, c, c, 3, 3, p, p, st, st, pt, pt)
        x, wt = torch.randn(n, h, w, c, device=dev), torch.randn(3, 3, c, 1, device=dev)
        y, dy = torch.empty(n, p, p, c, device=dev), torch.randn(n, p, p, c, device=dev)
        dx, dw = torch.empty_like(x), torch.empty_like(wt)
        ws = torch.empty(max(ops.dwconv_wgrad_workspace_floats(d), 4), device=dev)
        fns = dict(fwd=lambda: ops.dwconv_fwd(d, x, wt, y), dgrad=lambda: ops.dwconv_dgrad(d, dy, wt, False, dx),
                   wgrad=lambda: ops.dwconv_wgrad(d, x, dy, ws, dw))
        line = '%4dx%-4d C=%-4d s%d ' % (h, w, c, st)
        for name, fn in fns.items():
            ts = []
            for _ in range(2 if short else 5):
                flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts)[len(ts) // 2]
            gb = (x.numel() + y.numel()) * 4 / 1e9
            line += ' %s %.3f ms %5.0f GB/s' % (name, t, gb / t * 1e3)
        print(line, flush=True)


if __name__ == '__main__':
    main()
