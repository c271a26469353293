"""One launch of each tcgen05 conv kernel for ncu (tools/gpu_prof_conv.sh): operands in split-bf16 planes, as
inside the training step.  usage: python tools/prof_conv.py s3 [s1 ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pocketflow_b200 import ops  # noqa: E402

SHAPES = {'s3': (256, 14, 14, 256, 256, 3, 3, 1, 1), 's1': (256, 56, 56, 64, 256, 1, 1, 1, 0),
          's2': (256, 28, 28, 128, 128, 3, 3, 1, 1), 's1c': (256, 56, 56, 64, 64, 3, 3, 1, 1)}
dev = torch.device('cuda:0')
for name in sys.argv[1:]:
    n, h, w, c, k, r, s, st, pd = SHAPES[name]
    p = (h + 2 * pd - r) // st + 1
    d = ops.conv_desc(n, h, w, c, k, r, s, p, p, st, st, pd, pd)
    x = torch.randn(n, h, w, c, device=dev)
    wt = torch.randn(r, s, c, k, device=dev) * 0.05
    y = torch.empty(n, p, p, k, device=dev)
    dy = torch.randn(n, p, p, k, device=dev)
    dx = torch.empty_like(x)
    dw = torch.empty_like(wt)
    tw = ops.TcWeights(d, dev)
    tw.prepare(wt)
    xp, dyp = ops.Planes(x.numel(), dev), ops.Planes(dy.numel(), dev)
    ops.split_bf16(x, xp)
    ops.split_bf16(dy, dyp)
    ws = torch.empty(max(ops.conv2d_tc_wgrad_planes_workspace_floats(d), 4), device=dev)
    resid = torch.randn_like(y) if os.environ.get('RESIDUAL', '0') == '1' else None
    for _ in range(3):
        ops.conv2d_tc_fwd_planes(d, xp, tw, None, False, y, resid)
        ops.conv2d_tc_dgrad_planes(d, dyp, tw, False, dx)
        ops.conv2d_tc_wgrad_planes(d, xp, dyp, ws, dw)
    torch.cuda.synchronize()
