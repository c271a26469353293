"""BN kernel timing at ResNet-50 / batch-256 shapes (HBM-bound): stats(+range), apply+quant -> planes, backward.
usage: python tools/bench_bn.py [tag]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pocketflow_b200 import ops  # noqa: E402

SHAPES = [('s1 56x56x256', 256 * 56 * 56, 256), ('s1 56x56x64', 256 * 56 * 56, 64), ('s2 28x28x512', 256 * 28 * 28, 512),
          ('s3 14x14x1024', 256 * 14 * 14, 1024), ('s3 14x14x256', 256 * 14 * 14, 256), ('s4 7x7x2048', 256 * 7 * 7, 2048)]


def timeit(fn, iters=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device('cuda:0')
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for name, m, c in SHAPES:
        x = torch.randn(m, c, device=dev)
        dy = torch.randn(m, c, device=dev)
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
        mean, var, rstd, dga, dbe = [torch.empty(c, device=dev) for _ in range(5)]
        ws = torch.empty(5 * c * ops.BN_MAX_SPLITS, device=dev)
        slot = torch.zeros(1, 2, dtype=torch.int32, device=dev)
        pl, gp = ops.Planes(m * c, dev), ops.Planes(m * c, dev)
        n = m * c

        def stats():
            ops.minmax_reset(slot)
            ops.bn_train_stats_range(x, m, c, 1e-5, 0.9, mean, var, rstd, None, None, gamma, beta, 1, slot[0], ws)
        stats()
        fns = dict(stats=(stats, 4), apply_q=(lambda: ops.bn_apply_quant(x, m, c, mean, rstd, gamma, beta, 1, slot[0], 8, None, pl), 8),
                   bwd=(lambda: ops.bn_bwd(dy, x, m, c, mean, rstd, gamma, beta, 1, dga, dbe, None, False, ws, gp), 20))
        line = '%-16s' % name
        for k, (fn, bpe) in fns.items():
            def run():
                flush.zero_()
                fn()
            t = timeit(run) - timeit(lambda: flush.zero_())
            line += '  %s %.3f ms %5.0f GB/s' % (k, t, bpe * n / t / 1e6)
        print(line, flush=True)


if __name__ == '__main__':
    main()
