"""Generates tests/golden/hotpath_v1.npz from the oracle (oracle/pf_oracle.py).

The reference (TF 1.x) cannot be imported in this image, so these are NOT reference outputs:
they freeze the oracle restatement so that (a) the oracle cannot drift silently and (b) the GPU
tests have committed vectors to hit even if oracle/ is unavailable.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pf_oracle as O  # noqa: E402

F32 = np.float32


def main():
    rng = np.random.RandomState(20260922)
    g = {}
    # weights: HWIO kernels of the shapes the configs use (+ ragged ones)
    shapes = [(3, 3, 16, 16), (1, 1, 16, 32), (5, 5, 3, 7), (64, 10), (3, 3, 5, 1), (7,)]
    for i, shp in enumerate(shapes):
        fan_in = int(np.prod(shp[:-1])) if len(shp) > 1 else 1
        w = (rng.randn(*shp) * np.sqrt(2.0 / fan_in)).astype(F32)
        g['w%d' % i] = w
        for bits in (2, 4, 8):
            g['w%d_layer_b%d' % (i, bits)] = O.uniform_quantize(w, bits)
            g['w%d_channel_b%d' % (i, bits)] = O.uniform_quantize(w, bits, use_buckets=True, bucket_type='channel')
            g['w%d_split_b%d' % (i, bits)] = O.uniform_quantize(w, bits, use_buckets=True, bucket_type='split',
                                                                bucket_size=16)
    a = np.maximum(rng.randn(2, 6, 6, 8), 0).astype(F32)
    g['act'] = a
    g['act_b8'] = O.uniform_quantize(a, 8, mode='activation')
    g['act_b32'] = O.uniform_quantize(a, 32, mode='activation')
    # masks
    w = rng.randn(777).astype(F32)
    w[::13] = w[5]                                  # ties
    bk = rng.randn(777).astype(F32)
    mk = (rng.rand(777) > 0.3).astype(F32)
    g['ws_w'], g['ws_bkup'], g['ws_mask'] = w, bk, mk
    for r in (0.0, 0.3, 0.5, 0.9):
        v2, b2, m2, thr = O.ws_build_mask(w, bk, mk, r)
        tag = 'ws_r%02d' % int(r * 100)
        g[tag + '_w'], g[tag + '_bkup'], g[tag + '_mask'], g[tag + '_thr'] = v2, b2, m2, np.array([thr], F32)
    # optimizers
    p, acc, gr = rng.randn(101).astype(F32), rng.randn(101).astype(F32) * F32(.1), rng.randn(101).astype(F32)
    msk = (rng.rand(101) > 0.5).astype(F32)
    g['opt_w'], g['opt_acc'], g['opt_g'], g['opt_mask'] = p, acc, gr, msk
    g['mom_w'], g['mom_acc'] = O.momentum_step(p, acc, gr, 0.05, 0.9, mask=msk, wd=1e-4, grad_scale=0.5)
    m0, v0 = rng.randn(101).astype(F32) * F32(.01), np.abs(rng.randn(101)).astype(F32) * F32(.001)
    g['adam_m0'], g['adam_v0'] = m0, v0
    b1p, b2p = F32(0.9) * F32(0.9), F32(0.999) * F32(0.999)
    g['adam_w'], g['adam_m'], g['adam_v'] = O.adam_step(p, m0, v0, gr, 1e-3, b1p, b2p, wd=2e-4)
    # losses
    s, t = (rng.randn(16, 10) * 3).astype(F32), (rng.randn(16, 10) * 3).astype(F32)
    lab = np.eye(10, dtype=F32)[rng.randint(0, 10, 16)]
    g['ce_s'], g['ce_t'], g['ce_lab'] = s, t, lab
    lh, gh = O.softmax_cross_entropy(lab, s)
    ld, gd = O.distillation_loss(s, t, 4.0, 4.0)
    g['ce_hard'], g['ce_dst'], g['ce_grad'] = np.array([lh], F32), np.array([ld], F32), (gh + gd).astype(F32)
    # codebook
    wq = (rng.randn(3, 3, 8, 8) * 0.1).astype(F32)
    qx, c, idx = O.nonuniform_quantize(wq, 4)
    g['nuq_w'], g['nuq_q'], g['nuq_c'], g['nuq_idx'] = wq, qx, c, idx.astype(np.uint8)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hotpath_v1.npz')
    np.savez_compressed(out, **g)
    print('wrote', out, os.path.getsize(out), 'bytes,', len(g), 'arrays')


if __name__ == '__main__':
    main()
