"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs, against the committed golden vectors, and — at the configs' full sizes — through
size-independent properties.  Bar: bit-exact for quantized weights/activations, masks, thresholds
and optimizer state (every fp32 op individually rounded, like the oracle); 1e-5 relative for the
transcendental loss kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import pf_oracle as O
from pocketflow_b200 import ops

pytestmark = pytest.mark.gpu
F32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev():
    return torch.device('cuda:0')


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def he(rng, shape):
    fan_in = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
    return (rng.randn(*shape) * np.sqrt(2.0 / fan_in)).astype(F32)


SHAPES = [(3, 3, 16, 16), (1, 1, 16, 32), (3, 3, 64, 64), (5, 5, 32, 64), (1600, 256), (5, 5, 3, 7),
          (64, 10), (3, 3, 5, 1), (7,), (1,), (3, 3, 3, 1001), (1, 1, 256, 1024)]


@pytest.mark.parametrize('mode', ['layer', 'channel', 'split'])
@pytest.mark.parametrize('bits', [1, 2, 4, 8, 32])
def test_weight_fake_quant_bit_exact(mode, bits):
    rng = np.random.RandomState({'layer': 1, 'channel': 2, 'split': 3}[mode] * 100 + bits)
    ws = [he(rng, s) for s in SHAPES]
    ws[2][0, 0, :, 3] = 0.25                                   # a constant channel (alpha = 1e-10)
    src = [cu(w) for w in ws]
    dst = [torch.empty_like(s) for s in src]
    kw = dict(use_buckets=mode != 'layer', bucket_type='split' if mode == 'split' else 'channel',
              bucket_size=256)
    q = ops.UniformWeightQuantizer(src, dst, bits, **kw)
    q.forward()
    rngs = q.ranges()
    for i, (w, d) in enumerate(zip(ws, dst)):
        ref, alpha, beta = O.uniform_quantize(w, bits, return_scales=True, **kw)
        got = d.cpu().numpy()
        assert np.array_equal(got, ref), 'tensor %d %s mismatch' % (i, SHAPES[i])
        assert np.array_equal(np.atleast_1d(beta), rngs[i][0])
    assert q.bucket_storage_bits() == sum(O.bucket_storage_bits(c) for c in q.bucket_counts)


def test_weight_fake_quant_mixed_bits_inplace_and_split_small_bucket():
    rng = np.random.RandomState(5)
    ws = [he(rng, s) for s in SHAPES[:6]]
    bits = [2, 3, 8, 5, 8, 1]
    src = [cu(w) for w in ws]
    q = ops.UniformWeightQuantizer(src, src, bits, use_buckets=True, bucket_type='split', bucket_size=16)
    q.forward()
    for w, d, b in zip(ws, src, bits):
        ref = O.uniform_quantize(w, b, use_buckets=True, bucket_type='split', bucket_size=16)
        assert np.array_equal(d.cpu().numpy(), ref)


def test_weight_ste_backward_bit_exact():
    rng = np.random.RandomState(6)
    ws = [he(rng, s) for s in SHAPES[:5]]
    gs = [rng.randn(*w.shape).astype(F32) for w in ws]
    src = [cu(w) for w in ws]
    dst = [torch.empty_like(s) for s in src]
    q = ops.UniformWeightQuantizer(src, dst, 8, use_buckets=True, bucket_type='channel')
    q.forward()
    tg = [cu(g) for g in gs]
    q.ste_backward_(tg)
    for w, g, t in zip(ws, gs, tg):
        _, alpha, _ = O.uniform_quantize(w, 8, use_buckets=True, bucket_type='channel', return_scales=True)
        ref = O.uq_ste_grad(g.reshape(-1, w.shape[-1]), alpha, 8).reshape(w.shape)
        assert np.array_equal(t.cpu().numpy(), ref)
        np.testing.assert_allclose(t.cpu().numpy(), g, rtol=4e-7)


@pytest.mark.parametrize('n', [1, 3, 4, 1023, 4096, 100003, 256 * 32 * 32 * 16])
@pytest.mark.parametrize('bits', [8, 32, 4])
def test_activation_fake_quant_bit_exact(n, bits):
    rng = np.random.RandomState(n % 1000 + bits)
    a = np.maximum(rng.randn(n), 0).astype(F32)
    x = cu(a)
    y = ops.act_fake_quant(x, bits)
    assert np.array_equal(y.cpu().numpy(), O.uniform_quantize(a, bits, mode='activation'))
    ops.act_fake_quant(x, bits, out=x)                       # in place
    assert np.array_equal(x.cpu().numpy(), y.cpu().numpy())


def test_activation_negative_and_constant():
    for a in (np.full(1000, 3.5, F32), -np.abs(np.random.RandomState(1).randn(999)).astype(F32),
              np.zeros(64, F32)):
        y = ops.act_fake_quant(cu(a), 8)
        assert np.array_equal(y.cpu().numpy(), O.uniform_quantize(a, 8, mode='activation'))


@pytest.mark.parametrize('ratio', [0.0, 0.1, 0.5, 0.75, 0.999, 1.0])
def test_mask_build_bit_exact(ratio):
    rng = np.random.RandomState(int(ratio * 1000))
    shapes = [(3, 3, 64, 64), (1, 1, 64, 256), (2048, 10), (7,), (1,), (5, 5, 3, 32), (100003,)]
    ws = [he(rng, s) for s in shapes]
    ws[0].reshape(-1)[::7] = ws[0].reshape(-1)[3]             # ties
    ws[2][::2] = 0.0                                          # many zeros (previously pruned)
    bk = [rng.randn(*s).astype(F32) for s in shapes]
    mk = [(rng.rand(*s) > 0.4).astype(F32) for s in shapes]
    tw, tb, tm = [cu(w) for w in ws], [cu(b) for b in bk], [cu(m) for m in mk]
    mb = ops.MaskBuilder(tw, tb, tm)
    ratios = [ratio] * len(ws)
    mb.build(ratios)
    thr = mb.thr.cpu().numpy()
    for i in range(len(ws)):
        rv, rb, rm, rt = O.ws_build_mask(ws[i], bk[i], mk[i], ratio)
        assert thr[i] == rt, (i, thr[i], rt)
        assert np.array_equal(tm[i].cpu().numpy(), rm), 'mask %d' % i
        assert np.array_equal(tb[i].cpu().numpy(), rb)
        assert np.array_equal(tw[i].cpu().numpy().view(np.uint32), rv.view(np.uint32))   # even -0.0


def test_mask_build_twice_is_stable_and_schedule():
    rng = np.random.RandomState(9)
    w = he(rng, (3, 3, 32, 32))
    tw, tb, tm = cu(w), cu(w.copy()), torch.ones(w.shape, device=dev())
    mb = ops.MaskBuilder([tw], [tb], [tm])
    rw, rb, rm = w.copy(), w.copy(), np.ones_like(w)
    for step in (1500, 2500, 4000, 5000):
        r = O.ws_prune_ratio_dyn(step, 10000, 0.75)
        mb.build([r])
        rw, rb, rm, _ = O.ws_build_mask(rw, rb, rm, r)
        assert np.array_equal(tm.cpu().numpy(), rm)
    assert abs(float(1 - tm.mean()) - 0.75) < 1e-3


def test_select_desc_matches_sort():
    rng = np.random.RandomState(10)
    ts = [rng.randn(5000).astype(F32), -np.abs(rng.randn(333)).astype(F32), np.zeros(17, F32)]
    queries = [(0, 0), (0, 4999), (0, 2500), (1, 100), (2, 5), (0, 1234), (1, 0), (1, 332)]
    got = ops.select_desc([cu(t) for t in ts], queries).cpu().numpy()
    for (ti, r), g in zip(queries, got):
        assert g == np.sort(ts[ti])[::-1][r]


@pytest.mark.parametrize('n', [1, 5, 4096, 100003])
def test_momentum_and_adam_bit_exact(n):
    rng = np.random.RandomState(n)
    w, acc, g = rng.randn(n).astype(F32), (rng.randn(n) * .1).astype(F32), rng.randn(n).astype(F32)
    mask = (rng.rand(n) > 0.5).astype(F32)
    for use_mask, wd, gs in ((True, 1e-4, 0.125), (False, 0.0, 1.0)):
        tw, ta = cu(w), cu(acc)
        hp = cu(np.array([0.05, 0, 0, 0], F32))
        ops.momentum_step(tw, ta, cu(g), cu(mask) if use_mask else None, hp, 0.9, wd, gs)
        rw, ra = O.momentum_step(w, acc, g, 0.05, 0.9, mask=mask if use_mask else None, wd=wd, grad_scale=gs)
        assert np.array_equal(tw.cpu().numpy(), rw) and np.array_equal(ta.cpu().numpy(), ra)
    m0, v0 = (rng.randn(n) * .01).astype(F32), (np.abs(rng.randn(n)) * .001).astype(F32)
    b1p, b2p = F32(0.9) ** 3, F32(0.999) ** 3
    tw, tm, tv = cu(w), cu(m0), cu(v0)
    hp = cu(np.array([1e-3, b1p, b2p, 0], F32))
    ops.adam_step(tw, tm, tv, cu(g), hp, wd=2e-4, grad_scale=0.5)
    rw, rm, rv = O.adam_step(w, m0, v0, g, 1e-3, b1p, b2p, wd=2e-4, grad_scale=0.5)
    assert np.array_equal(tm.cpu().numpy(), rm) and np.array_equal(tv.cpu().numpy(), rv)
    assert np.array_equal(tw.cpu().numpy(), rw)


@pytest.mark.parametrize('n,k', [(256, 10), (256, 1001), (7, 3), (1, 5), (33, 100)])
@pytest.mark.parametrize('with_teacher', [True, False])
def test_softmax_ce_matches_oracle(n, k, with_teacher):
    rng = np.random.RandomState(n * 7 + k)
    s, t = (rng.randn(n, k) * 3).astype(F32), (rng.randn(n, k) * 3).astype(F32)
    lab = np.eye(k, dtype=F32)[rng.randint(0, k, n)]
    out, dl = ops.softmax_ce(cu(s), cu(lab), cu(t) if with_teacher else None, 4.0, 4.0)
    o = out.cpu().numpy()
    lh, gh = O.softmax_cross_entropy(lab, s)
    ref_g = gh
    assert abs(o[0] - lh) <= 1e-5 * abs(lh)                    # tolerance: 1e-5 relative (north star)
    if with_teacher:
        ld, gd = O.distillation_loss(s, t, 4.0, 4.0)
        assert abs(o[1] - ld) <= 1e-5 * abs(ld)
        ref_g = (gh + gd).astype(F32)
    else:
        assert o[1] == 0.0
    np.testing.assert_allclose(dl.cpu().numpy(), ref_g, rtol=1e-5, atol=1e-8)
    assert o[2] == O.accuracy(lab, s)
    top5 = np.mean([(np.sum(s[i] > s[i, np.argmax(lab[i])]) < 5) for i in range(n)])
    assert abs(o[3] - top5) < 1e-6


def test_softmax_ce_lenet_probabilities_as_logits():
    # LeNet feeds softmax OUTPUTS into the CE (nets/lenet_at_cifar10.py:66): any input is legal
    rng = np.random.RandomState(3)
    p = O.softmax((rng.randn(32, 10) * 2).astype(F32))
    lab = np.eye(10, dtype=F32)[rng.randint(0, 10, 32)]
    out, dl = ops.softmax_ce(cu(p), cu(lab))
    lh, gh = O.softmax_cross_entropy(lab, p)
    assert abs(out.cpu().numpy()[0] - lh) <= 1e-5 * lh
    np.testing.assert_allclose(dl.cpu().numpy(), gh, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize('n', [1, 1001, 25_000_000])
def test_l2_loss(n):
    rng = np.random.RandomState(4)
    v = (rng.randn(n) * 0.05).astype(F32)
    out = torch.zeros(4, device=dev())
    pw = torch.empty(ops.L2_PARTIALS, device=dev())
    ops.l2_loss(cu(v), 1e-4, out, pw)
    ref = 1e-4 * np.sum(v.astype(np.float64) ** 2) / 2
    assert abs(out.cpu().numpy()[0] - ref) <= 2e-6 * ref
    ops.l2_loss(cu(v), 1e-4, out, pw, accumulate=True)
    assert abs(out.cpu().numpy()[0] - 2 * ref) <= 2e-6 * 2 * ref


@pytest.mark.parametrize('bits', [1, 2, 4])
def test_codebook_quant_bit_exact(bits):
    rng = np.random.RandomState(bits)
    ws = [he(rng, s) for s in [(3, 3, 16, 16), (1, 1, 64, 64), (3, 3, 5, 7), (300,)]]
    src = [cu(w) for w in ws]
    dst = [torch.empty_like(s) for s in src]
    q = ops.CodebookWeightQuantizer(src, dst, bits, keep_index=True)
    q.quantile_init()
    q.forward()
    c = q.clusters.cpu().numpy()
    idx = q.idx.cpu().numpy()
    for i, w in enumerate(ws):
        rq, rc, ridx = O.nonuniform_quantize(w, bits)
        assert np.array_equal(c[i, :1 << bits], rc), 'codebook %d' % i
        assert np.array_equal(dst[i].cpu().numpy(), rq)
        o = q.idx_offsets[i]
        assert np.array_equal(idx[o:o + w.size], ridx.reshape(-1).astype(np.uint8))


@pytest.mark.parametrize('bits', [2, 4, 8])
def test_codebook_gradient_is_the_segment_sum(bits):
    """pf_nuq_cluster_grad (+ pf_nuq_weight_quant_ex, codebooks inside one flat parameter buffer): dL/dc_j =
    alpha * sum_{idx=j} g against the oracle's nuq_grads (learners/nonuniform_quantization/utils.py:303-306, :433)."""
    rng = np.random.RandomState(10 + bits)
    shapes = [(3, 3, 16, 16), (1, 1, 64, 64), (3, 3, 5, 7), (300,), (3, 3, 128, 130)]
    ws = [he(rng, s) for s in shapes]
    src = [cu(w) for w in ws]
    dst = [torch.empty_like(s) for s in src]
    k = 1 << bits
    base = torch.zeros(16 + len(ws) * (k + 4), dtype=torch.float32, device='cuda:0')     # codebooks at odd offsets
    views = [base[16 + i * (k + 4):16 + i * (k + 4) + k] for i in range(len(ws))]
    q = ops.CodebookWeightQuantizer(src, dst, bits, keep_index=True, cluster_views=views, cluster_base=base)
    q.quantile_init()
    q.forward()
    gs = [rng.randn(*s).astype(np.float32) for s in shapes]
    gdev = [cu(g) for g in gs]
    gbase = torch.full_like(base, 7.0)
    q.cluster_grad(gdev, gbase)
    gb = gbase.cpu().numpy()
    rngs = q.uq.ranges()
    for i, (w, g) in enumerate(zip(ws, gs)):
        rq, rc, ridx = O.nonuniform_quantize(w, bits)
        assert np.array_equal(views[i].cpu().numpy(), rc) and np.array_equal(dst[i].cpu().numpy(), rq)
        alpha = np.float32(np.float32(rngs[i][1][0] - rngs[i][0][0]) + np.float32(1e-10))
        _, gc = O.nuq_grads(g, ridx, k, alpha)
        got = gb[16 + i * (k + 4):16 + i * (k + 4) + k]
        assert np.abs(got - gc).max() <= 2e-6 * max(np.abs(gc).max(), 1e-20) * np.sqrt(w.size), i
    keep = np.ones(gb.size, bool)
    for i in range(len(ws)):
        keep[16 + i * (k + 4):16 + i * (k + 4) + k] = False
    assert np.all(gb[keep] == 7.0)                                       # nothing written outside the codebooks
    q.cluster_grad(gdev, gbase)
    assert np.array_equal(gbase.cpu().numpy(), gb)                       # deterministic


def test_against_committed_golden_vectors():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'hotpath_v1.npz'))
    for mode, kw in (('layer', dict()), ('channel', dict(use_buckets=True, bucket_type='channel')),
                     ('split', dict(use_buckets=True, bucket_type='split', bucket_size=16))):
        for bits in (2, 4, 8):
            src = [cu(g['w%d' % i]) for i in range(6)]
            dst = [torch.empty_like(s) for s in src]
            ops.UniformWeightQuantizer(src, dst, bits, **kw).forward()
            for i in range(6):
                assert np.array_equal(dst[i].cpu().numpy(), g['w%d_%s_b%d' % (i, mode, bits)])
    for bits in (8, 32):
        assert np.array_equal(ops.act_fake_quant(cu(g['act']), bits).cpu().numpy(), g['act_b%d' % bits])
    for r in (0.0, 0.3, 0.5, 0.9):
        tw, tb, tm = cu(g['ws_w']), cu(g['ws_bkup']), cu(g['ws_mask'])
        mb = ops.MaskBuilder([tw], [tb], [tm])
        mb.build([r])
        tag = 'ws_r%02d' % int(r * 100)
        assert np.array_equal(tm.cpu().numpy(), g[tag + '_mask'])
        assert np.array_equal(tw.cpu().numpy(), g[tag + '_w'])
        assert np.array_equal(tb.cpu().numpy(), g[tag + '_bkup'])
        assert mb.thr.cpu().numpy()[0] == g[tag + '_thr'][0]
    tw, ta = cu(g['opt_w']), cu(g['opt_acc'])
    ops.momentum_step(tw, ta, cu(g['opt_g']), cu(g['opt_mask']), cu(np.array([0.05, 0, 0, 0], F32)), 0.9, 1e-4, 0.5)
    assert np.array_equal(tw.cpu().numpy(), g['mom_w']) and np.array_equal(ta.cpu().numpy(), g['mom_acc'])
    tw, tm, tv = cu(g['opt_w']), cu(g['adam_m0']), cu(g['adam_v0'])
    hp = cu(np.array([1e-3, F32(0.9) * F32(0.9), F32(0.999) * F32(0.999), 0], F32))
    ops.adam_step(tw, tm, tv, cu(g['opt_g']), hp, wd=2e-4)
    assert np.array_equal(tw.cpu().numpy(), g['adam_w']) and np.array_equal(tv.cpu().numpy(), g['adam_v'])
    out, dl = ops.softmax_ce(cu(g['ce_s']), cu(g['ce_lab']), cu(g['ce_t']))
    o = out.cpu().numpy()
    assert abs(o[0] - g['ce_hard'][0]) <= 1e-5 * g['ce_hard'][0]
    assert abs(o[1] - g['ce_dst'][0]) <= 1e-5 * g['ce_dst'][0]
    np.testing.assert_allclose(dl.cpu().numpy(), g['ce_grad'], rtol=1e-5, atol=1e-8)
    src, dst = [cu(g['nuq_w'])], [torch.empty(g['nuq_w'].shape, device=dev())]
    q = ops.CodebookWeightQuantizer(src, dst, 4)
    q.quantile_init()
    q.forward()
    assert np.array_equal(dst[0].cpu().numpy(), g['nuq_q'])


# ------------------------------------------------------------------ full-size property tests
RESNET50_3x3 = [(3, 3, 64, 64), (3, 3, 128, 128), (3, 3, 256, 256), (3, 3, 512, 512)]


def test_fullsize_weight_quant_properties():
    """ResNet-50-sized tensors (too big for the per-element oracle loop to be the only check):
    <= 2^b levels per bucket, range preserved, and Q is idempotent up to 1 ulp of re-normalising."""
    torch.manual_seed(0)
    src = [torch.randn(s, device=dev()) * (2.0 / (9 * s[2])) ** 0.5 for s in RESNET50_3x3] + \
          [torch.randn(2048, 1001, device=dev()) * 0.03]
    dst = [torch.empty_like(s) for s in src]
    q = ops.UniformWeightQuantizer(src, dst, 4, use_buckets=True, bucket_type='channel')
    q.forward()
    for s, d in zip(src, dst):
        cout = s.shape[-1]
        s2, d2 = s.reshape(-1, cout), d.reshape(-1, cout)
        assert torch.equal(d2.min(0).values, s2.min(0).values)
        assert torch.all(d2.max(0).values <= s2.max(0).values + 1e-6)
        for c in (0, cout // 2, cout - 1):
            assert torch.unique(d2[:, c]).numel() <= 16
    # the biggest one against the oracle too (seconds)
    ref = O.uniform_quantize(src[3].cpu().numpy(), 4, use_buckets=True, bucket_type='channel')
    assert np.array_equal(dst[3].cpu().numpy(), ref)


def test_fullsize_activation_quant_property():
    n = 256 * 56 * 56 * 64                                       # a ResNet-50 stage-1 ReLU output @B=256
    x = torch.relu(torch.randn(n, device=dev()))
    y = ops.act_fake_quant(x, 8)
    assert torch.unique(y).numel() <= 256
    assert float(y.max()) == float(x.max()) and float(y.min()) == 0.0
    step = float(x.max()) / 255
    assert float((y - x).abs().max()) <= step / 2 * 1.0001
    sl = slice(12345678, 12345678 + 1_000_000)
    ref = O.uq_inv_scale((np.rint((((x[sl].cpu().numpy() - F32(0)) / F32(x.max().item() + 1e-10)) * F32(255)
                                   ).astype(F32)) / F32(255)).astype(F32), F32(x.max().item() + 1e-10), F32(0))
    assert np.array_equal(y[sl].cpu().numpy(), ref)


def test_fullsize_mask_density_property():
    torch.manual_seed(1)
    shapes = RESNET50_3x3 + [(1, 1, 1024, 2048), (2048, 1001)]
    ws = [torch.randn(s, device=dev()) for s in shapes]
    bk = [w.clone() for w in ws]
    mk = [torch.ones_like(w) for w in ws]
    mb = ops.MaskBuilder(ws, bk, mk)
    ranks = mb.build([0.5] * len(ws))
    for w, b, m, r, t in zip(ws, bk, mk, ranks, mb.thr.cpu().numpy()):
        kept = int(m.sum().item())
        assert kept == int((b.abs() > float(t)).sum().item())
        assert kept <= r and kept >= r - 2                      # continuous data: no ties expected
        assert int((w != 0).sum().item()) == kept
        srt = torch.sort(b.abs().reshape(-1), descending=True).values
        assert float(srt[r]) == float(t)                        # exact order statistic
