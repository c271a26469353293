"""Known-answer tests that pin the oracle.  The reference ships
no golden vectors for this path ("parity unpinned"), so every case here is
hand-derivable from the reference's op chain."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import pf_oracle as O

F32 = np.float32


def test_round_half_even_discriminator_1bit():
    # w=[0,.5,1], 1 bit: k=1, xn=[0,.5,1]; rint(.5)=0 (half-away would give 1)
    q = O.uniform_quantize(np.array([0, 0.5, 1], F32), 1)
    assert q.tolist() == [0.0, 0.0, 1.0]


def test_2bit_levels():
    # k=3: xn*3 = [0,.75,1.5,3] -> rint [0,1,2,3] -> /3
    q = O.uniform_quantize(np.array([0, 0.25, 0.5, 1], F32), 2)
    exp = (np.array([0, 1, 2, 3], F32) / F32(3)).astype(F32)
    np.testing.assert_array_equal(q, exp)


def test_constant_tensor_identity():
    w = np.full((3, 3, 2, 4), 0.37, F32)
    q, alpha, beta = O.uniform_quantize(w, 8, return_scales=True)
    assert alpha == F32(1e-10) and beta == F32(0.37)
    np.testing.assert_array_equal(q, w)


def test_k_values():
    assert O.uq_k(8) == F32(255.0)
    assert O.uq_k(4) == F32(15.0)
    assert O.uq_k(32) == F32(4294967296.0)


def test_bucketing_channel_layer_split():
    # 2x2x2x3 kernel (24 elems), distinct per-channel ranges
    w = np.zeros((2, 2, 2, 3), F32)
    rng = np.random.RandomState(0)
    for c, (lo, hi) in enumerate([(-1, 1), (0, 10), (-100, -50)]):
        w[..., c] = rng.uniform(lo, hi, size=(2, 2, 2))
    ql = O.uniform_quantize(w, 2)                               # layer: one range
    qc = O.uniform_quantize(w, 2, use_buckets=True, bucket_type='channel')
    for c in range(3):
        np.testing.assert_array_equal(qc[..., c], O.uniform_quantize(w[..., c], 2))
        assert len(np.unique(qc[..., c])) <= 4
    assert len(np.unique(ql)) <= 4
    # split with bucket_size 8: multiple=3 -> bucket j = flat[j::3]  (strided!)
    qs = O.uniform_quantize(w, 2, use_buckets=True, bucket_type='split', bucket_size=8)
    flat = w.reshape(-1)
    for j in range(3):
        np.testing.assert_array_equal(qs.reshape(-1)[j::3], O.uniform_quantize(flat[j::3], 2))


def test_split_bucket_padding_uses_last_element():
    w = np.arange(10, dtype=F32)          # bucket_size 4 -> pad 2 copies of 9, multiple=3
    xb, multiple, padded = O.split_bucket(w, 4)
    assert multiple == 3 and padded == 2 and xb.shape == (4, 3)
    assert xb[3].tolist() == [9.0, 9.0, 9.0]
    q = O.uniform_quantize(w, 8, use_buckets=True, bucket_type='split', bucket_size=4)
    assert q.shape == w.shape
    # bucket 1 = {1,4,7,9(pad)} -> max 9, not 7
    col = np.array([1, 4, 7, 9], F32)
    np.testing.assert_array_equal(q[[1, 4, 7]], O.uniform_quantize(col, 8)[:3])


def test_percentile_nearest_rule():
    # n=10 distinct, q=50 -> idx=rint(9*0.5)=rint(4.5)=4 (half-even) of the DESCENDING sort
    x = np.arange(10, dtype=F32)
    assert O.percentile_index(10, 50.0) == 4
    thr = O.percentile_nearest(x, 50.0)
    assert thr == 5.0
    assert int(np.sum(x > thr)) == 4        # 4 kept / 6 pruned
    assert O.percentile_index(10, 0.0) == 9 and O.percentile_index(10, 100.0) == 0


def test_mask_ties_pruned():
    w = np.array([3, -3, 3, 1, 2, -5, 4, 0.5], F32)
    var, bkup, mask, thr = O.ws_build_mask(w, np.zeros_like(w), np.ones_like(w), 0.5)
    # n=8, q=50: idx=rint(3.5)=4 ; |w| desc = [5,4,3,3,3,2,1,.5] -> thr=3 ; ties pruned
    assert thr == 3.0
    assert mask.tolist() == [0, 0, 0, 0, 0, 1, 1, 0]
    np.testing.assert_array_equal(var, w * mask)


def test_mask_bkup_semantics():
    w = np.array([0.0, 2.0, 0.0, 4.0], F32)       # currently pruned at 0 and 2
    bkup = np.array([1.5, 9.0, 0.125, 9.0], F32)
    mask = np.array([0, 1, 0, 1], F32)
    var, nb, nm, thr = O.ws_build_mask(w, bkup, mask, 0.25)
    assert nb.tolist() == [1.5, 2.0, 0.125, 4.0]   # live weights refresh the backup
    # n=4,q=25: idx=rint(3*.75)=rint(2.25)=2 ; desc [4,2,1.5,.125] -> thr 1.5
    assert thr == 1.5 and nm.tolist() == [0, 1, 0, 1]


def test_prune_ratio_schedule():
    nb = 1000
    assert O.ws_prune_ratio_dyn(100, nb, 0.5) == 0.0
    assert O.ws_prune_ratio_dyn(500, nb, 0.5) == F32(0.5)
    assert O.ws_prune_ratio_dyn(900, nb, 0.5) == F32(0.5)
    mid = O.ws_prune_ratio_dyn(300, nb, 0.5)
    assert mid == F32(F32(0.5) * F32(F32(1) - F32(np.power(F32(0.5), F32(3.0)))))


def test_distillation_closed_form_k2():
    s = np.array([[1.0, -1.0]], F32)
    t = np.array([[0.5, 0.25]], F32)
    T, w = 4.0, 4.0
    p = 1 / (1 + np.exp(-(0.5 - 0.25) / T))
    ls = -np.log(1 / (1 + np.exp(-(1.0 - -1.0) / T)))
    ls2 = -np.log(1 / (1 + np.exp((1.0 - -1.0) / T)))
    exp = w * (p * ls + (1 - p) * ls2)
    loss, g = O.distillation_loss(s, t, w, T)
    assert abs(loss - exp) < 1e-6
    ps = 1 / (1 + np.exp(-2.0 / T))
    np.testing.assert_allclose(g[0], [w / T * (ps - p), -w / T * (ps - p)], atol=1e-6)


def test_distillation_reduces_to_hard_ce():
    rng = np.random.RandomState(1)
    s = rng.randn(5, 7).astype(F32)
    lab = np.eye(7, dtype=F32)[rng.randint(0, 7, 5)]
    t = (lab * 200 - 100).astype(F32)                 # softmax(t) == one-hot in fp32
    l1, g1 = O.distillation_loss(s, t, 1.0, 1.0)
    l2, g2 = O.softmax_cross_entropy(lab, s)
    assert l1 == l2
    np.testing.assert_array_equal(g1, g2)


def test_nuq_1bit_quantiles():
    x = np.linspace(0, 1, 7).astype(F32)      # already normalised, sorted
    c = O.nuq_quantile_init(x, 2)
    # q=33.33: idx=rint(6*(1-1/3))=4 of desc -> x[2]; q=66.67: idx=rint(6/3)=2 -> x[4]
    assert c.tolist() == [x[2], x[4]]
    q, idx = O.nuq_assign(x, c)
    assert idx.tolist() == [0, 0, 0, 0, 1, 1, 1]   # x[3] equidistant -> FIRST index


def test_adam_first_step_closed_form():
    g = np.array([0.3, -2.0, 1e-3], F32)
    w0 = np.array([1.0, 2.0, 3.0], F32)
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    w, m, v = O.adam_step(w0, np.zeros(3, F32), np.zeros(3, F32), g, lr, F32(b1), F32(b2))
    exp = -lr * np.sqrt(1 - b2) / (1 - b1) * (1 - b1) * g.astype(np.float64) / \
        (np.sqrt((1 - b2) * g.astype(np.float64) ** 2) + eps)
    np.testing.assert_allclose((w - w0).astype(np.float64), exp, rtol=2e-4)
    np.testing.assert_allclose(m, 0.1 * g, rtol=1e-6)


def test_momentum_masked():
    w = np.array([1.0, 2.0], F32)
    w1, a1 = O.momentum_step(w, np.array([0.5, 0.5], F32), np.array([1.0, 1.0], F32), 0.1, 0.9,
                             mask=np.array([1.0, 0.0], F32))
    np.testing.assert_allclose(a1, [1.45, 0.45], rtol=1e-6)
    np.testing.assert_allclose(w1, [1 - 0.145, 2 - 0.045], rtol=1e-6)


def test_uq_schedule():
    lr, bnds, rates, steps = O.uq_bnds_decay_rates('resnet_20', 'cifar_10', 50000, 256, 1, 0.1, 128)
    assert steps == 195 * 60 == 11700 and bnds == [195 * 15, 195 * 40] and lr == 0.1
    assert O.piecewise_constant(0, bnds, [1, 2, 3]) == 1
    assert O.piecewise_constant(bnds[0], bnds, [1, 2, 3]) == 1
    assert O.piecewise_constant(bnds[0] + 1, bnds, [1, 2, 3]) == 2
    assert O.piecewise_constant(10 ** 9, bnds, [1, 2, 3]) == 3


def test_heurist_ratios_weighted_mean():
    n = [1000, 50000, 2000000]
    r = O.ws_heurist_ratios(n, 0.5)
    assert abs(np.sum(r * np.array(n)) / np.sum(n) - 0.5) < 1e-12


# ---------------------------------------------------------------- properties
@settings(max_examples=60, deadline=None)
@given(st.integers(1, 8), st.integers(2, 300), st.integers(0, 2 ** 31 - 1))
def test_prop_levels_and_idempotence(bits, n, seed):
    rng = np.random.RandomState(seed)
    w = rng.randn(n).astype(F32)
    q = O.uniform_quantize(w, bits)
    assert len(np.unique(q)) <= 2 ** bits
    assert q.min() >= w.min() - 1e-6 and q.max() <= w.max() + 1e-6
    # Q(Q(w)) == Q(w) up to the 1-ulp wobble of re-normalising the levels
    q2 = O.uniform_quantize(q, bits)
    np.testing.assert_allclose(q2, q, rtol=0, atol=4e-7 * max(1.0, float(np.abs(w).max())))


@settings(max_examples=60, deadline=None)
@given(st.integers(2, 400), st.floats(0.0, 0.99), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_prop_mask_density(n, ratio, seed, ties):
    rng = np.random.RandomState(seed)
    w = rng.randn(n).astype(F32)
    if ties:
        w = np.round(w * 2).astype(F32) / 2
    var, bkup, mask, thr = O.ws_build_mask(w, w.copy(), np.ones_like(w), ratio)
    idx = O.ws_mask_rank(n, ratio)
    a = np.sort(np.abs(w))[::-1]
    assert thr == a[idx]
    assert int(mask.sum()) == int(np.sum(a > a[idx]))
    assert int(mask.sum()) <= idx            # kept = strictly-greater count <= idx


def test_ste_grad_is_near_identity():
    g = np.random.RandomState(3).randn(1000).astype(F32)
    out = O.uq_ste_grad(g, F32(0.731), 8)
    np.testing.assert_allclose(out, g, rtol=3e-7)
