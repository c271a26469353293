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


# ---------------------------------------------------------------------------------------------------------------
# Values produced by the REFERENCE'S OWN host-side functions (tests/golden/make_golden_from_reference.py runs them from
# /root/reference under a stub tensorflow module): schedules and the 'heurist' pruning-ratio formula.
def _ref_gold():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_executed_v1.json')))


@pytest.mark.parametrize('which', ['uq', 'nuq'])
def test_finetune_schedules_match_the_reference_functions(which, monkeypatch):
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
    FLAGS.reset()
    import importlib
    import pocketflow_b200.datasets.cifar10_dataset  # noqa: F401  (dataset flags)
    mod = importlib.import_module('pocketflow_b200.learners.%s.learner' % ('uniform_quantization' if which == 'uq' else 'nonuniform_quantization'))
    gold = _ref_gold()['%s_setup_bnds_decay_rates' % which]
    assert len(gold) == 36
    for g in gold:
        FLAGS.nb_smpls_train, FLAGS.batch_size = g['nb_smpls_train'], g['batch_size']
        FLAGS.enbl_multi_gpu, FLAGS.enbl_warm_start = g['enbl_multi_gpu'], g['enbl_warm_start']
        FLAGS.lrn_rate_init, FLAGS.batch_size_norm = 1e-1, 128.0
        setattr(FLAGS, 'uql_quant_epochs' if which == 'uq' else 'nuql_quant_epochs', 60)
        monkeypatch.setattr(mgw, 'size', classmethod(lambda cls, w=g['world']: w))
        init_lr, bnds, rates, steps = mod.setup_bnds_decay_rates(g['model'], g['dataset'])
        assert float(init_lr) == g['init_lr'] and [int(b) for b in bnds] == g['bnds'], g
        assert [float(r) for r in rates] == g['decay_rates'] and int(steps) == g['finetune_steps'], g
    FLAGS.reset()


def test_piecewise_constant_lr_matches_the_reference_function():
    from pocketflow_b200.flags import FLAGS
    FLAGS.reset()
    import pocketflow_b200.datasets.cifar10_dataset  # noqa: F401
    from pocketflow_b200.utils import lrn_rate_utils as L
    for g in _ref_gold()['lrn_rate_piecewise_constant']:
        FLAGS.nb_smpls_train, FLAGS.nb_epochs_rat = g['nb_smpls_train'], g['nb_epochs_rat']
        FLAGS.lrn_rate_init, FLAGS.batch_size_norm = 1e-1, 128.0
        fn = L.setup_lrn_rate_piecewise_constant(None, g['batch_size'], g['idxs_epoch'], g['decay_rates'])
        bnds, vals = g['bnds'], g['vals']
        # tf.train.piecewise_constant: vals[0] for step <= bnds[0]; vals[i] for bnds[i-1] < step <= bnds[i]; vals[-1] after
        for i, b in enumerate(bnds):
            assert fn(b) == vals[i] and fn(b + 1) == vals[i + 1], (g, i)
        assert fn(0) == vals[0] and fn(bnds[-1] * 10) == vals[-1]
    FLAGS.reset()


def test_heurist_prune_ratios_match_the_reference_function():
    for g in _ref_gold()['ws_heurist_prune_ratios']:
        n = [int(np.prod(s)) for s in g['shapes']]
        got = O.ws_heurist_ratios(n, g['ws_prune_ratio'])
        assert np.array_equal(np.asarray(got, np.float64), np.asarray(g['ratios'], np.float64)), g


def test_quantized_op_selection_matches_the_reference_functions():
    """search_matmul_op / search_activation_op of the reference (run under the stub on the (type, name) lists of this
    repo's graphs) select exactly the ops this repo's UniformQuantization selects — teacher ops excluded, first and
    last matmul kept at full precision unless quantize_all_layers."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from graphs_for_golden import GRAPHS, build_graph
    from pocketflow_b200.learners.uniform_quantization.utils import UniformQuantization
    gold = _ref_gold()['uq_op_selection']
    assert len(gold) == 2 * len(GRAPHS)
    for g in gold:
        net, flags, dst = GRAPHS[g['graph']]
        graph = build_graph(net, flags, dst)
        assert len(graph.ops) == g['n_ops'], 'graph changed: regenerate tests/golden/ref_executed_v1.json'
        uq = UniformQuantization(graph, 256, True, 'channel')
        assert [o.name for o in uq.search_matmul_op(g['quantize_all_layers'])] == g['matmul'], g['graph']
        assert [o.name for o in uq.search_activation_op()] == g['activation'], g['graph']
    from pocketflow_b200.flags import FLAGS
    FLAGS.reset()


def test_dynamic_prune_ratio_matches_the_reference_function():
    """WeightSparseLearner.__calc_prune_ratio_dyn executed from the reference source (tf scalar ops mapped onto numpy
    float32) vs the oracle restatement: bit-identical float32 at the schedule's corner steps."""
    gold = _ref_gold()['ws_prune_ratio_dyn']
    assert len(gold) >= 30
    for g in gold:
        got = O.ws_prune_ratio_dyn(g['global_step'], g['nb_iters_train'], g['prune_ratio_fnl'])
        assert np.float32(got).tobytes().hex() == g['value_f32_hex'], (g, float(got))


def test_product_dynamic_prune_ratio_matches_the_reference_function():
    """The learner's own host-side schedule (not the oracle) against the same reference-generated values."""
    from pocketflow_b200.flags import FLAGS
    FLAGS.reset()
    import importlib
    W = importlib.import_module('pocketflow_b200.learners.weight_sparsification.learner')
    fn = getattr(W.WeightSparseLearner, '_WeightSparseLearner__calc_prune_ratio_dyn')
    import types
    for g in _ref_gold()['ws_prune_ratio_dyn']:
        fake = types.SimpleNamespace(nb_iters_train=g['nb_iters_train'])
        got = fn(fake, g['prune_ratio_fnl'], g['global_step'])
        assert np.float32(got).tobytes().hex() == g['value_f32_hex'], g
    FLAGS.reset()


def test_uniform_quantize_matches_the_reference_function_structure():
    """UniformQuantization.__uniform_quantize (with __scale / __inv_scale / __split_bucket / __channel_bucket) executed
    FROM THE REFERENCE SOURCE on numpy-backed tensors (every tf op mapped one-to-one onto the numpy float32 op; tf.round =
    np.rint) vs the oracle restatement: bit-identical on 120 cases — per-layer / per-channel / strided split buckets
    with last-element padding, activations, 1..32 bits, constant tensors."""
    import hashlib
    gold = _ref_gold()['uniform_quantize']
    assert len(gold) == 120
    for g in gold:
        rng = np.random.default_rng(g['seed'])
        x = (rng.standard_normal(tuple(g['shape'])) * rng.choice([1e-3, 1.0, 37.0])).astype(np.float32)
        if g['constant']:
            x[...] = x.flat[0]
        q = O.uniform_quantize(x, g['bits'], mode=g['mode'], use_buckets=g['use_buckets'], bucket_type=g['bucket_type'],
                               bucket_size=256)
        q = np.ascontiguousarray(q, np.float32)
        assert hashlib.sha256(q.tobytes()).hexdigest() == g['sha256'], (g, q.reshape(-1)[:3])


def test_mask_build_matches_the_reference_function_structure():
    """WeightSparseLearner.__build_masks executed FROM THE REFERENCE SOURCE on numpy-backed variables (percentile = the
    oracle's 'nearest' rule, a documented recollection) over several mask updates with simulated training in between vs
    the oracle's ws_build_mask + ws_prune_ratio_dyn: var / backup / mask bit-identical after every update."""
    import hashlib
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()   # noqa: E731
    for rec in _ref_gold()['ws_build_masks']:
        shape = tuple(rec['shape'])
        var = np.random.default_rng(rec['seed']).standard_normal(shape).astype(np.float32)
        var.reshape(-1)[1] = var.reshape(-1)[0]
        bkup, mask = var.copy(), np.ones(shape, np.float32)       # get_variable initializers: var's value, ones
        for u in rec['updates']:
            ratio = O.ws_prune_ratio_dyn(u['global_step'], rec['nb_iters_train'], rec['prune_ratio_fnl'])
            var, bkup, mask, _ = O.ws_build_mask(var, bkup, mask, ratio)
            assert (h(var), h(bkup), h(mask), int(mask.sum())) == (u['var'], u['bkup'], u['mask'], u['kept']), (rec['shape'], u)
            nz = np.random.default_rng(u['noise_seed']).standard_normal(shape).astype(np.float32) * np.float32(0.05)
            var = (var + nz * mask).astype(np.float32)


def test_nonuniform_quantize_matches_the_reference_function_structure():
    """NonUniformQuantization.__nonuni_quantize (scale, quantile init, nearest centroid * sign, inverse scale) executed
    FROM THE REFERENCE SOURCE on numpy tensors vs the oracle restatement: bit-identical."""
    import hashlib
    gold = _ref_gold()['nonuniform_quantize']
    assert len(gold) == 12
    for g in gold:
        rng = np.random.default_rng(g['seed'])
        x = (rng.standard_normal(tuple(g['shape'])) * rng.choice([1e-2, 1.0, 9.0])).astype(np.float32)
        q, _, _ = O.nonuniform_quantize(x, g['bits'])
        q = np.ascontiguousarray(q, np.float32)
        assert hashlib.sha256(q.tobytes()).hexdigest() == g['sha256'] and len(np.unique(q)) == g['distinct'], g


def test_distillation_loss_wiring_matches_the_reference_function():
    for g in _ref_gold()['distillation_loss']:
        rng = np.random.default_rng(g['seed'])
        s_ = (rng.standard_normal((g['n'], g['k'])) * 3).astype(np.float32)
        t_ = (rng.standard_normal((g['n'], g['k'])) * 5).astype(np.float32)
        loss, _ = O.distillation_loss(s_, t_, g['loss_w_dst'], g['tempr_dst'])
        assert np.float32(loss).tobytes().hex() == g['value_f32_hex'], (g, float(loss))


def test_l2_regularised_variables_match_the_reference_calc_loss():
    """ModelHelper.calc_loss of the reference (run under the stub on this repo's trainable-variable names): the set of
    variables that receive the L2 term (name filter; MobileNet's slim BatchNorm parameters ARE regularised, SURVEY
    A.6-9) and the default loss_w_dcy equal what this repo's ModelHelpers put into their LossSpec."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from graphs_for_golden import GRAPHS
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    for g in _ref_gold()['calc_loss_l2']:
        net, flags, _ = GRAPHS[g['graph']]
        FLAGS.reset()
        mod = importlib.reload(importlib.import_module('pocketflow_b200.nets.' + net))   # re-DEFINE this net's defaults
        for k, v in flags.items():
            setattr(FLAGS, k, v)
        mh = mod.ModelHelper()
        gr = G.Graph()
        with gr.as_default():
            with G.variable_scope('data'):
                im, lab = mh.build_dataset_train().get_next()
            with G.variable_scope('model'):
                out = mh.forward_train(im)
                tv = [v for v in gr.variables.values() if v.name.startswith('model/') and v.trainable]
                loss, _ = mh.calc_loss(lab, out, tv)
        assert len(tv) == g['n_trainable']
        mine = [(v.name, float(wd)) for v, wd in loss.l2.items()]
        assert sorted(n for n, _ in mine) == sorted(g['regularised']), g['graph']
        assert all(abs(wd - g['loss_w_dcy']) <= 1e-12 for _, wd in mine), (g['graph'], g['loss_w_dcy'])
    FLAGS.reset()


def _canonical_layers(graph_ops):
    """This repo's op list in the vocabulary of the recorded reference architecture."""
    out = []
    for op in graph_ops:
        t = op.type
        if t == 'Conv2D':
            (kh, kw), (sh, sw), (pt, pl) = op.attrs['ksize'], op.attrs['strides'], op.attrs['pad']
            out.append(('conv', op.output.shape[-1], kh, sh, pt, pl, op.output.shape[1], 'bias' in op.vars))
        elif t == 'FusedBatchNorm':
            out.append(('bn', float(op.attrs['momentum']), float(op.attrs['epsilon']), bool(op.attrs['training'])))
        elif t == 'Relu':
            out.append(('relu',))
        elif t == 'MaxPool':
            (kh, kw), (sh, sw), (pt, pl) = op.attrs['ksize'], op.attrs['strides'], op.attrs['pad']
            out.append(('maxpool', kh, sh, pt, pl, op.output.shape[1]))
        elif t == 'Add':
            out.append(('add',))
        elif t == 'Mean':
            out.append(('mean',))
        elif t == 'MatMul':
            out.append(('dense', op.output.shape[-1]))
    return out


def _same_pad(h, k, s):
    total = max((-(-h // s) - 1) * s + k - h, 0)
    return total // 2, -(-h // s)


@pytest.mark.parametrize('idx', range(8))
def test_resnet_architecture_matches_the_reference_source(idx):
    """The layer sequence the reference's resnet_model.py + forward_fn emit (recorded by executing them with symbolic
    tensors) vs the op list of this repo's graphs: every conv's filters / kernel / stride / effective padding / output
    size, BN momentum-epsilon-mode, ReLU, pooling, residual adds, mean, dense — same order, same parameters."""
    import importlib
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    g = _ref_gold()['resnet_architecture'][idx]
    FLAGS.reset()
    mod = importlib.reload(importlib.import_module('pocketflow_b200.nets.' + g['net']))
    FLAGS.resnet_size, FLAGS.nb_classes, FLAGS.batch_size = g['resnet_size'], g['nb_classes'], 2
    mh = mod.ModelHelper()
    gr = G.Graph()
    with gr.as_default():
        with G.variable_scope('data'):
            im, _ = mh.build_dataset_train().get_next()
        with G.variable_scope('model'):
            (mh.forward_train if g['is_train'] else mh.forward_eval)(im)
    mine = _canonical_layers(gr.ops)
    h = im.shape[1]
    want = []
    for r in g['layers']:
        if r[0] == 'conv':
            _, f, k, s, padding, pad, bias = r
            if padding == 'SAME':
                pt, oh = _same_pad(h, k, s)
                pl = pt
            else:
                pt, pl = pad[0], pad[2]
                oh = (h + pad[0] + pad[1] - k) // s + 1
            want.append(('conv', f, k, s, pt, pl, oh, bias))
            h_next = oh
        elif r[0] == 'bn':
            want.append(('bn', r[1], r[2], r[3]))
            h_next = h
        elif r[0] == 'maxpool':
            _, k, s, padding = r
            pt, oh = _same_pad(h, k, s) if padding == 'SAME' else (0, (h - k) // s + 1)
            want.append(('maxpool', k, s, pt, pt, oh))
            h_next = oh
        elif r[0] == 'dense':
            want.append(('dense', r[1]))
            h_next = h
        else:
            want.append((r[0],))
            h_next = h
        # shortcut convs run on the block INPUT: track the spatial size per op from this repo's own tensors instead
        h = h_next
    # spatial sizes of the reference records are re-derived per layer from this repo's graph (projection shortcuts branch
    # off the block input), so compare everything except the derived sizes first, then the sizes this repo produces
    strip = lambda L: [tuple(x[:4]) if x[0] == 'conv' else (tuple(x[:3]) if x[0] == 'maxpool' else x) for x in L]   # noqa: E731
    assert strip(mine) == strip(want), g['net']
    # effective padding and output size of every conv / pool: the reference's rule (explicit fixed padding + VALID, or SAME)
    # applied to the input size THIS repo's op actually sees
    my_ops = [op for op in gr.ops if op.type in ('Conv2D', 'MaxPool')]
    ref_ops = [r for r in g['layers'] if r[0] in ('conv', 'maxpool')]
    assert len(my_ops) == len(ref_ops)
    for op, r in zip(my_ops, ref_ops):
        hin = op.inputs[0].shape[1]
        k, s_ = (r[2], r[3]) if r[0] == 'conv' else (r[1], r[2])
        padding = r[4] if r[0] == 'conv' else r[3]
        if padding == 'SAME':
            pt, oh = _same_pad(hin, k, s_)
        else:
            pad = r[5]
            pt, oh = pad[0], (hin + pad[0] + pad[1] - k) // s_ + 1
        assert tuple(op.attrs['pad']) == (pt, pt) and op.output.shape[1] == oh and op.output.shape[2] == oh, (op.name, r)
    FLAGS.reset()


@pytest.mark.parametrize('idx', [0, 1])
def test_mobilenet_v1_architecture_matches_the_reference_source(idx):
    """utils/external/mobilenet_v1.py + forward_fn executed against a slim stub (arg_scope semantics, shape-tracking
    symbolic tensors) vs this repo's MobileNet-v1 graph: layer order, scopes, filters, kernels, strides, depthwise
    multiplier, BN decay 0.9997 / epsilon 1e-3 / mode, ReLU6, 7x7 average pool (= the global mean at 224x224), logits conv
    with bias.  Flagged deviation: the reference's Dropout_1b (keep_prob 0.999, training only) is the identity here."""
    import importlib
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    g = _ref_gold()['mobilenet_architecture'][idx]
    assert g['mobilenet_version'] == 1 and g['depth_mult'] == 1.0
    FLAGS.reset()
    mod = importlib.reload(importlib.import_module('pocketflow_b200.nets.mobilenet_at_ilsvrc12'))
    FLAGS.batch_size, FLAGS.nb_classes = 2, 1001
    mh = mod.ModelHelper()
    gr = G.Graph()
    with gr.as_default():
        with G.variable_scope('data'):
            im, _ = mh.build_dataset_train().get_next()
        with G.variable_scope('model'):
            (mh.forward_train if g['is_train'] else mh.forward_eval)(im)
    mine = []
    for op in gr.ops:
        scope = op.name.split('/')[-2] if '/' in op.name else ''
        if op.type == 'Conv2D':
            mine.append(['conv', scope, op.output.shape[-1], op.attrs['ksize'][0], op.attrs['strides'][0], 'bias' in op.vars])
        elif op.type == 'DepthwiseConv2dNative':
            mine.append(['dwconv', scope, op.attrs['ksize'][0], op.attrs['strides'][0]])
        elif op.type == 'FusedBatchNorm':
            mine.append(['bn', float(op.attrs['momentum']), float(op.attrs['epsilon']), bool(op.attrs['training'])])
        elif op.type == 'Relu6':
            mine.append(['relu6'])
        elif op.type == 'Mean':
            assert op.inputs[0].shape[1:3] == (7, 7)
            mine.append(['avgpool7'])
    want = []
    for r in g['layers']:
        if r[0] == 'conv':
            assert r[5] == 'SAME'
            want.append(['conv', r[1], r[2], r[3], r[4], r[6]])
        elif r[0] == 'dwconv':
            assert r[4] == 'SAME' and r[5] == 1
            want.append(['dwconv', r[1], r[2], r[3]])
        elif r[0] == 'bn':
            assert r[4] and r[5]                                   # center and scale
            want.append(['bn', r[1], r[2], r[3]])
        elif r[0] == 'avgpool':
            assert r[1] == [7, 7] and r[2] == 'VALID'
            want.append(['avgpool7'])
        elif r[0] == 'dropout':
            assert r[1] == 0.999                                   # identity here (flagged deviation), identity in eval anyway
        else:
            want.append([r[0]])
    assert mine == want
    FLAGS.reset()


def test_maskable_variable_selection_matches_the_reference_function():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from graphs_for_golden import GRAPHS, build_graph
    from pocketflow_b200.learners.weight_sparsification.utils import get_maskable_vars
    from pocketflow_b200.flags import FLAGS
    for g in _ref_gold()['ws_maskable_vars']:
        net, flags, dst = GRAPHS[g['graph']]
        graph = build_graph(net, flags, dst)
        tv = [v for v in graph.variables.values() if v.name.startswith('model/') and v.trainable]
        assert len(tv) == g['n_trainable']
        assert [v.name for v in get_maskable_vars(tv)] == g['maskable'], g['graph']
    FLAGS.reset()


def test_lenet_architecture_matches_the_reference_source():
    import importlib
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    FLAGS.reset()
    mod = importlib.reload(importlib.import_module('pocketflow_b200.nets.lenet_at_cifar10'))
    FLAGS.batch_size = 2
    FLAGS.nb_classes = 10          # the flag's default belongs to whichever dataset module was imported first
    mh = mod.ModelHelper()
    gr = G.Graph()
    with gr.as_default():
        with G.variable_scope('data'):
            im, _ = mh.build_dataset_train().get_next()
        with G.variable_scope('model'):
            mh.forward_train(im)
    mine = []
    for op in gr.ops:
        if op.type == 'Conv2D':
            assert tuple(op.attrs['pad']) == (0, 0)                                  # 'valid'
            mine.append(['conv', op.output.shape[-1], op.attrs['ksize'][0], op.attrs['strides'][0], 'VALID', 'bias' in op.vars])
        elif op.type == 'MaxPool':
            assert tuple(op.attrs['pad']) == (0, 0)
            mine.append(['maxpool', op.attrs['ksize'][0], op.attrs['strides'][0], 'VALID'])
        elif op.type == 'Relu':
            mine.append(['relu'])
        elif op.type == 'MatMul':
            mine.append(['dense', op.output.shape[-1]])
        elif op.type == 'Softmax':
            mine.append(['softmax'])
        elif op.type == 'Reshape':
            mine.append(['flatten'])
    assert mine == _ref_gold()['lenet_architecture']
    FLAGS.reset()


def test_model_helpers_flag_defaults_schedules_and_names_match_the_reference_source(monkeypatch):
    """Every nets/*_at_*.py of the reference, loaded under the stub: the flag defaults it declares, model / dataset names
    and what its setup_lrn_rate hands to the schedule (global batch, drop epochs, multipliers) and returns (# of
    iterations) — against this repo's ModelHelpers."""
    import importlib
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
    gold = _ref_gold()['net_helpers']
    assert [g['file'] for g in gold] == ['nets/lenet_at_cifar10.py', 'nets/resnet_at_cifar10.py', 'nets/resnet_at_ilsvrc12.py',
                                         'nets/mobilenet_at_ilsvrc12.py']
    world = {'size': 1}
    monkeypatch.setattr(mgw, 'size', classmethod(lambda cls: world['size']))
    from pocketflow_b200.nets import classification_helper as CH
    seen = []
    real = CH.setup_lrn_rate_piecewise_constant
    monkeypatch.setattr(CH, 'setup_lrn_rate_piecewise_constant',
                        lambda gs, bs, idxs, rates: (seen.append((bs, list(idxs), list(rates))), real(gs, bs, idxs, rates))[1])
    for g in gold:
        FLAGS.reset()
        mod = importlib.reload(importlib.import_module('pocketflow_b200.' + g['file'][:-3].replace('/', '.')))
        for name, default in g['flag_defaults'].items():
            assert getattr(FLAGS, name) == default, (g['file'], name)      # (DEFINE_float(…, 128) holds 128.0)
        if 'resnet' in g['file']:
            FLAGS.resnet_size = int(g['model_name'].split('_')[1])
        mh = mod.ModelHelper()
        assert (mh.model_name, mh.dataset_name) == (g['model_name'], g['dataset_name'])
        for s in g['schedules']:
            FLAGS.enbl_multi_gpu, FLAGS.batch_size = s['enbl_multi_gpu'], s['batch_size']
            FLAGS.nb_smpls_train, FLAGS.nb_epochs_rat = s['nb_smpls_train'], s['nb_epochs_rat']
            world['size'] = s['world']
            del seen[:]
            _, nb_iters = mh.setup_lrn_rate(None)
            assert seen == [(s['global_batch'], s['idxs_epoch'], s['decay_rates'])] and nb_iters == s['nb_iters'], (g['file'], s)
    FLAGS.reset()


def test_flag_defaults_match_the_reference_modules():
    """Every tf.app.flags.DEFINE_* of the reference modules on the path (collected by importing them under the stub)
    against the flag this repo's same-named module declares: same name, same default."""
    import ast
    import os
    from pocketflow_b200.flags import FLAGS
    gold = _ref_gold()['flag_defaults']
    assert len(gold) == 17 and sum(len(v) for v in gold.values()) >= 115
    # flags this build deliberately does not declare (the subsystem behind them is out of scope)
    absent = {
        'learners/channel_pruning_gpu/learner.py': set(),
        'rl_agents/ddpg/running_mean_std.py': {'ddpg_rms_eps'},          # state / return normalisation: dead code upstream
    }
    changed = {
        # the reference ships 'optimal' pruning ratios through files and a separate session; semantics kept, default kept
    }
    missing, different = [], []
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for ref_file, defaults in gold.items():
        if ref_file == 'rl_agents/ddpg/running_mean_std.py':
            continue
        # this repo's declarations, read from the source (several modules re-declare a flag with their own default,
        # so the process-global FLAGS only knows the last one imported)
        tree = ast.parse(open(os.path.join(root, 'pocketflow_b200', ref_file)).read())
        mine = {}
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and getattr(node.func, 'id', '').startswith('DEFINE_'):
                value = ast.literal_eval(node.args[1])
                mine[ast.literal_eval(node.args[0])] = float(value) if node.func.id == 'DEFINE_float' and value is not None else value
        for name, default in defaults.items():
            if name in absent.get(ref_file, ()):
                continue
            if name not in mine:
                missing.append((ref_file, name))
            elif mine[name] != default and (ref_file, name) not in changed:
                different.append((ref_file, name, default, mine[name]))
    FLAGS.reset()
    assert not different, different
    assert not missing, missing


def test_ws_training_loop_prunes_and_saves_when_the_reference_loop_does():
    """WeightSparseLearner.train of the reference was run with a recording session (golden 'ws_train_loop'); this
    repo's train() driven with recording stand-ins must rebuild the masks after the same iterations (including the one
    late application after ws_iter_ratio_end) and save at the same steps."""
    from types import SimpleNamespace
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.weight_sparsification.learner import WeightSparseLearner as L
    for g in _ref_gold()['ws_train_loop']:
        FLAGS.reset()
        FLAGS.ws_mask_update_step, FLAGS.ws_iter_ratio_beg = g['ws_mask_update_step'], g['ws_iter_ratio_beg']
        FLAGS.ws_iter_ratio_end, FLAGS.save_step, FLAGS.summ_step = g['ws_iter_ratio_end'], g['save_step'], g['summ_step']
        ev = dict(train=0, prune=[], save=[], monitor=[], evaluate=0)
        me = SimpleNamespace(sess_train=SimpleNamespace(store=SimpleNamespace(P=None, O=None)), nb_iters_train=g['nb_iters_train'],
                             evaluate=lambda: ev.__setitem__('evaluate', ev['evaluate'] + 1),
                             is_primary_worker=lambda scope='global': True)
        me.train_step = lambda: ev.__setitem__('train', ev['train'] + 1)
        me.prune = lambda: ev['prune'].append(ev['train'])
        me._WeightSparseLearner__save_model = lambda: ev['save'].append(ev['train'])
        me._WeightSparseLearner__monitor_progress = lambda i, t: ev['monitor'].append(i + 1)
        L.train(me)
        want = g['events']
        assert ev['train'] == want['train'] and ev['prune'] == want['prune'], g
        assert ev['save'] == want['save'] and ev['monitor'] == want['monitor'], g
        assert ev['evaluate'] == want['evaluate'], g                     # evaluated after every save (learner.py:131-140)
    FLAGS.reset()


def test_uq_training_loop_cadence_matches_the_reference_loop():
    """UniformQuantLearner.train of the reference, run with a recording session (golden 'uq_train_loop'): warm-start
    restore, barriers, logging / saving / evaluating steps — against this repo's train() with recording stand-ins."""
    from types import SimpleNamespace
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.uniform_quantization.learner import UniformQuantLearner as L
    for g in _ref_gold()['uq_train_loop']:
        FLAGS.reset()
        FLAGS.save_step, FLAGS.summ_step, FLAGS.enbl_warm_start = g['save_step'], g['summ_step'], g['enbl_warm_start']
        ev, n = [], {'train': 0}
        me = SimpleNamespace(finetune_steps=g['finetune_steps'], sess_train=SimpleNamespace(fetch_losses=lambda: {}),
                             train_step=lambda: n.__setitem__('train', n['train'] + 1),
                             auto_barrier=lambda: ev.append(['barrier', n['train']]),
                             evaluate=lambda: ev.append(['evaluate', n['train']]))
        me._UniformQuantLearner__restore_model = lambda is_train: ev.append(['restore', is_train, n['train']])
        me._UniformQuantLearner__save_model = lambda: ev.append(['save', n['train']])
        me._UniformQuantLearner__monitor_progress = lambda r, t, i: (ev.append(['monitor', i + 1]), t)[1]
        L.train(me)
        assert n['train'] == g['nb_train']
        assert ev == [e for e in g['events'] if e != 'init'], g          # ('init': variables are initialised at build time here)
    FLAGS.reset()


def test_full_prec_and_nuq_training_loop_cadence_matches_the_reference_loops():
    from types import SimpleNamespace
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.full_precision.learner import FullPrecLearner as FP
    from pocketflow_b200.learners.nonuniform_quantization.learner import NonUniformQuantLearner as NUQ
    import pocketflow_b200.datasets.cifar10_dataset  # noqa: F401  (declares batch_size for the progress lines)
    for g in _ref_gold()['fp_train_loop']:
        FLAGS.reset()
        FLAGS.save_step, FLAGS.summ_step = g['save_step'], g['summ_step']
        ev, n = [], {'train': 0}
        ex = SimpleNamespace(store=SimpleNamespace(P=None, O=None), fetch_losses=lambda: dict(loss=0.0))
        me = SimpleNamespace(sess_train=ex, nb_iters_train=g['nb_iters_train'], lrn_rate=lambda i: 0.0,
                             warm_start=lambda sess: ev.append('warm_start'), is_primary_worker=lambda scope='global': True,
                             train_step=lambda: n.__setitem__('train', n['train'] + 1),
                             evaluate=lambda: ev.append(['evaluate', n['train']]))
        me._FullPrecLearner__save_model = lambda: ev.append(['save', True, n['train']])
        FP.train(me)
        # one model, one format here: the reference's train-graph -> eval-graph hand-over (restore(False), save(False)) and
        # its log lines have no counterpart
        want = [e for e in g['events'] if e[0] not in ('monitor', 'restore') and e[:2] != ['save', False]]
        assert n['train'] == g['nb_train'] and ev == want, g
    for g in _ref_gold()['nuq_train_loop']:
        FLAGS.reset()
        FLAGS.save_step, FLAGS.summ_step = g['save_step'], g['summ_step']
        ev, n = [], {'train': 0}
        ex = SimpleNamespace(store=SimpleNamespace(P=None, O=None), fetch_losses=lambda: dict(loss=0.0, model_loss=0.0, acc_top1=0.0))
        me = SimpleNamespace(sess_train=ex, finetune_steps=g['finetune_steps'], lrn_rate=lambda i: 0.0,
                             is_primary_worker=lambda scope='global': True,
                             train_step=lambda: n.__setitem__('train', n['train'] + 1),
                             auto_barrier=lambda: ev.append(['barrier', n['train']]),
                             evaluate=lambda: ev.append(['evaluate', n['train']]))
        me._NonUniformQuantLearner__save_model = lambda: ev.append(['save', n['train']])
        NUQ.train(me)
        assert n['train'] == g['nb_train'] and ev == [e for e in g['events'] if e[0] != 'monitor'], g
    FLAGS.reset()


# ---------------------------------------------------------------------------- channel selection
def _cpg_gold():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'ref_executed_cpg_v1.json')))


def test_cpg_prox_step_matches_the_reference_op_chain():
    """oracle.cpg_prox_step / cpg_channel_mask against the output of the reference's own __build_layer_ops
    (learners/channel_pruning_gpu/learner.py:356-402) executed on numpy tensors (tests/golden/make_golden_cpg.py)."""
    import hashlib
    from oracle import pf_oracle as O
    for g in _cpg_gold()['prox_step']:
        rng = np.random.default_rng(g['seed'])
        w0 = rng.standard_normal(g['shape']).astype(np.float32)
        g0 = rng.standard_normal(g['shape']).astype(np.float32)
        if g['zero_channel'] is not None:
            w0[:, :, g['zero_channel'], :] = 0.0
            g0[:, :, g['zero_channel'], :] = 0.0
        keep = (rng.random(g['shape'][2]) > 0.3).astype(np.float32)
        out, norms, thr = O.cpg_prox_step(w0, g0, g['lrn_rate_pgd'], g['prune_perctl'])
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == g['sha256'], g
        assert [int(c) for c in np.where(np.abs(out).sum(axis=(0, 1, 3)) == 0)[0]] == g['zero_channels']
        assert [int(v) for v in keep] == g['mask_keep']
        masked = (g0 * (keep[None, None, :, None] * np.ones(g['shape'], np.float32))).astype(np.float32)
        assert hashlib.sha256(masked.tobytes()).hexdigest() == g['masked_grad_sha256']       # grad * mask (:438)
        m = O.cpg_channel_mask(out)
        assert sorted(set(np.where(m[0, 0, :, 0] == 0)[0].tolist())) == g['zero_channels']


def test_cpg_selection_loop_feeds_what_the_reference_loop_feeds():
    """ChannelPrunedGpuLearner.__choose_channels of the reference, run with a recording session that replays given
    regression losses (golden 'choose_channels'), against (a) the oracle's schedule and (b) this repo's
    choose_channels() driven with recording stand-ins for the device methods."""
    from types import SimpleNamespace
    from oracle import pf_oracle as O
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.channel_pruning_gpu.learner import ChannelPrunedGpuLearner as L
    gold = _cpg_gold()
    FLAGS.reset()
    for k, v in gold['flag_defaults'].items():
        assert getattr(FLAGS, k) == v, k
    for g in gold['choose_channels']:
        FLAGS.reset()
        n_it = int(g['cpg_nb_iters_layer'] / g['world'])
        ratios = [g['cpg_prune_ratio']] * g['nb_layers']
        if g['cpg_skip_ht_layers']:
            ratios[0] = ratios[-1] = 0.0
        # (a) the oracle's schedule, layer by layer
        want = []
        for layer, r in enumerate(ratios):
            if r == 0.0:
                continue
            for lr, pc in O.cpg_selection_schedule([g['reg_losses'][i % len(g['reg_losses'])] for i in range(n_it)], r, n_it):
                want.append(['prune', layer, lr, pc])
            want.append(['mask_updt', layer])
            want += [['finetune', layer]] * n_it
        assert len(want) == len(g['events'])
        for a, b in zip(want, g['events']):
            assert a[:2] == b[:2]
            if a[0] == 'prune':
                assert abs(a[2] - b[2]) <= 1e-7 * b[2]                       # fed through a float32 placeholder
                assert abs(a[3] - b[3]) <= 1e-6 * max(b[3], 1.0)
        # (b) this repo's host loop
        ev, cnt = [], {}

        def sel_prune(layer, lr, pc, ev=ev, cnt=cnt, g=g):
            i = cnt.get(layer, 0)
            cnt[layer] = i + 1
            ev.append(['prune', layer, lr, pc])
            return g['reg_losses'][i % len(g['reg_losses'])]
        me = SimpleNamespace(prune_ratios=ratios, nb_layers=g['nb_layers'], is_primary_worker=lambda scope='global': False,
                             sel_prune=sel_prune, sel_update_mask=lambda layer: ev.append(['mask_updt', layer]),
                             sel_finetune=lambda layer, it: (ev.append(['finetune', layer]), 1.0)[1],
                             sel_prune_ratio=lambda layer: 0.0)
        L.choose_channels(me, nb_iters_layer=n_it)
        assert len(ev) == len(g['events'])
        for a, b in zip(ev, g['events']):
            assert a[:2] == b[:2]
            if a[0] == 'prune':
                assert abs(a[2] - b[2]) <= 1e-7 * b[2] and abs(a[3] - b[3]) <= 1e-6 * max(b[3], 1.0)
    FLAGS.reset()


def test_codebook_gradient_kat():
    """Hand-derived: gradient_override_map {'Mul': 'Add', 'Sign': 'Identity'} around qx = gather(c, idx) * sign(x_n + 1e-6)
    (learners/nonuniform_quantization/utils.py:303-306) sends the upstream gradient unchanged to BOTH factors — a
    segment sum into the codebook, the identity into x_n; the inverse scale alpha * q + beta (:433) contributes alpha.
    w = [0, 1, 2, 4]: alpha = 4 (+1e-10), x_n = [0, .25, .5, 1]; codebook [0.2, 0.9] -> idx = [0, 0, 0, 1]
    (|.5 - .2| = .3 < |.5 - .9| = .4); g = [1, 2, 3, 5]:  dL/dc = alpha * [1 + 2 + 3, 5] = [24, 20],  dL/dw = g."""
    import torch
    from oracle import pf_oracle as O
    from oracle.step_oracle import codebook_quant
    w = np.array([0.0, 1.0, 2.0, 4.0], np.float32)
    c = np.array([0.2, 0.9], np.float32)
    g = np.array([1.0, 2.0, 3.0, 5.0], np.float32)
    q, _, idx = O.nonuniform_quantize(w, 1, c)
    assert idx.tolist() == [0, 0, 0, 1] and np.allclose(q, [0.8, 0.8, 0.8, 3.6], rtol=1e-6)
    gx, gc = O.nuq_grads(g, idx, 2, np.float32(4.0))
    assert np.allclose(gc, [24.0, 20.0]) and np.allclose(gx, g)
    wt = torch.tensor(w, requires_grad=True)
    ct = torch.tensor(c, requires_grad=True)
    out = codebook_quant(wt, ct)
    assert np.allclose(out.detach().numpy(), q, rtol=1e-6)
    out.backward(torch.tensor(g))
    assert np.allclose(ct.grad.numpy(), [24.0, 20.0], rtol=1e-6) and np.allclose(wt.grad.numpy(), g, rtol=1e-6)


def test_ws_layerwise_regression_is_wired_like_the_reference():
    """The pruning-ratio search's regression stage (learners/weight_sparsification/pr_optimizer.py:283-314): the ops
    get_ops_by_scope_n_patterns picks on this repo's graphs (golden 'ws_core_ops', from the reference's own function)
    against WeightSparseLearner.pr_core_ops; and what __build_layer_rg_ops builds, executed from the reference source:
    l2_loss(out_pruned - out_full), Adam at ws_lrn_rate_rg, gradient times mask."""
    import os
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from graphs_for_golden import GRAPHS, build_graph
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.weight_sparsification.learner import WeightSparseLearner as L
    from pocketflow_b200.learners.weight_sparsification.utils import get_maskable_vars
    gold = _cpg_gold()
    rg = gold['ws_layer_regression']
    assert rg['loss'] == 'l2_loss' and rg['loss_is_pruned_minus_full'] and rg['masked_grad_matches']
    FLAGS.reset()
    assert rg['lrn_rate'] == FLAGS.ws_lrn_rate_rg == 3e-2
    for e in gold['ws_core_ops']:
        net, flags, dst = GRAPHS[e['graph']]
        g = build_graph(net, flags, dst)
        model_name = 'mobilenet_v1' if e['graph'].startswith('mobilenet') else 'resnet_20'
        me = SimpleNamespace(model_name=model_name, model_scope='model', sess_train=SimpleNamespace(ops=g.ops))
        got = L.pr_core_ops(me)
        assert [o.name for o in got] == e['ops'], e['graph']
        # ... and they pair, by index, with the variables the learner masks (pr_optimizer.py:303)
        trainable = [v for v in g.variables.values() if v.name.startswith('model/') and v.trainable]
        assert [o.vars['kernel'].name for o in got] == [v.name for v in get_maskable_vars(trainable)], e['graph']
    FLAGS.reset()
