"""Golden values produced by the REFERENCE'S OWN host-side code, run in this container from /root/reference under a
stub `tensorflow` module (TensorFlow 1.x itself cannot be imported here).  Only pure-Python / numpy functions of the
path can run this way — the learning-rate / fine-tuning schedule helpers and the 'heurist' pruning-ratio formula; the
tensor arithmetic (fake-quant, masks, losses) lives inside TensorFlow ops and stays pinned by hand-derived KATs only.

  python tests/golden/make_golden_from_reference.py        ->  tests/golden/ref_executed_v1.json

The stub provides exactly what the imported reference modules touch at import / call time: tf.app.flags (DEFINE_* +
FLAGS), tf.logging.info, tf.train.piecewise_constant (records its arguments), tf.shape (identity on .shape)."""
import importlib.util
import json
import os
import sys
import types

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_executed_v1.json')


class Flags(object):
    pass


def make_tf_stub(flags):
    tf = types.ModuleType('tensorflow')
    tf.app = types.SimpleNamespace(flags=types.SimpleNamespace(FLAGS=flags))

    def define(name, default, doc=''):
        if not hasattr(flags, name):
            setattr(flags, name, default)
    for kind in ('DEFINE_integer', 'DEFINE_float', 'DEFINE_string', 'DEFINE_boolean', 'DEFINE_bool'):
        setattr(tf.app.flags, kind, define)
    tf.logging = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None)
    tf.train = types.SimpleNamespace(piecewise_constant=lambda step, bnds, vals: ('piecewise_constant', list(bnds), list(vals)))
    tf.shape = lambda v: v.shape
    return tf


def load(path, name, stubs):
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def main():
    flags = Flags()
    tf = make_tf_stub(flags)
    world = {'size': 1}
    mgw = types.SimpleNamespace(size=lambda: world['size'], rank=lambda: 0)
    blank = lambda **kw: types.SimpleNamespace(**kw)     # noqa: E731
    stubs = {
        'tensorflow': tf,
        'utils': types.ModuleType('utils'),
        'utils.multi_gpu_wrapper': blank(MultiGpuWrapper=mgw),
        'utils.misc_utils': blank(is_primary_worker=lambda *a: True),
        'learners': types.ModuleType('learners'),
        'learners.abstract_learner': blank(AbstractLearner=object),
        'learners.distillation_helper': blank(DistillationHelper=object),
        'learners.uniform_quantization': types.ModuleType('learners.uniform_quantization'),
        'learners.uniform_quantization.utils': blank(UniformQuantization=object),
        'learners.uniform_quantization.bit_optimizer': blank(BitOptimizer=object),
        'learners.nonuniform_quantization': types.ModuleType('learners.nonuniform_quantization'),
        'learners.nonuniform_quantization.utils': blank(NonUniformQuantization=object),
        'learners.nonuniform_quantization.bit_optimizer': blank(BitOptimizer=object),
        'learners.weight_sparsification': types.ModuleType('learners.weight_sparsification'),
        'learners.weight_sparsification.rl_helper': blank(RLHelper=object),
        'learners.weight_sparsification.utils': blank(get_maskable_vars=lambda *a: []),
        'rl_agents': types.ModuleType('rl_agents'),
        'rl_agents.ddpg': types.ModuleType('rl_agents.ddpg'),
        'rl_agents.ddpg.agent': blank(Agent=object),
    }
    uq = load('learners/uniform_quantization/learner.py', 'ref_uq_learner', stubs)
    nuq = load('learners/nonuniform_quantization/learner.py', 'ref_nuq_learner', stubs)
    lru = load('utils/lrn_rate_utils.py', 'ref_lrn_rate_utils', stubs)
    pro = load('learners/weight_sparsification/pr_optimizer.py', 'ref_pr_optimizer', stubs)

    gold = {'source': 'functions of /root/reference executed under a stub tensorflow module',
            'uq_setup_bnds_decay_rates': [], 'nuq_setup_bnds_decay_rates': [], 'lrn_rate_piecewise_constant': [],
            'ws_heurist_prune_ratios': []}
    base = dict(nb_smpls_train=50000, batch_size_norm=128.0, lrn_rate_init=1e-1, nb_epochs_rat=1.0,
                uql_quant_epochs=60, nuql_quant_epochs=60)
    cases = [('resnet_20', 'cifar_10', 50000), ('resnet_50', 'ilsvrc_12', 1281167), ('mobilenet_v1', 'ilsvrc_12', 1281167)]
    for model, dataset, nsmp in cases:
        for multi, size in ((False, 1), (True, 4), (True, 8)):
            for warm in (False, True):
                for bs in (128, 256):
                    for k, v in base.items():
                        setattr(flags, k, v)
                    flags.nb_smpls_train, flags.batch_size = nsmp, bs
                    flags.enbl_multi_gpu, flags.enbl_warm_start = multi, warm
                    world['size'] = size
                    for mod, key in ((uq, 'uq_setup_bnds_decay_rates'), (nuq, 'nuq_setup_bnds_decay_rates')):
                        init_lr, bnds, rates, steps = mod.setup_bnds_decay_rates(model, dataset)
                        gold[key].append(dict(model=model, dataset=dataset, nb_smpls_train=nsmp, batch_size=bs,
                                              enbl_multi_gpu=multi, world=size, enbl_warm_start=warm,
                                              init_lr=float(init_lr), bnds=[int(b) for b in bnds],
                                              decay_rates=[float(r) for r in rates], finetune_steps=int(steps)))
    world['size'] = 1
    for nsmp, bs, idxs, rates, rat in [(50000, 128, [100, 150, 200], [1.0, 0.1, 0.01, 0.001], 1.0),
                                       (50000, 256, [100, 150, 200], [1.0, 0.1, 0.01, 0.001], 0.5),
                                       (1281167, 256, [30, 60, 80, 90], [1.0, 0.1, 0.01, 0.001, 0.0001], 1.0),
                                       (1281167, 2048, [30, 60, 80, 90], [1.0, 0.1, 0.01, 0.001, 0.0001], 0.25)]:
        flags.nb_smpls_train, flags.nb_epochs_rat, flags.lrn_rate_init, flags.batch_size_norm = nsmp, rat, 1e-1, 128.0
        _, bnds, vals = lru.setup_lrn_rate_piecewise_constant(None, bs, idxs, rates)
        gold['lrn_rate_piecewise_constant'].append(dict(nb_smpls_train=nsmp, batch_size=bs, idxs_epoch=idxs, decay_rates=rates,
                                                        nb_epochs_rat=rat, bnds=[int(b) for b in bnds],
                                                        vals=[float(v) for v in vals]))
    heur = getattr(pro.PROptimizer, '_PROptimizer__calc_heurist_prune_ratios')
    for ratio, shapes in [(0.5, [(3, 3, 16, 16), (3, 3, 16, 32), (1, 1, 32, 64), (64, 10)]),
                          (0.75, [(7, 7, 3, 64), (1, 1, 64, 256), (3, 3, 64, 64), (1, 1, 2048, 512), (2048, 1001)])]:
        flags.ws_prune_ratio = ratio
        var = [types.SimpleNamespace(name='v%d:0' % i, shape=s) for i, s in enumerate(shapes)]
        fake = types.SimpleNamespace(sess=types.SimpleNamespace(run=lambda x: x), vars_full={'maskable': var})
        pairs = heur(fake)
        gold['ws_heurist_prune_ratios'].append(dict(ws_prune_ratio=ratio, shapes=[list(s) for s in shapes],
                                                    ratios=[float(r) for _, r in pairs]))
    # ---- which ops the reference's UniformQuantization selects (search_matmul_op / search_activation_op,
    # learners/uniform_quantization/utils.py:115-134), run on the (type, name) lists of the graphs this repo builds
    tf.constant = lambda *a, **k: 0
    tf.int32 = 'int32'
    contrib = types.ModuleType('tensorflow.contrib')
    contrib.graph_editor = types.SimpleNamespace()
    tf.contrib = contrib
    stubs2 = dict(stubs)
    stubs2.update({'tensorflow.contrib': contrib, 'tensorflow.contrib.graph_editor': contrib.graph_editor})
    uqu = load('learners/uniform_quantization/utils.py', 'ref_uq_utils', stubs2)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from graphs_for_golden import op_lists
    gold['uq_op_selection'] = []
    for name, ops_ in op_lists().items():
        fake_ops = [types.SimpleNamespace(type=t, name=n) for t, n in ops_]
        sess = types.SimpleNamespace(graph=types.SimpleNamespace(get_operations=lambda fo=fake_ops: fo))
        for qall in (False, True):
            q = uqu.UniformQuantization(sess, 256, True, 'channel')
            mm = [o.name for o in q.search_matmul_op(qall)]
            aa = [o.name for o in q.search_activation_op()]
            gold['uq_op_selection'].append(dict(graph=name, quantize_all_layers=qall, n_ops=len(ops_), matmul=mm, activation=aa))
    # ---- WeightSparseLearner.__calc_prune_ratio_dyn (learners/weight_sparsification/learner.py:296-312): float32
    # graph arithmetic on scalars; the stub maps tf.cast / minimum / maximum / pow one-to-one onto numpy float32 ops
    import numpy as np
    tf.float32 = np.float32
    tf.cast = lambda x, dt: dt(x)
    tf.minimum = lambda a, b: np.minimum(np.float32(a), np.float32(b))
    tf.maximum = lambda a, b: np.maximum(np.float32(a), np.float32(b))
    tf.pow = lambda a, b: np.power(np.float32(a), np.float32(b))
    stubs3 = dict(stubs)
    stubs3['learners.weight_sparsification.pr_optimizer'] = blank(PROptimizer=object)
    wsl = load('learners/weight_sparsification/learner.py', 'ref_ws_learner', stubs3)
    dyn = getattr(wsl.WeightSparseLearner, '_WeightSparseLearner__calc_prune_ratio_dyn')
    flags.ws_iter_ratio_beg, flags.ws_iter_ratio_end, flags.ws_prune_ratio_exp = 0.1, 0.5, 3.0
    gold['ws_prune_ratio_dyn'] = []
    for nb_iters, fnl in [(10000, 0.5), (97650, 0.75), (1234, 0.9)]:
        steps = sorted(set([0, 1, int(nb_iters * 0.1) - 1, int(nb_iters * 0.1), int(nb_iters * 0.1) + 1, int(nb_iters * 0.2),
                            int(nb_iters * 0.3), int(nb_iters * 0.37), int(nb_iters * 0.5) - 1, int(nb_iters * 0.5),
                            int(nb_iters * 0.5) + 1, nb_iters]))
        for st in steps:
            fake = types.SimpleNamespace(nb_iters_train=nb_iters, global_step=st)
            v = np.float32(fnl) * 1 if False else dyn(fake, np.float32(fnl))
            gold['ws_prune_ratio_dyn'].append(dict(nb_iters_train=nb_iters, prune_ratio_fnl=fnl, global_step=st,
                                                   value_f32_hex=np.float32(v).tobytes().hex(), value=float(np.float32(v))))
    # ---- UniformQuantization.__uniform_quantize ITSELF (learners/uniform_quantization/utils.py:163-289: __scale,
    # __inv_scale, __split_bucket, __channel_bucket) executed from the reference source on numpy-backed tensors.  The
    # tensor stub maps every op one-to-one onto the numpy float32 op (each individually rounded, no FMA); the only
    # semantic assumption is tf.round = round-half-to-even (np.rint).  This pins the STRUCTURE of the restatement: where
    # eps is added, the bucket reshapes, last-element padding, the op order of scale / inverse scale, k = 2^bits - 1.
    import builtins
    import hashlib

    class Dim(object):
        def __init__(self, v):
            self.value = int(v)

    class Shape(list):
        pass

    class T(object):
        def __init__(self, a):
            self.a = np.asarray(a, dtype=np.float32)

        def get_shape(self):
            return Shape(Dim(d) for d in self.a.shape)

        def __getitem__(self, i):
            return T(self.a[i])

        @staticmethod
        def _v(o):
            return o.a if isinstance(o, T) else np.float32(o)

        def __add__(self, o):
            return T(self.a + T._v(o))

        def __radd__(self, o):
            return T(T._v(o) + self.a)

        def __sub__(self, o):
            return T(self.a - T._v(o))

        def __rsub__(self, o):
            return T(T._v(o) - self.a)

        def __mul__(self, o):
            return T(self.a * T._v(o))

        def __rmul__(self, o):
            return T(T._v(o) * self.a)

        def __truediv__(self, o):
            return T(self.a / T._v(o))

        def __rtruediv__(self, o):
            return T(T._v(o) / self.a)

    class Ctx(object):
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def reshape(t, shape):
        return T(t.a.reshape([d.value if isinstance(d, Dim) else int(d) for d in shape]))
    tf.variable_scope = lambda *a, **k: Ctx()
    tf.get_variable_scope = lambda: types.SimpleNamespace(name='scope')
    tf.reduce_max = lambda w, axis=None: T(np.max(w.a, axis=axis))
    tf.reduce_min = lambda w, axis=None: T(np.min(w.a, axis=axis))
    tf.stop_gradient = lambda x: x
    tf.constant = lambda value=0, dtype=None: T(value) if dtype is np.float32 else 0
    tf.cast = lambda x, dt: x if isinstance(x, T) else T(np.float32(x))
    tf.round = lambda x: T(np.rint(x.a))
    tf.reshape = reshape
    tf.ones = lambda n: T(np.ones(int(n), np.float32))
    tf.concat = lambda ts, axis=0: T(np.concatenate([t.a for t in ts], axis=axis))
    uqu2 = load('learners/uniform_quantization/utils.py', 'ref_uq_utils2', stubs2)
    quant = getattr(uqu2.UniformQuantization, '_UniformQuantization__uniform_quantize')
    fake_graph = types.SimpleNamespace(gradient_override_map=lambda m: Ctx())
    gold['uniform_quantize'] = []
    cases = []
    for shape in [(3, 3, 8, 16), (1, 1, 64, 10), (5, 5, 3, 7), (64, 10), (3, 3, 3, 1), (2, 2, 2, 2)]:
        for bits in (1, 2, 4, 8, 32):
            for mode, use_b, btype in (('weight', False, 'channel'), ('weight', True, 'channel'), ('weight', True, 'split'),
                                       ('activation', False, 'channel')):
                cases.append((shape, bits, mode, use_b, btype))
    _print = builtins.print
    builtins.print = lambda *a, **k: None                      # the reference prints "Quantized: ..." per call
    try:
        for ci, (shape, bits, mode, use_b, btype) in enumerate(cases):
            rng = np.random.default_rng(1000 + ci)
            x = (rng.standard_normal(shape) * rng.choice([1e-3, 1.0, 37.0])).astype(np.float32)
            if ci % 7 == 3:
                x[...] = x.flat[0]                              # constant tensor: alpha = 1e-10
            obj = uqu2.UniformQuantization(types.SimpleNamespace(graph=fake_graph), 256, use_b, btype)
            q = quant(obj, T(x), bits, mode, 'p')
            out = np.ascontiguousarray(q.a, np.float32)
            assert out.shape == tuple(shape)
            gold['uniform_quantize'].append(dict(seed=1000 + ci, shape=list(shape), bits=bits, mode=mode, use_buckets=use_b,
                                                 bucket_type=btype, constant=(ci % 7 == 3),
                                                 sha256=hashlib.sha256(out.tobytes()).hexdigest(),
                                                 first=[float(v) for v in out.reshape(-1)[:3]]))
    finally:
        builtins.print = _print
    # ---- WeightSparseLearner.__build_masks (learners/weight_sparsification/learner.py:260-294) executed from the
    # reference source on numpy-backed variables: bkup refresh, threshold on |bkup|, strict '>', var = bkup * mask, and the
    # order the control dependencies impose.  tf.contrib.distributions.percentile is NOT available: the stub calls the
    # oracle's percentile_nearest (the 'nearest' index rule stays a documented recollection); everything around it is
    # the reference's own code.
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import pf_oracle as ORC

    class BT(object):                                       # boolean tensor
        def __init__(self, a):
            self.a = np.asarray(a, dtype=bool)

    class V(T):                                             # variable
        def __init__(self, name, a):
            T.__init__(self, a)
            self.name, self.shape = name, tuple(np.asarray(a).shape)

        def assign(self, t):
            self.a = np.array(t.a, dtype=np.float32)
            return self

        def initialized_value(self):
            return T(self.a.copy())
    T.__gt__ = lambda self, o: BT(self.a > T._v(o))
    variables = {}

    def get_variable(name, initializer=None, trainable=True):
        if name not in variables:
            variables[name] = V(name, np.array(initializer.a, dtype=np.float32))
        return variables[name]
    tf.get_variable = get_variable
    tf.ones = lambda shape: T(np.ones(shape if isinstance(shape, tuple) else int(shape), np.float32))
    tf.where = lambda c, a, b: T(np.where(c.a, a.a, b.a))
    tf.abs = lambda x: T(np.abs(x.a))
    tf.cast = lambda x, dt: T(x.a.astype(np.float32)) if isinstance(x, BT) else (x if isinstance(x, T) else T(np.float32(x)))
    tf.minimum = lambda a, b: T(np.minimum(T._v(a), T._v(b)))
    tf.maximum = lambda a, b: T(np.maximum(T._v(a), T._v(b)))
    tf.pow = lambda a, b: T(np.power(T._v(a), T._v(b)))
    tf.control_dependencies = lambda deps: Ctx()
    tf.group = lambda ops_: ops_
    contrib.distributions = types.SimpleNamespace(
        percentile=lambda x, q: T(ORC.percentile_nearest(x.a, np.float32(q.a if isinstance(q, T) else q))))
    wsl2 = load('learners/weight_sparsification/learner.py', 'ref_ws_learner2', stubs3)
    build = getattr(wsl2.WeightSparseLearner, '_WeightSparseLearner__build_masks')
    dyn2 = getattr(wsl2.WeightSparseLearner, '_WeightSparseLearner__calc_prune_ratio_dyn')
    gold['ws_build_masks'] = []
    for ci, (shape, fnl, nb_iters, steps) in enumerate([((3, 3, 8, 16), 0.5, 1000, (150, 300, 700)),
                                                        ((1, 1, 64, 32), 0.75, 1000, (100, 101, 499, 500)),
                                                        ((257,), 0.9, 400, (60, 120, 200)),
                                                        ((16, 10), 0.3, 1000, (250, 260))]):
        rng = np.random.default_rng(2000 + ci)
        variables.clear()
        var = V('model/w%d:0' % ci, rng.standard_normal(shape).astype(np.float32))
        var.a[var.a.reshape(-1).argsort()[:2].tolist()] if False else None
        flat = var.a.reshape(-1)
        flat[1] = flat[0]                                   # a tie in |w|
        fake = types.SimpleNamespace(mask_scope='mask', maskable_vars=[var], var_names_n_prune_ratios=[(var.name, fnl)],
                                     nb_iters_train=nb_iters, global_step=0)
        fake._WeightSparseLearner__calc_prune_ratio_dyn = types.MethodType(dyn2, fake)
        rec = dict(seed=2000 + ci, shape=list(shape), prune_ratio_fnl=fnl, nb_iters_train=nb_iters, updates=[])
        for st in steps:
            fake.global_step = st
            masks, _ = build(fake)
            mask, bkup = masks[0], variables[var.name.replace(':0', '_var_bkup')]
            rec['updates'].append(dict(global_step=st, noise_seed=3000 + st,
                                       var=hashlib.sha256(np.ascontiguousarray(var.a).tobytes()).hexdigest(),
                                       bkup=hashlib.sha256(np.ascontiguousarray(bkup.a).tobytes()).hexdigest(),
                                       mask=hashlib.sha256(np.ascontiguousarray(mask.a).tobytes()).hexdigest(),
                                       kept=int(mask.a.sum())))
            # "training" between mask updates: surviving weights move, pruned ones stay zero (masked gradients)
            nz = np.random.default_rng(3000 + st).standard_normal(shape).astype(np.float32) * np.float32(0.05)
            var.a = (var.a + nz * mask.a).astype(np.float32)
        gold['ws_build_masks'].append(rec)
    # ---- NonUniformQuantization.__nonuni_quantize (learners/nonuniform_quantization/utils.py:168-194 with __scale,
    # __quantile_init, __build_norm_quant_point, __inv_scale) executed from the reference source on numpy tensors
    # (no buckets, 'weights' mode — the built configuration).  percentile = the oracle's 'nearest' rule again; argmin,
    # gather, tile, sign, expand_dims map one-to-one onto numpy.
    T.shape = property(lambda self: tuple(self.a.shape))
    tf.int64 = 'int64'
    tf.cast = lambda x, dtype=None, dt=None: (int(x) if (dtype or dt) == 'int64' else
                                              (T(x.a.astype(np.float32)) if isinstance(x, BT) else
                                               (x if isinstance(x, T) else T(np.float32(x)))))
    tf.range = lambda n: list(range(int(n)))
    tf.map_fn = lambda fn, elems, dtype=None: T(np.stack([np.asarray(T._v(fn(e)), np.float32) for e in elems]))
    contrib.distributions = types.SimpleNamespace(
        percentile=lambda x, q, axis=None: T(ORC.percentile_nearest(x.a, float(q), axis=axis)))
    tf.get_variable = lambda name, validate_shape=True, initializer=None, trainable=True: T(np.array(initializer.a, np.float32))
    tf.ones = lambda n, dtype=None: [1] * int(n) if dtype == 'int64' else T(np.ones(int(n), np.float32))
    tf.concat = lambda ts, axis=0: ([int(v) for part in ts for v in part] if isinstance(ts[0], list)
                                    else T(np.concatenate([t.a for t in ts], axis=axis)))
    tf.expand_dims = lambda x, axis: T(np.expand_dims(x.a, axis))
    tf.tile = lambda x, reps: T(np.tile(x.a, reps))
    tf.argmin = lambda x, axis=-1: np.argmin(x.a, axis=axis)
    tf.gather = lambda c, idx: T(c.a[idx])
    tf.sign = lambda x: T(np.sign(x.a))
    tf.abs = lambda x: T(np.abs(x.a))
    nuqu = load('learners/nonuniform_quantization/utils.py', 'ref_nuq_utils', stubs2)
    nquant = getattr(nuqu.NonUniformQuantization, '_NonUniformQuantization__nonuni_quantize')
    gold['nonuniform_quantize'] = []
    builtins.print = lambda *a, **k: None
    try:
        ci = 0
        for shape in [(3, 3, 8, 16), (1, 1, 64, 10), (5, 5, 3, 7), (64, 10)]:
            for bits in (1, 2, 4):
                rng = np.random.default_rng(4000 + ci)
                x = (rng.standard_normal(shape) * rng.choice([1e-2, 1.0, 9.0])).astype(np.float32)
                obj = nuqu.NonUniformQuantization(types.SimpleNamespace(graph=fake_graph), 256, False, 'quantile', 'split')
                q = nquant(obj, T(x), bits, 'weight', 'p')
                out = np.ascontiguousarray(q.a, np.float32)
                assert out.shape == tuple(shape)
                gold['nonuniform_quantize'].append(dict(seed=4000 + ci, shape=list(shape), bits=bits,
                                                        sha256=hashlib.sha256(out.tobytes()).hexdigest(),
                                                        distinct=int(len(np.unique(out)))))
                ci += 1
    finally:
        builtins.print = _print
    # ---- DistillationHelper.calc_loss (learners/distillation_helper.py:86-103) from the reference source: the WIRING
    # (temperature on both sides, soft-label cross-entropy, weight, no T^2); softmax / softmax_cross_entropy themselves
    # are TF kernels and are supplied by the oracle's numpy versions.
    tf.nn = types.SimpleNamespace(softmax=lambda x: T(ORC.softmax(x.a)))
    tf.losses = types.SimpleNamespace(softmax_cross_entropy=lambda labels, logits: T(ORC.softmax_cross_entropy(labels.a, logits.a)[0]))
    tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None)
    dh = load('learners/distillation_helper.py', 'ref_dst_helper', stubs3)
    gold['distillation_loss'] = []
    for ci, (n, k, w, tt) in enumerate([(8, 10, 4.0, 4.0), (5, 1001, 4.0, 4.0), (3, 7, 1.0, 2.0), (16, 10, 0.5, 8.0)]):
        rng = np.random.default_rng(5000 + ci)
        s_ = (rng.standard_normal((n, k)) * 3).astype(np.float32)
        t_ = (rng.standard_normal((n, k)) * 5).astype(np.float32)
        flags.loss_w_dst, flags.tempr_dst = w, tt
        v = dh.DistillationHelper.calc_loss(T(s_), T(t_))
        gold['distillation_loss'].append(dict(seed=5000 + ci, n=n, k=k, loss_w_dst=w, tempr_dst=tt,
                                              value_f32_hex=np.float32(v.a).tobytes().hex(), value=float(v.a)))
    # ---- ModelHelper.calc_loss of every net on the path (a8): WHICH variables receive the L2 term (the name filter) and
    # the default loss_w_dcy, by running the reference's calc_loss on the trainable-variable NAMES of this repo's graphs
    from graphs_for_golden import GRAPHS, build_graph
    l2_seen = []
    tf.nn = types.SimpleNamespace(l2_loss=lambda v: (l2_seen.append(v.name), T(0.0))[1], softmax=tf.nn.softmax,
                                  in_top_k=lambda *a: BT(np.zeros(1, bool)))
    tf.add_n = lambda ts: T(sum(np.float32(t.a) for t in ts))
    tf.losses = types.SimpleNamespace(softmax_cross_entropy=lambda labels, logits: T(0.0))
    tf.argmax = lambda x, axis=1: T(np.zeros(1))
    tf.equal = lambda a, b: BT(np.zeros(1, bool))
    tf.reduce_mean = lambda x: T(0.0)
    contrib.slim = types.SimpleNamespace()
    stubs4 = dict(stubs3)
    stubs4.update({'tensorflow.contrib': contrib, 'tensorflow.contrib.slim': contrib.slim})
    stubs4.update({'nets': types.ModuleType('nets'), 'nets.abstract_model_helper': blank(AbstractModelHelper=object),
                   'datasets': types.ModuleType('datasets'), 'datasets.cifar10_dataset': blank(Cifar10Dataset=object),
                   'datasets.ilsvrc12_dataset': blank(Ilsvrc12Dataset=object),
                   'utils.external': types.ModuleType('utils.external'), 'utils.external.resnet_model': types.ModuleType('rm'),
                   'utils.external.mobilenet_v1': types.ModuleType('mv1'), 'utils.external.mobilenet_v2': types.ModuleType('mv2'),
                   'utils.lrn_rate_utils': blank(setup_lrn_rate_piecewise_constant=None, setup_lrn_rate_exponential_decay=None)})
    gold['calc_loss_l2'] = []
    for gname, ref_file in [('resnet20_cifar10_dst', 'nets/resnet_at_cifar10.py'), ('mobilenet_v1_ilsvrc12', 'nets/mobilenet_at_ilsvrc12.py'),
                            ('lenet_cifar10', 'nets/lenet_at_cifar10.py')]:
        if hasattr(flags, 'loss_w_dcy'):
            delattr(flags, 'loss_w_dcy')
        mod = load(ref_file, 'ref_' + gname, stubs4)
        net, fl, dst = GRAPHS[gname]
        graph = build_graph(net, fl, dst)
        tv = [types.SimpleNamespace(name=v.name) for v in graph.variables.values() if v.name.startswith('model/') and v.trainable]
        del l2_seen[:]
        mod.ModelHelper.calc_loss(types.SimpleNamespace(), T(np.zeros((1, 2))), T(np.zeros((1, 2))), tv)
        gold['calc_loss_l2'].append(dict(graph=gname, loss_w_dcy=float(flags.loss_w_dcy), n_trainable=len(tv),
                                         regularised=list(l2_seen)))
    # ---- the ResNet ARCHITECTURE the reference builds (utils/external/resnet_model.py driven by the forward_fn of
    # nets/resnet_at_cifar10.py / nets/resnet_at_ilsvrc12.py), recorded by executing the reference source with symbolic
    # tensors: every conv (filters, kernel, stride, explicit padding), batch-norm (momentum, epsilon, training), ReLU,
    # pooling, residual add, mean, dense — in order — to compare with the op list of this repo's graphs.
    rec = []

    class S(object):                                        # symbolic tensor: remembers pending explicit padding
        def __init__(self, pad=None):
            self.pad = pad

        def __add__(self, o):
            rec.append(('add',))
            return S()

    def layers_conv2d(inputs=None, filters=None, kernel_size=None, strides=1, padding='SAME', use_bias=True, **kw):
        rec.append(('conv', int(filters), int(kernel_size), int(strides), padding.upper(),
                    inputs.pad if inputs.pad else [0, 0, 0, 0], bool(use_bias)))
        return S()

    def layers_bn(inputs=None, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False, fused=None, **kw):
        rec.append(('bn', float(momentum), float(epsilon), bool(training), bool(center), bool(scale)))
        return S()

    def tf_pad(inputs, paddings):
        flat = [int(v) for row in paddings for v in row]
        assert flat[:2] == [0, 0] and flat[-2:] == [0, 0], 'NHWC padding expected'
        return S(pad=flat[2:6])                             # [top, bottom, left, right]
    tf.layers = types.SimpleNamespace(
        conv2d=layers_conv2d, batch_normalization=layers_bn,
        max_pooling2d=lambda inputs=None, pool_size=None, strides=None, padding='SAME', **kw: (
            rec.append(('maxpool', int(pool_size), int(strides), padding.upper())), S())[1],
        dense=lambda inputs=None, units=None, **kw: (rec.append(('dense', int(units))), S())[1])
    tf.pad = tf_pad
    tf.nn = types.SimpleNamespace(relu=lambda x: (rec.append(('relu',)), S())[1])
    tf.identity = lambda x, name=None: x
    tf.reduce_mean = lambda x, axes=None, keepdims=False: (rec.append(('mean', list(axes), bool(keepdims))), S())[1]
    tf.squeeze = lambda x, axes=None: x
    tf.transpose = lambda x, perm: x
    tf.variance_scaling_initializer = lambda *a, **k: None
    tf.test = types.SimpleNamespace(is_built_with_cuda=lambda: False)
    tf.float16 = 'float16'

    class VS(Ctx):
        pass
    tf.variable_scope = lambda *a, **k: VS()
    rm = load('utils/external/resnet_model.py', 'ref_resnet_model', stubs4)
    ext = types.ModuleType('utils.external')
    ext.resnet_model = rm
    stubs5 = dict(stubs4)
    stubs5.update({'utils.external': ext, 'utils.external.resnet_model': rm})
    gold['resnet_architecture'] = []
    for ref_file, size, classes in [('nets/resnet_at_cifar10.py', 20, 10), ('nets/resnet_at_cifar10.py', 32, 10),
                                    ('nets/resnet_at_ilsvrc12.py', 18, 1001), ('nets/resnet_at_ilsvrc12.py', 50, 1001)]:
        for attr in ('resnet_size', 'nb_classes'):
            if hasattr(flags, attr):
                delattr(flags, attr)
        mod = load(ref_file, 'ref_net_%d' % size, stubs5)
        flags.resnet_size, flags.nb_classes = size, classes
        for is_train in (True, False):
            del rec[:]
            mod.forward_fn(S(), is_train, 'channels_last')
            gold['resnet_architecture'].append(dict(net=os.path.basename(ref_file)[:-3], resnet_size=size, nb_classes=classes,
                                                    is_train=is_train, layers=[list(r) for r in rec]))
    # ---- the MobileNet-v1 ARCHITECTURE (utils/external/mobilenet_v1.py driven by nets/mobilenet_at_ilsvrc12.py:forward_fn),
    # recorded by executing the reference source against a stub of tf.contrib.slim with arg_scope semantics and
    # shape-tracking symbolic tensors
    mrec = []

    class S4(object):
        def __init__(self, shape):
            self.shape4 = list(shape)

        def get_shape(self):
            return types.SimpleNamespace(as_list=lambda: list(self.shape4))

    def out_hw(h, k, s, padding):
        return -(-h // s) if padding == 'SAME' else (h - k) // s + 1
    scope_stack = [{}]

    class ArgScope(object):
        def __init__(self, fns_or_scope, kw):
            cur = {k: dict(v) for k, v in scope_stack[-1].items()}
            if isinstance(fns_or_scope, dict):
                for k, v in fns_or_scope.items():
                    cur.setdefault(k, {}).update(v)
            else:
                for f in fns_or_scope:
                    cur.setdefault(f.__name__, {}).update(kw)
            self.scope = cur

        def __enter__(self):
            scope_stack.append(self.scope)
            return self.scope

        def __exit__(self, *a):
            scope_stack.pop()
            return False

    def scoped(fn):
        def wrapper(*a, **kw):
            merged = dict(scope_stack[-1].get(fn.__name__, {}))
            merged.update(kw)
            return fn(*a, **merged)
        wrapper.__name__ = fn.__name__
        return wrapper

    def _after(net, normalizer_fn, activation_fn):
        if normalizer_fn is not None:
            bn = scope_stack[-1].get('batch_norm', {})
            mrec.append(('bn', float(bn.get('decay', 0.999)), float(bn.get('epsilon', 0.001)), bool(bn.get('is_training', True)),
                         bool(bn.get('center', True)), bool(bn.get('scale', False))))
        if activation_fn is not None:
            mrec.append((activation_fn.__name__,))
        return net

    @scoped
    def conv2d(inputs, num_outputs, kernel_size, stride=1, padding='SAME', activation_fn=None, normalizer_fn=None, scope=None, **kw):
        n, h, w, c = inputs.shape4
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        mrec.append(('conv', scope, int(num_outputs), int(k), int(stride), padding, normalizer_fn is None))
        return _after(S4([n, out_hw(h, k, stride, padding), out_hw(w, k, stride, padding), num_outputs]), normalizer_fn, activation_fn)

    @scoped
    def separable_conv2d(inputs, num_outputs, kernel_size, depth_multiplier=1, stride=1, rate=1, padding='SAME', activation_fn=None,
                         normalizer_fn=None, scope=None, **kw):
        assert num_outputs is None and rate == 1
        n, h, w, c = inputs.shape4
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        mrec.append(('dwconv', scope, int(k), int(stride), padding, int(depth_multiplier)))
        return _after(S4([n, out_hw(h, k, stride, padding), out_hw(w, k, stride, padding), c * depth_multiplier]), normalizer_fn, activation_fn)

    def batch_norm(*a, **k):
        raise AssertionError('batch_norm is only used as normalizer_fn')

    @scoped
    def dropout(inputs, keep_prob=0.5, is_training=True, scope=None, **kw):
        mrec.append(('dropout', float(keep_prob), bool(is_training)))
        return inputs

    def avg_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None):
        n, h, w, c = inputs.shape4
        mrec.append(('avgpool', [int(v) for v in kernel_size], padding))
        return S4([n, out_hw(h, kernel_size[0], stride, padding), out_hw(w, kernel_size[1], stride, padding), c])

    def relu6(x):
        return x
    slim = types.SimpleNamespace(conv2d=conv2d, separable_conv2d=separable_conv2d, batch_norm=batch_norm, dropout=dropout,
                                 avg_pool2d=avg_pool2d, arg_scope=lambda f, **kw: ArgScope(f, kw))
    contrib.slim = slim
    contrib.layers = types.SimpleNamespace(softmax=lambda logits, scope=None: logits, l2_regularizer=lambda wd: ('l2', wd))
    tf.nn = types.SimpleNamespace(relu6=relu6, relu=lambda x: x)
    tf.GraphKeys = types.SimpleNamespace(UPDATE_OPS='update_ops')
    tf.truncated_normal_initializer = lambda stddev=1.0: ('tn', stddev)
    tf.variable_scope = lambda *a, **k: Ctx()
    tf.squeeze = lambda x, axes=None, name=None: x
    tf.reduce_mean = lambda x, axes=None, keep_dims=False, name=None: (mrec.append(('mean', list(axes))), S4([x.shape4[0], 1, 1, x.shape4[3]]))[1]
    stubs6 = dict(stubs4)
    stubs6.update({'tensorflow.contrib': contrib, 'tensorflow.contrib.slim': slim})
    mv1 = load('utils/external/mobilenet_v1.py', 'ref_mobilenet_v1', stubs6)
    ext2 = types.ModuleType('utils.external')
    ext2.mobilenet_v1, ext2.mobilenet_v2 = mv1, types.ModuleType('mv2')
    stubs6.update({'utils.external': ext2, 'utils.external.mobilenet_v1': mv1, 'utils.external.mobilenet_v2': ext2.mobilenet_v2})
    for attr in ('nb_classes', 'mobilenet_version', 'mobilenet_depth_mult'):
        if hasattr(flags, attr):
            delattr(flags, attr)
    mbn = load('nets/mobilenet_at_ilsvrc12.py', 'ref_mobilenet_net', stubs6)
    gold['mobilenet_architecture'] = []
    flags.nb_classes = 1001
    for is_train in (True, False):
        del mrec[:]
        mbn.forward_fn(S4([2, 224, 224, 3]), is_train)
        gold['mobilenet_architecture'].append(dict(is_train=is_train, mobilenet_version=int(flags.mobilenet_version),
                                                   depth_mult=float(flags.mobilenet_depth_mult),
                                                   layers=[list(r) for r in mrec]))
    # ---- get_maskable_vars (learners/weight_sparsification/utils.py:21-41): which trainable variables the
    # WeightSparseLearner masks, run on the trainable-variable names of this repo's graphs
    wsu = load('learners/weight_sparsification/utils.py', 'ref_ws_utils', stubs3)
    gold['ws_maskable_vars'] = []
    for gname in ('resnet20_cifar10_dst', 'mobilenet_v1_ilsvrc12', 'lenet_cifar10'):
        net, fl, dst = GRAPHS[gname]
        graph = build_graph(net, fl, dst)
        tv = [types.SimpleNamespace(name=v.name) for v in graph.variables.values() if v.name.startswith('model/') and v.trainable]
        gold['ws_maskable_vars'].append(dict(graph=gname, n_trainable=len(tv), maskable=[v.name for v in wsu.get_maskable_vars(tv)]))
    # ---- LeNet (nets/lenet_at_cifar10.py:forward_fn) with the same recorder
    lrec = []

    def l_conv(inputs, filters, kernel_size, strides=1, padding='valid', use_bias=True, **kw):
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        lrec.append(('conv', int(filters), int(k), int(strides), padding.upper(), bool(use_bias)))
        return S()

    def l_pool(inputs, pool_size, strides, padding='valid', **kw):
        k = pool_size if isinstance(pool_size, int) else pool_size[0]
        lrec.append(('maxpool', int(k), int(strides), padding.upper()))
        return S()
    tf.layers = types.SimpleNamespace(conv2d=l_conv, max_pooling2d=l_pool,
                                      flatten=lambda x, name=None: (lrec.append(('flatten',)), S())[1],
                                      dense=lambda x, units, name=None, **kw: (lrec.append(('dense', int(units))), S())[1])
    tf.nn = types.SimpleNamespace(relu=lambda x, name=None: (lrec.append(('relu',)), S())[1],
                                  softmax=lambda x, name=None: (lrec.append(('softmax',)), S())[1])
    for attr in ('nb_classes',):
        if hasattr(flags, attr):
            delattr(flags, attr)
    ln = load('nets/lenet_at_cifar10.py', 'ref_lenet_net', stubs5)
    flags.nb_classes = 10
    ln.forward_fn(S(), 'channels_last')
    gold['lenet_architecture'] = [list(r) for r in lrec]
    # ---- the RL bit search's bookkeeping (learners/uniform_quantization/rl_helper.py, bit_optimizer.py) and the DDPG
    # agent's numpy parts (rl_agents/ddpg/replay_buffer.py, noise.py), executed from the reference source
    import random as _random
    rlh = load('learners/uniform_quantization/rl_helper.py', 'ref_uq_rl_helper', stubs)
    fake_sess = types.SimpleNamespace(run=lambda x: np.array(x))
    layer_sets = {
        'lenet': [(5, 5, 32, 64), (1600, 256)],
        'resnet8': [(1, 1, 16, 16), (3, 3, 16, 16), (3, 3, 16, 16), (1, 1, 16, 32), (3, 3, 16, 32), (3, 3, 32, 32),
                    (1, 1, 32, 64), (3, 3, 32, 64), (3, 3, 64, 64)],
        'mixed': [(3, 3, 3, 8), (64, 10), (1, 1, 8, 128), (3, 3, 128, 4)],
    }
    gold['uq_rl_helper'] = []
    ci = 0
    for lname, shapes in layer_sets.items():
        nums = [int(np.prod(s)) for s in shapes]
        for eq_bits in (2.0, 3.0, 4.0, 6.5, 8.0):
            for (bmin, bmax) in ((2, 8), (1, 4)):
                if eq_bits < bmin:
                    continue
                for rand_layers in (False, True):
                    ci += 1
                    flags.uql_w_bit_min, flags.uql_w_bit_max = bmin, bmax
                    vars_list = [types.SimpleNamespace(shape=s) for s in shapes]
                    h = rlh.RLHelper(fake_sess, sum(nums) * eq_bits, nums, vars_list, random_layers=rand_layers)
                    rng = np.random.RandomState(7000 + ci)
                    rollouts = []
                    _random.seed(100 + ci)
                    for _ in range(3):
                        h.reset()
                        order = list(h.layer_idxs)
                        raw = rng.uniform(0, bmax - bmin, len(shapes))
                        out = [float(h.calc_w(np.array([[raw[k]]]), idx)[0][0]) for k, idx in enumerate(order)]
                        rollouts.append(dict(order=order, raw=raw.tolist(), bits=out, used=float(h.w_bits_used)))
                    gold['uq_rl_helper'].append(dict(
                        shapes=[list(s) for s in shapes], equivalent_bits=eq_bits, w_bit_min=bmin, w_bit_max=bmax,
                        random_layers=rand_layers, py_seed=100 + ci, s_dims=int(h.s_dims),
                        states=[h.calc_state(i)[0].tolist() for i in range(len(shapes))], rollouts=rollouts,
                        reward=h.calc_reward(0.625).tolist()))
    rb = load('rl_agents/ddpg/replay_buffer.py', 'ref_replay_buffer', stubs)
    gold['ddpg_replay_buffer'] = []
    for ci, (buf_size, chunks) in enumerate([(5, [2, 2, 2, 1, 3]), (4, [4, 1]), (7, [3, 3, 3, 3]), (3, [1, 1])]):
        rng = np.random.RandomState(8000 + ci)
        buf = rb.ReplayBuffer(3, 2, buf_size)
        trace = []
        for n in chunks:
            batch = [rng.randn(n, 3), rng.randn(n, 2), rng.randn(n, 1), (rng.rand(n, 1) < 0.3).astype(float), rng.randn(n, 3)]
            buf.append(*batch)
            trace.append(dict(n=n, idx_smpl=int(buf.idx_smpl), nb_smpls=int(buf.nb_smpls), ready=bool(buf.is_ready()),
                              states=buf.buffers['states'].tolist(), rewards=buf.buffers['rewards'].tolist()))
        gold['ddpg_replay_buffer'].append(dict(seed=8000 + ci, buf_size=buf_size, chunks=chunks, trace=trace))
    nz = load('rl_agents/ddpg/noise.py', 'ref_noise', stubs)
    gold['ddpg_noise'] = []
    for init, finl, nb in ((1.0, 1e-5, 200), (0.5, 1e-2, 30)):
        flags.ddpg_noise_std_init, flags.ddpg_noise_std_finl = init, finl
        flags.ddpg_noise_dst_finl, flags.ddpg_noise_adpt_rat = 1e-2, 1.03
        td = nz.TimeDecayNoiseSpec(nb)
        seq = []
        for _ in range(5):
            td.adapt()
            seq.append(td.stdev_curr)
        ad = nz.AdaptiveNoiseSpec()
        seq2 = []
        for dst in (0.5, 0.02, 0.001, 0.0, 0.3):
            ad.adapt(dst)
            seq2.append(ad.stdev_curr)
        gold['ddpg_noise'].append(dict(std_init=init, std_finl=finl, nb_rlouts=nb, tdecy=seq, adapt=seq2))
    # BitOptimizer's budget check and transition recording, called unbound on a minimal stand-in for `self`
    stubs_bo = dict(stubs)
    stubs_bo['learners.uniform_quantization.rl_helper'] = blank(RLHelper=object)
    stubs_bo['rl_agents.ddpg.agent'] = blank(Agent=object)
    bo = load('learners/uniform_quantization/bit_optimizer.py', 'ref_bit_optimizer', stubs_bo)
    recorded = []
    me = types.SimpleNamespace(statistics=dict(nb_matmuls=3, num_weights=[10, 20, 5]), total_bits=35 * 4.0, s_dims=9,
                               agent=types.SimpleNamespace(record=lambda *a: recorded.append([np.asarray(x).tolist() for x in a])))
    check = bo.BitOptimizer._BitOptimizer__check_bits
    me._BitOptimizer__check_bits = lambda bits: check(me, bits)
    gold['uq_bit_optimizer'] = dict(
        check_ok=float(check(me, [4, 3, 8])), arrange=bo.BitOptimizer._BitOptimizer__arrange_layer_bits(me, [2, 0, 1], [8.0, 4.0, 3.0])[0])
    try:
        check(me, [8, 8, 8])
        gold['uq_bit_optimizer']['check_over'] = 'no error'
    except ValueError as e:
        gold['uq_bit_optimizer']['check_over'] = str(e)
    sa = [(np.full((1, 9), float(i)), np.array([[float(2 + i)]])) for i in range(3)]
    bo.BitOptimizer._BitOptimizer__record_rollout_transitions(me, sa, 0.75 * np.ones((1, 1)))
    gold['uq_bit_optimizer']['transitions'] = recorded
    # ---- weight sparsification: RLHelper (learners/weight_sparsification/rl_helper.py) executed from the reference
    wsrl = load('learners/weight_sparsification/rl_helper.py', 'ref_ws_rl_helper', stubs)
    gold['ws_rl_helper'] = []
    ci = 0
    for lname, shapes in layer_sets.items():
        for target in (0.5, 0.75, 0.9):
            for reward_type in ('single-obj', 'multi-obj'):
                for skip in (False, True):
                    ci += 1
                    flags.ws_prune_ratio, flags.ws_reward_type = target, reward_type
                    vars_list = [types.SimpleNamespace(shape=s) for s in shapes]
                    h = wsrl.RLHelper(fake_sess, vars_list, skip)
                    rng = np.random.RandomState(9000 + ci)
                    rollouts = []
                    for _ in range(3):
                        acts = rng.uniform(0, 1, len(shapes))
                        states, ratios, error = [], [], None
                        try:
                            for idx in range(len(shapes)):
                                states.append(h.calc_state(idx)[0].tolist())
                                ratios.append(float(h.cvt_action_to_prune_ratio(idx, acts[idx])))
                        except AssertionError as e:          # the target cannot be reached (e.g. every layer skipped)
                            error = str(e)
                        rollouts.append(dict(actions=acts.tolist(), states=states, ratios=ratios, error=error,
                                             overall=float(h.calc_overall_prune_ratio()), reward=float(h.calc_reward(0.8))))
                    gold['ws_rl_helper'].append(dict(shapes=[list(s) for s in shapes], ws_prune_ratio=target, reward_type=reward_type,
                                                     skip_head_n_tail=skip, s_dims=int(h.s_dims), rollouts=rollouts))
    # ---- input pipelines: datasets/cifar10_dataset.py:parse_fn and utils/external/imagenet_preprocessing.py executed from
    # the reference source.  Random ops are stubs that take their draws from `ctl`; JPEG decoding is Pillow and
    # tf.image.resize_images is this repo's resize_bilinear (so the resize ARITHMETIC is not pinned here — its call
    # arguments, the target-size arithmetic, crop offsets, flip order, mean subtraction and the
    # sample_distorted_bounding_box parameters are).
    import base64
    import hashlib
    import io as _io
    from PIL import Image as _Image
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from pocketflow_b200.datasets.ilsvrc12_dataset import resize_bilinear as _resize

    class TT(np.ndarray):
        def get_shape(self):
            return types.SimpleNamespace(ndims=self.ndim)

        def set_shape(self, shape):
            assert list(self.shape) == list(shape), (self.shape, shape)

    def T(x, dtype=None):
        return np.asarray(x, dtype).view(TT)
    ctl = {}
    calls = []

    def _decode(buf):
        return T(np.asarray(_Image.open(_io.BytesIO(buf)).convert('RGB'), np.uint8))

    def _sdbb(shape, bounding_boxes=None, **kw):
        calls.append(('sample_distorted_bounding_box', {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in kw.items()},
                      np.asarray(bounding_boxes).shape))
        y, x, h, w = ctl['crop']
        return T([y, x, 0]), T([h, w, -1]), None

    def _crop_jpeg(buf, window, channels=3):
        y, x, h, w = [int(v) for v in window]
        return _decode(buf)[y:y + h, x:x + w]

    def _resize_images(image, size, method=None, align_corners=None):
        calls.append(('resize_images', [int(size[0]), int(size[1])], method, align_corners))
        return T(_resize(np.asarray(image), int(size[0]), int(size[1])))

    def _slice(t, begin, size):
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return t[idx]

    def _cast(x, dtype):
        return T(np.asarray(x).astype(dtype)) if np.ndim(x) else dtype(x)
    tfd = make_tf_stub(flags)
    tfd.float32, tfd.int32, tfd.uint8 = np.float32, np.int32, np.uint8
    tfd.constant = lambda v, dtype=None: T(v, dtype)
    tfd.shape = lambda t: np.array(t.shape)
    tfd.cast = _cast
    tfd.minimum = lambda a, b: min(a, b)
    tfd.unstack = lambda t: list(t)
    tfd.stack = lambda vals: T(vals)
    tfd.slice = _slice
    tfd.expand_dims = lambda t, axis: T(np.expand_dims(np.asarray(t, np.float32), axis))
    tfd.reshape = lambda t, shape: T(np.reshape(t, shape))
    tfd.transpose = lambda t, perm: T(np.transpose(t, perm))
    tfd.decode_raw = lambda s, dtype: T(np.frombuffer(s, dtype))
    tfd.one_hot = lambda i, k: T(np.eye(int(k), dtype=np.float32)[int(i)])
    tfd.random_crop = lambda t, size: t[ctl['oy']:ctl['oy'] + size[0], ctl['ox']:ctl['ox'] + size[1]]

    def _crop_or_pad(image, th, tw):
        h, w = image.shape[:2]
        assert th >= h and tw >= w
        out = np.zeros((th, tw, image.shape[2]), image.dtype)
        out[(th - h) // 2:(th - h) // 2 + h, (tw - w) // 2:(tw - w) // 2 + w] = image
        return T(out)
    tfd.image = types.SimpleNamespace(
        sample_distorted_bounding_box=_sdbb, extract_jpeg_shape=lambda b: np.array(_decode(b).shape),
        decode_and_crop_jpeg=_crop_jpeg, decode_jpeg=lambda b, channels=3: _decode(b),
        random_flip_left_right=lambda t: t[:, ::-1] if ctl['flip'] else t,
        resize_images=_resize_images, ResizeMethod=types.SimpleNamespace(BILINEAR='BILINEAR'),
        resize_image_with_crop_or_pad=_crop_or_pad)
    stubs_d = dict(stubs)
    stubs_d['tensorflow'] = tfd
    stubs_d['datasets'] = types.ModuleType('datasets')
    stubs_d['datasets.abstract_dataset'] = blank(AbstractDataset=object)
    ipp = load('utils/external/imagenet_preprocessing.py', 'ref_imagenet_preprocessing', stubs_d)
    cif = load('datasets/cifar10_dataset.py', 'ref_cifar10_dataset', stubs_d)

    def digest(a):
        a = np.ascontiguousarray(np.asarray(a, np.float32))
        return dict(shape=list(a.shape), sha256=hashlib.sha256(a.tobytes()).hexdigest(),
                    samples=[float(a.reshape(-1)[i]) for i in (0, a.size // 3, a.size - 1)])
    gold['imagenet_preprocessing'] = []
    rng = np.random.RandomState(31)
    for ci, (h, w) in enumerate([(300, 400), (375, 500), (256, 256), (500, 333), (90, 120)]):
        base_img = rng.randint(0, 256, (h // 8 + 1, w // 8 + 1, 3)).astype(np.uint8)
        b = _io.BytesIO()
        _Image.fromarray(np.kron(base_img, np.ones((8, 8, 1), np.uint8))[:h, :w]).save(b, format='JPEG', quality=90)
        jpeg = b.getvalue()
        rec = dict(jpeg_b64=base64.b64encode(jpeg).decode('ascii'), height=h, width=w)
        del calls[:]
        rec['eval'] = digest(ipp.preprocess_image(jpeg, np.zeros((1, 0, 4), np.float32), 224, 224, 3, is_training=False))
        rec['eval_calls'] = [list(c) for c in calls]
        rec['train'] = []
        for flip in (False, True):
            ch, cw = int(rng.randint(h // 3, h + 1)), int(rng.randint(w // 3, w + 1))
            cy, cx = int(rng.randint(0, h - ch + 1)), int(rng.randint(0, w - cw + 1))
            ctl.update(crop=(cy, cx, ch, cw), flip=flip)
            del calls[:]
            out = ipp.preprocess_image(jpeg, np.zeros((1, 2, 4), np.float32), 224, 224, 3, is_training=True)
            rec['train'].append(dict(crop=[cy, cx, ch, cw], flip=flip, out=digest(out), calls=[list(c) for c in calls]))
        gold['imagenet_preprocessing'].append(rec)
    gold['imagenet_constants'] = dict(means=[float(v) for v in ipp._CHANNEL_MEANS], resize_min=int(ipp._RESIZE_MIN))
    gold['cifar10_parse_fn'] = []
    flags.nb_classes = 10
    for ci in range(6):
        record = rng.randint(0, 256, 1 + 3 * 32 * 32).astype(np.uint8)
        record[0] = ci + 2
        img, lab = cif.parse_fn(record.tobytes(), is_train=False)
        item = dict(record_b64=base64.b64encode(record.tobytes()).decode('ascii'), label=np.asarray(lab).tolist(), eval=digest(img), train=[])
        for oy, ox, flip in ((0, 0, False), (8, 8, True), (3, 5, False), (4, 4, True)):
            ctl.update(oy=oy, ox=ox, flip=flip)
            img, _ = cif.parse_fn(record.tobytes(), is_train=True)
            item['train'].append(dict(oy=oy, ox=ox, flip=flip, out=digest(img)))
        gold['cifar10_parse_fn'].append(item)
    # ---- every ModelHelper's flag defaults, schedule constants and names (nets/*_at_*.py), from the reference source
    gold['net_helpers'] = []
    for ref_file, extra in [('nets/lenet_at_cifar10.py', {}), ('nets/resnet_at_cifar10.py', dict(resnet_size=20)),
                            ('nets/resnet_at_ilsvrc12.py', dict(resnet_size=50)), ('nets/mobilenet_at_ilsvrc12.py', {})]:
        fl = Flags()
        tfn = make_tf_stub(fl)
        seen = []
        st = dict(stubs4)
        st['tensorflow'] = tfn
        st['utils.lrn_rate_utils'] = blank(
            setup_lrn_rate_piecewise_constant=lambda gs, bs, idxs, rates: (seen.append((bs, list(idxs), list(rates))), 'lr')[1],
            setup_lrn_rate_exponential_decay=None)
        world_n = {'size': 1}
        st['utils.multi_gpu_wrapper'] = blank(MultiGpuWrapper=types.SimpleNamespace(size=lambda: world_n['size'], rank=lambda: 0))
        mod = load(ref_file, 'ref_helper_' + os.path.basename(ref_file)[:-3], st)
        defaults = {k: v for k, v in vars(fl).items()}
        for k, v in extra.items():
            setattr(fl, k, v)
        me = types.SimpleNamespace()
        rec = dict(file=ref_file, flag_defaults=defaults, model_name=mod.ModelHelper.model_name.fget(me),
                   dataset_name=mod.ModelHelper.dataset_name.fget(me), schedules=[])
        for multi, size, bs, nsmp, rat in [(False, 1, 128, 50000, 1.0), (True, 4, 64, 1281167, 1.0), (True, 8, 32, 1281167, 0.5)]:
            fl.enbl_multi_gpu, fl.batch_size, fl.nb_smpls_train, fl.nb_epochs_rat = multi, bs, nsmp, rat
            world_n['size'] = size
            del seen[:]
            lr, nb_iters = mod.ModelHelper.setup_lrn_rate(me, None)
            (gbs, idxs, rates), = seen
            rec['schedules'].append(dict(enbl_multi_gpu=multi, world=size, batch_size=bs, nb_smpls_train=nsmp, nb_epochs_rat=rat,
                                         global_batch=gbs, idxs_epoch=idxs, decay_rates=rates, nb_iters=int(nb_iters)))
        gold['net_helpers'].append(rec)
    # ---- the flag defaults every reference module on the path declares (tf.app.flags.DEFINE_* at import time)
    gold['flag_defaults'] = {}
    flag_files = ['learners/abstract_learner.py', 'learners/distillation_helper.py',
                  'learners/uniform_quantization/learner.py', 'learners/uniform_quantization/bit_optimizer.py',
                  'learners/nonuniform_quantization/learner.py', 'learners/nonuniform_quantization/bit_optimizer.py',
                  'learners/weight_sparsification/learner.py',
                  'learners/channel_pruning_gpu/learner.py', 'learners/full_precision/learner.py',
                  'datasets/abstract_dataset.py', 'datasets/cifar10_dataset.py', 'datasets/ilsvrc12_dataset.py',
                  'rl_agents/ddpg/agent.py', 'rl_agents/ddpg/actor_critic.py', 'rl_agents/ddpg/noise.py',
                  'rl_agents/ddpg/running_mean_std.py', 'utils/lrn_rate_utils.py']

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            if name[0].isupper():                      # something the module may subclass
                return type(name, (object,), {})
            return _Anything(name)

        def __call__(self, *a, **k):
            return _Anything('call')
    for ref_file in flag_files:
        fl = Flags()
        tff = make_tf_stub(fl)
        st = {k: _Anything(k) for k in list(stubs_d) + list(stubs4) + [
            'mpi4py', 'rl_agents.ddpg.actor_critic', 'rl_agents.ddpg.noise', 'rl_agents.ddpg.replay_buffer',
            'rl_agents.ddpg.running_mean_std', 'utils.external.imagenet_preprocessing', 'learners.full_precision',
            'learners.full_precision.learner', 'learners.channel_pruning_gpu', 'learners.weight_sparsification.pr_optimizer',
            'learners.nonuniform_quantization.rl_helper', 'learners.uniform_quantization.rl_helper', 'tensorflow.contrib']}
        tff.constant = lambda *a, **k: None
        tff.float32, tff.int32, tff.uint8 = np.float32, np.int32, np.uint8
        tff.contrib = _Anything('contrib')
        st['tensorflow'] = tff
        try:
            load(ref_file, 'ref_flags_' + ref_file.replace('/', '_')[:-3], st)
        except Exception as e:  # pylint: disable=broad-except
            print('flag_defaults: could not load', ref_file, repr(e)[:200])
            continue
        gold['flag_defaults'][ref_file] = dict(vars(fl))
    # ---- WeightSparseLearner.train (learners/weight_sparsification/learner.py:101-143): WHEN masks are rebuilt and the
    # model saved, from the reference's own loop driven with a recording session
    wsl = load('learners/weight_sparsification/learner.py', 'ref_ws_learner_loop', stubs3)
    gold['ws_train_loop'] = []
    for nb_iters, upd, beg, end, save, summ in [(1000, 100, 0.1, 0.5, 400, 250), (1000, 100, 0.0, 1.0, 10000, 100),
                                                (4000, 500, 0.1, 0.5, 1000, 100), (1000, 300, 0.25, 0.6, 10000, 100),
                                                (777, 50, 0.33, 0.34, 10000, 100), (1000, 100, 0.95, 0.99, 10000, 100)]:
        flags.ws_mask_update_step, flags.ws_iter_ratio_beg, flags.ws_iter_ratio_end = upd, beg, end
        flags.save_step, flags.summ_step, flags.enbl_multi_gpu = save, summ, False
        ev = dict(train=0, prune=[], save=[], evaluate=0, monitor=[])
        it = {'i': 0}

        def run(op, ev=ev, it=it):
            ops_ = op if isinstance(op, list) else [op]
            if 'train' in ops_:
                ev['train'] += 1
                it['i'] = ev['train']
            if 'prune' in ops_:
                assert 'init_opt' in ops_
                ev['prune'].append(it['i'])
            return [None, None, None]
        me = types.SimpleNamespace(sess_train=types.SimpleNamespace(run=run), init_op='init', bcast_op='bcast', train_op='train',
                                   summary_op='summary', log_op='log', prune_op='prune', init_opt_op='init_opt',
                                   nb_iters_train=nb_iters, is_primary_worker=lambda scope='global': True,
                                   evaluate=lambda ev=ev: ev.__setitem__('evaluate', ev['evaluate'] + 1))
        me._WeightSparseLearner__monitor_progress = lambda s, l, i, t, ev=ev: ev['monitor'].append(i + 1)
        me._WeightSparseLearner__save_model = lambda ev=ev, it=it: ev['save'].append(it['i'])
        wsl.WeightSparseLearner.train(me)
        gold['ws_train_loop'].append(dict(nb_iters_train=nb_iters, ws_mask_update_step=upd, ws_iter_ratio_beg=beg, ws_iter_ratio_end=end,
                                          save_step=save, summ_step=summ, events=ev))
    # ---- UniformQuantLearner.train (learners/uniform_quantization/learner.py:113-149): cadence of logging, saving,
    # evaluating and barriers; warm start restores first
    gold['uq_train_loop'] = []
    for steps, save, summ, warm in [(1000, 400, 250, False), (1200, 300, 100, True), (50, 10000, 100, False)]:
        flags.save_step, flags.summ_step, flags.enbl_warm_start, flags.enbl_multi_gpu = save, summ, warm, False
        ev = []
        n = {'train': 0}

        def run(op, feed_dict=None, ev=ev, n=n):
            ops_ = op if isinstance(op, list) else [op]
            if 'train' in ops_:
                n['train'] += 1
            elif ops_ == ['init']:
                ev.append('init')
            return [None, None, None]
        me = types.SimpleNamespace(sess_train=types.SimpleNamespace(run=run), finetune_steps=steps,
                                   ops=dict(init='init', bcast='bcast', train='train', summary='summary', log='log'),
                                   bit_placeholders=dict(w_train='w', a_train='a'), optimal_w_bit_list=[8], optimal_a_bit_list=[8],
                                   auto_barrier=lambda ev=ev, n=n: ev.append(('barrier', n['train'])),
                                   evaluate=lambda ev=ev, n=n: ev.append(('evaluate', n['train'])))
        me._UniformQuantLearner__restore_model = lambda is_train, ev=ev, n=n: ev.append(('restore', is_train, n['train']))
        me._UniformQuantLearner__save_model = lambda ev=ev, n=n: ev.append(('save', n['train']))
        me._UniformQuantLearner__monitor_progress = lambda s, l, t, i, ev=ev: (ev.append(('monitor', i + 1)), t)[1]
        uq.UniformQuantLearner.train(me)
        gold['uq_train_loop'].append(dict(finetune_steps=steps, save_step=save, summ_step=summ, enbl_warm_start=warm,
                                          nb_train=n['train'], events=[list(e) if isinstance(e, tuple) else e for e in ev]))
    # ---- FullPrecLearner.train / NonUniformQuantLearner.train: same cadence pins
    fpl = load('learners/full_precision/learner.py', 'ref_fp_learner_loop', stubs3)
    gold['fp_train_loop'] = []
    for steps, save, summ in [(900, 400, 250), (60, 10000, 100)]:
        flags.save_step, flags.summ_step, flags.enbl_multi_gpu = save, summ, False
        ev, n = [], {'train': 0}

        def run_fp(op, ev=ev, n=n):
            ops_ = op if isinstance(op, list) else [op]
            if 'train' in ops_:
                n['train'] += 1
            return [None, None, None]
        me = types.SimpleNamespace(sess_train=types.SimpleNamespace(run=run_fp), init_op='init', bcast_op='bcast', train_op='train',
                                   summary_op='summary', log_op='log', nb_iters_train=steps,
                                   warm_start=lambda sess, ev=ev: ev.append('warm_start'),
                                   is_primary_worker=lambda scope='global': True,
                                   evaluate=lambda ev=ev, n=n: ev.append(['evaluate', n['train']]))
        me._FullPrecLearner__monitor_progress = lambda s, l, i, t, ev=ev: ev.append(['monitor', i + 1])
        me._FullPrecLearner__save_model = lambda is_train, ev=ev, n=n: ev.append(['save', bool(is_train), n['train']])
        me._FullPrecLearner__restore_model = lambda is_train, ev=ev, n=n: ev.append(['restore', bool(is_train), n['train']])
        fpl.FullPrecLearner.train(me)
        gold['fp_train_loop'].append(dict(nb_iters_train=steps, save_step=save, summ_step=summ, nb_train=n['train'], events=ev))
    gold['nuq_train_loop'] = []
    for steps, save, summ in [(900, 400, 250), (60, 10000, 100)]:
        flags.save_step, flags.summ_step, flags.enbl_warm_start, flags.enbl_multi_gpu = save, summ, False, False
        ev, n = [], {'train': 0}

        def run_nuq(op, feed_dict=None, ev=ev, n=n):
            ops_ = op if isinstance(op, list) else [op]
            if 'train' in ops_:
                n['train'] += 1
            return [None, None, None]
        me = types.SimpleNamespace(sess_train=types.SimpleNamespace(run=run_nuq), finetune_steps=steps,
                                   ops=dict(non_cluster_init='nci', cluster_init='ci', bcast='bcast', train='train', summary='summary', log='log'),
                                   bit_placeholders=dict(w_train='w', a_train='a'), optimal_w_bit_list=[4], optimal_a_bit_list=[32],
                                   auto_barrier=lambda ev=ev, n=n: ev.append(['barrier', n['train']]),
                                   evaluate=lambda ev=ev, n=n: ev.append(['evaluate', n['train']]))
        me._NonUniformQuantLearner__save_model = lambda ev=ev, n=n: ev.append(['save', n['train']])
        me._NonUniformQuantLearner__monitor_progress = lambda s, l, t, i, ev=ev: (ev.append(['monitor', i + 1]), t)[1]
        nuq.NonUniformQuantLearner.train(me)
        gold['nuq_train_loop'].append(dict(finetune_steps=steps, save_step=save, summ_step=summ, nb_train=n['train'], events=ev))
    json.dump(gold, open(OUT, 'w'), indent=1)
    print('wrote', OUT, {k: len(v) for k, v in gold.items() if isinstance(v, list)})


if __name__ == '__main__':
    main()
