"""Golden values produced by the REFERENCE'S OWN host-side code, run in this container from /root/reference under a
stub `tensorflow` module (TensorFlow 1.x itself cannot be imported here).  Only pure-Python / numpy functions of the
path can run this way — the learning-rate / fine-tuning schedule helpers and the 'heurist' pruning-ratio formula; the
tensor arithmetic (fake-quant, masks, losses) lives inside TensorFlow ops and stays pinned by hand-derived KATs only.

  python tests/golden/make_golden_from_reference.py        ->  tests/golden/ref_host_schedules_v1.json

The stub provides exactly what the imported reference modules touch at import / call time: tf.app.flags (DEFINE_* +
FLAGS), tf.logging.info, tf.train.piecewise_constant (records its arguments), tf.shape (identity on .shape)."""
import importlib.util
import json
import os
import sys
import types

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_host_schedules_v1.json')


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
    from tests.golden.graphs_for_golden import op_lists
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
    json.dump(gold, open(OUT, 'w'), indent=1)
    print('wrote', OUT, {k: len(v) for k, v in gold.items() if isinstance(v, list)})


if __name__ == '__main__':
    main()
