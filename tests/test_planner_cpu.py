"""Planner invariants of engine.Executor, checked symbolically on the CPU (no kernels run):

* gradient-buffer sharing through residual Adds (in-place accumulation through identity shortcuts): when an op's
  backward runs, the buffer holding the gradient of its output contains exactly the contributions of that output's
  consumers — no more (a foreign writer ran too early), no fewer (a reader ran too early);
* dy operand planes: the BatchNorm chosen to emit a conv's dy planes is the LAST writer of that gradient before the
  conv's backward, and what it leaves in the buffer is the conv output's complete gradient;
* x operand planes: a BN output's fp32 copy is dropped only when every consumer reads planes.
"""
import pytest
import torch

from pocketflow_b200 import graph as G
from pocketflow_b200.engine import Executor
from pocketflow_b200.flags import FLAGS


def build(net, **flags):
    FLAGS.reset()
    import importlib
    mod = importlib.import_module('pocketflow_b200.nets.' + net)
    for k, v in flags.items():
        setattr(FLAGS, k, v)
    mh = mod.ModelHelper()
    g = G.Graph()
    with g.as_default():
        with G.variable_scope('data'):
            it = mh.build_dataset_train()
            im, lab = it.get_next()
        with G.variable_scope('model'):
            out = mh.forward_train(im)
            tv = [v for v in g.variables.values() if v.name.startswith('model/') and v.trainable]
            loss, _ = mh.calc_loss(lab, out, tv)
    return Executor(g, im, out, torch.device('cpu'), train=True, loss=loss, labels=lab,
                    optimizer=dict(kind='momentum', momentum=0.9))


def grad_inputs(ex, op):
    """Tensors whose gradient buffer op's backward WRITES (mirrors Executor.loss_and_backward)."""
    if op.type in ('Placeholder', 'Reshape', 'Identity') or op in ex.fused_into:
        return []
    ins = op.inputs if op.type == 'Add' else op.inputs[:1]
    out = []
    for t in ins:
        if t.op.type == 'Placeholder':
            continue
        if op.type == 'Add' and ex.gkey(t) is ex.gkey(op.output):
            continue                                   # shared buffer: the Add's backward is a no-op for this input
        out.append(t)
    return out


def expected_contributions(ex, t, memo):
    """The set of WRITER ops whose contributions make up dL/dt."""
    if t in memo:
        return memo[t]
    s = set()
    for c in ex._consumers(t):
        if c.type in ('Reshape', 'Identity') or c in ex.fused_into:
            s |= expected_contributions(ex, c.output, memo)              # pass-through: same gradient
        elif c.type == 'Add' and ex.gkey(t) is ex.gkey(c.output):
            s |= expected_contributions(ex, c.output, memo)              # identity: shares the Add output's gradient
        else:
            s.add(c)
    if t is ex.loss.ce[1] or (t in ex.alias and False):
        s.add('loss')
    memo[t] = s
    return s


@pytest.mark.parametrize('net,flags', [('resnet_at_cifar10', dict(resnet_size=20, batch_size=4)),
                                       ('resnet_at_ilsvrc12', dict(resnet_size=50, batch_size=2)),
                                       ('resnet_at_ilsvrc12', dict(resnet_size=18, batch_size=2)),
                                       ('mobilenet_at_ilsvrc12', dict(batch_size=2, nb_classes=1001)),
                                       ('lenet_at_cifar10', dict(batch_size=4))])
def test_gradient_buffers_hold_exactly_the_consumers_contributions(net, flags):
    ex = build(net, **flags)
    memo, state = {}, {}
    state[ex.gkey(ex.loss.ce[1])] = {'loss'}                             # softmax-CE writes dL/dlogits first
    ran = []
    dy_plane_state = {}
    for op in reversed(ex.ops):
        if op.type == 'Placeholder':
            continue
        k = ex.gkey(op.output)
        if k not in state:
            continue                                                     # no gradient flows here
        if not (op.type in ('Reshape', 'Identity') or op in ex.fused_into):
            want = expected_contributions(ex, op.output, memo)
            assert state[k] == want, '%s: buffer holds %s, expected %s' % (
                op.name, sorted(getattr(o, 'name', o) for o in state[k]), sorted(getattr(o, 'name', o) for o in want))
            if op in getattr(ex, 'conv_dy_planes', {}):
                # the planes were emitted by a BN backward that ran earlier: they must hold this same complete gradient
                bn = [b for b, pl in ex.bn_gplanes.items() if pl is ex.conv_dy_planes[op]][0]
                assert dy_plane_state[bn] == want, op.name
        for t in grad_inputs(ex, op):
            kk = ex.gkey(t)
            state[kk] = (state[kk] | {op}) if kk in state else {op}     # accumulate / first write
            if op.type == 'FusedBatchNorm' and op in getattr(ex, 'bn_gplanes', {}):
                dy_plane_state[op] = set(state[kk])                      # what the emitted planes contain
        ran.append(op)
    # every residual Add shares its buffer with all of its inputs (no copy kernels) in the ResNets
    adds = [op for op in ex.ops if op.type == 'Add']
    for a in adds:
        assert all(ex.gkey(t) is ex.gkey(a.output) for t in a.inputs), a.name


@pytest.mark.parametrize('net,flags', [('resnet_at_cifar10', dict(resnet_size=20, batch_size=4)),
                                       ('resnet_at_ilsvrc12', dict(resnet_size=50, batch_size=2)),
                                       ('mobilenet_at_ilsvrc12', dict(batch_size=2, nb_classes=1001))])
def test_fp32_copy_of_a_bn_output_is_dropped_only_when_every_consumer_reads_planes(net, flags):
    ex = build(net, **flags)
    assert ex.xplanes, 'no operand planes planned'
    for bn_op, planes in ex.xplanes.items():
        outs = [bn_op.output] + [c.output for c in ex._consumers(bn_op.output) if c in ex.fused_into]
        consumers = [c for t in outs for c in ex._consumers(t) if not (c in ex.fused_into and ex.fused_into[c] is bn_op)]
        all_planes = all(c in ex.tc_wgrad and c not in ex.im2col for c in consumers)
        assert ex.bn_need_f32[bn_op] == (not all_planes), bn_op.name
        assert planes.numel == bn_op.output.numel
    n_dy = len(ex.conv_dy_planes)
    assert n_dy > 0 and all(op in ex.tc_wgrad for op in ex.conv_dy_planes)


def test_unsafe_identity_sharing_is_refused():
    """A tensor that is consumed again AFTER the residual Add (so one of its gradient writers would run BEFORE the
    readers of the Add's gradient) must keep its own gradient buffer; the symbolic check still holds."""
    FLAGS.reset()
    import pocketflow_b200.datasets.cifar10_dataset  # noqa: F401  (flags)
    g = G.Graph()
    with g.as_default():
        im = G.placeholder((2, 8, 8, 64), 'images')
        lab = G.placeholder((2, 10), 'labels')
        with G.variable_scope('model'):
            a = G.conv2d(im, 64, 3, padding='same', use_bias=False)
            b = G.conv2d(G.relu(G.batch_normalization(a, training=True)), 64, 3, padding='same', use_bias=False)
            s = G.add(b, a)                                   # `a` feeds the BN, the Add ...
            c = G.conv2d(a, 64, 1, use_bias=False)            # ... and is consumed AGAIN later in forward order
            t = G.add(s, c)
            out = G.dense(G.reduce_mean_hw(G.relu(G.batch_normalization(t, training=True))), 10)
            loss = G.softmax_cross_entropy(lab, out)
    ex = Executor(g, im, out, torch.device('cpu'), train=True, loss=loss, labels=lab,
                  optimizer=dict(kind='momentum', momentum=0.9))
    add_s = s.op
    a_root = ex._root(a)
    assert ex.gkey(a_root) is not ex.gkey(add_s.output), 'unsafe in-place sharing was planned'
    assert ex.gkey(b) is ex.gkey(add_s.output)                # the single-consumer branch still shares
    # and the general invariant holds for this graph too
    memo, state = {}, {ex.gkey(ex.loss.ce[1]): {'loss'}}
    for op in reversed(ex.ops):
        if op.type == 'Placeholder' or ex.gkey(op.output) not in state:
            continue
        if not (op.type in ('Reshape', 'Identity') or op in ex.fused_into):
            assert state[ex.gkey(op.output)] == expected_contributions(ex, op.output, memo), op.name
        for x in grad_inputs(ex, op):
            kk = ex.gkey(x)
            state[kk] = (state[kk] | {op}) if kk in state else {op}
