"""Diagnostic: per-op forward and per-variable gradient differences between the CUDA step and the
CPU oracle step, for several learner configurations (not a collected test — it lives under tests/ because only tests may use the oracle; prints a report)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.step_oracle import StepOracle  # noqa: E402
from pocketflow_b200.flags import FLAGS  # noqa: E402

F32 = np.float32


def make(dst, buckets, w_bits, a_bits, quant_all=False):
    FLAGS.reset()
    from pocketflow_b200.nets import resnet_at_cifar10 as R
    from pocketflow_b200.learners.uniform_quantization.learner import UniformQuantLearner
    FLAGS.resnet_size, FLAGS.batch_size = 8, 16
    FLAGS.uql_weight_bits, FLAGS.uql_activation_bits = w_bits, a_bits
    FLAGS.uql_use_buckets, FLAGS.uql_bucket_type = buckets, 'channel'
    FLAGS.enbl_dst = dst
    return UniformQuantLearner(None, R.ModelHelper())


def run(tag, exact_ste=True, **kw):
    lrn = make(**kw)
    ex = lrn.sess_train
    if not exact_ste:
        ex._ste_grads = None
    teacher = StepOracle(ex.teacher.ops, ex.teacher.logits_t, lrn.images) if ex.teacher is not None else None
    orc = StepOracle(ex.ops, ex.logits_t, lrn.images, lrn.labels, ex.loss, ex.weight_quant, ex.act_quant, teacher)
    state = ex.store.state_dict()
    tstate = ex.teacher.store.state_dict() if ex.teacher is not None else None
    images, labels = lrn.iterator_train.next_batch()
    ex.buf[lrn.images].copy_(images)
    ex.buf[lrn.labels].copy_(labels)
    ex.run_step(1e-4)
    got = ex.fetch_losses()
    ref, new_state, grads = orc.step(state, images.numpy(), labels.numpy(), dict(kind='adam', slots={}), 1e-4,
                                     teacher_state=tstate)
    print('==== %s  loss gpu %.7f ref %.7f' % (tag, got['loss'], ref['loss']))
    # forward activations
    params = {k: torch.from_numpy(v.copy()) for k, v in state.items()}
    val = orc.forward(params, images, True)
    worst = []
    for op in ex.ops:
        if op.type in ('Placeholder',):
            continue
        if op in ex.fused_act:     # pre-activation value is not materialised
            continue
        a = ex.T(op.output).cpu().numpy()
        b = val[op.output.name].detach().numpy()
        err = np.abs(a - b).max() / (np.abs(b).max() + 1e-12)
        worst.append((err, op.name, op.type))
    worst.sort(reverse=True)
    print('  fwd worst:', ['%.2e %s' % (e, n.split('/')[-2] + '/' + n.split('/')[-1]) for e, n, t in worst[:4]])
    # activation gradients (oracle: autograd with retain_grad)
    for n in params:
        if n in [v.name for v in ex.store.train_vars]:
            params[n].requires_grad_(True)
    val = orc.forward(params, images, True)
    for t in val.values():
        if t.requires_grad:
            t.retain_grad()
    lab = labels
    logits = val[ex.loss.ce[1].name]
    hard = (-(lab * torch.log_softmax(logits, dim=-1)).sum(-1)).mean()
    hard.backward()
    if ex.loss.dst is None:
        trows = []
        for op in ex.ops:
            t = op.output
            if op.type == 'Placeholder' or t not in ex.gbuf or ex.gkey(t) is not t:
                continue
            rg = val[t.name].grad
            if rg is None:
                continue
            gg = ex.gbuf[t].cpu().numpy()
            trows.append((np.abs(gg - rg.numpy()).max() / (rg.abs().max().item() + 1e-20), op.name, op.type))
        for e, n, ty in trows:
            print('   dL/d[%s %s] err %.2e' % (ty, '/'.join(n.split('/')[-2:]), e))
    rows = []
    for v in ex.store.train_vars:
        g = ex.store.view(v, ex.G).cpu().numpy()
        r = grads[v.name]
        rows.append((np.abs(g - r).max() / (np.abs(r).max() + 1e-12), v.name, float(np.abs(r).max())))
    rows.sort(reverse=True)
    for e, n, m in rows[:6]:
        print('  grad err %.3e  |g|max %.3e  %s' % (e, m, n))


if __name__ == '__main__':
    run('w8 channel only, exact ste', dst=False, buckets=True, w_bits=8, a_bits=32)
    run('w8 LAYER only', dst=False, buckets=False, w_bits=8, a_bits=32)
