#!/usr/bin/env python
"""GPU smoke of the RL bit search through the real learner: ResNet-8 / synthetic CIFAR-10, a handful of roll-outs with a
few fine-tuning steps each.  Checks what the CPU tests cannot: that per-roll-out bit-widths reach the kernels
(quantized weights take at most 2^bits values per bucket), that restoring between roll-outs really restores, and that
the search leaves its best allocation in the executor."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import importlib  # noqa: E402

from pocketflow_b200.flags import FLAGS  # noqa: E402
from pocketflow_b200.nets import resnet_at_cifar10 as R  # noqa: E402


def _cifar_defaults():
    """Dataset / net modules declare their flag defaults at import: re-declare CIFAR-10's (another dataset module may
    have been imported since, e.g. by an earlier test in the same process)."""
    global R
    FLAGS.reset()
    import pocketflow_b200.datasets.cifar10_dataset as D
    importlib.reload(D)
    R = importlib.reload(R)

from pocketflow_b200.learners.uniform_quantization.learner import UniformQuantLearner  # noqa: E402


def main():
    import torch
    torch.manual_seed(0)            # the agent's exploration draws from torch's global generator unless it is seeded
    np.random.seed(0)
    _cifar_defaults()
    FLAGS.resnet_size, FLAGS.batch_size, FLAGS.batch_size_eval, FLAGS.nb_smpls_eval = 8, 32, 32, 64
    FLAGS.uql_enbl_rl_agent, FLAGS.uql_nb_rlouts, FLAGS.uql_equivalent_bits = True, 6, 5
    FLAGS.uql_tune_global_steps, FLAGS.uql_tune_disp_steps = 8, 4
    FLAGS.uql_use_buckets, FLAGS.uql_bucket_type = False, 'channel'
    lrn = UniformQuantLearner(None, R.ModelHelper())
    ex = lrn.sess_train
    bits = lrn.optimal_w_bit_list
    nums = lrn.statistics['num_weights']
    print('allocation:', bits, 'budget used %.3f' % (sum(b * n for b, n in zip(bits, nums)) / (5.0 * sum(nums))))
    assert ex.wq.bits == [int(b) for b in bits] and ex.step_count == 0
    assert sum(b * n for b, n in zip(bits, nums)) <= 5 * sum(nums)
    lrn.train_step()
    for op, b in zip(ex.wq_ops, bits):                   # per-layer quantization: at most 2^bits distinct values
        q = ex.store.view(op.vars['kernel'], ex.QW).detach().cpu().numpy()
        assert len(np.unique(q)) <= 2 ** int(b), (op.name, b, len(np.unique(q)))
    before = ex.store.state_dict()
    lrn.rl_restore()
    after = ex.store.state_dict()
    changed = sum(not np.array_equal(before[k], after[k]) for k in before)
    assert changed > 0 and ex.step_count == 0, 'rl_restore left the trained weights in place'
    print('rl smoke ok: %d tensors restored' % changed)


def ws_main():
    """The pruning-ratio search through the real WeightSparseLearner: 4 roll-outs, 6 fine-tuning steps each."""
    from pocketflow_b200.learners.weight_sparsification.learner import WeightSparseLearner, calc_prune_ratio
    import torch
    torch.manual_seed(0)
    np.random.seed(0)
    _cifar_defaults()
    FLAGS.resnet_size, FLAGS.batch_size, FLAGS.batch_size_eval = 8, 32, 32
    FLAGS.ws_prune_ratio, FLAGS.ws_prune_ratio_prtl = 0.5, 'optimal'
    FLAGS.ws_nb_rlouts, FLAGS.ws_nb_rlouts_min, FLAGS.ws_nb_iters_ft, FLAGS.ws_nb_iters_feval = 4, 2, 6, 2
    lrn = WeightSparseLearner(None, R.ModelHelper())
    ex = lrn.sess_train
    ratios = [r for _, r in lrn.var_names_n_prune_ratios]
    nums = [v.numel for v in lrn.maskable_vars]
    overall = sum(r * n for r, n in zip(ratios, nums)) / sum(nums)
    print('ratios:', ['%.2f' % r for r in ratios], 'overall %.3f' % overall)
    assert ratios[0] == 0.0 and ratios[-1] == 0.0 and overall >= 0.5 - 1e-6            # CIFAR-10: head / tail kept
    assert float(ex.MASK.min()) == 1.0 and ex.step_count == 0                          # the search left no trace
    assert calc_prune_ratio([ex.store.view(v) for v in lrn.maskable_vars]) < 0.01
    lrn.pr_prune(ratios)
    got = calc_prune_ratio([ex.store.view(v) for v in lrn.maskable_vars])
    assert abs(got - overall) < 0.02, (got, overall)
    print('ws smoke ok: pruned model at %.3f' % got)


if __name__ == '__main__':
    main()
    ws_main()
