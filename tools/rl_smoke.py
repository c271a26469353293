#!/usr/bin/env python
"""GPU smoke of the RL bit search through the real learner: ResNet-8 / synthetic CIFAR-10, a handful of roll-outs with a
few fine-tuning steps each.  Checks what the CPU tests cannot: that per-roll-out bit-widths reach the kernels
(quantized weights take at most 2^bits values per bucket), that restoring between roll-outs really restores, and that
the search leaves its best allocation in the executor."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pocketflow_b200.flags import FLAGS  # noqa: E402
from pocketflow_b200.nets import resnet_at_cifar10 as R  # noqa: E402
from pocketflow_b200.learners.uniform_quantization.learner import UniformQuantLearner  # noqa: E402


def main():
    FLAGS.reset()
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


if __name__ == '__main__':
    main()
