"""Utility functions for learning rates (/root/reference/utils/lrn_rate_utils.py:23-70).
The schedules are host-side scalars here: they return callables step -> learning rate."""
from ..flags import FLAGS


def piecewise_constant(bnds, vals):
    """tf.train.piecewise_constant: vals[i] for bnds[i-1] < step <= bnds[i]."""
    def fn(step):
        for b, v in zip(bnds, vals):
            if step <= b:
                return v
        return vals[-1]
    return fn


def setup_lrn_rate_piecewise_constant(global_step, batch_size, idxs_epoch, decay_rates):
    idxs_epoch = [idx_epoch * FLAGS.nb_epochs_rat for idx_epoch in idxs_epoch]
    lrn_rate_init = FLAGS.lrn_rate_init * batch_size / FLAGS.batch_size_norm
    nb_batches_per_epoch = float(FLAGS.nb_smpls_train) / batch_size
    bnds = [int(nb_batches_per_epoch * idx_epoch) for idx_epoch in idxs_epoch]
    vals = [lrn_rate_init * decay_rate for decay_rate in decay_rates]
    return piecewise_constant(bnds, vals)


def setup_lrn_rate_exponential_decay(global_step, batch_size, epoch_step, decay_rate):
    epoch_step *= FLAGS.nb_epochs_rat
    lrn_rate_init = FLAGS.lrn_rate_init * batch_size / FLAGS.batch_size_norm
    batch_step = int(FLAGS.nb_smpls_train * epoch_step / batch_size)
    return lambda step: lrn_rate_init * decay_rate ** (int(step) // batch_step)
