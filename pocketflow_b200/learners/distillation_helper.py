"""Helper for training with distillation loss (/root/reference/learners/distillation_helper.py:28-158)."""
import os

from .. import graph as G
from ..flags import FLAGS, DEFINE_float, DEFINE_string
from .abstract_learner import latest_checkpoint, load_checkpoint

DEFINE_float('loss_w_dst', 4.0, 'distillation loss\'s multiplier')
DEFINE_float('tempr_dst', 4.0, 'temperature in the distillation loss')
DEFINE_string('save_path_dst', './models_dst/model.ckpt', 'distillation model\'s save path')


class DistillationHelper(object):
    """Other learners use calc_logits() to build the (stop-gradient, eval-mode) teacher forward pass
    under scope 'distilled_model' and calc_loss() for the soft-label cross-entropy."""

    def __init__(self, sm_writer, model_helper, mpi_comm):
        self.model_scope = 'distilled_model'   # to distinguish from models created by other learners
        self.model_helper = model_helper
        self.mpi_comm = mpi_comm
        self.ckpt = None
        ckpt_dir = os.path.dirname(FLAGS.save_path_dst)
        fn = latest_checkpoint(ckpt_dir) if os.path.isdir(ckpt_dir) else None
        if fn is not None:
            self.ckpt = load_checkpoint(fn)
        elif FLAGS.data_dir_local:
            # real data and no teacher checkpoint: distilling from a random teacher is never what was asked for
            raise ValueError('--enbl_dst with real data needs a pre-trained teacher checkpoint in ' + ckpt_dir)
        # The reference downloads a pre-trained checkpoint here; on the synthetic benchmark path the
        # teacher keeps its own random initialisation (seed 2) — distribution, not accuracy, matters.

    def calc_logits(self, sess, images):
        """Teacher logits for `images` (N x K): a new eval-mode forward path under
        'distilled_model'; gradients never flow into it (the executor runs it forward-only)."""
        with G.variable_scope(self.model_scope):
            logits = self.model_helper.forward_eval(images)
        return logits

    def restore(self, store):
        """Initialise the teacher's weights from save_path_dst when a checkpoint exists (scope renamed
        model/ -> distilled_model/, distillation_helper.py:105-145)."""
        if self.ckpt is None:
            return False
        renamed = {self.model_scope + '/' + '/'.join(k.split('/')[1:]): v for k, v in self.ckpt.items()}
        found, total = store.load_state_dict(renamed, strict=False, require='all')
        print('distillation teacher restored from %s (%d of %d trainable variables)'
              % (os.path.dirname(FLAGS.save_path_dst), found, total))
        return True

    @classmethod
    def calc_loss(cls, logits_pri, logits_dst):
        """loss_w_dst * softmax_cross_entropy(softmax(t/T), s/T) (distillation_helper.py:86-103)."""
        return G.distillation_cross_entropy(logits_pri, logits_dst, FLAGS.loss_w_dst, FLAGS.tempr_dst)
