"""Channel pruning learner, GPU variant (/root/reference/learners/channel_pruning_gpu/learner.py:30-568).

Built here: the steady-state masked whole-network step (reference :404-443 — `g * mask` on every
Conv2D kernel + Momentum, identical to the weight-sparse step but with INPUT-CHANNEL masks,
:250-260) through pf_momentum_step.
Not built yet ("next"): the layer-wise selection phase (group-lasso proximal gradient
descent on the layer-output regression loss, :339-402, :445-518).  Until it is, channels are ranked by
the same statistic the reference's prox op thresholds — the L2 norm of the kernel over axes [0,1,3]
(:379-383) — and the lowest `cpg_prune_ratio` fraction of input channels of every interior layer is
masked; this is a flagged simplification of the SELECTION only, the per-step arithmetic is the
reference's."""
from timeit import default_timer as timer

import numpy as np
import torch

from ... import graph as G
from ...engine import Executor
from ...flags import FLAGS, DEFINE_string, DEFINE_float, DEFINE_boolean, DEFINE_integer
from ...utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from ..abstract_learner import AbstractLearner, save_checkpoint
from ..distillation_helper import DistillationHelper

DEFINE_string('cpg_save_path', './models_cpg/model.ckpt', 'CPG: model\'s save path')
DEFINE_string('cpg_save_path_eval', './models_cpg_eval/model.ckpt', 'CPG: model\'s save path for evaluation')
DEFINE_string('cpg_prune_ratio_type', 'uniform', 'CPG: pruning ratio type (\'uniform\' OR \'list\')')
DEFINE_float('cpg_prune_ratio', 0.5, 'CPG: uniform pruning ratio')
DEFINE_boolean('cpg_skip_ht_layers', True, 'CPG: skip head & tail layers for pruning')
DEFINE_string('cpg_prune_ratio_file', None, 'CPG: file path to the list of pruning ratios')
DEFINE_float('cpg_lrn_rate_pgd_init', 1e-10, 'CPG: proximal gradient descent\'s initial learning rate')
DEFINE_float('cpg_lrn_rate_pgd_incr', 1.4, 'CPG: proximal gradient descent\'s learning rate\'s increase ratio')
DEFINE_float('cpg_lrn_rate_pgd_decr', 0.7, 'CPG: proximal gradient descent\'s learning rate\'s decrease ratio')
DEFINE_float('cpg_lrn_rate_adam', 1e-2, 'CPG: Adam\'s initial learning rate')
DEFINE_integer('cpg_nb_iters_layer', 1000, 'CPG: # of iterations for layer-wise FT')


def calc_prune_ratio(tensors):
    nnz = sum(int(torch.count_nonzero(t).item()) for t in tensors)
    tot = sum(t.numel() for t in tensors)
    return np.float32(np.float32(1.0) - np.float32(nnz) / np.float32(tot))


class ChannelPrunedGpuLearner(AbstractLearner):  # pylint: disable=too-many-instance-attributes
    def __init__(self, sm_writer, model_helper):
        super(ChannelPrunedGpuLearner, self).__init__(sm_writer, model_helper)
        if FLAGS.enbl_dst:
            self.helper_dst = DistillationHelper(sm_writer, model_helper, self.mpi_comm)
        self.__build()
        self.__choose_channels()

    def train(self, nb_iters=None):
        ex = self.sess_train
        if FLAGS.enbl_multi_gpu:
            mgw.broadcast_global_variables([ex.store.P, ex.store.O])
        time_prev = timer()
        total = self.nb_iters_train if nb_iters is None else nb_iters
        for idx_iter in range(total):
            self.train_step()
            if (idx_iter + 1) % FLAGS.summ_step == 0 and self.is_primary_worker('global'):
                r = ex.fetch_losses()
                speed = FLAGS.batch_size * FLAGS.summ_step / (timer() - time_prev) * (mgw.size() if FLAGS.enbl_multi_gpu else 1)
                print('iter #%d: lr = %.4e | loss = %.4e | pr_msk = %.4e | speed = %.2f pics / sec'
                      % (idx_iter + 1, self.lrn_rate(idx_iter), r['loss'], self.pr_maskable(), speed))
                time_prev = timer()
            # save the model at certain steps (learner.py:171-175).  The reference barriers after EVERY iteration; the
            # gradient all-reduce already keeps the ranks in step, so only the iterations where the primary worker
            # does extra work need one (a per-step NCCL barrier would drain the device queue every step).
            if (idx_iter + 1) % FLAGS.save_step == 0:
                if self.is_primary_worker('global'):
                    self.__save_model()
                    self.evaluate()
                self.auto_barrier()
        if self.is_primary_worker('global'):
            self.__save_model()
            self.evaluate()

    def __save_model(self):
        ex = self.sess_train
        print('model saved to ' + save_checkpoint(FLAGS.cpg_save_path, ex.store.state_dict(), ex.step_count))

    def train_step(self):
        ex = self.sess_train
        self.h2d_bytes = self.feed(ex, self.iterator_train)
        ex.run_step(self.lrn_rate(ex.step_count), self.grad_allreduce())

    def evaluate(self, nb_iters=None):
        self.restore_for_eval(FLAGS.cpg_save_path)
        ex = self.sess_train
        out = []
        for _ in range(self.eval_nb_iters(nb_iters)):
            self.feed(ex, self.eval_iterator())
            ex.forward_eval_loss()
            out.append(ex.fetch_losses()['loss'])
        return float(np.mean(out)), float(self.pr_maskable())

    def pr_maskable(self):
        return calc_prune_ratio([self.sess_train.store.view(v) for v in self.maskable_vars])

    def __build(self):
        self.graph_train = G.Graph()
        with self.graph_train.as_default():
            with G.variable_scope(self.data_scope):
                self.iterator_train = self.build_dataset_train()
                images, labels = self.iterator_train.get_next()
            self.images, self.labels = images, labels
            logits_dst = self.helper_dst.calc_logits(None, images) if FLAGS.enbl_dst else None
            with G.variable_scope(self.model_scope):
                logits = self.forward_train(images)
                loss, metrics = self.calc_loss(labels, logits, self.trainable_vars)
                if FLAGS.enbl_dst:
                    loss += self.helper_dst.calc_loss(logits, logits_dst)
                self.lrn_rate, self.nb_iters_train = self.setup_lrn_rate(None)
        # maskable = trainable variables read by ops named .../Conv2D (depthwise excluded) (:52-66)
        conv_ops = [op for op in self.graph_train.ops if op.name.endswith('/Conv2D')
                    and op.name.startswith(self.model_scope + '/')]
        self.maskable_vars = [op.vars['kernel'] for op in conv_ops]
        self.maskable_var_names = [v.name for v in self.maskable_vars]
        world = mgw.size() if FLAGS.enbl_multi_gpu else 1
        teacher = None
        if FLAGS.enbl_dst:
            teacher = Executor(self.graph_train, images, logits_dst, self.device, train=False, seed=2)
            self.helper_dst.restore(teacher.store)
        self.sess_train = Executor(self.graph_train, images, logits, self.device, train=True, loss=loss, labels=labels,
                                   optimizer=dict(kind='momentum', momentum=FLAGS.momentum),
                                   maskable=self.maskable_vars, teacher=teacher, seed=1, grad_scale=1.0 / world)
        if teacher is not None:
            teacher.buf[images] = self.sess_train.buf[images]
            self.sess_train.share_im2col_from(teacher)

    def __choose_channels(self):
        """Input-channel masks (:250-260): mask[:, :, c, :] = 0 for pruned input channels; the variable is
        zeroed accordingly and every later step multiplies the gradient by the mask."""
        ex = self.sess_train
        nb = len(self.maskable_vars)
        if FLAGS.cpg_prune_ratio_type == 'uniform':
            ratios = [FLAGS.cpg_prune_ratio] * nb
            if FLAGS.cpg_skip_ht_layers:
                ratios[0] = ratios[-1] = 0.0
        elif FLAGS.cpg_prune_ratio_type == 'list':
            ratios = list(np.loadtxt(FLAGS.cpg_prune_ratio_file, delimiter=','))
            assert len(ratios) == nb
        else:
            raise ValueError('unrecognized pruning ratio type: ' + FLAGS.cpg_prune_ratio_type)
        self.prune_ratios = ratios
        for v, ratio in zip(self.maskable_vars, ratios):
            w = ex.store.view(v)
            mask = ex.store.view(v, ex.MASK)
            cin = v.shape[2] if len(v.shape) == 4 else v.shape[0]
            nb_prune = int(round(cin * ratio))
            if nb_prune == 0:
                continue
            w4 = w if w.dim() == 4 else w.view(1, 1, *w.shape)
            norm = torch.sqrt((w4 * w4).sum(dim=(0, 1, 3)))            # var_norm of the prox op (:379)
            idx = torch.argsort(norm, stable=True)[:nb_prune]
            m4 = mask if mask.dim() == 4 else mask.view(1, 1, *mask.shape)
            m4[:, :, idx, :] = 0.0
            w.mul_(mask)
