"""What every image-classification ModelHelper of the reference repeats (nets/*_at_*.py): two dataset objects, a
train / eval forward that differ only in a flag, softmax cross-entropy plus an L2 term over (a filtered subset of) the
trainable variables, and a piecewise-constant learning-rate schedule that sees the GLOBAL batch size.  The concrete
helpers state only what differs: the dataset, the network, which variables the L2 term skips, the metrics, the
schedule constants and the model's name."""
from .. import graph as G
from ..flags import FLAGS
from ..utils.lrn_rate_utils import setup_lrn_rate_piecewise_constant
from ..utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from .abstract_model_helper import AbstractModelHelper


class ClassificationModelHelper(AbstractModelHelper):
    DATASET = None                  # dataset class (constructed once for training, once for evaluation)
    DATASET_NAME = None
    NB_EPOCHS = None                # schedule: total epochs, epochs at which the rate drops, rate multipliers
    IDXS_EPOCH = None
    DECAY_RATES = None
    L2_SKIPS = None                 # variables whose name contains this substring stay out of the L2 term
    TOP5 = False                    # report acc_top1 / acc_top5 instead of a single 'accuracy'

    def __init__(self, data_format='channels_last'):
        super(ClassificationModelHelper, self).__init__(data_format)
        self.dataset_train = self.DATASET(is_train=True)
        self.dataset_eval = self.DATASET(is_train=False)

    # ---- data
    def build_dataset_train(self, enbl_trn_val_split=False):
        return self.dataset_train.build(enbl_trn_val_split)

    def build_dataset_eval(self):
        return self.dataset_eval.build()

    # ---- network
    def network(self, inputs, is_train):
        raise NotImplementedError

    def forward_train(self, inputs):
        return self.network(inputs, True)

    def forward_eval(self, inputs):
        return self.network(inputs, False)

    # ---- objective
    def regularised(self, trainable_vars):
        if self.L2_SKIPS is None:
            return list(trainable_vars)
        return [v for v in trainable_vars if self.L2_SKIPS not in v.name]

    def metrics(self, labels, outputs):
        if not self.TOP5:
            return {'accuracy': G.accuracy(labels, outputs)}
        return {'acc_top1': G.accuracy(labels, outputs), 'acc_top5': G.in_top_k_accuracy(labels, outputs, 5)}

    def calc_loss(self, labels, outputs, trainable_vars):
        loss = G.softmax_cross_entropy(labels, outputs)
        loss += FLAGS.loss_w_dcy * G.add_n([G.l2_loss(v) for v in self.regularised(trainable_vars)])
        return loss, self.metrics(labels, outputs)

    # ---- schedule
    def setup_lrn_rate(self, global_step):
        world = mgw.size() if FLAGS.enbl_multi_gpu else 1
        batch_size = FLAGS.batch_size * world
        lrn_rate = setup_lrn_rate_piecewise_constant(global_step, batch_size, self.IDXS_EPOCH, self.DECAY_RATES)
        nb_iters = int(FLAGS.nb_smpls_train * self.NB_EPOCHS * FLAGS.nb_epochs_rat / batch_size)
        return lrn_rate, nb_iters

    @property
    def dataset_name(self):
        return self.DATASET_NAME
