"""Model helper for creating a MobileNet model for the ILSVRC-12 dataset
(/root/reference/nets/mobilenet_at_ilsvrc12.py:29-160); version 1 only (v2 is out of scope)."""
from .. import graph as G
from ..flags import FLAGS, DEFINE_integer, DEFINE_float
from ..datasets.ilsvrc12_dataset import Ilsvrc12Dataset
from ..utils.lrn_rate_utils import setup_lrn_rate_piecewise_constant
from ..utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from .abstract_model_helper import AbstractModelHelper
from . import mobilenet_v1 as MobileNetV1

DEFINE_integer('mobilenet_version', 1, 'MobileNet\'s version (1 or 2)')
DEFINE_float('mobilenet_depth_mult', 1.0, 'MobileNet\'s depth multiplier')
DEFINE_float('nb_epochs_rat', 1.0, '# of training epochs\'s ratio')
DEFINE_float('lrn_rate_init', 0.045, 'initial learning rate')
DEFINE_float('batch_size_norm', 96, 'normalization factor of batch size')
DEFINE_float('momentum', 0.9, 'momentum coefficient')
DEFINE_float('loss_w_dcy', 4e-5, 'weight decaying loss\'s coefficient')


def forward_fn(inputs, is_train):
    if FLAGS.mobilenet_version != 1:
        raise ValueError('invalid MobileNet version: {} (only v1 is built)'.format(FLAGS.mobilenet_version))
    return MobileNetV1.mobilenet_v1(inputs, num_classes=FLAGS.nb_classes, is_training=is_train,
                                    depth_multiplier=FLAGS.mobilenet_depth_mult)


class ModelHelper(AbstractModelHelper):
    def __init__(self, data_format='channels_last'):
        assert data_format == 'channels_last', 'MobileNet only supports \'channels_last\' data format'
        super(ModelHelper, self).__init__(data_format)
        self.dataset_train = Ilsvrc12Dataset(is_train=True)
        self.dataset_eval = Ilsvrc12Dataset(is_train=False)

    def build_dataset_train(self, enbl_trn_val_split=False):
        return self.dataset_train.build(enbl_trn_val_split)

    def build_dataset_eval(self):
        return self.dataset_eval.build()

    def forward_train(self, inputs):
        return forward_fn(inputs, is_train=True)

    def forward_eval(self, inputs):
        return forward_fn(inputs, is_train=False)

    def calc_loss(self, labels, outputs, trainable_vars):
        loss = G.softmax_cross_entropy(labels, outputs)
        # slim names its batch-norm scope 'BatchNorm': this filter does NOT exclude gamma/beta, so they are
        # regularised — mirrored (mobilenet_at_ilsvrc12.py:107-109)
        loss_filter = lambda var: 'batch_normalization' not in var.name
        loss += FLAGS.loss_w_dcy * G.add_n([G.l2_loss(var) for var in trainable_vars if loss_filter(var)])
        acc_top1, acc_top5 = G.accuracy(labels, outputs), G.in_top_k_accuracy(labels, outputs, 5)
        metrics = {'accuracy': acc_top5, 'acc_top1': acc_top1, 'acc_top5': acc_top5}
        return loss, metrics

    def setup_lrn_rate(self, global_step):
        batch_size = FLAGS.batch_size * (1 if not FLAGS.enbl_multi_gpu else mgw.size())
        nb_epochs = 100
        idxs_epoch = [30, 60, 80, 90]
        decay_rates = [1.0, 0.1, 0.01, 0.001, 0.0001]
        lrn_rate = setup_lrn_rate_piecewise_constant(global_step, batch_size, idxs_epoch, decay_rates)
        nb_iters = int(FLAGS.nb_smpls_train * nb_epochs * FLAGS.nb_epochs_rat / batch_size)
        return lrn_rate, nb_iters

    @property
    def model_name(self):
        return 'mobilenet_v%d' % FLAGS.mobilenet_version

    @property
    def dataset_name(self):
        return 'ilsvrc_12'
