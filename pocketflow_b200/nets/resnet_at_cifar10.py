"""Model helper for creating a ResNet model for the CIFAR-10 dataset
(/root/reference/nets/resnet_at_cifar10.py:29-135)."""
from .. import graph as G
from ..flags import FLAGS, DEFINE_integer, DEFINE_float
from ..datasets.cifar10_dataset import Cifar10Dataset
from ..utils.lrn_rate_utils import setup_lrn_rate_piecewise_constant
from ..utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from .abstract_model_helper import AbstractModelHelper
from . import resnet_model as ResNet

DEFINE_integer('resnet_size', 20, '# of layers in the ResNet model')
DEFINE_float('nb_epochs_rat', 1.0, '# of training epochs\'s ratio')
DEFINE_float('lrn_rate_init', 1e-1, 'initial learning rate')
DEFINE_float('batch_size_norm', 128, 'normalization factor of batch size')
DEFINE_float('momentum', 0.9, 'momentum coefficient')
DEFINE_float('loss_w_dcy', 2e-4, 'weight decaying loss\'s coefficient')


def forward_fn(inputs, is_train, data_format):
    nb_blocks = (FLAGS.resnet_size - 2) // 6
    model = ResNet.Model(FLAGS.resnet_size, False, FLAGS.nb_classes, 16, 3, 1, None, None,
                         [nb_blocks] * 3, [1, 2, 2], data_format=data_format)
    return model(inputs, is_train)


class ModelHelper(AbstractModelHelper):
    def __init__(self, data_format='channels_last'):
        super(ModelHelper, self).__init__(data_format)
        self.dataset_train = Cifar10Dataset(is_train=True)
        self.dataset_eval = Cifar10Dataset(is_train=False)

    def build_dataset_train(self, enbl_trn_val_split=False):
        return self.dataset_train.build(enbl_trn_val_split)

    def build_dataset_eval(self):
        return self.dataset_eval.build()

    def forward_train(self, inputs):
        return forward_fn(inputs, is_train=True, data_format=self.data_format)

    def forward_eval(self, inputs):
        return forward_fn(inputs, is_train=False, data_format=self.data_format)

    def calc_loss(self, labels, outputs, trainable_vars):
        loss = G.softmax_cross_entropy(labels, outputs)
        loss_filter = lambda var: 'batch_normalization' not in var.name
        loss += FLAGS.loss_w_dcy * G.add_n([G.l2_loss(var) for var in trainable_vars if loss_filter(var)])
        metrics = {'accuracy': G.accuracy(labels, outputs)}
        return loss, metrics

    def setup_lrn_rate(self, global_step):
        nb_epochs = 250
        idxs_epoch = [100, 150, 200]
        decay_rates = [1.0, 0.1, 0.01, 0.001]
        batch_size = FLAGS.batch_size * (1 if not FLAGS.enbl_multi_gpu else mgw.size())
        lrn_rate = setup_lrn_rate_piecewise_constant(global_step, batch_size, idxs_epoch, decay_rates)
        nb_iters = int(FLAGS.nb_smpls_train * nb_epochs * FLAGS.nb_epochs_rat / batch_size)
        return lrn_rate, nb_iters

    @property
    def model_name(self):
        return 'resnet_%d' % FLAGS.resnet_size

    @property
    def dataset_name(self):
        return 'cifar_10'
