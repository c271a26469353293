"""Model helper for creating a LeNet-like model for the CIFAR-10 dataset
(/root/reference/nets/lenet_at_cifar10.py:28-135).  Note the forward pass ends in softmax, so both
cross-entropy terms see probabilities as logits — mirrored."""
from .. import graph as G
from ..flags import FLAGS, DEFINE_float
from ..datasets.cifar10_dataset import Cifar10Dataset
from ..utils.lrn_rate_utils import setup_lrn_rate_piecewise_constant
from ..utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from .abstract_model_helper import AbstractModelHelper

DEFINE_float('nb_epochs_rat', 1.0, '# of training epochs\'s ratio')
DEFINE_float('lrn_rate_init', 1e-2, 'initial learning rate')
DEFINE_float('batch_size_norm', 128, 'normalization factor of batch size')
DEFINE_float('momentum', 0.9, 'momentum coefficient')
DEFINE_float('loss_w_dcy', 5e-4, 'weight decaying loss\'s coefficient')


def forward_fn(inputs, data_format):
    if data_format != 'channels_last':
        raise ValueError('NHWC (channels_last) only')
    inputs = G.conv2d(inputs, 32, [5, 5], name='conv1')
    inputs = G.relu(inputs, name='relu1')
    inputs = G.max_pooling2d(inputs, [2, 2], 2, name='pool1')
    inputs = G.conv2d(inputs, 64, [5, 5], name='conv2')
    inputs = G.relu(inputs, name='relu2')
    inputs = G.max_pooling2d(inputs, [2, 2], 2, name='pool2')
    inputs = G.flatten(inputs, name='flatten')
    inputs = G.dense(inputs, 256, name='fc3')
    inputs = G.relu(inputs, name='relu3')
    inputs = G.dense(inputs, FLAGS.nb_classes, name='fc4')
    inputs = G.softmax(inputs, name='softmax')
    return inputs


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
        return forward_fn(inputs, self.data_format)

    def forward_eval(self, inputs):
        return forward_fn(inputs, self.data_format)

    def calc_loss(self, labels, outputs, trainable_vars):
        loss = G.softmax_cross_entropy(labels, outputs)
        loss += FLAGS.loss_w_dcy * G.add_n([G.l2_loss(var) for var in trainable_vars])
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
        return 'lenet'

    @property
    def dataset_name(self):
        return 'cifar_10'
