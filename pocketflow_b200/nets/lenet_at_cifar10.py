"""LeNet-like model on CIFAR-10 behind the ModelHelper plugin surface (/root/reference/nets/lenet_at_cifar10.py:28-135).
The forward pass ends in a softmax, so both cross-entropy terms see probabilities where logits are expected
 — kept, it is what the reference trains."""
from .. import graph as G
from ..flags import FLAGS, DEFINE_float
from ..datasets.cifar10_dataset import Cifar10Dataset
from .classification_helper import ClassificationModelHelper

DEFINE_float('nb_epochs_rat', 1.0, 'scales the number of training epochs')
DEFINE_float('lrn_rate_init', 1e-2, 'learning rate at batch size batch_size_norm')
DEFINE_float('batch_size_norm', 128, 'batch size the initial learning rate is quoted for')
DEFINE_float('momentum', 0.9, 'momentum of the SGD optimizer')
DEFINE_float('loss_w_dcy', 5e-4, 'weight of the L2 term')

CONV_STAGES = ((32, 'conv1', 'relu1', 'pool1'), (64, 'conv2', 'relu2', 'pool2'))


def forward_fn(inputs, data_format):
    """two x (5x5 VALID conv + bias, ReLU, 2x2 max-pool), flatten, dense 256 + ReLU, dense nb_classes, softmax."""
    if data_format != 'channels_last':
        raise ValueError('NHWC (channels_last) only')
    net = inputs
    for filters, conv, relu, pool in CONV_STAGES:
        net = G.max_pooling2d(G.relu(G.conv2d(net, filters, [5, 5], name=conv), name=relu), [2, 2], 2, name=pool)
    net = G.relu(G.dense(G.flatten(net, name='flatten'), 256, name='fc3'), name='relu3')
    return G.softmax(G.dense(net, FLAGS.nb_classes, name='fc4'), name='softmax')


class ModelHelper(ClassificationModelHelper):
    DATASET, DATASET_NAME = Cifar10Dataset, 'cifar_10'
    NB_EPOCHS, IDXS_EPOCH, DECAY_RATES = 250, [100, 150, 200], [1.0, 0.1, 0.01, 0.001]
    L2_SKIPS = None                                   # every trainable variable is regularised (:105-107)

    def network(self, inputs, is_train):
        return forward_fn(inputs, self.data_format)

    @property
    def model_name(self):
        return 'lenet'
