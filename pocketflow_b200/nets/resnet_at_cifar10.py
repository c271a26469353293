"""ResNet-(6n+2) on CIFAR-10 behind the ModelHelper plugin surface (/root/reference/nets/resnet_at_cifar10.py:29-135)."""
from ..flags import FLAGS, DEFINE_integer, DEFINE_float
from ..datasets.cifar10_dataset import Cifar10Dataset
from .classification_helper import ClassificationModelHelper
from . import resnet_model as ResNet

DEFINE_integer('resnet_size', 20, 'depth of the ResNet (6n + 2)')
DEFINE_float('nb_epochs_rat', 1.0, 'scales the number of training epochs')
DEFINE_float('lrn_rate_init', 1e-1, 'learning rate at batch size batch_size_norm')
DEFINE_float('batch_size_norm', 128, 'batch size the initial learning rate is quoted for')
DEFINE_float('momentum', 0.9, 'momentum of the SGD optimizer')
DEFINE_float('loss_w_dcy', 2e-4, 'weight of the L2 term')


def forward_fn(inputs, is_train, data_format):
    """v1 basic blocks, 16 filters, 3x3 stride-1 stem without pooling, three stages of (size - 2) / 6 blocks."""
    per_stage = (FLAGS.resnet_size - 2) // 6
    model = ResNet.Model(FLAGS.resnet_size, False, FLAGS.nb_classes, 16, 3, 1, None, None,
                         [per_stage] * 3, [1, 2, 2], data_format=data_format)
    return model(inputs, is_train)


class ModelHelper(ClassificationModelHelper):
    DATASET, DATASET_NAME = Cifar10Dataset, 'cifar_10'
    NB_EPOCHS, IDXS_EPOCH, DECAY_RATES = 250, [100, 150, 200], [1.0, 0.1, 0.01, 0.001]
    L2_SKIPS = 'batch_normalization'                  # BN gains / offsets are not regularised (:103-107)

    def network(self, inputs, is_train):
        return forward_fn(inputs, is_train=is_train, data_format=self.data_format)

    @property
    def model_name(self):
        return 'resnet_%d' % FLAGS.resnet_size
