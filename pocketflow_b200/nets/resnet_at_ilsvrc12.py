"""ResNet-18/34/50/101/152/200 on ILSVRC-12 behind the ModelHelper plugin surface
(/root/reference/nets/resnet_at_ilsvrc12.py:29-165)."""
from ..flags import FLAGS, DEFINE_integer, DEFINE_float
from ..datasets.ilsvrc12_dataset import Ilsvrc12Dataset
from .classification_helper import ClassificationModelHelper
from . import resnet_model as ResNet

DEFINE_integer('resnet_size', 18, 'depth of the ResNet')
DEFINE_float('nb_epochs_rat', 1.0, 'scales the number of training epochs')
DEFINE_float('lrn_rate_init', 1e-1, 'learning rate at batch size batch_size_norm')
DEFINE_float('batch_size_norm', 256, 'batch size the initial learning rate is quoted for')
DEFINE_float('momentum', 0.9, 'momentum of the SGD optimizer')
DEFINE_float('loss_w_dcy', 1e-4, 'weight of the L2 term')

BLOCKS_PER_STAGE = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3],
                    200: [3, 24, 36, 3]}


def get_block_sizes(resnet_size):
    if resnet_size not in BLOCKS_PER_STAGE:
        raise ValueError('invalid # of layers for ResNet: {}'.format(resnet_size))
    return BLOCKS_PER_STAGE[resnet_size]


def forward_fn(inputs, is_train, data_format):
    """64 filters, 7x7 stride-2 stem + 3x3 stride-2 max-pool, four stages; bottleneck blocks from depth 50 on."""
    model = ResNet.Model(FLAGS.resnet_size, FLAGS.resnet_size >= 50, FLAGS.nb_classes, 64, 7, 2, 3, 2,
                         get_block_sizes(FLAGS.resnet_size), [1, 2, 2, 2], data_format=data_format)
    return model(inputs, is_train)


class ModelHelper(ClassificationModelHelper):
    DATASET, DATASET_NAME = Ilsvrc12Dataset, 'ilsvrc_12'
    NB_EPOCHS, IDXS_EPOCH, DECAY_RATES = 100, [30, 60, 80, 90], [1.0, 0.1, 0.01, 0.001, 0.0001]
    L2_SKIPS = 'batch_normalization'
    TOP5 = True

    def network(self, inputs, is_train):
        return forward_fn(inputs, is_train=is_train, data_format=self.data_format)

    @property
    def model_name(self):
        return 'resnet_%d' % FLAGS.resnet_size
