"""MobileNet-v1 on ILSVRC-12 behind the ModelHelper plugin surface
(/root/reference/nets/mobilenet_at_ilsvrc12.py:29-160); version 2 is out of scope."""
from .. import graph as G
from ..flags import FLAGS, DEFINE_integer, DEFINE_float
from ..datasets.ilsvrc12_dataset import Ilsvrc12Dataset
from .classification_helper import ClassificationModelHelper
from . import mobilenet_v1 as MobileNetV1

DEFINE_integer('mobilenet_version', 1, 'MobileNet version (only 1 is built)')
DEFINE_float('mobilenet_depth_mult', 1.0, 'channel multiplier of every layer')
DEFINE_float('nb_epochs_rat', 1.0, 'scales the number of training epochs')
DEFINE_float('lrn_rate_init', 0.045, 'learning rate at batch size batch_size_norm')
DEFINE_float('batch_size_norm', 96, 'batch size the initial learning rate is quoted for')
DEFINE_float('momentum', 0.9, 'momentum of the SGD optimizer')
DEFINE_float('loss_w_dcy', 4e-5, 'weight of the L2 term')


def forward_fn(inputs, is_train):
    if FLAGS.mobilenet_version != 1:
        raise ValueError('invalid MobileNet version: {} (only v1 is built)'.format(FLAGS.mobilenet_version))
    return MobileNetV1.mobilenet_v1(inputs, num_classes=FLAGS.nb_classes, is_training=is_train,
                                    depth_multiplier=FLAGS.mobilenet_depth_mult)


class ModelHelper(ClassificationModelHelper):
    DATASET, DATASET_NAME = Ilsvrc12Dataset, 'ilsvrc_12'
    NB_EPOCHS, IDXS_EPOCH, DECAY_RATES = 100, [30, 60, 80, 90], [1.0, 0.1, 0.01, 0.001, 0.0001]
    # the filter names TF-layers' scope; slim calls its batch-norm scope 'BatchNorm', so gamma / beta ARE regularised
    # here — what the reference does (mobilenet_at_ilsvrc12.py:107-109)
    L2_SKIPS = 'batch_normalization'

    def __init__(self, data_format='channels_last'):
        assert data_format == 'channels_last', 'MobileNet only supports \'channels_last\' data format'
        super(ModelHelper, self).__init__(data_format)

    def network(self, inputs, is_train):
        return forward_fn(inputs, is_train=is_train)

    def metrics(self, labels, outputs):
        """'accuracy' is the top-5 figure here (mobilenet_at_ilsvrc12.py:110-113)."""
        top1, top5 = G.accuracy(labels, outputs), G.in_top_k_accuracy(labels, outputs, 5)
        return {'accuracy': top5, 'acc_top1': top1, 'acc_top5': top5}

    @property
    def model_name(self):
        return 'mobilenet_v%d' % FLAGS.mobilenet_version
