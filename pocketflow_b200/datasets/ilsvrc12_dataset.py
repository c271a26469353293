"""ILSVRC-12 dataset (/root/reference/datasets/ilsvrc12_dataset.py:27-32): flags and shapes;
synthetic 224x224x3 batches with 1001 classes."""
from ..flags import FLAGS, DEFINE_integer
from .abstract_dataset import AbstractDataset

DEFINE_integer('nb_classes', 1001, '# of classes')
DEFINE_integer('nb_smpls_train', 1281167, '# of samples for training')
DEFINE_integer('nb_smpls_val', 10000, '# of samples for validation')
DEFINE_integer('nb_smpls_eval', 50000, '# of samples for evaluation')
DEFINE_integer('batch_size', 64, 'batch size per GPU for training')
DEFINE_integer('batch_size_eval', 100, 'batch size for evaluation')

IMAGE_HEI, IMAGE_WID, IMAGE_CHN = 224, 224, 3


class Ilsvrc12Dataset(AbstractDataset):
    def __init__(self, is_train):
        super(Ilsvrc12Dataset, self).__init__(is_train)
        self.batch_size = FLAGS.batch_size if is_train else FLAGS.batch_size_eval
        self.image_shape = (IMAGE_HEI, IMAGE_WID, IMAGE_CHN)
        self.nb_classes = FLAGS.nb_classes
