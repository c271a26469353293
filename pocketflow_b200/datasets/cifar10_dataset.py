"""CIFAR-10 dataset (/root/reference/datasets/cifar10_dataset.py:26-104): flags and shapes.
The configs run on synthetic CIFAR-10-shaped batches (32x32x3, 10 classes); the binary-file
reader + pad/crop/flip augmentation is a "next" row."""
from ..flags import FLAGS, DEFINE_integer
from .abstract_dataset import AbstractDataset

DEFINE_integer('nb_classes', 10, '# of classes')
DEFINE_integer('nb_smpls_train', 50000, '# of samples for training')
DEFINE_integer('nb_smpls_val', 5000, '# of samples for validation')
DEFINE_integer('nb_smpls_eval', 10000, '# of samples for evaluation')
DEFINE_integer('batch_size', 128, 'batch size per GPU for training')
DEFINE_integer('batch_size_eval', 100, 'batch size for evaluation')

IMAGE_HEI, IMAGE_WID, IMAGE_CHN = 32, 32, 3


class Cifar10Dataset(AbstractDataset):
    def __init__(self, is_train):
        super(Cifar10Dataset, self).__init__(is_train)
        self.batch_size = FLAGS.batch_size if is_train else FLAGS.batch_size_eval
        self.image_shape = (IMAGE_HEI, IMAGE_WID, IMAGE_CHN)
        self.nb_classes = FLAGS.nb_classes
