"""CIFAR-10 dataset (/root/reference/datasets/cifar10_dataset.py:26-104).

Without --data_dir_local the configs run on synthetic CIFAR-10-shaped batches (32x32x3, 10 classes).  With it, the
binary distribution (data_batch_{1..5}.bin / test_batch.bin: records of 1 label byte + 3x32x32 pixel bytes, planar
CHW) is read, standardised with the reference's per-channel mean / std (:39-41,62) and — for training — augmented
the reference's way (:65-68): zero-pad to 40x40, random 32x32 crop, random horizontal flip.  Host-side numpy; the
batches land in the iterator's rotating pinned buffers and reach the GPU through the learner's staged H2D copy.
Deviation (flagged): tf.data shuffles with a 1024-record buffer over interleaved files; here every epoch is a
full permutation of the rank's records."""
import glob
import os

import numpy as np

from ..flags import FLAGS, DEFINE_integer
from ..utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from .abstract_dataset import AbstractDataset

DEFINE_integer('nb_classes', 10, '# of classes')
DEFINE_integer('nb_smpls_train', 50000, '# of samples for training')
DEFINE_integer('nb_smpls_val', 5000, '# of samples for validation')
DEFINE_integer('nb_smpls_eval', 10000, '# of samples for evaluation')
DEFINE_integer('batch_size', 128, 'batch size per GPU for training')
DEFINE_integer('batch_size_eval', 100, 'batch size for evaluation')

IMAGE_HEI, IMAGE_WID, IMAGE_CHN = 32, 32, 3
LABEL_BYTES = 1
IMAGE_BYTES = IMAGE_CHN * IMAGE_HEI * IMAGE_WID
RECORD_BYTES = LABEL_BYTES + IMAGE_BYTES
IMAGE_AVE = np.array([125.3, 123.0, 113.9], np.float32)
IMAGE_STD = np.array([63.0, 62.1, 66.7], np.float32)
PAD = 4                                                   # resize_image_with_crop_or_pad(+8) = 4 zero pixels per side


def read_records(path):
    """(uint8 labels [n], uint8 images [n,32,32,3]) of one CIFAR-10 binary file."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size % RECORD_BYTES:
        raise ValueError('%s: size %d is not a multiple of the %d-byte record' % (path, raw.size, RECORD_BYTES))
    rec = raw.reshape(-1, RECORD_BYTES)
    images = rec[:, LABEL_BYTES:].reshape(-1, IMAGE_CHN, IMAGE_HEI, IMAGE_WID).transpose(0, 2, 3, 1)
    return rec[:, 0].copy(), np.ascontiguousarray(images)


def standardize(images_u8):
    """(x - mean) / std per channel, in the reference's op order (true fp32 division)."""
    return (images_u8.astype(np.float32) - IMAGE_AVE) / IMAGE_STD


def augment(images, rng):
    """zero-pad to 40x40, random 32x32 crop, random left-right flip — per image."""
    n = images.shape[0]
    padded = np.zeros((n, IMAGE_HEI + 2 * PAD, IMAGE_WID + 2 * PAD, IMAGE_CHN), np.float32)
    padded[:, PAD:PAD + IMAGE_HEI, PAD:PAD + IMAGE_WID] = images
    oy = rng.integers(0, 2 * PAD + 1, size=n)
    ox = rng.integers(0, 2 * PAD + 1, size=n)
    flip = rng.random(n) < 0.5
    out = np.empty_like(images)
    for i in range(n):
        crop = padded[i, oy[i]:oy[i] + IMAGE_HEI, ox[i]:ox[i] + IMAGE_WID]
        out[i] = crop[:, ::-1] if flip[i] else crop
    return out


class RecordStream(object):
    """Endless shuffled stream over in-memory records: a fresh permutation per epoch, batches may straddle epochs."""

    def __init__(self, labels, images, nb_classes, is_train, seed):
        self.labels, self.images, self.k, self.is_train = labels, images, nb_classes, is_train
        self.rng = np.random.default_rng(seed)
        self.order, self.pos = self.rng.permutation(len(labels)), 0

    def __call__(self, b):
        idx = np.empty(b, np.int64)
        got = 0
        while got < b:
            take = min(b - got, len(self.order) - self.pos)
            idx[got:got + take] = self.order[self.pos:self.pos + take]
            got, self.pos = got + take, self.pos + take
            if self.pos == len(self.order):
                self.order, self.pos = self.rng.permutation(len(self.labels)), 0
        img = standardize(self.images[idx])
        if self.is_train:
            img = augment(img, self.rng)
        lab = np.zeros((b, self.k), np.float32)
        lab[np.arange(b), self.labels[idx]] = 1.0
        return np.ascontiguousarray(img, np.float32), lab


class Cifar10Dataset(AbstractDataset):
    def __init__(self, is_train):
        super().__init__(is_train)   # zero-arg form: survives importlib.reload of this module
        self.batch_size = FLAGS.batch_size if is_train else FLAGS.batch_size_eval
        self.image_shape = (IMAGE_HEI, IMAGE_WID, IMAGE_CHN)
        self.nb_classes = FLAGS.nb_classes

    def _file_generators(self, enbl_trn_val_split):
        if FLAGS.data_disk != 'local':
            raise ValueError('unrecognized data disk: ' + FLAGS.data_disk)     # (HDFS access is cluster glue: out of scope)
        pattern = os.path.join(FLAGS.data_dir_local, 'data_batch_*.bin' if self.is_train else 'test_batch.bin')
        files = sorted(glob.glob(pattern))
        if not files:
            raise FileNotFoundError('no CIFAR-10 binary files match ' + pattern)
        rank, size = (mgw.rank(), mgw.size()) if self.enbl_shard else (0, 1)
        files = files[rank::size] or files                   # file-level sharding (abstract_dataset.py:80-81)
        parts = [read_records(f) for f in files]
        labels = np.concatenate([p[0] for p in parts])
        images = np.concatenate([p[1] for p in parts])
        seed = 4321 + 7919 * rank + (0 if self.is_train else 1)
        if self.is_train and enbl_trn_val_split:
            nv = min(FLAGS.nb_smpls_val // size, len(labels) // 2)            # take(nb_smpls_val) / skip(nb_smpls_val)
            return [RecordStream(labels[nv:], images[nv:], self.nb_classes, True, seed),
                    RecordStream(labels[:nv], images[:nv], self.nb_classes, True, seed + 1)]
        return [RecordStream(labels, images, self.nb_classes, self.is_train, seed)]
