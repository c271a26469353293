#!/usr/bin/env python
"""GPU check of --enbl_device_preprocess end to end: the same TFRecord shards are streamed twice with the same seeds —
once through the host pipeline (fp32 batches), once packed (uint8 crops + descriptors) through
AbstractLearner._feed_packed and the pf_preprocess_images kernel — and the step's image placeholder must hold
bit-identical batches.  Prints the H2D bytes per step of both paths."""
import io
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pocketflow_b200 import graph as G  # noqa: E402
from pocketflow_b200.flags import FLAGS  # noqa: E402
from pocketflow_b200.datasets import ilsvrc12_dataset as D  # noqa: E402
from pocketflow_b200.learners.abstract_learner import AbstractLearner  # noqa: E402
from pocketflow_b200.utils import tf_record as R  # noqa: E402


def jpeg(h, w, seed):
    from PIL import Image
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, (h // 8 + 1, w // 8 + 1, 3)).astype(np.uint8)
    b = io.BytesIO()
    Image.fromarray(np.kron(base, np.ones((8, 8, 1), np.uint8))[:h, :w]).save(b, format='JPEG', quality=92)
    return b.getvalue()


def batches(data_dir, packed, is_train, n, batch):
    FLAGS.reset()
    FLAGS.data_dir_local, FLAGS.batch_size, FLAGS.batch_size_eval, FLAGS.nb_classes = data_dir, batch, batch, 1001
    FLAGS.buffer_size, FLAGS.nb_threads, FLAGS.prefetch_size = 16, 4, 2
    FLAGS.enbl_device_preprocess = packed
    dev = torch.device('cuda:0')
    with G.Graph().as_default():
        ds = D.Ilsvrc12Dataset(is_train)
        ds.batch_size, ds.nb_classes = batch, 1001
        it = ds.build()
        images, labels = it.get_next()
    ex = SimpleNamespace(buf={images: torch.full(images.shape, float('nan'), device=dev),
                              labels: torch.zeros(labels.shape, device=dev)})
    me = SimpleNamespace(_feed_packed=lambda *a: AbstractLearner._feed_packed(None, *a))
    os.environ['PF_INPUT_PREFETCH'] = '0'
    out, nbytes = [], []
    for _ in range(n):
        nbytes.append(AbstractLearner.feed(me, ex, it))
        torch.cuda.synchronize()
        out.append((ex.buf[images].cpu().numpy().copy(), ex.buf[labels].cpu().numpy().copy()))
    return out, nbytes


def main():
    d = tempfile.mkdtemp()
    for shard in range(4):
        recs = []
        for i in range(16):
            k = 16 * shard + i
            recs.append(R.encode_example({'image/encoded': jpeg(200 + 5 * k, 500 - 4 * k, k), 'image/class/label': [1 + k],
                                          'image/object/bbox/ymin': [0.1], 'image/object/bbox/xmin': [0.2],
                                          'image/object/bbox/ymax': [0.9], 'image/object/bbox/xmax': [0.8]}))
        R.write_records(os.path.join(d, 'train-%05d-of-00004' % shard), recs)
    R.write_records(os.path.join(d, 'validation-00000-of-00001'),
                    [R.encode_example({'image/encoded': jpeg(300 + 3 * i, 280 + 7 * i, 900 + i), 'image/class/label': [500 + i]})
                     for i in range(16)])
    for is_train in (True, False):
        host, hb = batches(d, False, is_train, 6, 8)
        dev, db = batches(d, True, is_train, 6, 8)
        for (hi, hl), (di, dl) in zip(host, dev):
            assert np.array_equal(hl, dl), 'labels differ'
            assert np.array_equal(hi, di), 'images differ: max |d| = %g' % np.nanmax(np.abs(hi - di))
        print('%s: 6 batches bit-identical; H2D bytes/step host %d, device-preprocess %d (%.2fx less)'
              % ('train' if is_train else 'eval', hb[0], int(np.mean(db)), hb[0] / np.mean(db)))
    print('preproc e2e ok')


if __name__ == '__main__':
    main()
