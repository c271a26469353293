"""CIFAR-10 binary reader + augmentation (host side): record format, standardisation constants, pad /
crop / flip semantics, epoch shuffling, train/val split, file-level rank sharding."""
import os

import numpy as np
import pytest

from pocketflow_b200.flags import FLAGS


def write_bins(tmp_path, n_files=2, per_file=40, seed=0):
    rng = np.random.default_rng(seed)
    truth = []
    for f in range(n_files):
        lab = rng.integers(0, 10, size=per_file).astype(np.uint8)
        img = rng.integers(0, 256, size=(per_file, 3, 32, 32)).astype(np.uint8)          # planar CHW, as on disk
        rec = np.concatenate([lab[:, None], img.reshape(per_file, -1)], axis=1)
        rec.tofile(os.path.join(tmp_path, 'data_batch_%d.bin' % (f + 1)))
        truth.append((lab, img))
    lab = rng.integers(0, 10, size=30).astype(np.uint8)
    img = rng.integers(0, 256, size=(30, 3, 32, 32)).astype(np.uint8)
    np.concatenate([lab[:, None], img.reshape(30, -1)], axis=1).tofile(os.path.join(tmp_path, 'test_batch.bin'))
    return truth, (lab, img)


@pytest.fixture
def cifar(tmp_path):
    FLAGS.reset()
    import importlib
    import pocketflow_b200.datasets.cifar10_dataset as C
    importlib.reload(C)
    truth, test = write_bins(str(tmp_path))
    FLAGS.data_dir_local = str(tmp_path)
    yield C, truth, test
    FLAGS.reset()


def test_record_parsing_and_standardisation(cifar):
    C, truth, test = cifar
    lab, img = C.read_records(os.path.join(FLAGS.data_dir_local, 'data_batch_1.bin'))
    assert np.array_equal(lab, truth[0][0]) and img.shape == (40, 32, 32, 3)
    assert np.array_equal(img, truth[0][1].transpose(0, 2, 3, 1))                       # CHW on disk -> HWC
    x = C.standardize(img[:2])
    ref = (img[:2].astype(np.float32) - np.array([125.3, 123.0, 113.9], np.float32)) / np.array([63.0, 62.1, 66.7], np.float32)
    assert np.array_equal(x, ref)
    with open(os.path.join(FLAGS.data_dir_local, 'bad.bin'), 'wb') as f:
        f.write(b'123')
    with pytest.raises(ValueError):
        C.read_records(os.path.join(FLAGS.data_dir_local, 'bad.bin'))


def test_eval_stream_covers_every_record_once_per_epoch(cifar):
    C, truth, (lab, img) = cifar
    FLAGS.batch_size_eval = 10
    it = C.Cifar10Dataset(is_train=False).build()
    assert it.stream
    seen = []
    for _ in range(3):                                                                   # 30 test records = one epoch
        x, y = it.next_batch()
        assert x.shape == (10, 32, 32, 3) and y.shape == (10, 10) and np.all(y.numpy().sum(1) == 1)
        seen.append((x.numpy().copy(), y.numpy().argmax(1)))
    ref = C.standardize(img.transpose(0, 2, 3, 1))
    xs = np.concatenate([s[0] for s in seen])
    ys = np.concatenate([s[1] for s in seen])
    # every record appears exactly once, un-augmented, with its own label
    used = set()
    for i in range(30):
        j = [k for k in range(30) if k not in used and ys[i] == lab[k] and np.array_equal(xs[i], ref[k])]
        assert j, i
        used.add(j[0])


def test_training_augmentation_is_pad_crop_flip(cifar):
    C, truth, _ = cifar
    rng = np.random.default_rng(3)
    base = C.standardize(truth[0][1][:16].transpose(0, 2, 3, 1))
    out = C.augment(base, rng)
    assert out.shape == base.shape
    padded = np.zeros((16, 40, 40, 3), np.float32)
    padded[:, 4:36, 4:36] = base
    kinds = set()
    for i in range(16):
        found = False
        for oy in range(9):
            for ox in range(9):
                crop = padded[i, oy:oy + 32, ox:ox + 32]
                if np.array_equal(out[i], crop):
                    found, k = True, ('plain', oy != 4 or ox != 4)
                elif np.array_equal(out[i], crop[:, ::-1]):
                    found, k = True, ('flip', oy != 4 or ox != 4)
                if found:
                    break
            if found:
                break
        assert found, 'image %d is not a (flipped) 32x32 crop of its zero-padded original' % i
        kinds.add(k[0])
    assert kinds == {'plain', 'flip'}                                                    # both happen in 16 draws


def test_train_val_split_and_determinism(cifar):
    C, truth, _ = cifar
    FLAGS.batch_size, FLAGS.nb_smpls_val = 8, 20
    trn, val = C.Cifar10Dataset(is_train=True).build(enbl_trn_val_split=True)
    assert len(trn.generator.labels) == 60 and len(val.generator.labels) == 20
    all_lab = np.concatenate([t[0] for t in truth])
    assert np.array_equal(val.generator.labels, all_lab[:20]) and np.array_equal(trn.generator.labels, all_lab[20:])
    a = [trn.next_batch()[0].numpy().copy() for _ in range(3)]
    trn2, _ = C.Cifar10Dataset(is_train=True).build(enbl_trn_val_split=True)
    b = [trn2.next_batch()[0].numpy().copy() for _ in range(3)]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))                               # same seed -> same stream


def test_rotating_pinned_buffers_are_refilled(cifar):
    C, _, _ = cifar
    FLAGS.batch_size = 4
    it = C.Cifar10Dataset(is_train=True).build()
    first = [it.next_batch()[0].numpy().copy() for _ in range(it.pool_size)]
    again = it.next_batch()[0].numpy()                                                   # reuses buffer 0 with NEW data
    assert not np.array_equal(again, first[0])


def test_learner_evaluates_on_the_eval_split_when_real_data_is_configured(cifar):
    """AbstractLearner.eval_iterator: with --data_dir_local the evaluation pass reads test_batch.bin (at the step's
    batch size, into the step's input placeholders); without it the synthetic training pool is reused."""
    import torch
    from types import SimpleNamespace
    from pocketflow_b200 import graph as G
    from pocketflow_b200.learners.abstract_learner import AbstractLearner
    from pocketflow_b200.nets import resnet_at_cifar10 as R
    import importlib
    importlib.reload(R)
    C, truth, (test_lab, test_img) = cifar

    class Probe(AbstractLearner):
        def train(self):
            pass

        def evaluate(self):
            pass

    FLAGS.batch_size, FLAGS.batch_size_eval = 6, 100
    lrn = Probe(None, R.ModelHelper())
    lrn.graph_train = G.Graph()
    with lrn.graph_train.as_default():
        lrn.iterator_train = lrn.build_dataset_train()
        images, labels = lrn.iterator_train.get_next()
    ex = SimpleNamespace(buf={images: torch.zeros(images.shape), labels: torch.zeros(labels.shape)})
    it = lrn.eval_iterator()
    assert it is not lrn.iterator_train and it is lrn.eval_iterator() and it.batch_size == 6
    seen = []
    for _ in range(5):                                         # 30 test records = 5 batches of 6: one epoch
        nbytes = lrn.feed(ex, it)
        assert nbytes == 6 * 32 * 32 * 3 * 4 + 6 * 10 * 4
        seen.append((ex.buf[images].numpy().copy(), ex.buf[labels].numpy().argmax(1)))
    got_labels = np.sort(np.concatenate([s[1] for s in seen]))
    np.testing.assert_array_equal(got_labels, np.sort(test_lab))               # exactly the evaluation split
    want = C.standardize(test_img.transpose(0, 2, 3, 1))                       # and un-augmented
    got = np.concatenate([s[0] for s in seen])
    assert sorted(np.round(got.reshape(30, -1).sum(1), 2).tolist()) == sorted(np.round(want.reshape(30, -1).sum(1), 2).tolist())
    FLAGS.data_dir_local = None
    assert lrn.eval_iterator() is lrn.iterator_train


def test_parse_fn_matches_the_executed_reference_source(cifar):
    """datasets/cifar10_dataset.py:parse_fn executed from the reference source with numpy stand-ins (controlled crop
    offsets / flip): record layout, standardisation (true fp32 division), +8 zero padding, crop, flip."""
    import base64
    import hashlib
    import json
    C, _, _ = cifar
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'ref_executed_v1.json')))['cifar10_parse_fn']
    assert len(gold) == 6

    def digest(a):
        a = np.ascontiguousarray(np.asarray(a, np.float32))
        return list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()

    class Draws(object):
        """stands in for numpy's Generator inside `augment`: hands out the golden's crop offsets and flip."""

        def __init__(self, oy, ox, flip):
            self.ints, self.flip = [np.array([oy]), np.array([ox])], flip

        def integers(self, lo, hi, size):
            return self.ints.pop(0)

        def random(self, n):
            return np.array([0.25 if self.flip else 0.75])
    for g in gold:
        path = os.path.join(FLAGS.data_dir_local, 'one_record.bin')
        open(path, 'wb').write(base64.b64decode(g['record_b64']))
        lab, img = C.read_records(path)
        onehot = np.zeros(10, np.float32)
        onehot[lab[0]] = 1.0
        assert onehot.tolist() == g['label']
        std = C.standardize(img)
        assert digest(std[0]) == (g['eval']['shape'], g['eval']['sha256'])
        for tr in g['train']:
            out = C.augment(std, Draws(tr['oy'], tr['ox'], tr['flip']))
            assert digest(out[0]) == (tr['out']['shape'], tr['out']['sha256']), tr
