"""ILSVRC-12 input pipeline without TensorFlow: TFRecord framing, tf.train.Example parsing,
the reference's eval / training preprocessing and the shuffled, sharded, prefetched stream."""
import io
import os
import struct

import numpy as np
import pytest

from pocketflow_b200.utils import tf_record as R
from pocketflow_b200.utils.tf_bundle import crc32c, mask_crc

PIL = pytest.importorskip('PIL.Image')


def _jpeg(h, w, seed, mode='RGB'):
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, (h // 8 + 1, w // 8 + 1, 3)).astype(np.uint8)
    img = PIL.fromarray(np.kron(base, np.ones((8, 8, 1), np.uint8))[:h, :w])
    if mode != 'RGB':
        img = img.convert(mode)
    b = io.BytesIO()
    img.save(b, format='JPEG', quality=92)
    return b.getvalue()


def _example(jpeg, label, boxes=()):
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    return R.encode_example({
        'image/encoded': jpeg, 'image/class/label': [label], 'image/class/text': b'n0000',
        'image/object/bbox/ymin': boxes[:, 0], 'image/object/bbox/xmin': boxes[:, 1],
        'image/object/bbox/ymax': boxes[:, 2], 'image/object/bbox/xmax': boxes[:, 3]})


def test_record_framing_bytes_and_corruption(tmp_path):
    p = str(tmp_path / 'a.tfrecord')
    payloads = [b'', b'x', bytes(range(256)) * 5]
    R.write_records(p, payloads)
    raw = open(p, 'rb').read()
    # first record, spelled out: length 0, crc of the 8 length bytes, no data, crc of the empty string
    assert raw[:16] == struct.pack('<Q', 0) + struct.pack('<I', mask_crc(crc32c(bytes(8)))) + struct.pack('<I', mask_crc(0))
    assert list(R.read_records(p)) == payloads
    for pos in (3, 9, 20, len(raw) - 2):
        bad = bytearray(raw)
        bad[pos] ^= 4
        open(p, 'wb').write(bad)
        with pytest.raises(ValueError):
            list(R.read_records(p))
    open(p, 'wb').write(raw[:-5])
    with pytest.raises(ValueError, match='truncated'):
        list(R.read_records(p))


def test_example_parsing_packed_and_unpacked():
    ex = R.encode_example({'image/encoded': b'\xff\xd8jpeg', 'image/class/label': [7], 'f': np.array([0.5, -2.0], np.float32),
                           'neg': np.array([-3, 1 << 40]), 'empty': np.zeros(0, np.float32)})
    f = R.parse_example(ex)
    assert f['image/encoded'] == [b'\xff\xd8jpeg'] and f['image/class/label'].tolist() == [7]
    assert f['f'].dtype == np.float32 and f['f'].tolist() == [0.5, -2.0]
    assert f['neg'].tolist() == [-3, 1 << 40] and len(f['empty']) == 0
    # the same message with UNPACKED repeated scalars, written out by hand:
    # Example{features{feature{key:"l" value{int64_list{value:5 value:6}}} feature{key:"x" value{float_list{value:1.5}}}}}
    int_list = bytes([0x08, 5, 0x08, 6])
    flt_list = bytes([0x0d]) + struct.pack('<f', 1.5)
    def entry(k, feat):
        body = bytes([0x0a, len(k)]) + k + bytes([0x12, len(feat)]) + feat
        return bytes([0x0a, len(body)]) + body
    feats = entry(b'l', bytes([0x1a, len(int_list)]) + int_list) + entry(b'x', bytes([0x12, len(flt_list)]) + flt_list)
    f = R.parse_example(bytes([0x0a, len(feats)]) + feats)
    assert f['l'].tolist() == [5, 6] and f['x'].tolist() == [1.5]


def test_legacy_bilinear_resize_known_answers():
    from pocketflow_b200.datasets.ilsvrc12_dataset import resize_bilinear
    ramp = np.arange(4, dtype=np.float32).reshape(1, 4, 1) * np.ones((3, 1, 2), np.float32)
    # x2 up-sampling without half-pixel centres: even outputs hit a source pixel, odd ones the mid-point to the next
    # pixel, and the last one clamps
    up = resize_bilinear(ramp, 3, 8)
    np.testing.assert_array_equal(up[0, :, 0], [0, 0.5, 1, 1.5, 2, 2.5, 3, 3])
    # x2 down-sampling picks every other pixel (no averaging) — the hallmark of the TF1 kernel
    img = np.random.RandomState(0).rand(8, 6, 3).astype(np.float32)
    np.testing.assert_array_equal(resize_bilinear(img, 4, 3), img[::2, ::2])
    np.testing.assert_array_equal(resize_bilinear(img, 8, 6), img)
    # a non-integer ratio against the formula spelled out per output pixel
    out = resize_bilinear(img, 5, 4)
    for y in range(5):
        for x in range(4):
            sy, sx = np.float32(y) * (np.float32(8) / np.float32(5)), np.float32(x) * (np.float32(6) / np.float32(4))
            y0, x0 = int(np.floor(sy)), int(np.floor(sx))
            y1, x1 = min(int(np.ceil(sy)), 7), min(int(np.ceil(sx)), 5)
            ly, lx = np.float32(sy - y0), np.float32(sx - x0)
            top = img[y0, x0] + (img[y0, x1] - img[y0, x0]) * lx
            bot = img[y1, x0] + (img[y1, x1] - img[y1, x0]) * lx
            np.testing.assert_array_equal(out[y, x], top + (bot - top) * ly)


def test_eval_preprocessing_follows_the_reference_steps():
    from pocketflow_b200.datasets import ilsvrc12_dataset as D
    assert D.smallest_size_at_least(375, 500) == (256, 341)            # 500 * (256 / 375) = 341.33 -> 341
    assert D.smallest_size_at_least(500, 375) == (341, 256)
    assert D.smallest_size_at_least(256, 256) == (256, 256)
    j = _jpeg(300, 400, 1)
    out = D.preprocess_image(j, np.zeros((0, 4)), False)
    assert out.shape == (224, 224, 3) and out.dtype == np.float32
    dec = D.decode_jpeg(j)
    assert dec.shape == (300, 400, 3) and D.jpeg_shape(j) == (300, 400)
    nh, nw = D.smallest_size_at_least(300, 400)
    ref = D.resize_bilinear(dec, nh, nw)
    ref = ref[(nh - 224) // 2:(nh - 224) // 2 + 224, (nw - 224) // 2:(nw - 224) // 2 + 224] - np.array([123.68, 116.78, 103.94], np.float32)
    np.testing.assert_array_equal(out, ref)
    grey = D.decode_jpeg(_jpeg(64, 48, 2, mode='L'))
    assert grey.shape == (64, 48, 3) and np.array_equal(grey[..., 0], grey[..., 1])


def test_distorted_bounding_box_respects_its_constraints():
    from pocketflow_b200.datasets.ilsvrc12_dataset import sample_distorted_bounding_box
    rng = np.random.default_rng(0)
    H, W = 375, 500
    box = np.array([[0.2, 0.3, 0.7, 0.9]], np.float32)
    bx0, by0, bx1, by1 = int(0.3 * W), int(0.2 * H), int(0.9 * W), int(0.7 * H)
    areas, aspects = [], []
    for _ in range(400):
        y, x, h, w = sample_distorted_bounding_box(H, W, box, rng)
        assert 0 <= y and 0 <= x and y + h <= H and x + w <= W and h > 0 and w > 0
        inter = max(min(x + w, bx1) - max(x, bx0), 0) * max(min(y + h, by1) - max(y, by0), 0)
        assert inter / float((bx1 - bx0) * (by1 - by0)) >= 0.1
        areas.append(h * w / float(H * W))
        aspects.append(w / float(h))
    assert 0.05 - 1e-3 <= min(areas) and max(areas) <= 1.0
    assert 0.74 <= min(aspects) and max(aspects) <= 1.35                # rounding to whole pixels
    assert np.std(areas) > 0.15 and min(areas) < 0.15 and max(areas) > 0.8      # the whole area range is used
    # no boxes: the image itself is the object; impossible constraints: the whole image comes back
    y, x, h, w = sample_distorted_bounding_box(H, W, np.zeros((0, 4)), rng)
    assert h * w >= 0.05 * H * W - H
    assert sample_distorted_bounding_box(H, W, box, rng, min_object_covered=1.0, area_range=(0.05, 0.06)) == (0, 0, H, W)


def test_dataset_stream_reads_shards_and_preprocesses(tmp_path):
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    import importlib
    D = importlib.import_module('pocketflow_b200.datasets.ilsvrc12_dataset')
    d = str(tmp_path)
    labels = {}
    for shard in range(3):
        recs = []
        for i in range(5):
            idx = shard * 5 + i
            labels[idx] = 1 + idx
            recs.append(_example(_jpeg(240 + 8 * idx, 320 - 8 * idx, idx), 1 + idx, [[0.1, 0.1, 0.9, 0.9]] if idx % 2 else []))
        R.write_records(os.path.join(d, 'train-%05d-of-00003' % shard), recs)
    R.write_records(os.path.join(d, 'validation-00000-of-00001'), [_example(_jpeg(256, 300, 99), 42)])
    FLAGS.reset()
    FLAGS.data_dir_local, FLAGS.batch_size, FLAGS.batch_size_eval, FLAGS.nb_classes = d, 4, 2, 1001
    FLAGS.buffer_size, FLAGS.nb_threads, FLAGS.prefetch_size, FLAGS.nb_smpls_val = 6, 2, 2, 5
    gr = G.Graph()
    with gr.as_default():
        it = D.Ilsvrc12Dataset(is_train=True).build()
        seen = []
        for _ in range(15):                                  # 60 samples = 4 passes over the 15 records
            img, lab = it.next_batch()
            assert tuple(img.shape) == (4, 224, 224, 3) and tuple(lab.shape) == (4, 1001)
            assert np.isfinite(img.numpy()).all() and (lab.numpy().sum(1) == 1).all()
            seen.extend(lab.numpy().argmax(1).tolist())
        counts = np.bincount(seen, minlength=17)[1:16]
        assert counts.min() >= 3 and counts.max() <= 5 and counts.sum() == 60          # every record, every pass
        assert seen[:15] != sorted(seen[:15])                                        # shuffled
        # train / validation split of the training files: disjoint, nb_smpls_val records on the validation side
        trn, val = D.Ilsvrc12Dataset(is_train=True).build(enbl_trn_val_split=True)
        tl = set(np.concatenate([trn.next_batch()[1].numpy().argmax(1) for _ in range(10)]).tolist())
        vl = set(np.concatenate([val.next_batch()[1].numpy().argmax(1) for _ in range(10)]).tolist())
        assert len(vl) == 5 and len(tl) == 10 and not (tl & vl)
        ev = D.Ilsvrc12Dataset(is_train=False).build()
        img, lab = ev.next_batch()
        assert tuple(img.shape) == (2, 224, 224, 3) and lab.numpy().argmax(1).tolist() == [42, 42]
        want = D.preprocess_image(_jpeg(256, 300, 99), np.zeros((0, 4)), False)
        np.testing.assert_array_equal(img.numpy()[0], want)
    FLAGS.data_dir_local = os.path.join(d, 'nothing_here')
    with pytest.raises(FileNotFoundError):
        D.Ilsvrc12Dataset(is_train=True).build()
    FLAGS.reset()


def test_descriptor_form_equals_the_host_preprocessing():
    """crop_and_descriptor + preprocess_from_descriptor (the numpy statement of the device kernel pf_preprocess_images)
    == preprocess_image, bit for bit: training (random crop, flip BEFORE the asymmetric resize) and evaluation (256
    short side + central window)."""
    from pocketflow_b200.datasets import ilsvrc12_dataset as D
    assert D.IMG_DESC.itemsize == 40
    box = np.array([[0.1, 0.2, 0.8, 0.9]], np.float32)
    flips = 0
    for seed in range(12):
        j = _jpeg(200 + 17 * seed, 330 - 13 * seed, seed)
        want = D.preprocess_image(j, box, True, np.random.default_rng(seed))
        crop, d = D.crop_and_descriptor(j, box, True, np.random.default_rng(seed))
        flips += int(d['flip'])
        assert crop.dtype == np.uint8 and crop.shape == (int(d['h']), int(d['w']), 3) and (int(d['rh']), int(d['rw'])) == (224, 224)
        np.testing.assert_array_equal(D.preprocess_from_descriptor(crop, d), want)
        want = D.preprocess_image(j, box, False)
        crop, d = D.crop_and_descriptor(j, box, False)
        assert int(d['flip']) == 0 and min(int(d['rh']), int(d['rw'])) == 256
        np.testing.assert_array_equal(D.preprocess_from_descriptor(crop, d), want)
    assert 0 < flips < 12


def test_device_kernel_source_run_on_the_host_matches_bit_for_bit(tmp_path):
    """csrc/pf_preproc.cu keeps its per-value arithmetic and its index decomposition in __host__ __device__ functions;
    compiled with -DPF_PREPROC_HOST_TEST (no kernel, no launcher, a plain loop instead) the SAME source runs here and
    must reproduce the host pipeline exactly.  What this leaves to the GPU run: the launch configuration and the
    device intrinsics' rounding (__f*_rn = IEEE, like the host ops compiled with -ffp-contract=off)."""
    import ctypes
    import shutil
    import subprocess
    from pocketflow_b200.datasets import ilsvrc12_dataset as D
    if shutil.which('g++') is None:
        pytest.skip('no host compiler')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / 'libpreproc_host.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-x', 'c++', '-D__host__=', '-D__device__=', '-DPF_PREPROC_HOST_TEST',
                           '-ffp-contract=off', '-fPIC', '-shared', '-I', os.path.join(root, 'include'), '-o', lib,
                           os.path.join(root, 'pocketflow_b200', 'csrc', 'pf_preproc.cu')])
    fn = ctypes.CDLL(lib).pf_test_preprocess_host
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                   ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
    box = np.array([[0.1, 0.2, 0.8, 0.9]], np.float32)
    for is_training in (True, False):
        crops, descs, want = [], [], []
        offset = 0
        for seed in range(7):
            j = _jpeg(150 + 23 * seed, 310 - 19 * seed, seed)
            crop, d = D.crop_and_descriptor(j, box, is_training, np.random.default_rng(seed))
            d['offset'] = offset
            offset += crop.size
            crops.append(crop.reshape(-1))
            descs.append(d)
            want.append(D.preprocess_image(j, box, is_training, np.random.default_rng(seed)))
        packed = np.concatenate(crops)
        table = np.stack(descs)
        out = np.full((7, 224, 224, 3), np.nan, np.float32)
        fn(packed.ctypes.data, table.ctypes.data, 7, 224, 224, 123.68, 116.78, 103.94, out.ctypes.data)
        np.testing.assert_array_equal(out, np.stack(want))


def _host_preprocess_lib(tmp_path):
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / 'libpreproc_host.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-x', 'c++', '-D__host__=', '-D__device__=', '-DPF_PREPROC_HOST_TEST',
                           '-ffp-contract=off', '-fPIC', '-shared', '-I', os.path.join(root, 'include'), '-o', lib,
                           os.path.join(root, 'pocketflow_b200', 'csrc', 'pf_preproc.cu')])
    fn = ctypes.CDLL(lib).pf_test_preprocess_host
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                   ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
    return fn


def test_packed_stream_through_feed_equals_the_host_pipeline(tmp_path, monkeypatch):
    """--enbl_device_preprocess end to end on the CPU: PackedBatchIterator -> AbstractLearner.feed ->
    ops.preprocess_images (here: the kernel's source compiled for the host) fills the step's image placeholder with
    exactly the batches the host pipeline produces from the same records and seeds."""
    import shutil
    import torch
    from types import SimpleNamespace
    from pocketflow_b200 import graph as G, ops
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.datasets import ilsvrc12_dataset as D
    from pocketflow_b200.datasets.abstract_dataset import PackedBatchIterator
    from pocketflow_b200.learners.abstract_learner import AbstractLearner
    if shutil.which('g++') is None:
        pytest.skip('no host compiler')
    host_fn = _host_preprocess_lib(tmp_path)

    def on_host(crops, desc, out, mean=(123.68, 116.78, 103.94)):
        assert crops.dtype == torch.uint8 and desc.numel() == out.shape[0] * 40 and out.is_contiguous()
        host_fn(crops.data_ptr(), desc.data_ptr(), out.shape[0], out.shape[1], out.shape[2], mean[0], mean[1], mean[2],
                out.data_ptr())
        return out
    monkeypatch.setattr(ops, 'preprocess_images', on_host)
    d = str(tmp_path / 'data')
    os.makedirs(d)
    for shard in range(2):
        R.write_records(os.path.join(d, 'train-%05d-of-00002' % shard),
                        [_example(_jpeg(180 + 16 * i, 260 - 8 * i, 10 * shard + i), 1 + 10 * shard + i,
                                  [[0.2, 0.2, 0.9, 0.8]] if i % 2 else []) for i in range(6)])
    R.write_records(os.path.join(d, 'validation-00000-of-00001'), [_example(_jpeg(300, 280, 77 + i), 500 + i) for i in range(4)])

    class Probe(AbstractLearner):
        def train(self):
            pass

        def evaluate(self):
            pass

    def batches(packed, is_train, n):
        FLAGS.reset()
        FLAGS.data_dir_local, FLAGS.batch_size, FLAGS.batch_size_eval, FLAGS.nb_classes = d, 4, 4, 1001
        FLAGS.buffer_size, FLAGS.nb_threads, FLAGS.prefetch_size = 5, 2, 1
        FLAGS.enbl_device_preprocess = packed
        with G.Graph().as_default():
            ds = D.Ilsvrc12Dataset(is_train)
            ds.batch_size, ds.nb_classes = 4, 1001
            it = ds.build()
            assert isinstance(it, PackedBatchIterator) == packed
            images, labels = it.get_next()
        ex = SimpleNamespace(buf={images: torch.full(images.shape, float('nan')), labels: torch.zeros(labels.shape)})
        out, nbytes = [], []
        for _ in range(n):
            nbytes.append(AbstractLearner.feed(SimpleNamespace(_feed_packed=lambda *a: AbstractLearner._feed_packed(None, *a)),
                                               ex, it))
            out.append((ex.buf[images].numpy().copy(), ex.buf[labels].numpy().copy()))
        return out, nbytes
    for is_train in (True, False):
        host, host_bytes = batches(False, is_train, 5)
        dev, dev_bytes = batches(True, is_train, 5)
        for (hi, hl), (di, dl) in zip(host, dev):
            np.testing.assert_array_equal(dl, hl)
            np.testing.assert_array_equal(di, hi)
        assert all(b == 4 * 224 * 224 * 3 * 4 + 4 * 1001 * 4 for b in host_bytes)
        assert all(b < 0.6 * host_bytes[0] for b in dev_bytes)                  # uint8 crops: far fewer bytes per step
    FLAGS.reset()


def _digest(a):
    import hashlib
    a = np.ascontiguousarray(np.asarray(a, np.float32))
    return list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()


def test_preprocessing_matches_the_executed_reference_source():
    """utils/external/imagenet_preprocessing.py:preprocess_image was executed (tests/golden/make_golden_from_reference.py)
    with numpy stand-ins for the TensorFlow ops — Pillow for the JPEG decode, this repo's resize_bilinear for
    tf.image.resize_images, controlled draws for the random ops.  Pinned here: the resize target arithmetic, crop
    offsets, crop -> flip -> resize order, mean subtraction, and the sample_distorted_bounding_box parameters."""
    import base64
    import json
    from pocketflow_b200.datasets import ilsvrc12_dataset as D
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'ref_executed_v1.json')))
    assert gold['imagenet_constants'] == dict(means=[float(np.float64(v)) for v in (123.68, 116.78, 103.94)], resize_min=D.RESIZE_MIN)
    assert [float(v) for v in D.CHANNEL_MEANS] == [float(np.float32(v)) for v in gold['imagenet_constants']['means']]
    import inspect
    sig = inspect.signature(D.sample_distorted_bounding_box).parameters
    for rec in gold['imagenet_preprocessing']:
        jpeg = base64.b64decode(rec['jpeg_b64'])
        assert D.jpeg_shape(jpeg) == (rec['height'], rec['width'])
        out = D.preprocess_image(jpeg, np.zeros((0, 4), np.float32), False)
        assert _digest(out) == (rec['eval']['shape'], rec['eval']['sha256'])
        (name, size, method, align), = rec['eval_calls']
        assert name == 'resize_images' and tuple(size) == D.smallest_size_at_least(rec['height'], rec['width'])
        assert method == 'BILINEAR' and align is False
        for tr in rec['train']:
            cy, cx, ch, cw = tr['crop']
            # the same crop and flip draw, through this repo's functions
            crop = D.decode_jpeg(jpeg, (cy, cx, ch, cw))
            img = D.resize_bilinear(crop[:, ::-1] if tr['flip'] else crop, 224, 224) - D.CHANNEL_MEANS
            assert _digest(img) == (tr['out']['shape'], tr['out']['sha256'])
            d = np.zeros((), D.IMG_DESC)
            d['h'], d['w'], d['rh'], d['rw'], d['flip'] = ch, cw, 224, 224, int(tr['flip'])
            assert _digest(D.preprocess_from_descriptor(crop, d)) == (tr['out']['shape'], tr['out']['sha256'])
            sdbb = tr['calls'][0]
            assert sdbb[0] == 'sample_distorted_bounding_box' and sdbb[1]['use_image_if_no_bounding_boxes'] is True
            assert sdbb[1]['min_object_covered'] == sig['min_object_covered'].default
            assert tuple(sdbb[1]['aspect_ratio_range']) == sig['aspect_ratio_range'].default
            assert tuple(sdbb[1]['area_range']) == sig['area_range'].default
            assert sdbb[1]['max_attempts'] == sig['max_attempts'].default
            assert tr['calls'][1] == ['resize_images', [224, 224], 'BILINEAR', False]


def test_example_and_record_round_trip_property(tmp_path):
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import given, settings, strategies as st, HealthCheck
    key = st.text(alphabet=st.sampled_from(list('abcdefghijklmnopqrstuvwxyz/_')), min_size=1, max_size=20)
    value = st.one_of(st.lists(st.binary(max_size=300), max_size=3),
                      st.lists(st.integers(-2 ** 62, 2 ** 62), max_size=6).map(lambda v: np.asarray(v, np.int64)),
                      st.lists(st.floats(width=32, allow_nan=False), max_size=6).map(lambda v: np.asarray(v, np.float32)))
    counter = {'n': 0}

    @settings(max_examples=50, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(st.dictionaries(key, value, max_size=5), min_size=0, max_size=6))
    def check(examples):
        counter['n'] += 1
        path = str(tmp_path / ('f%d.tfrecord' % counter['n']))
        R.write_records(path, [R.encode_example(e) for e in examples])
        back = [R.parse_example(rec) for rec in R.read_records(path)]
        assert len(back) == len(examples)
        for want, got in zip(examples, back):
            assert sorted(got) == sorted(want)
            for k, v in want.items():
                if isinstance(v, list):
                    assert got[k] == v
                else:
                    assert got[k].dtype == v.dtype and got[k].tolist() == v.tolist()
    check()


def test_make_tfrecords_tool_feeds_the_dataset(tmp_path, capsys):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('make_tfrecords', os.path.join(root, 'tools', 'make_tfrecords.py'))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    src = tmp_path / 'jpegs'
    for c, cls in enumerate(['n01', 'n02', 'n03']):
        (src / cls).mkdir(parents=True)
        for i in range(4):
            (src / cls / ('img_%d.JPEG' % i)).write_bytes(_jpeg(64 + 8 * i, 80, 10 * c + i))
    out = str(tmp_path / 'records')
    assert tool.main([str(src), out, 'train', '2']) == 0
    assert '12 images of 3 classes in 2 shards' in capsys.readouterr().out
    assert sorted(os.listdir(out)) == ['train-00000-of-00002', 'train-00001-of-00002']
    labels = []
    for f in sorted(os.listdir(out)):
        for rec in R.read_records(os.path.join(out, f)):
            ex = R.parse_example(rec)
            assert ex['image/encoded'][0][:2] == b'\xff\xd8' and ex['image/class/text'][0] in (b'n01', b'n02', b'n03')
            labels.append(int(ex['image/class/label'][0]))
    assert sorted(labels) == [1] * 4 + [2] * 4 + [3] * 4
