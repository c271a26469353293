"""TensorFlow V2 checkpoint ("tensor bundle") interchange — SURVEY.md §8(f) rank 1.

The reader is checked against (i) published CRC-32C known answers, (ii) an index file assembled byte by byte in this
test from the format description (independent of the writer), (iii) writer -> reader round trips at block boundaries,
(iv) corruption: every flipped byte must raise, never return wrong numbers."""
import os
import struct

import numpy as np
import pytest

from pocketflow_b200.utils import tf_bundle as B


def _crc_bitwise(data, crc=0):
    crc ^= 0xffffffff
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc ^ 0xffffffff


def test_crc32c_known_answers():
    # RFC 3720 B.4 and the classic check value
    assert B.crc32c(b'123456789') == 0xE3069283
    assert B.crc32c(bytes(32)) == 0x8A9136AA
    assert B.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert B.crc32c(bytes(range(32))) == 0x46DD794E
    assert B.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert B.crc32c(b'') == 0


def test_crc32c_lanes_agree_with_the_bitwise_definition_and_extend():
    rng = np.random.RandomState(0)
    for n in (1, 255, 511, 512, 513, 4097, 70001, 300007):
        d = rng.randint(0, 256, n).astype(np.uint8).tobytes()
        want = _crc_bitwise(d) if n < 80000 else None
        got = B.crc32c(d)
        if want is not None:
            assert got == want, n
        k = n // 3
        assert B.crc32c(d[k:], B.crc32c(d[:k])) == got, n            # Extend
    big = rng.randint(0, 256, 3 << 20).astype(np.uint8)
    assert B.crc32c(big) == B.crc32c(big[1 << 20:], B.crc32c(big[:1 << 20]))


def test_crc_mask_is_the_documented_rotation():
    for c in (0, 1, 0xffffffff, 0xE3069283, 0x12345678):
        m = B.mask_crc(c)
        assert m == ((((c >> 15) | (c << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff
        assert B.unmask_crc(m) == c
    assert B.mask_crc(0) == 0xa282ead8


def _hand_block(entries, restarts):
    """entries: [(shared, key_delta, value)] exactly as they lie in the file."""
    out = b''
    for shared, delta, value in entries:
        out += bytes([shared, len(delta), len(value)]) + delta + value        # all lengths < 128: one-byte varints
    for r in restarts:
        out += struct.pack('<I', r)
    return out + struct.pack('<I', len(restarts))


def _with_trailer(block):
    return block + b'\0' + struct.pack('<I', B.mask_crc(_crc_bitwise(block + b'\0')))


def test_reader_on_an_index_assembled_by_hand(tmp_path):
    """Two variables, written out from the format description alone (prefix compression, restart array, block
    trailers, index block, empty meta-index block, footer), then read back through BundleReader."""
    a = np.arange(6, dtype='<f4').reshape(2, 3)
    s = np.array(7, dtype='<i8')
    data = a.tobytes() + s.tobytes()
    # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    header = bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])
    # BundleEntryProto{dtype: DT_FLOAT, shape{dim{size:2} dim{size:3}}, size: 24, crc32c}
    ea = bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03, 0x28, 24, 0x35]) + \
        struct.pack('<I', B.mask_crc(_crc_bitwise(a.tobytes())))
    # BundleEntryProto{dtype: DT_INT64, shape{}, offset: 24, size: 8, crc32c}
    es = bytes([0x08, 0x09, 0x12, 0x00, 0x20, 24, 0x28, 8, 0x35]) + struct.pack('<I', B.mask_crc(_crc_bitwise(s.tobytes())))
    # keys "", "model/a", "model/step": the third shares the 6-byte prefix "model/" with the second
    data_block = _hand_block([(0, b'', header), (0, b'model/a', ea), (6, b'step', es)], [0])
    f = _with_trailer(data_block)
    meta_off = len(f)
    meta_block = _hand_block([], [0])
    f += _with_trailer(meta_block)
    index_off = len(f)
    index_block = _hand_block([(0, b'model/step', bytes([0, len(data_block)]))], [0])
    f += _with_trailer(index_block)
    footer = bytes([meta_off, len(meta_block), index_off, len(index_block)])
    footer += bytes(40 - len(footer)) + bytes([0x57, 0xfb, 0x80, 0x8b, 0x24, 0x75, 0x47, 0xdb])
    f += footer
    prefix = str(tmp_path / 'model.ckpt')
    open(prefix + '.index', 'wb').write(f)
    open(prefix + '.data-00000-of-00001', 'wb').write(data)
    r = B.BundleReader(prefix)
    assert r.header == {'num_shards': 1, 'endianness': 0, 'producer': 1, 'min_consumer': 0}
    assert r.keys() == ['model/a', 'model/step']
    assert r.shape_and_dtype('model/a') == ((2, 3), np.float32)
    np.testing.assert_array_equal(r.get_tensor('model/a'), a)
    got = r.get_tensor('model/step')
    assert got.shape == () and got.dtype == np.int64 and int(got) == 7
    # and the writer produces these very bytes for the same content
    w = B.BundleWriter(str(tmp_path / 'again.ckpt'))
    w.add('model/a', a)
    w.add('model/step', s)
    w.finish()
    assert open(str(tmp_path / 'again.ckpt.index'), 'rb').read() == f
    assert open(str(tmp_path / 'again.ckpt.data-00000-of-00001'), 'rb').read() == data


def _random_tensors(rng, n, prefix='model/resnet_model/'):
    out = {}
    for i in range(n):
        kind = i % 5
        name = '%s%s_%d/%s' % (prefix, ('conv2d', 'batch_normalization', 'dense')[i % 3], i, ('kernel', 'gamma', 'Adam_1')[i % 3])
        if kind == 0:
            out[name] = rng.randn(3, 3, 1 + i % 7, 2 + i % 5).astype(np.float32)
        elif kind == 1:
            out[name] = rng.randn(1 + i % 9).astype(np.float32)
        elif kind == 2:
            out[name] = np.asarray(rng.randint(-5, 5), np.int64)
        elif kind == 3:
            out[name] = rng.randint(0, 2, (4, i % 3)).astype(np.bool_)          # includes empty tensors
        else:
            out[name] = rng.randn(2, 2).astype(np.float64)
    return out


@pytest.mark.parametrize('n,block', [(1, 262144), (40, 262144), (40, 64), (300, 512)])
def test_round_trip_over_single_and_many_blocks(tmp_path, n, block, monkeypatch):
    rng = np.random.RandomState(n)
    t = _random_tensors(rng, n)
    real = B.build_table
    monkeypatch.setattr(B, 'build_table', lambda items: real(items, block_size=block))
    prefix = B.save(str(tmp_path / 'm' / 'model.ckpt'), t, global_step=17)
    assert prefix.endswith('model.ckpt-17')
    assert B.latest_checkpoint(str(tmp_path / 'm')) == prefix
    back = B.load(prefix)
    assert sorted(back) == sorted(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and back[k].shape == t[k].shape, k
        np.testing.assert_array_equal(back[k], t[k])


def test_checkpoint_state_file_keeps_history(tmp_path):
    d = str(tmp_path)
    for step in (1, 2, 3):
        B.save(os.path.join(d, 'model.ckpt'), {'v': np.float32(step)}, global_step=step)
    latest, hist = B.read_checkpoint_state(d)
    assert latest == 'model.ckpt-3' and hist == ['model.ckpt-1', 'model.ckpt-2', 'model.ckpt-3']
    text = open(os.path.join(d, 'checkpoint')).read().splitlines()
    assert text[0] == 'model_checkpoint_path: "model.ckpt-3"'
    assert text[1:] == ['all_model_checkpoint_paths: "model.ckpt-%d"' % s for s in (1, 2, 3)]
    assert float(B.load(B.latest_checkpoint(d))['v']) == 3.0


def test_any_corruption_raises(tmp_path):
    rng = np.random.RandomState(3)
    prefix = B.save(str(tmp_path / 'model.ckpt'), _random_tensors(rng, 12))
    index = bytearray(open(prefix + '.index', 'rb').read())
    # every byte the format uses: all of the file except the meta-index block (never read by the bundle reader), its
    # handle and the footer padding
    footer = bytes(index[-48:])
    moff, p = B.get_varint(footer, 0)
    msize, p_meta_end = B.get_varint(footer, p)
    _, p = B.get_varint(footer, p_meta_end)
    _, p_handles_end = B.get_varint(footer, p)
    unused = set(range(moff, moff + msize + 5)) | set(range(len(index) - 48, len(index) - 48 + p_meta_end)) | \
        set(range(len(index) - 48 + p_handles_end, len(index) - 8))
    for pos in range(len(index)):
        if pos in unused:
            continue
        bad = bytearray(index)
        bad[pos] ^= 0x40
        open(prefix + '.index', 'wb').write(bad)
        with pytest.raises((ValueError, NotImplementedError, KeyError, UnicodeDecodeError)):
            B.BundleReader(prefix).tensors()
    open(prefix + '.index', 'wb').write(index)
    B.BundleReader(prefix).tensors()
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[len(data) // 2] ^= 1
    open(prefix + '.data-00000-of-00001', 'wb').write(data)
    with pytest.raises(ValueError, match='CRC-32C mismatch'):
        B.BundleReader(prefix).tensors()
    open(prefix + '.data-00000-of-00001', 'wb').write(data[:-3])               # truncated data file
    with pytest.raises(ValueError):
        B.BundleReader(prefix, verify=False).tensors()


def test_unsupported_content_is_refused_not_guessed(tmp_path):
    prefix = str(tmp_path / 'model.ckpt')
    w = B.BundleWriter(prefix)
    with pytest.raises(TypeError):
        w.add('s', np.array(['a', 'b']))
    w.add('v', np.zeros(3, np.float32))
    with pytest.raises(ValueError):
        w.add('v', np.zeros(3, np.float32))
    w.finish()
    # a snappy-compressed block type must be reported as such
    f = bytearray(open(prefix + '.index', 'rb').read())
    f[_first_trailer(f)] = 1
    open(prefix + '.index', 'wb').write(f)
    with pytest.raises(NotImplementedError, match='compressed'):
        B.BundleReader(prefix, verify=False)


def _first_trailer(f):
    """offset of the compression-type byte of the first data block = its size, read from the index block."""
    footer = bytes(f[-48:])
    pos = 0
    _, pos = B.get_varint(footer, pos)
    _, pos = B.get_varint(footer, pos)
    ioff, pos = B.get_varint(footer, pos)
    isize, pos = B.get_varint(footer, pos)
    (_, hv), = list(B._block_entries(bytes(f[ioff:ioff + isize])))
    off, p = B.get_varint(hv, 0)
    size, _ = B.get_varint(hv, p)
    return off + size


def test_learner_checkpoint_helpers_speak_both_formats(tmp_path):
    """save_checkpoint / latest_checkpoint / load_checkpoint (the Saver calls of the reference's learners,
    learners/full_precision/learner.py:172-186) with the variable names of a real graph."""
    import time
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.abstract_learner import save_checkpoint, load_checkpoint, latest_checkpoint
    from pocketflow_b200.nets import resnet_at_cifar10 as R
    FLAGS.reset()
    FLAGS.batch_size, FLAGS.resnet_size = 2, 8
    mh, gr = R.ModelHelper(), G.Graph()
    with gr.as_default():
        with G.variable_scope('data'):
            im, _ = mh.build_dataset_train().get_next()
        with G.variable_scope('model'):
            mh.forward_train(im)
    rng = np.random.RandomState(0)
    state = {v.name: rng.randn(*v.shape).astype(np.float32) for v in gr.variables.values()}
    d = str(tmp_path / 'models')
    path = os.path.join(d, 'model.ckpt')
    fn_npz = save_checkpoint(path, state, 10)
    assert fn_npz.endswith('model.ckpt-10.npz') and latest_checkpoint(d) == fn_npz
    time.sleep(0.02)
    FLAGS.ckpt_format = 'tf'
    fn_tf = save_checkpoint(path, state, 20)
    assert fn_tf.endswith('model.ckpt-20') and latest_checkpoint(d) == fn_tf
    names = B.BundleReader(fn_tf).keys()
    assert 'global_step' in names and 'model/resnet_model/conv2d/kernel' in names
    assert 'model/resnet_model/batch_normalization/moving_variance' in names and not any(':' in n for n in names)
    for fn in (fn_npz, fn_tf):
        back = load_checkpoint(fn)
        for k, v in state.items():
            np.testing.assert_array_equal(back[k], v)
    assert int(load_checkpoint(fn_tf)['global_step:0']) == 20
    # an archive whose state file names the trainer's absolute path still resolves next to the state file
    open(os.path.join(d, 'checkpoint'), 'w').write('model_checkpoint_path: "/somewhere/else/model.ckpt-20"\n')
    assert latest_checkpoint(d) == fn_tf
    FLAGS.ckpt_format = 'zip'
    with pytest.raises(ValueError):
        save_checkpoint(path, state, 30)
    FLAGS.reset()


def test_ckpt_tool_converts_both_ways(tmp_path, capsys):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'ckpt_tool', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'ckpt_tool.py'))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.abstract_learner import save_checkpoint, load_checkpoint
    FLAGS.reset()
    rng = np.random.RandomState(5)
    state = {'model/conv2d/kernel:0': rng.randn(3, 3, 2, 4).astype(np.float32), 'model/dense/bias:0': rng.randn(10).astype(np.float32)}
    npz = save_checkpoint(str(tmp_path / 'a' / 'model.ckpt'), state, 7)
    assert tool.main(['to-tf', npz, str(tmp_path / 'b' / 'model.ckpt')]) == 0
    assert B.BundleReader(str(tmp_path / 'b' / 'model.ckpt')).keys() == ['model/conv2d/kernel', 'model/dense/bias']
    assert tool.main(['to-npz', str(tmp_path / 'b' / 'model.ckpt'), str(tmp_path / 'c' / 'model.ckpt')]) == 0
    back = load_checkpoint(str(tmp_path / 'c' / 'model.ckpt.npz'))
    assert sorted(back) == sorted(state)
    for k in state:
        np.testing.assert_array_equal(back[k], state[k])
    capsys.readouterr()
    assert tool.main(['list', str(tmp_path / 'b' / 'model.ckpt')]) == 0
    out = capsys.readouterr().out
    assert 'model/conv2d/kernel' in out and '(3, 3, 2, 4)' in out and '2 tensors, 82 values' in out


def test_round_trip_property(tmp_path):
    """Randomised names / dtypes / shapes / block sizes (hypothesis): whatever the writer accepts comes back identical."""
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import given, settings, strategies as st, HealthCheck
    dtypes = [np.float32, np.float64, np.int32, np.int64, np.uint8, np.int8, np.int16, np.bool_, np.float16, np.uint16]
    name = st.text(alphabet=st.sampled_from(list('abcdefghijklmnopqrstuvwxyz_/0123456789')), min_size=1, max_size=40)
    shape = st.lists(st.integers(0, 5), min_size=0, max_size=4)
    entry = st.tuples(name, st.sampled_from(range(len(dtypes))), shape, st.integers(0, 2 ** 31 - 1))
    counter = {'n': 0}

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(entry, min_size=1, max_size=25, unique_by=lambda e: e[0]), st.sampled_from([48, 200, 4096, 262144]),
           st.sampled_from([1, 2, 16]))
    def check(entries, block_size, restart_interval):
        counter['n'] += 1
        tensors = {}
        for nm, di, shp, seed in entries:
            rng = np.random.RandomState(seed)
            tensors[nm] = (rng.randn(*shp) * 50).astype(dtypes[di]) if shp else np.asarray(rng.randn() * 50).astype(dtypes[di])
        prefix = str(tmp_path / ('p%d' % counter['n']) / 'model.ckpt')
        w = B.BundleWriter(prefix)
        for nm in sorted(tensors):
            w.add(nm, tensors[nm])
        real = B.build_table
        B.build_table = lambda items: real(items, block_size=block_size, restart_interval=restart_interval)
        try:
            w.finish()
        finally:
            B.build_table = real
        r = B.BundleReader(prefix)
        assert r.keys() == sorted(tensors)
        for nm, t in tensors.items():
            got = r.get_tensor(nm)
            assert got.dtype == t.dtype and got.shape == t.shape
            np.testing.assert_array_equal(got, t)
    check()
    assert counter['n'] >= 40
