"""Pure-Python reader / writer of TensorFlow's V2 checkpoint format (the "tensor bundle":
``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` + the ``checkpoint`` state file), so that the
pre-trained ``model.ckpt`` files the reference downloads (learners/abstract_learner.py:105-125) can be imported and
the models this build trains can be handed to the reference's own tools (tf.train.Saver.restore at
learners/full_precision/learner.py:172-186, learners/uniform_quantization/learner.py:372-392,
learners/distillation_helper.py:79-82).  SURVEY.md §8(f) rank 1.

TensorFlow is not importable in this image, so the format is restated from its published layout
(tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*, tensorflow/core/protobuf/tensor_bundle.proto):

* ``.index`` is a LevelDB-style sorted string table: prefix-compressed key/value blocks, each followed by a
  5-byte trailer (compression type, masked CRC-32C), an index block of block handles, an (empty) meta-index block and a
  48-byte footer ending in the magic 0xdb4775248b80fb57.
* key ``""`` -> BundleHeaderProto {num_shards, endianness, version}; every other key is a variable name ->
  BundleEntryProto {dtype, shape, shard_id, offset, size, masked crc32c of the bytes}.
* ``.data-SSSSS-of-NNNNN`` holds the raw little-endian tensor bytes at [offset, offset + size).

What is verified here: writer -> reader round trips, CRC-32C known answers, the masked-CRC arithmetic, hand-assembled
index files.  What is NOT: a file written by TensorFlow itself (none exists in this container) — the first import of a
real checkpoint is the acceptance test, and every structural check below raises with the offending offset.
Snappy-compressed blocks (never produced by BundleWriter, which sets kNoCompression) and partitioned variables
(``slices``) are rejected loudly rather than guessed at.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
HEADER_KEY = b''

# tensorflow/core/framework/types.proto
DT_TO_NP = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
            10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
NP_TO_DT = {np.dtype(v): k for k, v in DT_TO_NP.items()}


# ----------------------------------------------------------------------------------------------- CRC-32C
def _make_table():
    poly = 0x82F63B78                      # Castagnoli, reflected
    t = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (poly if c & 1 else 0)
        t[i] = c
    return t


_T = _make_table()
_TL = [int(v) for v in _T]


def _raw_update(state, data):
    """The CRC register after `data` (bytes-like), starting from `state`; no pre/post inversion."""
    tl = _TL
    for b in data:
        state = tl[(state ^ b) & 0xff] ^ (state >> 8)
    return state


def _zero_operator(nbytes):
    """Columns of the GF(2) matrix of 'shift the register through nbytes zero bytes' (square and multiply)."""
    def apply(cols, v):
        r, j = 0, 0
        while v:
            if v & 1:
                r ^= cols[j]
            v >>= 1
            j += 1
        return r
    one = [_raw_update(1 << j, b'\0') for j in range(32)]
    result = [1 << j for j in range(32)]
    power = one
    while nbytes:
        if nbytes & 1:
            result = [apply(power, c) for c in result]
        power = [apply(power, c) for c in power]
        nbytes >>= 1
    return result


LANE_BYTES = 256
CHUNK_BYTES = 1 << 21
_FOLD = []          # byte-sliced tables of the 'LANE_BYTES zero bytes' operator, built on first use


def _fold_tables():
    if not _FOLD:
        cols = _zero_operator(LANE_BYTES)
        tabs = np.zeros((4, 256), np.uint32)
        for k in range(4):
            for j in range(8):
                bit = 1 << j
                tabs[k, bit:2 * bit] = tabs[k, :bit] ^ np.uint32(cols[8 * k + j])
        _FOLD.extend([int(v) for v in tabs[k]] for k in range(4))
    return _FOLD


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli) of `data`, continuing from a previous value `crc` (tensorflow/core/lib/hash/crc32c.h
    Extend/Value).  Inputs of more than a few lanes run as LANE_BYTES-long lanes in lock-step with numpy and are
    folded with the zero-shift operator (the register update is GF(2)-linear in (state, data))."""
    buf = np.frombuffer(memoryview(data).cast('B'), np.uint8) if not isinstance(data, np.ndarray) \
        else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    state = (crc ^ 0xffffffff) & 0xffffffff
    while buf.size > CHUNK_BYTES:                              # cache-sized pieces, chained through the register
        state = crc32c(buf[:CHUNK_BYTES], state ^ 0xffffffff) ^ 0xffffffff
        buf = buf[CHUNK_BYTES:]
    lanes = buf.size // LANE_BYTES
    if lanes >= 4:
        body = np.ascontiguousarray(buf[:lanes * LANE_BYTES].reshape(lanes, LANE_BYTES).T)
        st = np.zeros(lanes, np.uint32)
        st[0] = state
        for j in range(LANE_BYTES):
            st = _T[(st ^ body[j]) & 0xff] ^ (st >> 8)
        t0, t1, t2, t3 = _fold_tables()
        lane_states = st.tolist()
        acc = lane_states[0]
        for v in lane_states[1:]:
            acc = t0[acc & 0xff] ^ t1[(acc >> 8) & 0xff] ^ t2[(acc >> 16) & 0xff] ^ t3[acc >> 24] ^ v
        state = acc
        buf = buf[lanes * LANE_BYTES:]
    state = _raw_update(state, buf.tobytes())
    return state ^ 0xffffffff


def mask_crc(crc):
    """crc32c::Mask — CRCs stored next to the data they cover are rotated and offset."""
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - 0xa282ead8) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ----------------------------------------------------------------------------------------------- varints / protobuf
def put_varint(n):
    if n < 0:
        n += 1 << 64                       # protobuf int64: two's complement, ten bytes
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7f) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def get_varint(buf, pos):
    shift, val = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError('truncated varint at offset %d' % pos)
        b = buf[pos]
        pos += 1
        val |= (b & 0x7f) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise ValueError('varint longer than 10 bytes at offset %d' % pos)


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def proto_fields(buf):
    """[(field number, wire type, value)] of one serialized message (varint / fixed64 / bytes / fixed32)."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = get_varint(buf, pos)
        elif wt == 1:
            v, pos = struct.unpack_from('<Q', buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = get_varint(buf, pos)
            if pos + ln > len(buf):
                raise ValueError('length-delimited field %d overruns the message' % num)
            v, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            v, pos = struct.unpack_from('<I', buf, pos)[0], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d (field %d)' % (wt, num))
        out.append((num, wt, v))
    return out


def _f_varint(num, v):
    return put_varint(num << 3) + put_varint(v)


def _f_bytes(num, b):
    return put_varint((num << 3) | 2) + put_varint(len(b)) + b


def encode_header(num_shards=1):
    """BundleHeaderProto: num_shards = 1, endianness = LITTLE (0, omitted), version {producer: 1}."""
    return _f_varint(1, num_shards) + _f_bytes(3, _f_varint(1, 1))


def decode_header(buf):
    h = {'num_shards': 0, 'endianness': 0, 'producer': 0, 'min_consumer': 0}
    for num, _, v in proto_fields(buf):
        if num == 1:
            h['num_shards'] = v
        elif num == 2:
            h['endianness'] = v
        elif num == 3:
            for n2, _, v2 in proto_fields(v):
                if n2 == 1:
                    h['producer'] = v2
                elif n2 == 2:
                    h['min_consumer'] = v2
    return h


def encode_entry(dtype, shape, shard_id, offset, size, masked_crc):
    """BundleEntryProto; proto3 omits zero scalars, the (possibly empty) shape message is always present."""
    dims = b''.join(_f_bytes(2, _f_varint(1, int(d))) for d in shape)
    out = _f_varint(1, dtype) + _f_bytes(2, dims)
    if shard_id:
        out += _f_varint(3, shard_id)
    if offset:
        out += _f_varint(4, offset)
    if size:
        out += _f_varint(5, size)
    if masked_crc:
        out += put_varint((6 << 3) | 5) + struct.pack('<I', masked_crc)
    return out


def decode_entry(buf):
    e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': 0, 'slices': 0,
         'unknown_rank': False}
    for num, _, v in proto_fields(buf):
        if num == 1:
            e['dtype'] = v
        elif num == 2:
            dims = []
            for n2, _, v2 in proto_fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in proto_fields(v2):
                        if n3 == 1:
                            size = _signed64(v3)
                    dims.append(size)
                elif n2 == 3:
                    e['unknown_rank'] = bool(v2)
            e['shape'] = tuple(dims)
        elif num == 3:
            e['shard_id'] = v
        elif num == 4:
            e['offset'] = _signed64(v)
        elif num == 5:
            e['size'] = _signed64(v)
        elif num == 6:
            e['crc32c'] = v
        elif num == 7:
            e['slices'] += 1
    return e


# ----------------------------------------------------------------------------------------------- sorted string table
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b''

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            lim = min(len(key), len(self.last_key))
            while shared < lim and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value))
        self.buf += key[shared:] + value
        self.last_key = key
        self.counter += 1

    def size(self):
        return len(self.buf) + 4 * (len(self.restarts) + 1)

    def empty(self):
        return not self.buf

    def finish(self):
        return bytes(self.buf) + struct.pack('<%dI' % len(self.restarts), *self.restarts) + \
            struct.pack('<I', len(self.restarts))


def _handle(offset, size):
    return put_varint(offset) + put_varint(size)


def build_table(items, block_size=262144, restart_interval=16):
    """Serialized table holding `items` = [(key bytes, value bytes)] in strictly increasing key order."""
    out = bytearray()

    def write_block(contents):
        h = (len(out), len(contents))
        trailer_type = b'\0'                                         # kNoCompression
        out.extend(contents + trailer_type + struct.pack('<I', mask_crc(crc32c(contents + trailer_type))))
        return h

    index = _BlockBuilder(1)
    data = _BlockBuilder(restart_interval)
    prev = None
    for key, value in items:
        if prev is not None and not key > prev:
            raise ValueError('table keys must be strictly increasing: %r after %r' % (key, prev))
        data.add(key, value)
        prev = key
        if data.size() >= block_size:
            index.add(prev, _handle(*write_block(data.finish())))
            data = _BlockBuilder(restart_interval)
    if not data.empty():
        index.add(prev, _handle(*write_block(data.finish())))
    meta_h = write_block(_BlockBuilder(restart_interval).finish())
    index_h = write_block(index.finish())
    footer = _handle(*meta_h) + _handle(*index_h)
    footer += b'\0' * (FOOTER_LEN - 8 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


def _read_block(buf, offset, size, verify):
    end = offset + size + BLOCK_TRAILER_LEN
    if offset < 0 or end > len(buf):
        raise ValueError('block handle (%d, %d) outside the %d-byte table' % (offset, size, len(buf)))
    contents, ctype = buf[offset:offset + size], buf[offset + size]
    if verify:
        want = unmask_crc(struct.unpack_from('<I', buf, offset + size + 1)[0])
        got = crc32c(buf[offset:offset + size + 1])
        if want != got:
            raise ValueError('block at %d: CRC-32C mismatch (stored %08x, computed %08x)' % (offset, want, got))
    if ctype != 0:
        raise NotImplementedError('block at %d is compressed (type %d); BundleWriter writes uncompressed tables'
                                  % (offset, ctype))
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise ValueError('block shorter than its restart count')
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 * (n_restarts + 1)
    if end < 0:
        raise ValueError('block restart array (%d entries) larger than the block' % n_restarts)
    pos, key = 0, b''
    while pos < end:
        shared, pos = get_varint(block, pos)
        unshared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > end:
            raise ValueError('corrupt block entry at offset %d' % pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(buf, verify=True):
    """[(key, value)] of a serialized table, in file order."""
    if len(buf) < FOOTER_LEN:
        raise ValueError('table of %d bytes has no footer' % len(buf))
    footer = buf[len(buf) - FOOTER_LEN:]
    if struct.unpack_from('<Q', footer, FOOTER_LEN - 8)[0] != TABLE_MAGIC:
        raise ValueError('not a TensorFlow table file (bad magic number)')
    pos = 0
    _, pos = get_varint(footer, pos)          # meta-index handle: unused by the bundle format
    _, pos = get_varint(footer, pos)
    ioff, pos = get_varint(footer, pos)
    isize, pos = get_varint(footer, pos)
    items = []
    for _, hv in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p = get_varint(hv, 0)
        bsize, _ = get_varint(hv, p)
        items.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return items


# ----------------------------------------------------------------------------------------------- bundle
def data_filename(prefix, shard, num_shards):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


class BundleWriter:
    """tensorflow::BundleWriter: ``add`` tensors, then ``finish``.  One shard, little-endian."""

    def __init__(self, prefix):
        self.prefix = prefix
        os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
        self._tmp = data_filename(prefix, 0, 1) + '.tempstate'
        self._data = open(self._tmp, 'wb')
        self._offset = 0
        self._entries = {}

    def add(self, name, array):
        key = name.encode('utf-8')
        if not key or key in self._entries:
            raise ValueError('empty or duplicate tensor name %r' % name)
        a = np.asarray(array)
        if a.dtype not in NP_TO_DT:
            raise TypeError('%s: dtype %s has no checkpoint representation here' % (name, a.dtype))
        raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()      # C order, little-endian
        self._entries[key] = encode_entry(NP_TO_DT[np.dtype(a.dtype.name)], a.shape, 0, self._offset, len(raw),
                                          mask_crc(crc32c(raw)))
        self._data.write(raw)
        self._offset += len(raw)

    def finish(self):
        self._data.close()
        os.replace(self._tmp, data_filename(self.prefix, 0, 1))
        items = [(HEADER_KEY, encode_header(1))] + sorted(self._entries.items())
        tmp = self.prefix + '.index.tempstate'
        with open(tmp, 'wb') as f:
            f.write(build_table(items))
        os.replace(tmp, self.prefix + '.index')


class BundleReader:
    """tensorflow::BundleReader / tf.train.load_checkpoint: ``keys()``, ``shape_and_dtype(name)``, ``get_tensor``."""

    def __init__(self, prefix, verify=True):
        self.prefix = prefix
        self.verify = verify
        with open(prefix + '.index', 'rb') as f:
            items = read_table(f.read(), verify)
        if not items or items[0][0] != HEADER_KEY:
            raise ValueError('%s.index: first entry is not the bundle header' % prefix)
        self.header = decode_header(items[0][1])
        if self.header['endianness'] != 0:
            raise NotImplementedError('big-endian checkpoint')
        if self.header['min_consumer'] > 1:
            raise NotImplementedError('checkpoint needs bundle reader version %d' % self.header['min_consumer'])
        if self.header['num_shards'] < 1:
            raise ValueError('bundle header declares %d shards' % self.header['num_shards'])
        self.entries = {k.decode('utf-8'): decode_entry(v) for k, v in items[1:]}
        self._files = {}

    def keys(self):
        return sorted(self.entries)

    def has_tensor(self, name):
        return name in self.entries

    def shape_and_dtype(self, name):
        e = self.entries[name]
        return e['shape'], DT_TO_NP.get(e['dtype'])

    def _shard(self, shard_id):
        if shard_id not in self._files:
            if not 0 <= shard_id < self.header['num_shards']:
                raise ValueError('shard %d of %d' % (shard_id, self.header['num_shards']))
            path = data_filename(self.prefix, shard_id, self.header['num_shards'])
            # (a shard that holds only empty tensors is a zero-byte file, which cannot be mapped)
            self._files[shard_id] = np.memmap(path, dtype=np.uint8, mode='r') if os.path.getsize(path) else \
                np.zeros(0, np.uint8)
        return self._files[shard_id]

    def get_tensor(self, name):
        if name not in self.entries:
            raise KeyError('%s: no tensor %r (has %d tensors)' % (self.prefix, name, len(self.entries)))
        e = self.entries[name]
        if e['slices']:
            raise NotImplementedError('%s is a partitioned variable (%d slices)' % (name, e['slices']))
        if e['dtype'] not in DT_TO_NP:
            raise NotImplementedError('%s: DataType %d is not a fixed-size numeric type' % (name, e['dtype']))
        dt = np.dtype(DT_TO_NP[e['dtype']])
        count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if e['unknown_rank'] or count * dt.itemsize != e['size']:
            raise ValueError('%s: %d bytes stored for shape %s of %s' % (name, e['size'], e['shape'], dt))
        shard = self._shard(e['shard_id'])
        if e['offset'] < 0 or e['offset'] + e['size'] > shard.size:
            raise ValueError('%s: bytes [%d, +%d) outside the %d-byte data file' % (name, e['offset'], e['size'],
                                                                                 shard.size))
        raw = np.array(shard[e['offset']:e['offset'] + e['size']])
        if self.verify and e['crc32c']:
            got, want = crc32c(raw), unmask_crc(e['crc32c'])
            if got != want:
                raise ValueError('%s: CRC-32C mismatch (stored %08x, computed %08x)' % (name, want, got))
        return raw.view(dt.newbyteorder('<')).astype(dt, copy=False).reshape(e['shape'])

    def tensors(self):
        return {k: self.get_tensor(k) for k in self.keys()}


# ----------------------------------------------------------------------------------------------- Saver-level helpers
def _quote(s):
    return '"' + s.replace('\\', '\\\\').replace('"', '\\"') + '"'


def update_checkpoint_state(save_dir, prefix_path, keep=5):
    """The text-format CheckpointState file tf.train.Saver maintains next to the bundles ('checkpoint')."""
    state = os.path.join(save_dir, 'checkpoint')
    rel = os.path.basename(prefix_path) if os.path.dirname(os.path.abspath(prefix_path)) == os.path.abspath(save_dir) \
        else prefix_path
    history = [p for p in read_checkpoint_state(save_dir)[1] if p != rel][-(keep - 1):] if keep > 1 else []
    history.append(rel)
    with open(state + '.tmp', 'w') as f:
        f.write('model_checkpoint_path: %s\n' % _quote(rel))
        for p in history:
            f.write('all_model_checkpoint_paths: %s\n' % _quote(p))
    os.replace(state + '.tmp', state)


def read_checkpoint_state(save_dir):
    """(latest prefix or None, [all prefixes]) as recorded in `save_dir`/checkpoint (paths as written there)."""
    state = os.path.join(save_dir, 'checkpoint')
    latest, history = None, []
    if not os.path.exists(state):
        return latest, history
    with open(state) as f:
        for line in f:
            key, _, val = line.partition(':')
            val = val.strip()
            if len(val) >= 2 and val[0] == '"' and val[-1] == '"':
                val = val[1:-1].replace('\\"', '"').replace('\\\\', '\\')
            if key.strip() == 'model_checkpoint_path':
                latest = val
            elif key.strip() == 'all_model_checkpoint_paths':
                history.append(val)
    return latest, history


def latest_checkpoint(save_dir):
    """tf.train.latest_checkpoint: the prefix named by the state file, if its index file exists."""
    latest, _ = read_checkpoint_state(save_dir)
    if latest is None:
        return None
    prefix = latest if os.path.isabs(latest) else os.path.join(save_dir, latest)
    if not os.path.exists(prefix + '.index'):
        # archives carry the absolute path of the machine that trained them: look for the same name next to the
        # state file
        prefix = os.path.join(save_dir, os.path.basename(latest))
    return prefix if os.path.exists(prefix + '.index') else None


def save(prefix, tensors, global_step=None):
    """tf.train.Saver.save: write {name: array} as `prefix`[-global_step] and update the state file."""
    if global_step is not None:
        prefix = '%s-%d' % (prefix, global_step)
    w = BundleWriter(prefix)
    for name in sorted(tensors):
        w.add(name, tensors[name])
    w.finish()
    update_checkpoint_state(os.path.dirname(os.path.abspath(prefix)), prefix)
    return prefix


def load(prefix, verify=True):
    return BundleReader(prefix, verify).tensors()
