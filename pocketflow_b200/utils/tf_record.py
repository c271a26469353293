"""TFRecord files and tf.train.Example messages without TensorFlow — what `tf.data.TFRecordDataset` +
`tf.parse_single_example` do for the reference's ILSVRC-12 pipeline (datasets/ilsvrc12_dataset.py:39-76).

Record framing (tensorflow/core/lib/io/record_writer.h):
    uint64 length | uint32 masked_crc32c(length bytes) | byte data[length] | uint32 masked_crc32c(data)
Example (tensorflow/core/example/{example,feature}.proto):
    Example{features = 1}; Features{map<string, Feature> feature = 1}; Feature{oneof: BytesList bytes_list = 1,
    FloatList float_list = 2, Int64List int64_list = 3}, each list `repeated ... value = 1` (packed or not).
Restated from the published format (TensorFlow is absent here); CRC-32C and the varint / protobuf helpers are shared
with utils/tf_bundle.py."""
import struct

import numpy as np

from .tf_bundle import crc32c, mask_crc, unmask_crc, proto_fields, put_varint, get_varint, _signed64


def read_records(path, verify=True):
    """Yield the payload of every record of one TFRecord file; a torn or corrupt record raises.
    verify=True checks both CRCs (what TensorFlow's reader does); 'length' only the 12-byte header's, which still
    catches lost framing and costs nothing — the payload CRC in pure Python runs at ~100 MB/s."""
    with open(path, 'rb') as f:
        pos = 0
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError('%s: truncated record header at byte %d' % (path, pos))
            length, len_crc = struct.unpack('<QI', head)
            if verify and unmask_crc(len_crc) != crc32c(head[:8]):
                raise ValueError('%s: corrupt record length at byte %d' % (path, pos))
            body = f.read(length + 4)
            if len(body) < length + 4:
                raise ValueError('%s: truncated record at byte %d (%d of %d bytes)' % (path, pos, len(body), length + 4))
            data = body[:length]
            if verify is True and unmask_crc(struct.unpack_from('<I', body, length)[0]) != crc32c(data):
                raise ValueError('%s: corrupt record data at byte %d' % (path, pos))
            pos += 16 + length
            yield data


def write_records(path, payloads):
    """Write an iterable of bytes as one TFRecord file (tests, dataset conversion tools)."""
    with open(path, 'wb') as f:
        for data in payloads:
            head = struct.pack('<Q', len(data))
            f.write(head + struct.pack('<I', mask_crc(crc32c(head))) + data + struct.pack('<I', mask_crc(crc32c(data))))


def _decode_list(buf, kind):
    vals = []
    for num, wt, v in proto_fields(buf):
        if num != 1:
            continue
        if kind == 1:                                        # bytes
            vals.append(v)
        elif kind == 2:                                      # float: packed (bytes) or one fixed32 each
            if wt == 2:
                vals.extend(np.frombuffer(v, '<f4').tolist())
            else:
                vals.append(struct.unpack('<f', struct.pack('<I', v))[0])
        else:                                                # int64: packed varints or one varint each
            if wt == 2:
                p = 0
                while p < len(v):
                    x, p = get_varint(v, p)
                    vals.append(_signed64(x))
            else:
                vals.append(_signed64(v))
    if kind == 1:
        return vals
    return np.asarray(vals, np.float32 if kind == 2 else np.int64)


def parse_example(serialized):
    """{feature name: [bytes, ...] | float32 array | int64 array} of one serialized tf.train.Example."""
    out = {}
    for num, _, features in proto_fields(serialized):
        if num != 1:
            continue
        for n2, _, entry in proto_fields(features):
            if n2 != 1:
                continue
            key, feature = None, b''
            for n3, _, v in proto_fields(entry):
                if n3 == 1:
                    key = v.decode('utf-8')
                elif n3 == 2:
                    feature = v
            if key is None:
                raise ValueError('Example feature map entry without a key')
            value = []
            for kind, _, lst in proto_fields(feature):
                if kind in (1, 2, 3):
                    value = _decode_list(lst, kind)
            out[key] = value
    return out


def _ld(num, payload):
    return put_varint((num << 3) | 2) + put_varint(len(payload)) + payload


def encode_example(features):
    """Serialized tf.train.Example from {name: bytes | [bytes] | float array | int array} (packed lists, keys in
    sorted order)."""
    entries = b''
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, list) and all(isinstance(x, (bytes, bytearray)) for x in v):
            feature = _ld(1, b''.join(_ld(1, bytes(x)) for x in v))
        else:
            a = np.asarray(v)
            if a.dtype.kind == 'f':
                feature = _ld(2, _ld(1, a.astype('<f4').tobytes()) if a.size else b'')
            elif a.dtype.kind in 'iub':
                feature = _ld(3, _ld(1, b''.join(put_varint(int(x)) for x in a.reshape(-1))) if a.size else b'')
            else:
                raise TypeError('feature %r: unsupported value type %s' % (key, a.dtype))
        entries += _ld(1, _ld(1, key.encode('utf-8')) + _ld(2, feature))
    return _ld(1, entries)
