#!/usr/bin/env python
"""Inspect and convert checkpoints: TensorFlow V2 bundles (what the reference's savers and model archives hold) and
this build's .npz files.

  python tools/ckpt_tool.py list  models/model.ckpt-1000          # names, dtypes, shapes (either format)
  python tools/ckpt_tool.py to-tf models/model.ckpt-1000.npz out/model.ckpt
  python tools/ckpt_tool.py to-npz models/model.ckpt out/model.ckpt   # writes out/model.ckpt.npz
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pocketflow_b200.utils import tf_bundle  # noqa: E402


def read_any(path):
    """{name without ':0': array} from a .npz written by the learners or from a bundle prefix."""
    if path.endswith('.npz'):
        d = np.load(path)
        out = {}
        for k in d.files:
            name = k.replace('|', '/')
            out[name[:-2] if name.endswith(':0') else name] = d[k]
        return out
    return tf_bundle.load(path)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('command', choices=['list', 'to-tf', 'to-npz'])
    ap.add_argument('src')
    ap.add_argument('dst', nargs='?')
    a = ap.parse_args(argv)
    tensors = read_any(a.src)
    if a.command == 'list':
        total = 0
        for name in sorted(tensors):
            t = tensors[name]
            total += t.size
            print('%-72s %-8s %s' % (name, t.dtype, tuple(t.shape)))
        print('%d tensors, %d values' % (len(tensors), total))
        return 0
    if a.dst is None:
        ap.error('%s needs a destination' % a.command)
    os.makedirs(os.path.dirname(os.path.abspath(a.dst)), exist_ok=True)
    if a.command == 'to-tf':
        print('wrote ' + tf_bundle.save(a.dst, tensors))
    else:
        fn = a.dst if a.dst.endswith('.npz') else a.dst + '.npz'
        np.savez(fn, **{(k + ':0').replace('/', '|'): v for k, v in tensors.items()})
        print('wrote ' + fn)
    return 0


if __name__ == '__main__':
    sys.exit(main())
