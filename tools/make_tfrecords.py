#!/usr/bin/env python
"""Pack a directory tree of JPEG files into the TFRecord shards Ilsvrc12Dataset (and the reference) read:
`<prefix>-00000-of-000NN` holding tf.train.Example messages with 'image/encoded', 'image/class/label' (1-based index of
the sorted class directories; 0 is the background class of the 1001-way head), 'image/class/text' (directory name).

  python tools/make_tfrecords.py /data/ilsvrc12/train /data/tfrecords train 1024
  python tools/make_tfrecords.py /data/ilsvrc12/val   /data/tfrecords validation 128

Layout expected: <root>/<class directory>/<image>.JPEG (what the ImageNet tarballs unpack to for training; sort the
validation images into class directories first).  Files are distributed round-robin after a seeded shuffle."""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pocketflow_b200.utils import tf_record as R  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('root')
    ap.add_argument('out_dir')
    ap.add_argument('prefix', choices=['train', 'validation'])
    ap.add_argument('nb_shards', type=int)
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args(argv)
    classes = sorted(d for d in os.listdir(a.root) if os.path.isdir(os.path.join(a.root, d)))
    if not classes:
        ap.error('no class directories under ' + a.root)
    files = []
    for idx, cls in enumerate(classes):
        for f in sorted(os.listdir(os.path.join(a.root, cls))):
            if f.lower().endswith(('.jpeg', '.jpg')):
                files.append((os.path.join(a.root, cls, f), idx + 1, cls))
    random.Random(a.seed).shuffle(files)
    os.makedirs(a.out_dir, exist_ok=True)
    for shard in range(a.nb_shards):
        def examples(shard=shard):
            for path, label, text in files[shard::a.nb_shards]:
                with open(path, 'rb') as fh:
                    yield R.encode_example({'image/encoded': fh.read(), 'image/class/label': [label],
                                            'image/class/text': text.encode('utf-8')})
        R.write_records(os.path.join(a.out_dir, '%s-%05d-of-%05d' % (a.prefix, shard, a.nb_shards)), examples())
    print('%d images of %d classes in %d shards under %s' % (len(files), len(classes), a.nb_shards, a.out_dir))
    return 0


if __name__ == '__main__':
    sys.exit(main())
