"""Host-side logic of the data-parallel path on CPU: world_size-2 `gloo` process groups exercise the
MultiGpuWrapper surface (utils/multi_gpu_wrapper.py here; /root/reference/utils/multi_gpu_wrapper.py:30-98),
rank-sharded synthetic data, the one-flat-buffer gradient exchange with the Horovod average folded
into grad_scale, and the batch-size dependent schedules."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pf_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from pocketflow_b200.utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
    mgw.init(backend='gloo')
    try:
        ret[rank] = fn(rank, world, mgw)
    finally:
        dist.destroy_process_group()


def run_ranks(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _collectives(rank, world, mgw):
    assert mgw.size() == world and mgw.rank() == rank and mgw.local_rank() == rank
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    mgw.allreduce_flat_(flat)
    p = torch.full((5,), float(rank))
    mgw.broadcast_global_variables([p], 0)
    mgw.barrier()
    return flat.numpy().tolist(), p.numpy().tolist()


def test_allreduce_broadcast_barrier():
    out = run_ranks(_collectives)
    for flat, p in out:
        assert flat == (np.arange(10) * 3.0).tolist()      # sum over ranks 1x + 2x
        assert p == [0.0] * 5                               # rank 0's values everywhere


def _dp_step(rank, world, mgw):
    """Two replicas, different data, one flat all-reduce, optimizer with grad_scale = 1/world:
    replicas stay bit-identical and equal the oracle's update with the averaged gradient."""
    rng = np.random.RandomState(0)
    w0 = rng.randn(64).astype(np.float32)
    acc0 = np.zeros(64, np.float32)
    g_all = [np.random.RandomState(100 + r).randn(64).astype(np.float32) for r in range(world)]
    flat = torch.from_numpy(g_all[rank].copy())
    mgw.allreduce_flat_(flat)
    w1, a1 = O.momentum_step(w0, acc0, flat.numpy(), 0.1, 0.9, wd=1e-4, grad_scale=1.0 / world)
    return w1.tolist(), (g_all[0] + g_all[1]).tolist(), flat.numpy().tolist()


def test_data_parallel_update_is_replicated_and_averaged():
    out = run_ranks(_dp_step)
    assert out[0][0] == out[1][0]                           # replicas in sync, bit for bit
    assert out[0][2] == out[0][1] == out[1][2]              # the collective is a plain sum
    # grad_scale=1/2 of the sum == Horovod's average for power-of-two worlds (exact in fp32)
    g_avg = (np.array(out[0][1], np.float32) / np.float32(2)).astype(np.float32)
    w_ref, _ = O.momentum_step(np.random.RandomState(0).randn(64).astype(np.float32), np.zeros(64, np.float32),
                               g_avg, 0.1, 0.9, wd=1e-4)
    assert np.array_equal(np.array(out[0][0], np.float32), w_ref)


def _sharded_data(rank, world, mgw):
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.datasets.cifar10_dataset import Cifar10Dataset
    from pocketflow_b200 import graph as G
    FLAGS.reset()
    FLAGS.enbl_multi_gpu, FLAGS.batch_size = True, 4
    it = Cifar10Dataset(is_train=True).build()
    img, lab = it.next_batch()
    # schedules see the GLOBAL batch (nets/resnet_at_cifar10.py:120-122, utils/lrn_rate_utils.py:40)
    from pocketflow_b200.nets import resnet_at_cifar10 as R
    FLAGS.batch_size = 128
    lr_fn, nb_iters = R.ModelHelper().setup_lrn_rate(None)
    return float(img.sum()), lab.argmax(1).tolist(), lr_fn(0), nb_iters


def test_rank_sharded_data_and_scaled_schedule():
    out = run_ranks(_sharded_data)
    assert out[0][0] != out[1][0] and out[0][1] != out[1][1]      # every rank draws its own slice
    for _, _, lr0, nb in out:
        assert abs(lr0 - 0.1 * 256 / 128) < 1e-12                  # LR x world (batch_size_norm 128)
        assert nb == int(50000 * 250 / 256)                        # iterations / world


def test_single_process_wrapper_defaults():
    from pocketflow_b200.utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
    assert mgw.size() == 1 and mgw.rank() == 0
    t = torch.ones(3)
    assert mgw.allreduce_flat_(t) is t and t.tolist() == [1, 1, 1]
    env = {k: os.environ.pop(k) for k in ('RANK',) if k in os.environ}
    try:
        with pytest.raises(NameError):
            mgw.init()
    finally:
        os.environ.update(env)


class _BitTuner(object):
    """Stand-in learner for the RL bit search: records what every rank is asked to do."""
    device = 'cpu'

    def __init__(self):
        self.bits_seen = []

    def rl_restore(self):
        pass

    def rl_set_bits(self, w_bits, a_bits):
        self.bits_seen.append(list(w_bits))
        self.bits = list(w_bits)

    def rl_finetune(self, nb_steps, disp_steps):
        self.steps = nb_steps

    def rl_evaluate(self):
        acc = float(np.mean(1.0 - 2.0 ** (-np.asarray(self.bits, float) / 2.0)))
        return 1.0 - acc, acc, acc


def _bit_search(rank, world, mgw):
    """learners/uniform_quantization/bit_optimizer.py:137-190 on two ranks: rank 0 searches, both ranks fine-tune with
    the same broadcast bit-widths (no arranged_layer_bits.txt round trip) and end with the same allocation."""
    import random
    from types import SimpleNamespace
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.learners.uniform_quantization.bit_optimizer import BitOptimizer
    FLAGS.reset()
    FLAGS.enbl_multi_gpu, FLAGS.uql_enbl_rl_agent, FLAGS.uql_nb_rlouts = True, True, 6
    FLAGS.uql_tune_global_steps, FLAGS.uql_equivalent_bits = 100, 5
    random.seed(rank)                                  # the ranks' own random streams differ: only rank 0's matters
    shapes = [(3, 3, 4, 8), (3, 3, 8, 8), (8, 10)]
    tuner = _BitTuner()
    bo = BitOptimizer('cifar_10', [SimpleNamespace(shape=s) for s in shapes],
                      dict(nb_matmuls=3, nb_activations=2, num_weights=[int(np.prod(s)) for s in shapes]),
                      tuner=tuner, barrier_fn=mgw.barrier, seed=rank)
    w_bits, a_bits = bo.run()
    return w_bits, a_bits, tuner.bits_seen, tuner.steps, len(bo.reward_list)


def test_rl_bit_search_broadcasts_rank0_choices():
    (w0, a0, seen0, steps0, n0), (w1, a1, seen1, steps1, n1) = run_ranks(_bit_search)
    assert w0 == w1 and a0 == a1 == [32, 32] and seen0 == seen1 and len(seen0) == 6
    assert steps0 == steps1 == 50                       # uql_tune_global_steps / world size
    assert n0 == 6 and n1 == 0                          # rewards, agent and replay live on the primary worker only
    assert all(2 <= b <= 8 for b in w0)


def _sharded_files(rank, world, mgw):
    """filenames.shard(size, rank) (datasets/abstract_dataset.py:80-81) on real files: CIFAR-10 binaries and ILSVRC-12
    TFRecord shards — every rank streams its own files only."""
    from pocketflow_b200.flags import FLAGS
    from pocketflow_b200.datasets.cifar10_dataset import Cifar10Dataset
    root = os.environ['PF_TEST_DATA']
    FLAGS.reset()
    FLAGS.enbl_multi_gpu, FLAGS.batch_size, FLAGS.data_dir_local, FLAGS.nb_classes = True, 5, os.path.join(root, 'cifar'), 10
    it = Cifar10Dataset(is_train=True).build()
    cifar_labels = sorted(set(np.concatenate([it.next_batch()[1].numpy().argmax(1) for _ in range(8)]).tolist()))
    import importlib
    D = importlib.import_module('pocketflow_b200.datasets.ilsvrc12_dataset')
    FLAGS.data_dir_local, FLAGS.batch_size, FLAGS.nb_classes = os.path.join(root, 'ilsvrc'), 2, 1001
    FLAGS.nb_threads, FLAGS.prefetch_size, FLAGS.buffer_size = 1, 1, 4
    ds = D.Ilsvrc12Dataset(is_train=True)
    ds.batch_size, ds.nb_classes = 2, 1001              # (the flag defaults belong to the first dataset module imported)
    it = ds.build()
    ilsvrc_labels = sorted(set(np.concatenate([it.next_batch()[1].numpy().argmax(1) for _ in range(6)]).tolist()))
    return cifar_labels, ilsvrc_labels


def test_rank_sharded_real_files(tmp_path):
    pytest.importorskip('PIL.Image')
    import io
    from PIL import Image
    from pocketflow_b200.utils import tf_record as R
    cifar, ilsvrc = tmp_path / 'cifar', tmp_path / 'ilsvrc'
    cifar.mkdir()
    ilsvrc.mkdir()
    rng = np.random.RandomState(0)
    for f in range(2):                                  # file f holds only the labels {2f, 2f + 1}
        lab = (2 * f + rng.randint(0, 2, 20)).astype(np.uint8)
        img = rng.randint(0, 256, (20, 3 * 32 * 32)).astype(np.uint8)
        np.concatenate([lab[:, None], img], axis=1).tofile(str(cifar / ('data_batch_%d.bin' % (f + 1))))
    for f in range(2):                                  # shard f holds only the labels {10f + 1 .. 10f + 4}
        recs = []
        for i in range(4):
            b = io.BytesIO()
            Image.fromarray(rng.randint(0, 256, (40, 48, 3)).astype(np.uint8)).save(b, format='JPEG')
            recs.append(R.encode_example({'image/encoded': b.getvalue(), 'image/class/label': [10 * f + 1 + i]}))
        R.write_records(str(ilsvrc / ('train-%05d-of-00002' % f)), recs)
    os.environ['PF_TEST_DATA'] = str(tmp_path)
    try:
        (c0, i0), (c1, i1) = run_ranks(_sharded_files)
    finally:
        os.environ.pop('PF_TEST_DATA')
    assert c0 == [0, 1] and c1 == [2, 3]
    assert i0 == [1, 2, 3, 4] and i1 == [11, 12, 13, 14]


class _RatioTuner(object):
    device = 'cpu'

    def __init__(self):
        self.seen, self.retrain = [], None

    def pr_prune(self, prune_ratios):
        self.ratios = np.asarray(prune_ratios, float)
        self.seen.append(self.ratios.round(6).tolist())

    def pr_retrain(self, nb_iters_rg, nb_iters_ft):
        self.retrain = (nb_iters_rg, nb_iters_ft)

    def pr_evaluate(self):
        acc = 1.0 - float(np.mean(self.ratios ** 2))
        return 1.0 - acc, {'acc_top1': acc - 0.1, 'acc_top5': acc}


def _ratio_search(rank, world, mgw):
    """learners/weight_sparsification/pr_optimizer.py:411-470 on two ranks: ratios and rewards come from rank 0 by
    broadcast (no ./ws.prune.ratios / ./ws.reward files); iteration counts are divided by the world size."""
    from types import SimpleNamespace
    from pocketflow_b200.flags import FLAGS
    import pocketflow_b200.learners.weight_sparsification.learner  # noqa: F401  (declares the ws_* flags)
    from pocketflow_b200.learners.weight_sparsification.pr_optimizer import PROptimizer
    FLAGS.reset()
    FLAGS.enbl_multi_gpu, FLAGS.ws_prune_ratio_prtl, FLAGS.ws_prune_ratio = True, 'optimal', 0.5
    FLAGS.ws_nb_rlouts, FLAGS.ws_nb_rlouts_min, FLAGS.ws_nb_iters_rg, FLAGS.ws_nb_iters_ft = 5, 2, 20, 401
    shapes = [(3, 3, 3, 8), (3, 3, 8, 16), (3, 3, 16, 16), (16, 10)]
    mvars = [SimpleNamespace(name='model/v%d:0' % i, shape=s, numel=int(np.prod(s))) for i, s in enumerate(shapes)]
    tuner = _RatioTuner()
    opt = PROptimizer(mvars, 'ilsvrc_12', tuner=tuner, seed=10 + rank)
    out = opt.run()
    return [r for _, r in out], tuner.seen, tuner.retrain, len(opt.rewards), hasattr(opt, 'agent')


def test_rl_ratio_search_broadcasts_rank0_choices():
    (r0, seen0, rt0, n0, a0), (r1, seen1, rt1, n1, a1) = run_ranks(_ratio_search)
    assert r0 == r1 and seen0 == seen1 and len(seen0) == 5
    assert rt0 == rt1 == (10, 201)                       # ceil(20 / 2), ceil(401 / 2)
    assert (n0, a0) == (5, True) and (n1, a1) == (0, False)      # the agent and the rewards exist on rank 0 only
    n = np.array([216, 1152, 2304, 160], float)
    assert float(np.sum(np.array(r0) * n) / n.sum()) >= 0.5 - 1e-9
