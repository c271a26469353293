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
    P0, O0 = ex.store.P.clone(), ex.store.O.clone()
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
    # 4. the collective went through the C ABI (pf_allreduce_flat on the library's own NCCL communicator), in two
    # buckets overlapped with the backward pass; a second learner that sums the whole buffer in ONE all-reduce after the
    # backward pass (PF_AR_BUCKETS=1) gets the same gradients and parameters from the same state and mini-batch
    assert mgw.comm() is not None, 'the step did not use pf_allreduce_flat'
    bk = ex._bucket_plan()
    assert bk is not None and 0 < bk['split'] < bk['end'] <= ex.G.numel(), bk
    os.environ['PF_AR_BUCKETS'] = '1'
    lrn1 = UniformQuantLearner(None, R.ModelHelper())
    ex1 = lrn1.sess_train
    assert ex1._bucket_plan() is None            # (decided once per executor, while the variable is set)
    del os.environ['PF_AR_BUCKETS']
    ex1.store.P.copy_(P0)
    ex1.store.O.copy_(O0)
    ex1.teacher.store.P.copy_(ex.teacher.store.P)
    ex1.teacher.store.O.copy_(ex.teacher.store.O)
    ex1.buf[lrn1.images].copy_(ex.buf[lrn.images])
    ex1.buf[lrn1.labels].copy_(ex.buf[lrn.labels])
    ex1.run_step(lrn1.lrn_rate(0), mgw.allreduce_flat_)
    torch.cuda.synchronize()
    tol = 0.0 if world == 2 else 1e-6           # a + b is commutative; more ranks: NCCL's order may depend on the size
    assert (ex1.G - g).abs().max() <= tol * g.abs().max(), 'bucketed all-reduce differs from the single one'
    assert (ex1.store.P - p).abs().max() <= tol * p.abs().max()
    # 5. the captured CUDA graph with the bucketed collective inside replays to the same result as the eager step
    ex.capture(mgw.allreduce_flat_)              # (its warm-up runs one real step: restore the state afterwards)
    ex.store.P.copy_(P0)
    ex.store.O.copy_(O0)
    ex.reset_optimizer_state()
    ex.run_step(lrn.lrn_rate(0), mgw.allreduce_flat_)
    torch.cuda.synchronize()
    assert (ex.G - g).abs().max() <= tol * g.abs().max(), 'graph replay differs'
    assert (ex.store.P - p).abs().max() <= tol * p.abs().max()
    if rank == 0:
        print('mgpu_check ok: world=%d, |G|max=%.3e' % (world, float(g.abs().max())), flush=True)
    dist.barrier()


if __name__ == '__main__':
    main()
