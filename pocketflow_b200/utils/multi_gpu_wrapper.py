"""Wrapper for multi-GPU training — the MultiGpuWrapper surface of the reference
(/root/reference/utils/multi_gpu_wrapper.py:30-98) over torch.distributed instead of Horovod/TF-Plus.

One process per GPU (torchrun); `init()` joins the process group (NCCL on GPUs, gloo on CPU for the
host-logic tests); the gradient exchange itself is ONE flat all-reduce per step."""
import os

import torch
import torch.distributed as dist


class MultiGpuWrapper(object):
    _initialized_here = False

    def __init__(self):
        pass

    @classmethod
    def init(cls, backend=None):
        if dist.is_available() and dist.is_initialized():
            return
        if 'RANK' not in os.environ:
            raise NameError('module <mgw> not imported')   # same failure mode as the reference without Horovod
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group(backend=backend)
        cls._initialized_here = True

    @classmethod
    def size(cls):
        return dist.get_world_size() if dist.is_initialized() else 1

    @classmethod
    def rank(cls):
        return dist.get_rank() if dist.is_initialized() else 0

    @classmethod
    def local_size(cls):
        return int(os.environ.get('LOCAL_WORLD_SIZE', cls.size()))

    @classmethod
    def local_rank(cls):
        return int(os.environ.get('LOCAL_RANK', cls.rank()))

    @classmethod
    def allreduce_flat_(cls, flat):
        """Sum `flat` (one contiguous fp32 buffer holding every gradient) over all ranks, in place.
        The Horovod average (sum / size) is folded into the optimizer kernel's grad_scale."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return flat

    @classmethod
    def broadcast_global_variables(cls, tensors, root_rank=0):
        """mgw.broadcast_global_variables(0): rank-0 state to every rank (list of flat buffers)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            for t in tensors:
                dist.broadcast(t, src=root_rank)

    @classmethod
    def barrier(cls):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
