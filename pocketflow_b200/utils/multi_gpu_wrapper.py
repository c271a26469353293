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

    _comm = None           # the step's own NCCL communicator behind the C ABI (pf_comm_init)

    @classmethod
    def comm(cls):
        """The communicator pf_allreduce_flat runs on: created once per process from a unique id that rank 0 makes and
        torch.distributed (the plumbing) hands to the other ranks.  None on CPU / gloo (host-logic tests) and when
        PF_COMM=torch asks for torch.distributed's own all-reduce."""
        if cls._comm is None and dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available() \
                and dist.get_backend() == 'nccl' and os.environ.get('PF_COMM', 'pf') != 'torch':
            import ctypes
            from .. import lib as _lib
            L = _lib.load()
            buf = ctypes.create_string_buffer(128)
            if dist.get_rank() == 0:
                _lib.check(L.pf_comm_unique_id(buf), 'pf_comm_unique_id')
            idt = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
            dist.broadcast(idt, src=0)
            buf = ctypes.create_string_buffer(bytes(idt.cpu().numpy().tobytes()), 128)
            handle = ctypes.c_void_p()
            _lib.check(L.pf_comm_init(buf, dist.get_world_size(), dist.get_rank(), ctypes.byref(handle)), 'pf_comm_init')
            cls._comm = handle
        return cls._comm

    @classmethod
    def allreduce_flat_(cls, flat):
        """Sum `flat` (a contiguous fp32 range of the flat gradient buffer) over all ranks, in place, on the CURRENT
        stream: pf_allreduce_flat (the C ABI's NCCL all-reduce) on GPUs, torch.distributed on CPU / gloo.
        The Horovod average (sum / size) is folded into the optimizer kernel's grad_scale."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            comm = cls.comm() if flat.is_cuda else None
            if comm is not None:
                from .. import lib as _lib
                _lib.check(_lib.load().pf_allreduce_flat(comm, flat.data_ptr(), flat.numel(),
                                                         torch.cuda.current_stream().cuda_stream), 'pf_allreduce_flat')
            else:
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
