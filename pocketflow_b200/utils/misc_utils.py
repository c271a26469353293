"""Miscellaneous utility functions (/root/reference/utils/misc_utils.py:25-52)."""
from ..flags import FLAGS
from .multi_gpu_wrapper import MultiGpuWrapper as mgw


def auto_barrier(mpi_comm=None):
    """Barrier for multi-GPU training, no-op for single-GPU training."""
    if FLAGS.enbl_multi_gpu:
        mgw.barrier()


def is_primary_worker(scope='global'):
    """Whether this is the primary worker of all nodes (global) or of the current node (local)."""
    if scope == 'global':
        return True if not FLAGS.enbl_multi_gpu else mgw.rank() == 0
    elif scope == 'local':
        return True if not FLAGS.enbl_multi_gpu else mgw.local_rank() == 0
    else:
        raise ValueError('unrecognized worker scope: ' + scope)
