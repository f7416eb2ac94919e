"""Shot partitioning across ranks and the single all-reduce of one objective evaluation.

One process per GPU (torchrun); shots are independent (reference shot loop: fwi.py:183-199), so the
only exchange is the sum of [grad | illum | fval] at the end of ``fwi_obj_multi``.  The reference's
own parallel variant (dask, fwi.py:207-234) is dead code; this is its working replacement.
"""
import os

__all__ = ['rank', 'world_size', 'local_shots', 'all_reduce_sum', 'init_from_env']


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def local_shots(nshots, r=None, w=None):
    """Shots of rank ``r``: round-robin ``i % W == r``."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return list(range(r, nshots, w))


def all_reduce_sum(tensor):
    """In-place sum over ranks (NCCL for CUDA tensors, gloo for CPU tensors); no-op single-rank."""
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(tensor, op=d.ReduceOp.SUM)
    return tensor


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK)."""
    import torch
    import torch.distributed as dist
    if int(os.environ.get('WORLD_SIZE', '1')) <= 1 or dist.is_initialized():
        return
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend=backend)
