"""torch.distributed plumbing for the multi-GPU path (one process per GPU). Only rendezvous data travels here: the
64-byte peer-memory handles of sgb_comm_get_handle and a barrier; the PCG halo gathers and reductions go through
NVLink peer memory inside the kernels."""
from __future__ import annotations


def exchange_blobs(blob: bytes):
    """all-gather one small bytes object per rank, rank order (works with gloo on CPU and nccl on GPUs)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [blob]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, blob)
    return out


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
