"""detectron2.utils.comm subset used by the hot path (reference: trainer.py:440,470; utils/comm.py:7-13)."""
import torch
import torch.distributed as dist


def _ready():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if _ready() else 1


def get_rank():
    return dist.get_rank() if _ready() else 0


def get_local_rank():
    import os
    return int(os.environ.get("LOCAL_RANK", 0)) if _ready() else 0


def is_main_process():
    return get_rank() == 0


def synchronize():
    if _ready() and dist.get_world_size() > 1:
        dist.barrier()


def gather(data, dst=0):
    """world 1 -> [data]; else every rank's object gathered on rank `dst` (others get [])."""
    if get_world_size() == 1:
        return [data]
    out = [None] * get_world_size() if get_rank() == dst else None
    dist.gather_object(data, out, dst=dst)
    return out if out is not None else []


def reduce_sum(tensor):
    """ubteacher/utils/comm.py:7-13 — SUM all-reduce, identity for a single process."""
    if get_world_size() < 2:
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
