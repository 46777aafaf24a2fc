"""One process per GPU: torch.distributed (gloo) is the control plane — it broadcasts the NCCL unique id that
libcmpt_b200's own communicator is built from, and provides the barrier / max-over-ranks used for timing.
The data path (allreduce of the Gram-Schmidt coefficients, SpMV halo exchange) runs inside the library."""
import os

import numpy as np


def init():
    import torch.distributed as td

    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        td.init_process_group(backend="gloo", rank=int(os.environ.get("RANK", "0")),
                              world_size=int(os.environ.get("WORLD_SIZE", "1")))
    return td


def make_context(local_rank=None):
    """Context of this rank: device = LOCAL_RANK, NCCL id created on rank 0 and broadcast over gloo."""
    import torch

    from .solvers import Context

    td = init()
    rank, world = td.get_rank(), td.get_world_size()
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if world == 1:
        return Context(local_rank)
    if rank == 0:
        ident = np.frombuffer(Context.nccl_unique_id(), dtype=np.uint8).copy()
    else:
        ident = np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(ident)
    td.broadcast(t, src=0)
    return Context(local_rank, rank=rank, nranks=world, nccl_id=t.numpy().tobytes())


def barrier():
    init().barrier()


def all_max(x):
    import torch

    td = init()
    t = torch.tensor([float(x)], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def all_sum(x):
    import torch

    td = init()
    t = torch.tensor([float(x)], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return float(t.item())


def gather_objects(obj):
    """list over ranks of a small picklable object (diagnostics)"""
    td = init()
    out = [None] * td.get_world_size()
    td.all_gather_object(out, obj)
    return out


def row_range(n, rank=None, world=None):
    td = init()
    rank = td.get_rank() if rank is None else rank
    world = td.get_world_size() if world is None else world
    return (rank * n) // world, ((rank + 1) * n) // world
