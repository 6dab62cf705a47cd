"""Data parallelism over utterances: one process per GPU, full replicas, ONE collective per step -- a sum
all-reduce (NCCL over NVLink 5 / NVSwitch) of the flat fp32 gradient buffer, averaged inside the fused
clip+Adam kernel (`grad_scale`).  The reference has no distributed code at all; clip-by-value is
applied after averaging so that N GPUs x B utterances equals one GPU with N*B (architectures.py:125-127).
The autoregressive synthesis loop does not shard: replicas only."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29400")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(batch, rank, world):
    """Contiguous utterance shard of a host batch dict (global batch -> per-rank batch)."""
    out = {}
    for k, v in batch.items():
        n = len(v)
        assert n % world == 0, "global batch must divide by the number of ranks"
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def allreduce_gradients(flat_grad, group=None):
    """Sum-all-reduce the flat gradient buffer in place; returns the scale (1/world) to fold into clip+Adam."""
    if not dist.is_initialized():
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)


def broadcast_parameters(store, group=None, src=0):
    """Make every replica start from rank `src`'s values (same seed gives this for free; explicit for restores)."""
    if dist.is_initialized():
        dist.broadcast(store.flat, src=src, group=group)
        store.version += 1
