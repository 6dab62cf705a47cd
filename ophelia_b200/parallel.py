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


class GradBuckets(object):
    """Contiguous slices of the flat gradient buffer that are all-reduced as soon as the backward pass has produced them
    (on a communication stream, overlapping the rest of the backward pass) instead of in one collective at the end.

    `bounds` = variable-name prefixes in flat-buffer (creation) order that START a new bucket, e.g.
    ["Text2Mel/TextEnc/", "Text2Mel/TextEnc/HC_10/", "Text2Mel/AudioEnc/", "Text2Mel/AudioDec/"]; bucket i covers the
    variables from its first match up to the first match of bucket i+1."""

    def __init__(self, store, bounds, group=None):
        names = list(store.offsets)
        starts = []
        for b in bounds:
            first = next(n for n in names if n.startswith(b))
            starts.append(store.offsets[first])
        assert starts == sorted(starts) and starts[0] == 0, "bucket prefixes must follow the creation order from the start"
        ends = starts[1:] + [store.numel]
        self.slices = [store.grad_flat[a:b] for a, b in zip(starts, ends)]
        self.group = group
        self.works = []
        self.done = [False] * len(self.slices)
        cuda = store.grad_flat.is_cuda
        self.comm = torch.cuda.Stream(device=store.grad_flat.device) if cuda else None

    def launch(self, i, after=()):
        """All-reduce bucket i once the streams in `after` have finished what is enqueued on them so far."""
        if not dist.is_initialized() or self.done[i]:
            return
        self.done[i] = True
        if self.comm is None:
            self.works.append(dist.all_reduce(self.slices[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        for s_ in after:
            self.comm.wait_stream(s_)
        with torch.cuda.stream(self.comm):
            self.works.append(dist.all_reduce(self.slices[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Launch whatever is left, make the current stream wait for every bucket; returns the 1/world scale."""
        if not dist.is_initialized():
            return 1.0
        cur = torch.cuda.current_stream() if self.comm is not None else None
        for i in range(len(self.slices)):
            self.launch(i, after=(cur,) if cur is not None else ())
        for w in self.works:
            w.wait()
        if self.comm is not None:
            cur.wait_stream(self.comm)
        self.works = []
        self.done = [False] * len(self.slices)
        return 1.0 / dist.get_world_size(self.group)


def broadcast_parameters(store, group=None, src=0):
    """Make every replica start from rank `src`'s values (same seed gives this for free; explicit for restores)."""
    if dist.is_initialized():
        dist.broadcast(store.flat, src=src, group=group)
        store.version += 1
