"""Data-parallel run of the training driver (test tooling, launched by torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29431 tests/dp_train_check.py DIR

Rank 0 writes a tiny corpus in the reference's layout under DIR, every rank runs ophelia_b200.train.train(hp, 't2m') over its
shard of the batches (train.py:main_work flow with one process per GPU, gradients all-reduced over NCCL), and the replicas'
weights are compared at the end: they must be bit-identical on every rank, and different from the initial ones."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    from helpers import make_corpus
    from ophelia_b200 import tf_checkpoint
    from ophelia_b200 import train as drv
    from ophelia_b200.configuration import load_config
    from ophelia_b200.parallel import init_from_env
    root = sys.argv[1]
    rank, world, local = init_from_env("nccl")
    assert world >= 2, "launch with torchrun on at least 2 GPUs"
    torch.cuda.set_device(local)
    if rank == 0:
        make_corpus(root, n_utts=38, n_valid=3, max_epochs=2, save_every_n_epochs=1, decay_lr=False)
    dist.barrier()
    hp = load_config(os.path.join(root, "tiny.cfg"))
    score = drv.train(hp, 't2m')
    from ophelia_b200.architectures import _default_stores
    store = next(iter(_default_stores.values()))
    flat = store.flat.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered[1:])
    step = int(store.global_step.item())
    if rank == 0:
        ck = tf_checkpoint.read_checkpoint(hp.logdir + "-t2m/model_epoch_2", names=["global_step"])
        print("DP_TRAIN_CHECK world=%d replicas_identical=%s global_step=%d checkpoint_step=%d score=%.3f" %
              (world, same, step, int(np.asarray(ck["global_step"]).reshape(-1)[0]), score), flush=True)
    dist.barrier()
    assert same and step > 0
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
