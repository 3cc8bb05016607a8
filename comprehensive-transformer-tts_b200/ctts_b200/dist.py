"""Multi-GPU plumbing of the inference path: independent shards, no data-path collective.

One process per GPU (torchrun); every rank synthesises its own batches of utterances.  The only communication is the
bookkeeping the benchmark contract asks for: a barrier on both sides of the timed region and a MAX all-reduce of the
per-rank device time.  The backend is NCCL on GPUs and gloo in the CPU tests (tests/test_dist_gloo.py).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_of(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` utterance batches for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def barrier(world):
    if world > 1:
        dist.barrier()


def max_over_ranks(value, device, world):
    """The slowest rank's time: what the whole job waits for."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device, world):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
