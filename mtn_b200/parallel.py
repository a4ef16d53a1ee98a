"""Multi-GPU plumbing for the MTN hot path: one process per GPU (torchrun), pure data
parallelism over independent dialogue samples.  The forward pass has NO
data-path collective -- every rank runs the full model on its own shard of the batch; the
only communication is bookkeeping (global token counts, max-over-ranks timing), which works
on NCCL (GPU) and gloo (CPU tests) alike.
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n samples for `rank` (first n % world ranks get one
    more).  cfg3 of BASELINE.json: 256 dialogues over 8 ranks -> rows [32r, 32r+32)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(inputs, rank, world):
    """Slice every tensor (or list of tensors) of a synthetic-input dict along dim 0."""
    any_t = next(v for v in inputs.values() if torch.is_tensor(v))
    lo, hi = shard_range(any_t.shape[0], rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) else [f[lo:hi] for f in v]) for k, v in inputs.items()}


def _reduce(x, op, device):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def all_sum(x, device="cpu"):
    return _reduce(x, dist.ReduceOp.SUM, device)


def all_max(x, device="cpu"):
    return _reduce(x, dist.ReduceOp.MAX, device)


def global_tokens(trg_y, pad, device="cpu"):
    """Non-pad target tokens over ALL ranks (train.py:41-48 counts b.ntokens per batch; a sharded
    job must sum the shards)."""
    return all_sum(int((trg_y != pad).sum()), device)
