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


class SmPartition(object):
    """Two disjoint groups of SMs on the current device (CUDA green contexts): ``small`` SMs for the engine's side chain --
    the Query-Aware Auto-Encoder branch, a latency-bound sequence of small kernels that never reads the target stream --
    and the rest for the target path.  Kernels launched on a stream of a group run on that group's SMs only, so the two
    chains run CONCURRENTLY instead of taking turns on the whole machine (each target-path kernel is a single wave of 128
    CTAs on a 148-SM device: 20 SMs idle).  ``main`` / ``side``: torch streams of the big / small group; ``extra(i)``:
    further streams of the big group.  Persistent kernels size their grids to the launching stream's SM count
    (csrc/host.cu stream_sm_count).  Needs cuda-python (cuda.bindings); graph-capturable."""

    def __init__(self, small=16, device=0):
        from cuda.bindings import driver as cu
        self._cu = cu
        torch.cuda.init()
        torch.zeros(1, device="cuda:%d" % device)
        dev = self._ck(cu.cuDeviceGet(device))
        res = self._ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
        groups, _, rem = self._ck(cu.cuDevSmResourceSplitByCount(1, res, 0, int(small)))
        self.sms_small, self.sms_big = int(groups[0].sm.smCount), int(rem.sm.smCount)
        self._ctx, self._raw = {}, []
        for name, r in (("small", groups[0]), ("big", rem)):
            desc = self._ck(cu.cuDevResourceGenerateDesc([r], 1))
            self._ctx[name] = self._ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        self.main = self._stream("big")
        self.side = self._stream("small")
        self._extra = []

    @staticmethod
    def _ck(r):
        err, rest = r[0], r[1:]
        if int(err) != 0:
            raise RuntimeError("CUDA driver error %s" % (err,))
        return rest[0] if len(rest) == 1 else rest

    def _stream(self, group):
        cu = self._cu
        s = self._ck(cu.cuGreenCtxStreamCreate(self._ctx[group], cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        self._raw.append(s)
        return torch.cuda.ExternalStream(int(s))

    def extra(self, i):
        while len(self._extra) <= i:
            self._extra.append(self._stream("big"))
        return self._extra[i]
