"""CPU tier, world_size 2 over gloo: the N>1 host logic of the data-parallel path (sharding,
global token count, max-over-ranks timing).  The forward has no data-path collective."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mtn_oracle as O
    from mtn_b200 import parallel
    cfg = {"vocab": 50, "ft_sizes": [8, 4]}
    full = O.synth_inputs(cfg, B=6, Q=5, C=5, H=7, T=9, Lv=[6, 3], seed=1)
    mine = parallel.shard_batch(full, rank, world)
    lo, hi = parallel.shard_range(6, rank, world)
    assert torch.equal(mine["trg"], full["trg"][lo:hi]) and torch.equal(mine["fts"][1], full["fts"][1][lo:hi])
    tok = parallel.global_tokens(mine["trg_y"], 1)
    tmax = parallel.all_max(10.0 + rank)
    q.put((rank, tok, int((full["trg_y"] != 1).sum()), tmax, mine["trg"].shape[0]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, tok, total, tmax, n in res:
        assert tok == total                 # shards sum to the global token count on every rank
        assert tmax == 11.0                 # max over ranks
        assert n == 3
