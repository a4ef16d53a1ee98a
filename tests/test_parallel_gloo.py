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


# ---------------------------------------------------------------------------------------------------------------
# Data-parallel TRAINING recipe: shards + GLOBAL normalisers + ONE all-reduce(SUM) of the flat gradient
# must equal the single-process gradient of the whole batch.  CPU tier: the recipe itself with the oracle's autograd
# (world_size 2, gloo).  GPU tier (below): the same through TrainStep's kernels, two processes sharing one GPU (gloo
# moves the CUDA gradient buffer through the host -- NCCL cannot put two ranks on one device).
# ---------------------------------------------------------------------------------------------------------------
CFG_DP = {"N": 1, "d_model": 128, "d_ff": 512, "h": 4, "vocab": 100, "ft_sizes": [2048, 128], "auto_encoder_ft": "query",
          "diff_encoder": True}


def _dp_oracle_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import mtn_oracle as O
    from mtn_b200 import parallel
    sd = O.init_state_dict(CFG_DP, 11)
    full = O.synth_inputs(CFG_DP, B=4, Q=8, C=8, H=16, T=8, Lv=[16, 8], seed=5)
    mine = parallel.shard_batch(full, rank, world)
    ntok = parallel.global_tokens(mine["trg_y"], 1)
    nq = parallel.all_sum(int((mine["query"] != 1).sum()))
    _, g = O.loss_and_grads(sd, CFG_DP, mine["query"], mine["his"], mine["cap"], mine["trg"], mine["trg_y"], mine["fts"],
                            norm=ntok, ae_norm=nq)
    names = sorted(g)
    flat = torch.cat([g[k].reshape(-1) for k in names])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                      # the one gradient collective
    if rank == 0:
        _, gf = O.loss_and_grads(sd, CFG_DP, full["query"], full["his"], full["cap"], full["trg"], full["trg_y"], full["fts"])
        ref = torch.cat([gf[k].reshape(-1) for k in names])
        q.put(float((flat - ref).norm() / ref.norm()))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 500
    ps = [ctx.Process(target=_dp_oracle_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    err = q.get(timeout=300)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert err <= 1e-5, err                                          # SURVEY 8e equivalence bar


def _dp_gpu_worker(rank, world, port, q, backend="gloo", batch=4, overlap=False):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if backend == "nccl":                   # one rank per GPU, the production arrangement
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:                                   # gloo: both ranks share GPU 0 (gradient buffer staged through the host)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
    import mtn_oracle as O
    from mtn_b200 import mtn, parallel
    from mtn_b200.trainer import TrainStep
    sd = O.init_state_dict(CFG_DP, 11)
    full = O.synth_inputs(CFG_DP, B=batch, Q=8, C=8, H=16, T=8, Lv=[16, 8], seed=5)

    def grads_of(batch, ntok, nq, use_dist):
        model = mtn.make_model(100, 100, N=1, d_model=128, d_ff=512, h=4, dropout=0.0, ft_sizes=[2048, 128],
                               diff_encoder=True, auto_encoder_ft="query")
        model.load_state_dict(sd, strict=True)
        model = model.cuda()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        ts = TrainStep(model, 100, graph=False, optimizer=_NoOpt(), overlap=overlap and use_dist)
        if not use_dist:
            ts.world = 1
        snap = {}
        ts.opt.hook = lambda: snap.update(flat=ts.flat.clone())      # the gradient the optimizer would consume
        ts.eager({k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in batch.items()}, ntok, nq)
        return snap["flat"].cpu()

    mine = parallel.shard_batch(full, rank, world)
    dev = "cuda" if backend == "nccl" else "cpu"
    if backend == "nccl":
        # normalisers counted ON THE DEVICE inside the step (summed over the ranks by TrainStep), prefix all-reduce
        # overlapped with the rest of the backward: the production path
        g_dp = grads_of(mine, None, None, True)
    else:
        ntok = parallel.global_tokens(mine["trg_y"], 1, dev)
        nq = parallel.all_sum(int((mine["query"] != 1).sum()), dev)
        g_dp = grads_of(mine, ntok, nq, True)
    if rank == 0:
        g_one = grads_of(full, int((full["trg_y"] != 1).sum()), int((full["query"] != 1).sum()), False)
        q.put(float((g_dp - g_one).norm() / g_one.norm()))
    dist.barrier()
    dist.destroy_process_group()


class _NoOpt(object):
    """Stand-in optimizer: records instead of updating (so both runs start from the same weights)."""
    hook = None

    def step(self):
        if self.hook:
            self.hook()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_two_process_trainstep_allreduce_equals_single_process_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 13) % 500
    ps = [ctx.Process(target=_dp_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    err = q.get(timeout=300)
    [p.join(120) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert err <= 2e-3, err          # two f16-operand passes over different batch splits: rounding noise only


@pytest.mark.gpu
@pytest.mark.parametrize("overlap", [False, True])
def test_nccl_ranks_trainstep_equals_single_process(overlap):
    """SURVEY 8e on hardware: N ranks, one per GPU, NCCL -- shards of one global batch, token counts summed on the
    device, the gradient exchanged in one all-reduce (default) or in two overlapped ones (overlap=True) -- must give
    the single-process gradient of the whole batch (f16-operand rounding noise only).  Needs >= 2 GPUs (the driver's
    1-GPU box skips it)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs, found %d" % n)
    world = min(n, 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 29) % 500
    ps = [ctx.Process(target=_dp_gpu_worker, args=(r, world, port + int(overlap), q, "nccl", 2 * world, overlap)) for r in range(world)]
    [p.start() for p in ps]
    err = q.get(timeout=600)
    [p.join(180) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    print("NCCL %d-rank gradient vs single process: %.2e" % (world, err))
    assert err <= 2e-3, err
