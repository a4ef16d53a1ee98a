"""torchrun check (N >= 2 GPUs): the NVLS-fused gradient reduction (multimem.red in the backward epilogues) against
the plain NCCL all-reduce of the same step -- same model, same shards, dropout off, recording optimizer."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import mtn_oracle as O  # noqa: E402
from mtn_b200 import mtn, parallel  # noqa: E402
from mtn_b200.trainer import TrainStep  # noqa: E402

CFG = {"N": 2, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 200, "ft_sizes": [2048, 128], "auto_encoder_ft": "query",
       "diff_encoder": True}
sd = O.init_state_dict(CFG, 3)
full = O.synth_inputs(CFG, B=4 * world, Q=16, C=24, H=70, T=140, Lv=[140, 40], seed=5)
mine = parallel.shard_batch(full, rank, world)
ntok = parallel.all_sum(int((mine["trg_y"] != 1).sum()), dev)
nq = parallel.all_sum(int((mine["query"] != 1).sum()), dev)
batch = {k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v]) for k, v in mine.items()}


class Rec(object):
    def __init__(self):
        self.flat = None

    def step(self):
        self.flat = ts.flat.clone()


res = {}
for mode in ("0", "1"):
    os.environ["MTN_B200_NVLS"] = mode
    model = mtn.make_model(200, 200, N=2, d_model=512, d_ff=2048, h=8, dropout=0.0, ft_sizes=[2048, 128], diff_encoder=True,
                           auto_encoder_ft="query")
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    rec = Rec()
    ts = TrainStep(model, 200, graph=False, optimizer=rec)
    ts.eager(batch, ntok, nq)
    torch.cuda.synchronize()
    res[mode] = (rec.flat, ts.nvls is not None)
    del ts, model
a, b = res["0"][0], res["1"][0]
err = float((a - b).norm() / a.norm())
print("rank %d: nvls active %s, |g| %.4e, rel diff NVLS-fused vs NCCL all-reduce %.3e" % (rank, res["1"][1], float(a.norm()), err),
      flush=True)
assert res["1"][1] and err < 1e-5, err
dist.barrier()
dist.destroy_process_group()
