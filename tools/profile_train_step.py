"""One eager TRAINING step of the bench workload (forward + loss + backward + Adam, dropout 0.1) between
cudaProfilerStart/Stop, for ncu (see tools/profile_round_train.sh)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import mtn_oracle as O  # noqa: E402
from mtn_b200 import mtn  # noqa: E402
from mtn_b200.trainer import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tgt-len", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()
torch.manual_seed(7)
C = bench.CFG
model = mtn.make_model(C["vocab"], C["vocab"], N=C["N"], d_model=C["d_model"], d_ff=C["d_ff"], h=C["h"],
                       ft_sizes=C["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").cuda()
inp = bench.synth(O, args.batch, args.tgt_len, 1000)
d = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items()}
ntok, nq = int((inp["trg_y"] != 1).sum()), int((inp["query"] != 1).sum())
ts = TrainStep(model, C["vocab"], graph=False)
ts.eager(d, ntok, nq); ts.eager(d, ntok, nq)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts.eager(d, ntok, nq)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
