"""One eager forward of the bench workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:gemm_f16_tc -c 4 -o gpurun_out/prof_gemm python tools/profile_step.py
"""
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
from mtn_b200.data_utils import Batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tgt-len", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--preset", default="cfg2", choices=["cfg2", "cfg5"])
args = ap.parse_args()
if args.preset == "cfg5":          # BASELINE configs[4]: N=12 d=1024 h=16, video_len=1024, batch 4 per GPU
    bench.CFG.update({"N": 12, "d_model": 1024, "d_ff": 4096, "h": 16})
    bench.SHAPE.update({"B": 4, "Lv": [1024, 256]})
    if args.batch == 32:
        args.batch = 4
torch.manual_seed(7)
C = bench.CFG
model = mtn.make_model(C["vocab"], C["vocab"], N=C["N"], d_model=C["d_model"], d_ff=C["d_ff"], h=C["h"],
                       ft_sizes=C["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").cuda().eval()
inp = bench.synth(O, args.batch, args.tgt_len, 1000)
d = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items()}


def step():
    b = Batch(d["query"], d["his"], None, [f.permute(1, 0, 2) for f in d["fts"]], d["cap"], d["trg"], d["trg_y"], 1)
    with torch.no_grad():
        return model.forward(b)


step(); step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
