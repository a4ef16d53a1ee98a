"""One KV-cached decoding step (BASELINE configs[3]: batch 64, prefix length 10) between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv \
      python tools/profile_decode_step.py"""
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

torch.manual_seed(7)
C, S = bench.CFG, bench.SHAPE
model = mtn.make_model(C["vocab"], C["vocab"], N=C["N"], d_model=C["d_model"], d_ff=C["d_ff"], h=C["h"],
                       ft_sizes=C["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").cuda().eval()
inp = O.synth_inputs(C, B=64, Q=S["Q"], C=S["C"], H=S["H"], T=4, Lv=S["Lv"], seed=5001)
d = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items()}
with torch.no_grad():
    b = Batch(d["query"], d["his"], None, [f.permute(1, 0, 2) for f in d["fts"]], d["cap"], None, None, 1)
    q, vid, cap, his, ae = model.encode(b.query, b.query_mask, b.his, b.his_mask, b.cap, b.cap_mask, b.fts, b.fts_mask)
    st = model.decode_begin(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask, ae, 20)
    tok = torch.full((64,), 2, dtype=torch.int64, device="cuda")
    for t in range(10):
        tok = model.decode_step_argmax(st, tok, t)       # greedy decoding's step: embedding + the cluster kernel
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tok = model.decode_step_argmax(st, tok, 10)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
