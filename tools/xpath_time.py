"""Time the target path alone (memory stage cached) vs the whole forward, CUDA-graph replayed.
usage: python tools/xpath_time.py [tgt_len]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench, mtn_oracle as O
from mtn_b200 import mtn
from mtn_b200.data_utils import Batch

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = bench.CFG
torch.manual_seed(7)
model = mtn.make_model(C["vocab"], C["vocab"], N=C["N"], d_model=C["d_model"], d_ff=C["d_ff"], h=C["h"],
                       ft_sizes=C["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").cuda().eval()
inp = bench.synth(O, 32, T, 1000)
d = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items()}
with torch.no_grad():
    b = Batch(d["query"], d["his"], None, [f.permute(1, 0, 2) for f in d["fts"]], d["cap"], d["trg"], d["trg_y"], 1)
    mem = model.encode(b.query, b.query_mask, b.his, b.his_mask, b.cap, b.cap_mask, b.fts, b.fts_mask)
    q, vid, cap, his, ae = mem

    def dec():
        return model.decode(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask, b.trg, b.trg_mask, ae)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dec(); dec()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = dec()                 # memory stage is cached: target path only
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    g.replay()
e1.record(); torch.cuda.synchronize()
print("T=%d target path alone: %.3f ms" % (T, e0.elapsed_time(e1) / 50))
