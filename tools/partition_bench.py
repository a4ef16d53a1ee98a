"""Forward of the bench workload (cfg2, B = 32, T = 256) as one CUDA graph with / without SM partitioning
(parallel.SmPartition: the Query-Aware Auto-Encoder side chain on a small SM group, the target path on the rest).
Usage: python tools/partition_bench.py [small_sm_counts ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
import mtn_oracle as O
from mtn_b200 import mtn
from mtn_b200.graph import GraphedForward
from mtn_b200.parallel import SmPartition

torch.manual_seed(7)
C = bench.CFG
model = mtn.make_model(C["vocab"], C["vocab"], N=C["N"], d_model=C["d_model"], d_ff=C["d_ff"], h=C["h"],
                       ft_sizes=C["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").cuda().eval()
inp = bench.synth(O, 32, 256, 1000)
d = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items()}

def timed(g, n=50):
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g0 = GraphedForward(model, d)
t0 = timed(g0)
ref = [g0.out.clone()] + [a.clone() for a in g0.ae]
print("no partition: %.3f ms per forward" % t0, flush=True)
for small in [int(x) for x in (sys.argv[1:] or ["16"])]:
    part = SmPartition(small)
    g = GraphedForward(model, d, partition=part)
    t = timed(g)
    same = all(torch.equal(a, b) for a, b in zip(ref, [g.out] + list(g.ae)))
    print("partition %d + %d SMs: %.3f ms per forward (%.2fx), outputs bit-identical: %s" % (part.sms_small, part.sms_big, t, t0 / t, same), flush=True)
