"""Feed-forward sublayer (behind the LayerNorm): two linear launches vs the one-kernel form (csrc/ffn_fused.cu), per shape,
CUDA-graph replay of 20 calls.  Usage: python tools/ffn_bench.py [out_file]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib as L
L.lib()
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
d, dff = 512, 2048
torch.manual_seed(0)
for rows in (8192, 4096, 2048, 640, 64):
    x = torch.randn(rows, d, device="cuda")
    xn = torch.randn(rows, d, device="cuda").half()
    w1, w2 = (torch.randn(dff, d, device="cuda") * 0.05).half(), (torch.randn(d, dff, device="cuda") * 0.03).half()
    b1, b2 = torch.randn(dff, device="cuda") * 0.1, torch.randn(d, device="cuda") * 0.1
    hid = torch.empty(rows, dff, device="cuda", dtype=torch.float16)

    def seq():
        L.linear(xn, w1, b1, act=L.ACT_RELU, out_f16=hid)
        L.linear(hid, w2, b2, addend=x, out_f32=x)

    def fused():
        L.ffn_fused(xn, x, w1, b1, w2, b2)

    res = []
    for fn in (seq, fused):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        x.normal_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / 100)
        x.normal_()
    fl = 4.0 * rows * d * dff
    print("rows=%5d  two launches %6.1f us (%4.0f TFLOP/s)  one kernel %6.1f us (%4.0f TFLOP/s algorithmic)  %.2fx" %
          (rows, res[0], fl / res[0] / 1e6, res[1], fl / res[1] / 1e6, res[0] / res[1]), file=out, flush=True)
