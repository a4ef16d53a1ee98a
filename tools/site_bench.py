"""Time one attention site (hoisted K/V): launch sequence LayerNorm -> Q GEMM -> core -> out-proj vs LayerNorm -> fused
site kernel, per shape, CUDA-graph replay of 20 calls.  Usage: python tools/site_bench.py [out_file]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib as L
L.lib()
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
dev = "cuda"
torch.manual_seed(0)
SHAPES = [(32, 256, 64), (32, 256, 256), (32, 256, 512), (64, 64, 512), (64, 64, 256), (32, 20, 64), (32, 20, 256),
          (64, 1, 64), (64, 1, 256), (64, 5, 64), (64, 5, 256)]
d, h, dk = 512, 8, 64
for (B, Lq, Lk) in SHAPES:
    x = torch.randn(B * Lq, d, device=dev)
    a2, b2 = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    kv = (torch.randn(B * Lk, 2 * d, device=dev)).half()
    wq, wo = (torch.randn(d, d, device=dev) * 0.05).half(), (torch.randn(d, d, device=dev) * 0.05).half()
    bq, bo = torch.randn(d, device=dev) * 0.1, torch.randn(d, device=dev) * 0.1
    lens = torch.randint(Lk // 2, Lk + 1, (B,))
    mask = (torch.arange(Lk)[None, :] < lens[:, None]).view(B, 1, Lk).to(dev)
    bits = L.mask_pack(mask)
    xn = torch.empty(B * Lq, d, device=dev, dtype=torch.float16)
    qb = torch.empty(B * Lq, d, device=dev, dtype=torch.float16)
    ob = torch.empty(B * Lq, d, device=dev, dtype=torch.float16)

    def seq():
        L.layernorm(x, a2, b2, 1e-6, out_f16=xn)
        L.linear(xn, wq, bq, out_f16=qb)
        L.attn_core(qb, kv[:, :d], kv[:, d:], B, h, Lq, Lk, dk, ob, mask_bits=bits)
        L.linear(ob, wo, bo, addend=x, out_f32=x)

    def fused():
        L.layernorm(x, a2, b2, 1e-6, out_f16=xn)
        L.attn_site_fused(xn, x, wq, bq, wo, bo, kv, 0, d, B, h, Lq, Lk, mask_bits=bits)

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
    fl = 4 * B * Lq * d * d + 4 * B * Lq * Lk * d
    print("B=%3d Lq=%3d Lk=%3d  sequence %6.1f us  fused %6.1f us  (%.2fx)  fused %.0f TFLOP/s" %
          (B, Lq, Lk, res[0], res[1], res[0] / res[1], fl / res[1] / 1e6), file=out, flush=True)
