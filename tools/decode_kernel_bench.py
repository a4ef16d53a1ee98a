"""Per-kernel latency of the few-row decode kernels inside a dependent chain (CUDA-graph replay of 40 chained launches),
next to the tcgen05 kernels on the same shapes.  Usage: python tools/decode_kernel_bench.py [out_file]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib as L
L.lib()
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
dev = "cuda"
torch.manual_seed(0)
B, d, h, dff = 64, 512, 8, 2048


def timeit(fn, n=40):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)


x = torch.randn(B, d, device=dev)
a2, b2 = torch.ones(d, device=dev), torch.zeros(d, device=dev)
xn = torch.randn(B, d, device=dev).half()
hid = torch.randn(B, dff, device=dev).half()
# many distinct weight sets so that the chain streams weights like a real step (6 layers x 13 matrices)
Wq = [(torch.randn(d, d, device=dev) * 0.05).half() for _ in range(40)]
W1 = [(torch.randn(dff, d, device=dev) * 0.05).half() for _ in range(40)]
W2 = [(torch.randn(d, dff, device=dev) * 0.05).half() for _ in range(40)]
bq, b1 = torch.zeros(d, device=dev), torch.zeros(dff, device=dev)
q16 = torch.empty(B, d, device=dev, dtype=torch.float16)
h16 = torch.empty(B, dff, device=dev, dtype=torch.float16)
kv256 = [torch.randn(B * 256, 2 * d, device=dev).half() for _ in range(8)]
kv64 = [torch.randn(B * 64, 2 * d, device=dev).half() for _ in range(8)]
o16 = torch.empty(B, d, device=dev, dtype=torch.float16)
mask = torch.ones(B, 1, 256, dtype=torch.bool, device=dev); bits256 = L.mask_pack(mask)
bits64 = L.mask_pack(mask[:, :, :64].contiguous())
cnt = [0]


def nxt(lst):
    cnt[0] += 1
    return lst[cnt[0] % len(lst)]


for rows in (True, False):
    L.ROWS_KERNELS = rows
    tag = "few-row kernels" if rows else "tcgen05 kernels"
    res = {}
    res["layernorm"] = timeit(lambda: L.layernorm(x, a2, b2, 1e-6, out_f16=xn))
    res["linear 64x512x512 (f16 out)"] = timeit(lambda: L.linear(xn, nxt(Wq), bq, out_f16=q16))
    res["linear 64x512x512 + residual"] = timeit(lambda: L.linear(xn, nxt(Wq), bq, addend=x, out_f32=x))
    res["linear 64x2048x512 relu"] = timeit(lambda: L.linear(xn, nxt(W1), b1, act=L.ACT_RELU, out_f16=h16))
    res["linear 64x512x2048 + residual"] = timeit(lambda: L.linear(hid, nxt(W2), bq, addend=x, out_f32=x))
    if rows:
        res["LN + linear 64x512x512 (one launch)"] = timeit(lambda: L.rows_ln_linear(x, a2, b2, 1e-6, nxt(Wq), bias=bq, out_f16=q16))
        res["LN + linear 64x2048x512 (one launch)"] = timeit(lambda: L.rows_ln_linear(x, a2, b2, 1e-6, nxt(W1), bias=b1, act=1, out_f16=h16))
    res["attention Lq=1 Lk=256"] = timeit(lambda: (lambda kv: L.attn_core(xn, kv[:, :d], kv[:, d:], B, h, 1, 256, 64, o16, mask_bits=bits256))(nxt(kv256)))
    res["attention Lq=1 Lk=64"] = timeit(lambda: (lambda kv: L.attn_core(xn, kv[:, :d], kv[:, d:], B, h, 1, 64, 64, o16, mask_bits=bits64))(nxt(kv64)))
    res["attention Lq=1 Lk=12 (self cache)"] = timeit(lambda: L.attn_core(xn, kv64[0][:B * 12, :d], kv64[0][:B * 12, d:], B, h, 1, 12, 64, o16))
    for k, v in res.items():
        print("%-18s %-40s %6.2f us per launch in a dependent chain" % (tag, k, v), file=out, flush=True)
