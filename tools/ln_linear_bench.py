"""Fused LayerNorm + projection (csrc/ln_gemm.cu) against the two launches it replaces, on the shapes of the
decoder's dependent chains.  Each variant is captured as a chain of `reps` DEPENDENT rounds
(LN -> projection -> residual-updating projection back onto the stream, the pattern of one attention site
without the core) into a CUDA graph and timed with CUDA events; the difference per round is what the fusion
buys per sublayer.  Usage: python tools/ln_linear_bench.py [out_file]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib  # noqa: E402

SHAPES = [  # (M, N, d, act, what)
    (8192, 512, 512, 0, "Q projection, target stream T=256 (B=32)"),
    (8192, 1536, 512, 0, "packed Q|K|V, target self-attention T=256"),
    (8192, 2048, 512, 1, "FFN w_1, target stream T=256"),
    (2048, 512, 512, 0, "Q projection, QAE stream (B=32, Q=64)"),
    (2048, 1536, 512, 0, "packed Q|K|V, QAE self-attention"),
    (2048, 2048, 512, 1, "FFN w_1, QAE stream"),
    (640, 512, 512, 0, "Q projection, T=20 (B=32)"),
    (1280, 512, 512, 0, "Q projection, greedy step t=20 (B=64)"),
    (1280, 2048, 512, 1, "FFN w_1, greedy step t=20 (B=64)"),
    (64, 512, 512, 0, "Q projection, greedy step t=1 (B=64)"),
    (4096, 512, 512, 0, "Q projection, 4096 rows"),
    (16384, 512, 512, 0, "Q projection, 16384 rows"),
]


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (3 * reps))
    return best


def phases(out):
    """clock64() stamps of CTA (0, 0): where one launch spends its time (cycles of the SM clock)."""
    import ctypes as C
    ts = torch.zeros(8, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().mtn_ln_linear_debug_timestamps(C.c_void_p(ts.data_ptr())))
    names = ["entry", "after griddepcontrol.wait", "LayerNorm done (worker warp 0)", "MMA warp passes a_ready",
             "last MMA of n-tile 0 issued", "epilogue sees accumulator 0", "epilogue done (worker warp 0)", "exit"]
    for M, N, d in ((8192, 512, 512), (64, 512, 512), (8192, 2048, 512), (2048, 512, 512)):
        x = torch.randn(M, d, device="cuda")
        a, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
        W = (torch.randn(N, d, device="cuda") / d ** 0.5).half()
        bias = torch.randn(N, device="cuda")
        y = torch.empty(M, N, device="cuda", dtype=torch.float16)
        for _ in range(3):
            _lib.ln_linear(x, a, b, 1e-6, W, bias=bias, out_f16=y)
        torch.cuda.synchronize()
        t = ts.cpu().tolist()
        print("phases M=%d N=%d d=%d (cycles since entry):" % (M, N, d), file=out)
        for n, v in zip(names, t):
            print("   %-36s %8d" % (n, v - t[0]), file=out)
    _lib.check(_lib.lib().mtn_ln_linear_debug_timestamps(None))


def main():
    _lib.lib()
    if len(sys.argv) > 2 and sys.argv[2] == "--phases":
        with open(sys.argv[1], "w") as f:
            phases(f)
        return
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    reps = 20
    print("%-48s %6s %5s | %9s %9s %9s  (us per dependent round; back = the residual projection alone)" %
          ("shape", "M", "N", "two", "fused", "back"), file=out)
    for M, N, d, act, what in SHAPES:
        x = torch.randn(M, d, device="cuda")
        a, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
        W = (torch.randn(N, d, device="cuda") / d ** 0.5).half()
        Wb = (0.01 * torch.randn(d, N, device="cuda") / N ** 0.5).half()
        bias, bias_b = torch.randn(N, device="cuda"), torch.zeros(d, device="cuda")
        xn = torch.empty(M, d, device="cuda", dtype=torch.float16)
        y = torch.empty(M, N, device="cuda", dtype=torch.float16)

        def back():
            _lib.linear(y, Wb, bias_b, addend=x, out_f32=x)

        def two():
            _lib.layernorm(x, a, b, 1e-6, out_f16=xn)
            _lib.linear(xn, W, bias, act=act, out_f16=y)
            back()

        def fused():
            _lib.ln_linear(x, a, b, 1e-6, W, bias=bias, act=act, out_f16=y)
            back()

        t2, tf, tb = timed(two, reps), timed(fused, reps), timed(back, reps)
        print("%-48s %6d %5d | %9.2f %9.2f %9.2f" % (what, M, N, t2, tf, tb), file=out)
        out.flush()


if __name__ == "__main__":
    main()
