"""Per-shape throughput of the tcgen05 linear kernel: each shape is captured `reps` times into a
CUDA graph (no host gaps) and timed with CUDA events.  Shapes = the distinct GEMMs of one bench step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib  # noqa: E402

SHAPES = [  # (M, N, K, what)
    (16384, 6144, 512, "hoisted K/V rgb, all layers"),
    (8192, 6144, 512, "hoisted K/V history / vggish"),
    (2048, 6144, 512, "hoisted K/V caption / query"),
    (16384, 512, 2048, "video encoder rgb"),
    (8192, 2048, 512, "FFN w_1 (target)"),
    (8192, 512, 2048, "FFN w_2 (target) + residual"),
    (8192, 1536, 512, "self-attn QKV (target)"),
    (8192, 512, 512, "Q proj / out proj (target) + residual"),
    (2048, 2048, 512, "FFN w_1 (QAE)"),
    (2048, 512, 2048, "FFN w_2 (QAE)"),
    (2048, 1536, 512, "QKV (QAE)"),
    (2048, 1024, 512, "K/V of ae_i"),
    (2048, 512, 512, "Q/out proj (QAE)"),
]


def main():
    _lib.lib()
    reps = 20
    if len(sys.argv) > 1:          # "--one i": a single eager launch of shape i (for ncu)
        M, N, K, what = SHAPES[int(sys.argv[2])]
        A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda") if N == 512 else None
        o16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if res is None else None
        o32 = torch.empty(M, N, device="cuda") if res is not None else None
        for _ in range(3):
            _lib.linear(A, W, bias, addend=res, out_f32=o32, out_f16=o16)
        torch.cuda.synchronize()
        return
    print("%-40s %8s %8s %8s | %9s %9s" % ("shape", "M", "N", "K", "us/launch", "TFLOP/s"))
    for M, N, K, what in SHAPES:
        A = torch.randn(M, K, device="cuda").half()
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        bias = torch.randn(N, device="cuda")
        # residual shapes update the f32 stream IN PLACE (x += A W^T + b), as the engine does
        o32 = torch.randn(M, N, device="cuda") if N == 512 else None
        res = o32
        o16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if res is None else None

        def run():
            _lib.linear(A, W, bias, addend=res, out_f32=o32, out_f16=o16)
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
        print("%-40s %8d %8d %8d | %9.2f %9.1f" % (what, M, N, K, us, 2.0 * M * N * K / us / 1e6))


if __name__ == "__main__":
    main()
