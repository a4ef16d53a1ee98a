"""Per-shape time of the tcgen05 attention core: each shape is captured `reps` times into a CUDA graph (no host
gaps) and timed with CUDA events.  Shapes = the attention sites of one bench step (cfg2) + the north-star site.
Reports TFLOP/s (4 B h Lq Lk d_k) and exponentials per clock per SM (the MUFU ceiling is 16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib  # noqa: E402

SHAPES = [  # (B, h, Lq, Lk, d_k, mask, what)
    (32, 8, 256, 512, 64, "keypad", "north-star cross site"),
    (32, 8, 256, 512, 64, "none", "north-star cross site, no mask"),
    (32, 8, 256, 256, 64, "causal", "target self-attention"),
    (32, 8, 256, 256, 64, "keypad", "target -> history"),
    (32, 8, 256, 64, 64, "keypad", "target -> caption / query / ae_i"),
    (32, 8, 64, 512, 64, "keypad", "ae -> rgb"),
    (32, 8, 64, 256, 64, "keypad", "ae -> vggish"),
    (64, 8, 64, 64, 64, "keypad", "ae self (both modalities)"),
    (32, 8, 20, 256, 64, "keypad", "T=20 target -> history"),
    (16, 16, 256, 1024, 64, "keypad", "cfg5 cross site"),
]


def main():
    if os.environ.get("MTN_B200_LIB"):      # A/B runs against another build of the library
        _lib.LIB_PATH = os.environ["MTN_B200_LIB"]
    _lib.lib()
    reps = 20
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = 1965.0
    print("%-36s %3s %3s %5s %5s | %9s %9s %9s" % ("site", "B", "h", "Lq", "Lk", "us/launch", "TFLOP/s", "exp/clk/SM"))
    one = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[1] == "--one" else None   # single eager launches (for ncu)
    for B, h, Lq, Lk, dk, kind, what in (SHAPES if one is None else [SHAPES[one]]):
        d = h * dk
        g = torch.Generator(device="cuda").manual_seed(1)
        q = torch.randn(B * Lq, d, device="cuda", generator=g).half()
        kv = torch.randn(B * Lk, 2 * d, device="cuda", generator=g).half()
        out = torch.empty(B * Lq, d, device="cuda", dtype=torch.float16)
        mask = None
        if kind == "keypad":       # ragged valid lengths U[Lk/2, Lk], sample 0 full
            lens = torch.randint(Lk // 2, Lk + 1, (B,), generator=torch.Generator().manual_seed(2))
            lens[0] = Lk
            mask = (torch.arange(Lk)[None, :] < lens[:, None]).view(B, 1, Lk).cuda()
        elif kind == "causal":
            mask = torch.tril(torch.ones(Lq, Lk, dtype=torch.bool)).expand(B, Lq, Lk).contiguous().cuda()
        bits = _lib.mask_pack(mask) if mask is not None else None

        def run():
            _lib.attn_core(q, kv[:, :d], kv[:, d:], B, h, Lq, Lk, dk, out, mask_bits=bits)
        run(); torch.cuda.synchronize()
        if one is not None:
            run(); run(); torch.cuda.synchronize()
            return
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps):
                run()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps * 1e3)
        fl = 4.0 * B * h * Lq * Lk * dk
        exps = float(B * h * Lq * Lk)
        print("%-36s %3d %3d %5d %5d | %9.2f %9.1f %9.2f" % (what, B, h, Lq, Lk, best, fl / best * 1e-6,
                                                           exps / (best * 1e-6 * mhz * 1e6 * sm)))


if __name__ == "__main__":
    main()
