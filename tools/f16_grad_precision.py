"""How much gradient error do f16 tensor-core operands cost by themselves?  CPU experiment (test infrastructure):
the oracle's training step with every nn.Linear replaced by an autograd Function that rounds its operands to f16
in the forward AND the incoming gradient / operands in the backward (f32 accumulation), i.e. the arithmetic
contract of the sm_100a GEMM kernels without any of their code.  Prints the per-tensor normwise gradient error
against the pure-f32 oracle -- the precision floor the CUDA path can be held to."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mtn_oracle as O  # noqa: E402
from test_oracle_grads import grad_errors  # noqa: E402

r16 = lambda t: t.half().float()


class Linear16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x16, w16 = r16(x), r16(w)
        ctx.save_for_backward(x16, w16)
        return x16 @ w16.t() + b

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        s = 2.0 ** (8 - np.frexp(float(dy.abs().max()) + 1e-300)[1])
        dy16 = r16(dy * s) / s
        dx = dy16 @ w16
        dw = dy16.reshape(-1, dy16.shape[-1]).t() @ x16.reshape(-1, x16.shape[-1])
        return dx, dw, dy.reshape(-1, dy.shape[-1]).sum(0)


def main():
    cfg = {"N": 2, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 200, "ft_sizes": [2048, 128],
           "auto_encoder_ft": "query", "diff_encoder": True}
    sd = O.init_state_dict(cfg, 3)
    inp = O.synth_inputs(cfg, B=4, Q=16, C=24, H=70, T=12, Lv=[140, 40], seed=5)
    args = (sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
    loss32, g32 = O.loss_and_grads(*args)
    orig = O.linear
    O.linear = lambda x, w, b: Linear16.apply(x, w, b)
    try:
        loss16, g16 = O.loss_and_grads(*args)
    finally:
        O.linear = orig
    errs = grad_errors(g16, g32)
    v = np.array(sorted(errs.values()))
    print("loss f32 %.6f  f16-operand linears %.6f" % (loss32, loss16))
    print("gradient error of f16-operand linears vs f32: median %.2e  p90 %.2e  max %.2e" % (np.median(v), v[int(.9 * len(v))], v[-1]))
    for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:6]:
        print("  %-60s %.2e" % (k, e))


if __name__ == "__main__":
    main()
