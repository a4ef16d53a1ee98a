"""GPU tier, backward kernels (ABI v3) through the C ABI, compared with torch autograd through the CPU oracle's
arithmetic on the same (f16-rounded where the kernel rounds) operands, and -- model level -- with the gradients of
the unmodified reference (golden digest) and of the oracle's training step.

Tolerances (normwise ||g - ref|| / ||ref|| unless noted):
  GEMM forms vs on-device check kernel (identical rounding)  2e-5
  row kernels (f32 arithmetic)                                1e-5
  attention backward (P and dS rounded to f16 for the MMAs)   3e-3
  model-level parameter gradients vs the f32 reference: 3e-2 per tensor, median 8e-3 -- the floor of f16 (11-bit
  significand, = TF32) GEMM operands, measured WITHOUT any kernel by tools/f16_grad_precision.py (median 4.9e-3, max
  2.0e-2 for the d=512 case below).  The largest per-ENTRY deviations come from ReLU activations within rounding
  distance of zero (the mask flips), which is why the digest's sampled entries get a wider bar than the norms.
"""
import numpy as np
import pytest
import torch

import golden_util as G
import mtn_oracle as O
from test_oracle_grads import check_against_digest, golden_grad_case, grad_errors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from mtn_b200 import _lib
    _lib.lib()
    return _lib


def dev(t):
    return t.cuda()


def rnd16(t):
    return t.half().float()


# ------------------------------------------------------------------ general GEMM: operand forms, split-K, epilogue
GEMM_CASES = [
    # M, N, K, a_mn, b_mn
    (200, 264, 136, 0, 0), (200, 264, 136, 0, 1), (200, 264, 136, 1, 1),
    (128, 64, 64, 0, 1), (128, 64, 64, 1, 1), (520, 512, 2048, 0, 1), (512, 512, 4100, 1, 1),
    (2048, 512, 777, 1, 1), (104, 512, 333, 1, 1), (64, 2048, 192, 0, 1),
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", GEMM_CASES)
def test_gemm_forms_vs_check_kernel(L, M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + a_mn + 2 * b_mn)
    A = (torch.randn((K, M) if a_mn else (M, K), generator=g)).half().cuda()
    B = (torch.randn((K, N) if b_mn else (N, K), generator=g)).half().cuda()
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    ref = Af @ Bf.t()
    alpha = torch.tensor([0.5], device="cuda")
    out, chk = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    L.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, alpha=alpha, out_f32=out)
    L.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, alpha=alpha, out_f32=chk, _check_kernel=True)
    torch.cuda.synchronize()
    assert G.rel_err(out.cpu(), chk.cpu()) < 2e-5, "layout bug (vs on-device check kernel)"
    assert G.rel_err(out.cpu(), 0.5 * ref.cpu()) < 2e-5
    # split-K accumulation on top of existing contents
    acc = torch.full((M, N), 3.0, device="cuda")
    L.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, alpha=alpha, accumulate=True, out_f32=acc)
    torch.cuda.synchronize()
    assert G.rel_err((acc - 3.0).cpu(), 0.5 * ref.cpu()) < 5e-5


def test_gemm_relu_mask_and_pre_add(L):
    g = torch.Generator().manual_seed(5)
    M, N, K = 300, 256, 192
    A, B = torch.randn(M, K, generator=g).half().cuda(), torch.randn(K, N, generator=g).half().cuda()
    mask = torch.randn(M, N, generator=g).relu().half().cuda()
    out16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
    L.gemm(A, B, M, N, K, b_mn=True, relu_mask=mask, out_f16=out16)
    ref = (A.float() @ B.float()) * (mask.float() > 0)
    assert G.rel_err(out16.float().cpu(), ref.cpu()) < 6e-4
    # forward linear: out16 = relu(.) before the positional addend, out32 after it
    W = torch.randn(N, K, generator=g).half().cuda()
    bias, pe = torch.randn(N, generator=g).cuda(), torch.randn(50, N, generator=g).cuda()
    o32, o16 = torch.empty(M, N, device="cuda"), torch.empty(M, N, dtype=torch.float16, device="cuda")
    L.linear(A, W, bias, act=L.ACT_RELU, addend=pe, add_period=50, out_f32=o32, out_f16=o16, out16_pre_add=True)
    pre = (A.float() @ W.float().t() + bias).relu()
    assert G.rel_err(o16.float().cpu(), pre.cpu()) < 6e-4
    assert G.rel_err(o32.cpu(), (pre + pe.repeat(6, 1)).cpu()) < 2e-5


@pytest.mark.parametrize("M,N,K", [(77, 128, 128), (8192, 512, 512), (2048, 2048, 512), (640, 3000, 512)])
def test_linear_dgrad_wgrad_vs_autograd(L, M, N, K):
    """y = x W^T: dX = dY W, dW = dY^T X on f16-rounded operands, incl. the device-side gradient scale."""
    g = torch.Generator().manual_seed(M + N + K)
    x = rnd16(torch.randn(M, K, generator=g))
    w = rnd16(torch.randn(N, K, generator=g) * 0.05)
    dy = torch.randn(M, N, generator=g) * 1e-6            # tiny gradients: would underflow f16 unscaled
    dy32 = dy.cuda()
    S2 = L.grad_scale([dy32])
    S = float(S2[0])
    assert 128.0 <= float(dy.abs().max()) * S < 256.0 and abs(float(S2[1]) * S - 1) < 1e-6
    dy16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
    db = torch.zeros(N, device="cuda")
    L.cast_colsum(dy32, dst_f16=dy16, colsum=db, scale=S2[0:1], alpha=S2[1:2])
    dyr = dy16.float().cpu() / S                            # what the tensor core sees, un-scaled
    assert G.rel_err(dyr, dy) < 6e-4
    assert G.rel_err(db.cpu(), dy.sum(0)) < 1e-4
    dx = torch.empty(M, K, device="cuda")
    L.linear_dgrad(dy16, w.half().cuda(), alpha=S2[1:2], out_f32=dx)
    dW = torch.zeros(N, K, device="cuda")
    db2 = db.clone()                              # bias gradient from the same kernel (extra MMA against ones), accumulated
    L.linear_wgrad(dy16, x.half().cuda(), dW, alpha=S2[1:2], dbias=db2)
    torch.cuda.synchronize()
    assert G.rel_err(dx.cpu(), dyr @ w) < 3e-5
    assert G.rel_err(dW.cpu(), dyr.t() @ x) < 3e-5
    assert G.rel_err(db2.cpu(), db.cpu() + dyr.sum(0)) < 3e-5


# ------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("rows,d", [(1, 4), (7, 128), (33, 512), (5, 1024), (3, 96), (5000, 512)])
def test_layernorm_bwd(L, rows, d):
    g = torch.Generator().manual_seed(rows * 100 + d)
    x = (torch.randn(rows, d, generator=g) * 3 + 1).requires_grad_(True)
    a = (1 + 0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    dy = torch.randn(rows, d, generator=g)
    dres = torch.randn(rows, d, generator=g)
    O.layer_norm(x, a, b, 1e-6).backward(dy)
    dx = dres.clone().cuda()
    da, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    L.layernorm_bwd(dev(x.detach()), dev(a.detach()), 1e-6, dev(dy), dx, dres=dx, da_2=da, db_2=db)
    torch.cuda.synchronize()
    assert G.rel_err(dx.cpu() - dres, x.grad) < 1e-5
    assert G.rel_err(da.cpu(), a.grad) < 2e-5 and G.rel_err(db.cpu(), b.grad) < 2e-5
    # scaled entry: dy * S in, parameter gradients / S out
    S2 = torch.tensor([64.0, 1 / 64.0], device="cuda")
    dx2 = torch.empty(rows, d, device="cuda")
    da.zero_(); db.zero_()
    L.layernorm_bwd(dev(x.detach()), dev(a.detach()), 1e-6, dev(dy), dx2, da_2=da, db_2=db, dy_scale=S2[0:1],
                    param_alpha=S2[1:2])
    assert G.rel_err(dx2.cpu() / 64.0, x.grad) < 1e-5 and G.rel_err(da.cpu(), a.grad) < 2e-5
    if d in (128, 512, 1024):      # fused hand-off to the next backward step: f16 copy of dx and its column sums
        dx3, dx16 = dres.clone().cuda(), torch.empty(rows, d, dtype=torch.float16, device="cuda")
        cs = torch.ones(d, device="cuda")
        L.layernorm_bwd(dev(x.detach()), dev(a.detach()), 1e-6, dev(dy), dx3, dres=dx3, dx_f16=dx16, dx_colsum=cs,
                        param_alpha=S2[1:2])
        assert torch.equal(dx3, dx) and torch.equal(dx16, dx.half())
        assert G.rel_err(cs.cpu() - 1, dx.cpu().sum(0) / 64.0) < 2e-5


def test_cast_colsum_scale_f32_delta(L):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(333, 520, generator=g)
    m = torch.randn(333, 520, generator=g).half()
    y16 = torch.empty(333, 520, dtype=torch.float16, device="cuda")
    cs = torch.ones(520, device="cuda")
    sc, al = torch.tensor([4.0], device="cuda"), torch.tensor([0.25], device="cuda")
    L.cast_colsum(dev(x), dst_f16=y16, colsum=cs, scale=sc, alpha=al, relu_mask=dev(m))
    ref = x * 4 * (m.float() > 0)
    assert torch.equal(y16.cpu(), ref.half())
    assert G.rel_err(cs.cpu() - 1, ref.sum(0) * 0.25) < 1e-5
    cs2 = torch.zeros(520, device="cuda")
    L.cast_colsum(y16, colsum=cs2)                              # f16 source, column sums only
    assert G.rel_err(cs2.cpu(), y16.float().cpu().sum(0)) < 1e-5
    out = torch.ones(333 * 520, device="cuda")
    L.scale_f32(dev(x).view(-1), al, out, accumulate=True)
    assert G.rel_err(out.cpu(), 1 + 0.25 * x.reshape(-1)) < 1e-6
    # attention row term
    B, Lq, h, dk = 3, 37, 8, 64
    dO, Ot = torch.randn(B * Lq, h * dk, generator=g).half(), torch.randn(B * Lq, h * dk, generator=g).half()
    delta = torch.empty(B, h, Lq, device="cuda")
    L.attn_delta(dev(dO), dev(Ot), B, Lq, h, dk, delta)
    ref = (dO.float() * Ot.float()).view(B, Lq, h, dk).sum(-1).permute(0, 2, 1)
    assert G.rel_err(delta.cpu(), ref) < 1e-5


@pytest.mark.parametrize("with_ln", [True, False])
def test_embed_bwd(L, with_ln):
    g = torch.Generator().manual_seed(17)
    V, d, B, Lq = 50, 128, 4, 9
    ids = torch.randint(0, V, (B, Lq), generator=g)
    lut = torch.randn(V, d, generator=g).requires_grad_(True)
    pe = O.sinusoid_pe(d)[0, :64].contiguous()
    a = (1 + 0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    dy = torch.randn(B, Lq, d, generator=g)
    y = lut[ids] * (d ** 0.5) + pe[:Lq]
    if with_ln:
        y = O.layer_norm(y, a, b, 1e-6)
    y.backward(dy)
    dlut = torch.zeros(V, d, device="cuda")
    da, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    L.embed_bwd(dev(ids), dev(lut.detach()), dev(pe), d ** 0.5, dev(dy), dlut,
                ln=(dev(a.detach()), None, 1e-6) if with_ln else None, da_2=da if with_ln else None,
                db_2=db if with_ln else None)
    torch.cuda.synchronize()
    assert G.rel_err(dlut.cpu(), lut.grad) < 1e-5
    if with_ln:
        assert G.rel_err(da.cpu(), a.grad) < 2e-5 and G.rel_err(db.cpu(), b.grad) < 2e-5


def test_log_softmax_and_label_smoothing_bwd(L):
    g = torch.Generator().manual_seed(23)
    rows, V, V8 = 40, 203, 208
    z = (torch.randn(rows, V, generator=g) * 2).requires_grad_(True)
    dy = torch.randn(rows, V, generator=g)
    y = torch.log_softmax(z, -1)
    y.backward(dy)
    dz = torch.full((rows, V8), 7.0, device="cuda")
    L.log_softmax_bwd(dev(y.detach()), dev(dy), V, dz)
    assert G.rel_err(dz[:, :V].cpu(), z.grad) < 1e-5 and float(dz[:, V:].abs().max()) == 0
    for tgt_case in ("mixed", "lone_pad_row0"):
        tgt = torch.randint(2, V, (rows,), generator=g)
        if tgt_case == "mixed":
            tgt[5] = 1; tgt[17] = 1
        else:
            tgt[0] = 1                              # index sum 0: the reference does NOT zero this row
        z2 = torch.randn(rows, V, generator=g).requires_grad_(True)
        loss = O.label_smoothing_loss(torch.log_softmax(z2, -1), tgt, V, 1, 0.1) * 0.37
        loss.backward()
        zp = torch.zeros(rows, V8)
        zp[:, :V] = z2.detach()
        dz2 = torch.empty(rows, V8, device="cuda")
        gout = torch.tensor([2.0], device="cuda")
        L.label_smoothing_loss_bwd(dev(zp), V, dev(tgt), 1, 0.1, dz2, gscale=0.37 / 2.0, gout=gout)
        assert G.rel_err(dz2[:, :V].cpu(), z2.grad) < 2e-5, tgt_case
        assert float(dz2[:, V:].abs().max()) == 0


# ------------------------------------------------------------------ attention core backward
ATT_CASES = [
    # B, h, Lq, Lk, dk, mask kind
    (2, 4, 8, 16, 32, "keypad"), (2, 8, 64, 64, 64, "keypad"), (3, 8, 130, 300, 64, "keypad"),
    (2, 8, 256, 256, 64, "causal"), (2, 4, 20, 20, 32, "causal"), (2, 8, 40, 512, 64, "none"),
    (2, 8, 33, 70, 64, "allmasked_row"), (1, 8, 256, 512, 64, "keypad"),
    (2, 8, 256, 64, 64, "keypad"),       # one key tile, two query tiles: dQ stored as f16 per tile
    (24, 8, 130, 300, 64, "keypad"),     # 576 work items on 148 persistent CTAs: several items per CTA, double-buffered K/V
    (40, 4, 70, 64, 32, "causal_sq"),    # 160 single-tile items, d_k = 32
]


def _attn_ref(q, k, v, mask, B, h, Lq, Lk, dk):
    """autograd through the oracle's attention on (B, h, L, dk) views of the f16-rounded operands."""
    def heads(t, Lx):
        return t.view(B, Lx, h, dk).transpose(1, 2)
    o, _ = O.attention(heads(q, Lq), heads(k, Lk), heads(v, Lk), None if mask is None else mask.unsqueeze(1))
    return o.transpose(1, 2).reshape(B * Lq, h * dk)


@pytest.mark.parametrize("B,h,Lq,Lk,dk,kind", ATT_CASES)
def test_attn_core_bwd_vs_autograd(L, B, h, Lq, Lk, dk, kind):
    g = torch.Generator().manual_seed(B * 1000 + Lq * 10 + Lk)
    d = h * dk
    q = rnd16(torch.randn(B * Lq, d, generator=g)).requires_grad_(True)
    k = rnd16(torch.randn(B * Lk, d, generator=g)).requires_grad_(True)
    v = rnd16(torch.randn(B * Lk, d, generator=g)).requires_grad_(True)
    mask = None
    if kind == "keypad":
        mask = torch.ones(B, 1, Lk, dtype=torch.bool)
        for b in range(1, B):
            mask[b, 0, int(Lk * 0.6) + b:] = False
    elif kind == "causal":
        mask = O.subsequent_mask(Lq).expand(B, -1, -1).clone()
        mask[B - 1, :, Lk - 3:] = False
    elif kind == "causal_sq":            # rectangular "causal-like" mask (Lq != Lk): lower-triangular band
        mask = (torch.arange(Lk)[None, :] <= torch.arange(Lq)[:, None] + 3).expand(B, -1, -1).clone()
    elif kind == "allmasked_row":
        mask = torch.ones(B, 1, Lk, dtype=torch.bool)
        mask[1] = False                                   # every key masked: uniform softmax, dQ = dK = 0, dV != 0
    o = _attn_ref(q, k, v, mask, B, h, Lq, Lk, dk)
    dO = rnd16(torch.randn(B * Lq, d, generator=g))
    o.backward(dO)
    # --- device
    q16, k16, v16, dO16 = (t.detach().half().cuda() for t in (q, k, v, dO))
    bits = L.mask_pack(mask.cuda()) if mask is not None else None
    o16 = torch.empty(B * Lq, d, dtype=torch.float16, device="cuda")
    stats = torch.empty(B, h, Lq, 2, device="cuda")
    L.attn_core(q16, k16, v16, B, h, Lq, Lk, dk, o16, mask_bits=bits, stats=stats)
    assert G.rel_err(o16.float().cpu(), o.detach()) < 1e-3
    delta = torch.empty(B, h, Lq, device="cuda")
    L.attn_delta(dO16, o16, B, Lq, h, dk, delta)
    dq = torch.zeros(B * Lq, d, device="cuda")
    dk_ = torch.full((B * Lk, d), 9.0, dtype=torch.float16, device="cuda")
    dv_ = torch.full((B * Lk, d), 9.0, dtype=torch.float16, device="cuda")
    L.attn_core_bwd(q16, k16, v16, dO16, stats, delta, B, h, Lq, Lk, dk, dq, dk_, dv_, mask_bits=bits)
    torch.cuda.synchronize()
    errs = (G.rel_err(dq.cpu(), q.grad), G.rel_err(dk_.float().cpu(), k.grad), G.rel_err(dv_.float().cpu(), v.grad))
    print("attn bwd %s: dq %.2e dk %.2e dv %.2e" % ((B, h, Lq, Lk, dk, kind), *errs))
    assert max(errs) < 3e-3, errs
    if kind == "allmasked_row":
        assert float(dq[Lq:2 * Lq].abs().max()) == 0 and float(dk_[Lk:2 * Lk].abs().max()) == 0
        assert float(dv_[Lk:2 * Lk].abs().max()) > 0
    if Lk <= 128:          # one key tile: dQ can be stored as f16 directly (no f32 accumulation buffer)
        dq16 = torch.full((B * Lq, d), 9.0, dtype=torch.float16, device="cuda")
        L.attn_core_bwd(q16, k16, v16, dO16, stats, delta, B, h, Lq, Lk, dk, dq16, dk_, dv_, mask_bits=bits)
        assert torch.equal(dq16, dq.half())


# ------------------------------------------------------------------ model level
def _train_grads(mtn, du, cfg, sd, inp, smoothing=0.1):
    from mtn_b200 import label_smoothing
    model = mtn.make_model(cfg["vocab"], cfg["vocab"], N=cfg["N"], d_model=cfg["d_model"], d_ff=cfg["d_ff"], h=cfg["h"],
                           dropout=0.1, ft_sizes=cfg["ft_sizes"], diff_encoder=True, auto_encoder_ft="query")
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = lambda t: t.cuda()
    b = du.Batch(g(inp["query"]), g(inp["his"]), None, [g(f).permute(1, 0, 2).contiguous() for f in inp["fts"]],
                 g(inp["cap"]), g(inp["trg"]), g(inp["trg_y"]), 1)
    out, ae = model.forward(b)                                                    # train.py:33
    crit = label_smoothing.LabelSmoothing(size=cfg["vocab"], padding_idx=1, smoothing=smoothing)
    lc = du.SimpleLossCompute(model.generator, None, crit, opt=None)
    loss = lc.loss(out, b.trg_y, int(b.ntokens), ae, b.query, int((b.query != 1).sum()))   # train.py:37-39
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}


def _vs_f16_contract(grads, args, tag):
    """Against the oracle whose linears round their operands to f16 like the kernels (same arithmetic contract, no
    shared code): what is left is accumulation order and the attention core's P / dS rounding."""
    with O.f16_operand_linears():
        _, g16 = O.loss_and_grads(*args)
    errs = grad_errors(grads, g16)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print("train step %s vs f16-operand oracle: median %.2e worst %s" % (tag, float(np.median(list(errs.values()))), worst))
    assert max(errs.values()) < 3e-2, worst     # two f16-operand implementations differ by rounding noise of the same size


def test_training_step_vs_reference_golden_gradients():
    """BASELINE configs[0] family (N=1, d=128, h=4, d_k=32), ragged batch with an all-pad history row: loss and every
    parameter gradient against the UNMODIFIED reference's training step (digest) and the oracle's autograd."""
    from mtn_b200 import mtn, data_utils
    z, cfg, sd, inp = golden_grad_case()
    loss, grads = _train_grads(mtn, data_utils, cfg, sd, inp)
    norm = float((inp["trg_y"] != 1).sum())
    assert abs(loss * norm - float(z["loss_times_norm"])) <= 5e-3 * abs(float(z["loss_times_norm"]))
    oloss, og = O.loss_and_grads(sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
    errs = grad_errors(grads, og)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("train step d=128: loss %.6f vs %.6f; median %.2e worst grad errors: %s"
          % (loss, oloss, float(np.median(list(errs.values()))), worst))
    assert max(errs.values()) < 3e-2 and float(np.median(list(errs.values()))) < 8e-3, worst
    check_against_digest(z, grads, 8e-2)
    _vs_f16_contract(grads, (sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"]), "d=128")


def test_training_step_d512_vs_oracle():
    """cfg2 family (N=2, d=512, h=8, d_k=64), ragged batch."""
    from mtn_b200 import mtn, data_utils
    cfg = {"N": 2, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 200, "ft_sizes": [2048, 128],
           "auto_encoder_ft": "query", "diff_encoder": True}
    sd = O.init_state_dict(cfg, 3)
    inp = O.synth_inputs(cfg, B=4, Q=16, C=24, H=70, T=12, Lv=[140, 40], seed=5)
    loss, grads = _train_grads(mtn, data_utils, cfg, sd, inp)
    oloss, og = O.loss_and_grads(sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
    errs = grad_errors(grads, og)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("train step d=512: loss %.6f vs %.6f; median %.2e worst grad errors: %s"
          % (loss, oloss, float(np.median(list(errs.values()))), worst))
    assert abs(loss - oloss) <= 5e-3 * abs(oloss)
    assert max(errs.values()) < 3e-2 and float(np.median(list(errs.values()))) < 8e-3, worst
    _vs_f16_contract(grads, (sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"]), "d=512")


# ------------------------------------------------------------------ optimizer
def test_fused_adam_matches_torch_adam_and_noam_schedule(L):
    from mtn_b200.data_utils import NoamOpt
    g = torch.Generator().manual_seed(31)
    n = 4096 + 8
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone().cuda())
    topt = NoamOpt(512, 1, 40, torch.optim.Adam([ref], lr=0, betas=(0.9, 0.98), eps=1e-9))     # train.py:190-191
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p16 = torch.empty(n, dtype=torch.float16, device="cuda")
    state = torch.tensor([0.0, 0.9, 0.98, 1e-9, 1.0, 1.0, 0.0, 0.0], device="cuda")
    for it in range(5):
        grad = torch.randn(n, generator=g).cuda() * (10.0 ** (it - 2))
        ref.grad = grad.clone()
        topt.step()
        gbuf = grad.clone()
        L.adam_advance(state, (1.0, 512.0, 40.0))
        L.adam_step(p, gbuf, m, v, state, p_f16=p16, zero_grad=True)
        assert abs(float(state[0]) - topt.rate()) <= 1e-6 * topt.rate()
        assert float(gbuf.abs().max()) == 0.0
        assert G.rel_err(p.cpu(), ref.detach().cpu()) < 2e-6, it
        assert torch.equal(p16, p.half())


def test_reference_training_loop_three_steps_vs_oracle():
    """The reference's own loop (train.py:28-39, 190-209): make_model, LabelSmoothing, NoamOpt(Adam), SimpleLossCompute
    with opt -- driven exactly like train.py drives it, dropout 0 -- against the same three Adam steps taken with the
    CPU oracle's gradients.  Checks the loss trajectory (so the parameters after each update) end to end."""
    from mtn_b200 import mtn, data_utils, label_smoothing
    z, cfg, sd, inp = golden_grad_case()
    model = mtn.make_model(100, 100, N=1, d_model=128, d_ff=512, h=4, dropout=0.0, ft_sizes=[2048, 128], diff_encoder=True,
                           auto_encoder_ft="query")
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    criterion = label_smoothing.LabelSmoothing(size=100, padding_idx=1, smoothing=0.1)
    opt = data_utils.NoamOpt(128, 1, 4, torch.optim.Adam(model.parameters(), lr=0, betas=(0.9, 0.98), eps=1e-9))
    lc = data_utils.SimpleLossCompute(model.generator, model.auto_encoder_generator, criterion, opt=opt, l=1.0)
    g = lambda t: t.cuda()
    b = data_utils.Batch(g(inp["query"]), g(inp["his"]), None, [g(f).permute(1, 0, 2).contiguous() for f in inp["fts"]],
                         g(inp["cap"]), g(inp["trg"]), g(inp["trg_y"]), 1)
    got = []
    for _ in range(3):
        out, ae_out = model.forward(b)                                              # train.py:33
        ntokens_query = (b.query != 1).data.sum()                                   # train.py:38
        got.append(float(lc(out, b.trg_y, b.ntokens, ae_out, b.query, ntokens_query)))    # train.py:39 (loss * norm)
    # oracle: same optimizer on the state_dict tensors
    psd = {k: v.clone() for k, v in sd.items()}
    names = [k for k in psd if not k.endswith(".pe")]
    params = [torch.nn.Parameter(psd[k]) for k in names]
    oopt = data_utils.NoamOpt(128, 1, 4, torch.optim.Adam(params, lr=0, betas=(0.9, 0.98), eps=1e-9))
    norm = float((inp["trg_y"] != 1).sum())
    ref = []
    for _ in range(3):
        cur = dict(psd)
        cur.update({k: p.data for k, p in zip(names, params)})
        loss, grads = O.loss_and_grads(cur, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
        for k, p in zip(names, params):
            p.grad = grads[k]
        oopt.step()
        ref.append(loss * norm)
    print("reference loop losses", got, "oracle", ref)
    assert ref[2] < ref[0] and got[2] < got[0]
    for a, r in zip(got, ref):
        assert abs(a - r) <= 2e-3 * abs(r), (got, ref)


# ------------------------------------------------------------------ configuration variants of the training path
VARIANTS = [
    # auto_encoder_ft='caption': sublayers 2/3 swap, the QAE branch reads the caption (mtn.py:187-194), AE loss vs b.cap
    ({"N": 1, "d_model": 128, "d_ff": 256, "h": 4, "vocab": 50, "ft_sizes": [64], "auto_encoder_ft": "caption",
      "diff_encoder": True}, dict(B=3, Q=6, C=10, H=12, T=5, Lv=[17])),
    # diff_encoder=False: the auto-encoder streams start from the query memory itself (mtn.py:200-201, 205-208): its
    # gradient collects the K/V terms AND both streams' input gradients
    ({"N": 2, "d_model": 128, "d_ff": 256, "h": 4, "vocab": 50, "ft_sizes": [64, 32], "auto_encoder_ft": "query",
      "diff_encoder": False}, dict(B=3, Q=6, C=10, H=12, T=5, Lv=[17, 9])),
    # the stress configuration's family: d_model=1024, h=16 (d_k = 64), multi-tile sequences
    ({"N": 1, "d_model": 1024, "d_ff": 4096, "h": 16, "vocab": 120, "ft_sizes": [2048, 128], "auto_encoder_ft": "query",
      "diff_encoder": True}, dict(B=2, Q=20, C=30, H=150, T=140, Lv=[300, 70])),
]


@pytest.mark.parametrize("cfg,shape", VARIANTS, ids=["caption", "shared_ae", "d1024"])
def test_training_step_variants_vs_oracle(cfg, shape):
    from mtn_b200 import mtn, data_utils, label_smoothing
    sd = O.init_state_dict(cfg, 8)
    inp = O.synth_inputs(cfg, seed=9, **shape)
    model = mtn.make_model(cfg["vocab"], cfg["vocab"], N=cfg["N"], d_model=cfg["d_model"], d_ff=cfg["d_ff"], h=cfg["h"],
                           dropout=0.0, ft_sizes=cfg["ft_sizes"], diff_encoder=cfg["diff_encoder"],
                           auto_encoder_ft=cfg["auto_encoder_ft"])
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = lambda t: t.cuda()
    b = data_utils.Batch(g(inp["query"]), g(inp["his"]), None, [g(f).permute(1, 0, 2).contiguous() for f in inp["fts"]],
                         g(inp["cap"]), g(inp["trg"]), g(inp["trg_y"]), 1)
    out, ae = model.forward(b)
    ae_y = b.cap if cfg["auto_encoder_ft"] == "caption" else b.query                 # train.py:34-39
    lc = data_utils.SimpleLossCompute(model.generator, None, label_smoothing.LabelSmoothing(cfg["vocab"], 1, 0.1), opt=None)
    loss = lc.loss(out, b.trg_y, int(b.ntokens), ae, ae_y, int((ae_y != 1).sum()))
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    oloss, og = O.loss_and_grads(sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
    errs = grad_errors(grads, og)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print("variant %s: loss %.5f vs %.5f; median %.2e worst %s" % (cfg["auto_encoder_ft"], float(loss.detach()), oloss,
                                                                  float(np.median(list(errs.values()))), worst))
    assert abs(float(loss.detach()) - oloss) <= 5e-3 * abs(oloss)
    # tiny tensors (d=128, d_ff=256, a handful of tokens): a single ReLU flip moves an FFN gradient by percents, so the
    # per-tensor bar is wider than at d=512; the tensors that the variant-specific code paths feed are held tighter
    key = [k for k in errs if k.startswith("query_embed") or k.startswith("query_encoder.norm.0") or
           k.startswith("tgt_embed") or "src_attn" in k or "cap_attn" in k]
    print("   variant-specific tensors:", {k: round(errs[k], 4) for k in key[:8]})
    assert max(errs[k] for k in key) < 2.5e-2, {k: errs[k] for k in key}
    assert max(errs.values()) < 8e-2 and float(np.median(list(errs.values()))) < 1.5e-2, worst


class _RecordingOpt(object):
    """Optimizer stand-in: snapshots the gradients TrainStep hands to the optimizer instead of updating."""

    def __init__(self, model):
        self.model, self.grads = model, None

    def step(self):
        self.grads = {k: p.grad.detach().clone() for k, p in self.model.named_parameters()}


def test_trainstep_caption_variant_device_normalisers():
    """TrainStep must take the auto-encoder target and normaliser from auto_encoder_ft (b.cap / ntokens_cap
    for 'caption', train.py:34-36) -- caption and query lengths differ here, so the old hard-wired b.query would not even
    have matched the row count -- and count the normalisers on the device; loss and gradients vs the oracle's step."""
    from mtn_b200 import mtn
    from mtn_b200.trainer import TrainStep
    cfg, shape = VARIANTS[0]
    assert cfg["auto_encoder_ft"] == "caption" and shape["C"] != shape["Q"]
    sd = O.init_state_dict(cfg, 8)
    inp = O.synth_inputs(cfg, seed=9, **shape)
    model = mtn.make_model(cfg["vocab"], cfg["vocab"], N=cfg["N"], d_model=cfg["d_model"], d_ff=cfg["d_ff"], h=cfg["h"],
                           dropout=0.0, ft_sizes=cfg["ft_sizes"], diff_encoder=cfg["diff_encoder"],
                           auto_encoder_ft=cfg["auto_encoder_ft"])
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    for m in model.modules():                       # (MultiHeadedAttention keeps the reference's default p = 0.1)
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    rec = _RecordingOpt(model)
    ts = TrainStep(model, cfg["vocab"], graph=False, optimizer=rec)
    batch = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items()}
    loss = float(ts.eager(batch))                      # normalisers: None -> counted on the device
    oloss, og = O.loss_and_grads(sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
    assert [float(x) for x in ts.norms] == [float((inp["trg_y"] != 1).sum()), float((inp["cap"] != 1).sum())]
    errs = grad_errors(rec.grads, og)
    print("TrainStep caption variant: loss %.5f vs %.5f, gradient median %.2e max %.2e"
          % (loss, oloss, float(np.median(list(errs.values()))), max(errs.values())))
    assert abs(loss - oloss) <= 5e-3 * abs(oloss)
    assert max(errs.values()) < 8e-2 and float(np.median(list(errs.values()))) < 1.5e-2
    # a second batch with other token counts through the same object: its own normalisers
    inp2 = O.synth_inputs(cfg, seed=10, **shape)
    batch2 = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp2.items()}
    ts.eager(batch2)
    assert [float(x) for x in ts.norms] == [float((inp2["trg_y"] != 1).sum()), float((inp2["cap"] != 1).sum())]
