"""GPU tier: training-mode dropout inside the kernels (reference: nn.Dropout in attention() mtn.py:229-230, the FFN
mtn.py:280, SublayerConnection mtn.py:127, PositionalEncoding mtn.py:309).  torch's random stream cannot be
reproduced, so parity is: (1) every kernel draws EXACTLY the decisions of the documented counter-based contract
(oracle/philox.py, pinned to the published Philox known answers), (2) with those decisions as an explicit mask the
forward and the backward match torch autograd of the oracle's arithmetic, (3) model level: determinism under a seed,
keep statistics, and forward/backward consistency of a whole training step (directional derivative)."""
import numpy as np
import pytest
import torch

import golden_util as G
import mtn_oracle as O
import philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from mtn_b200 import _lib
    _lib.lib()
    return _lib


def seed_t(v):
    return torch.tensor([v], dtype=torch.int64, device="cuda")


def mask_of(seed, site, th, rows, cols):
    return torch.from_numpy(philox.keep_mask(seed, site, th, rows * cols).reshape(rows, cols))


def rnd16(t):
    return t.half().float()


@pytest.mark.parametrize("after_add", [False, True])
def test_linear_dropout_matches_contract(L, after_add):
    g = torch.Generator().manual_seed(1)
    M, N, K, p = 300, 264, 128, 0.25
    A, W = rnd16(torch.randn(M, K, generator=g)), rnd16(torch.randn(N, K, generator=g))
    bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    drop = L.drop_cfg(seed_t(991), 3, p)
    th = drop[2]
    keep = mask_of(991, 3, th, M, N).float() / (1 - th / 65536.0)
    pre = (A @ W.t() + bias).relu()
    ref = (pre + add) * keep if after_add else pre * keep + add
    out = torch.empty(M, N, device="cuda")
    L.linear(A.half().cuda(), W.half().cuda(), bias.cuda(), act=L.ACT_RELU, addend=add.cuda(), out_f32=out, drop=drop,
             drop_after_add=after_add)
    chk = torch.empty(M, N, device="cuda")
    L.linear(A.half().cuda(), W.half().cuda(), bias.cuda(), act=L.ACT_RELU, addend=add.cuda(), out_f32=chk, drop=drop,
             drop_after_add=after_add, _check_kernel=True)
    assert G.rel_err(out.cpu(), ref) < 2e-5 and G.rel_err(chk.cpu(), ref) < 2e-5
    # f16-only epilogue (FFN hidden layer)
    o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
    L.linear(A.half().cuda(), W.half().cuda(), bias.cuda(), act=L.ACT_RELU, out_f16=o16, drop=drop)
    assert G.rel_err(o16.float().cpu(), pre * keep) < 6e-4
    assert torch.equal(o16.cpu() > 0, (pre * keep).half() > 0)


def test_cast_colsum_and_layernorm_handoff_dropout(L):
    g = torch.Generator().manual_seed(2)
    rows, d, p = 70, 512, 0.1
    drop = L.drop_cfg(seed_t(5), 9, p)
    th = drop[2]
    keep = mask_of(5, 9, th, rows, d).float() / (1 - th / 65536.0)
    x = torch.randn(rows, d, generator=g)
    y16 = torch.empty(rows, d, dtype=torch.float16, device="cuda")
    cs = torch.zeros(d, device="cuda")
    L.cast_colsum(x.cuda(), dst_f16=y16, colsum=cs, drop=drop)
    assert torch.equal(y16.cpu(), (x * keep).half()) and G.rel_err(cs.cpu(), (x * keep).sum(0)) < 1e-5
    # LayerNorm backward: dx itself is NOT thinned, the hand-off copy and its column sums are
    xin, a, dy, dres = (torch.randn(rows, d, generator=g) for _ in range(4))
    a = a[0].contiguous()
    dx0 = dres.clone().cuda()
    L.layernorm_bwd(xin.cuda(), a.cuda(), 1e-6, dy.cuda(), dx0, dres=dx0)
    dx1, dx16, cs = dres.clone().cuda(), torch.empty(rows, d, dtype=torch.float16, device="cuda"), torch.zeros(d, device="cuda")
    L.layernorm_bwd(xin.cuda(), a.cuda(), 1e-6, dy.cuda(), dx1, dres=dx1, dx_f16=dx16, dx_colsum=cs, drop=drop)
    assert torch.equal(dx0, dx1)
    assert torch.equal(dx16.cpu(), (dx0.cpu() * keep).half()) and G.rel_err(cs.cpu(), (dx0.cpu() * keep).sum(0)) < 1e-5


@pytest.mark.parametrize("with_ln", [True, False])
def test_embed_dropout_fwd_bwd(L, with_ln):
    g = torch.Generator().manual_seed(3)
    V, d, B, Lq, p = 40, 128, 3, 11, 0.3
    ids = torch.randint(0, V, (B, Lq), generator=g)
    lut = torch.randn(V, d, generator=g).requires_grad_(True)
    pe = O.sinusoid_pe(d)[0, :32].contiguous()
    a = (1 + 0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    drop = L.drop_cfg(seed_t(77), 0, p)
    keep = mask_of(77, 0, drop[2], B * Lq, d).view(B, Lq, d).float() / (1 - drop[2] / 65536.0)
    y = (lut[ids] * d ** 0.5 + pe[:Lq]) * keep                    # PositionalEncoding: dropout(x + pe), mtn.py:309
    if with_ln:
        y = O.layer_norm(y, a, b, 1e-6)
    dy = torch.randn(B, Lq, d, generator=g)
    y.backward(dy)
    out = torch.empty(B, Lq, d, device="cuda")
    ln = (a.detach().cuda(), b.detach().cuda(), 1e-6) if with_ln else None
    L.embed(ids.cuda(), lut.detach().cuda(), pe.cuda(), d ** 0.5, ln=ln, out_f32=out, drop=drop)
    assert G.rel_err(out.cpu(), y.detach()) < 2e-6
    dlut, da, db = torch.zeros(V, d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    L.embed_bwd(ids.cuda(), lut.detach().cuda(), pe.cuda(), d ** 0.5, dy.cuda(), dlut, ln=ln,
                da_2=da if with_ln else None, db_2=db if with_ln else None, drop=drop)
    assert G.rel_err(dlut.cpu(), lut.grad) < 1e-5
    if with_ln:
        assert G.rel_err(da.cpu(), a.grad) < 2e-5


ATT = [(2, 4, 20, 37, 32, "keypad"), (2, 8, 130, 300, 64, "keypad"), (2, 8, 256, 256, 64, "causal"),
       (2, 8, 64, 64, 64, "allmasked_row")]


@pytest.mark.parametrize("B,h,Lq,Lk,dk,kind", ATT)
def test_attention_dropout_fwd_bwd_vs_autograd_with_explicit_mask(L, B, h, Lq, Lk, dk, kind):
    g = torch.Generator().manual_seed(B + Lq + Lk)
    d, p = h * dk, 0.1
    q = rnd16(torch.randn(B * Lq, d, generator=g)).requires_grad_(True)
    k = rnd16(torch.randn(B * Lk, d, generator=g)).requires_grad_(True)
    v = rnd16(torch.randn(B * Lk, d, generator=g)).requires_grad_(True)
    if kind == "keypad":
        mask = torch.ones(B, 1, Lk, dtype=torch.bool)
        mask[1, 0, int(Lk * 0.6):] = False
    elif kind == "causal":
        mask = O.subsequent_mask(Lq).expand(B, -1, -1).clone()
    else:
        mask = torch.ones(B, 1, Lk, dtype=torch.bool)
        mask[1] = False
    drop = L.drop_cfg(seed_t(4242), 17, p)
    th = drop[2]
    Lk32 = (Lk + 31) // 32 * 32
    keep = mask_of(4242, 17, th, B * h * Lq, Lk32).view(B, h, Lq, Lk32)[..., :Lk].float() / (1 - th / 65536.0)

    def heads(t, Lx):
        return t.view(B, Lx, h, dk).transpose(1, 2)
    s = heads(q, Lq) @ heads(k, Lk).transpose(-2, -1) / dk ** 0.5
    s = s.masked_fill(mask.unsqueeze(1) == 0, -1e9)
    pr = torch.softmax(s, -1) * keep                                   # mtn.py:228-230
    o = (pr @ heads(v, Lk)).transpose(1, 2).reshape(B * Lq, d)
    dO = rnd16(torch.randn(B * Lq, d, generator=g))
    o.backward(dO)
    q16, k16, v16, dO16 = (t.detach().half().cuda() for t in (q, k, v, dO))
    bits = L.mask_pack(mask.cuda())
    o16 = torch.empty(B * Lq, d, dtype=torch.float16, device="cuda")
    stats = torch.empty(B, h, Lq, 2, device="cuda")
    L.attn_core(q16, k16, v16, B, h, Lq, Lk, dk, o16, mask_bits=bits, stats=stats, drop=drop)
    assert G.rel_err(o16.float().cpu(), o.detach()) < 1.5e-3
    delta = torch.empty(B, h, Lq, device="cuda")
    L.attn_delta(dO16, o16, B, Lq, h, dk, delta)
    dq = torch.zeros(B * Lq, d, device="cuda")
    dk_ = torch.empty(B * Lk, d, dtype=torch.float16, device="cuda")
    dv_ = torch.empty(B * Lk, d, dtype=torch.float16, device="cuda")
    L.attn_core_bwd(q16, k16, v16, dO16, stats, delta, B, h, Lq, Lk, dk, dq, dk_, dv_, mask_bits=bits, drop=drop)
    errs = (G.rel_err(dq.cpu(), q.grad), G.rel_err(dk_.float().cpu(), k.grad), G.rel_err(dv_.float().cpu(), v.grad))
    print("attn dropout %s: dq %.2e dk %.2e dv %.2e" % ((B, h, Lq, Lk, dk, kind), *errs))
    assert max(errs) < 4e-3, errs


# ------------------------------------------------------------------ model level
def _model_and_batch(p):
    from mtn_b200 import mtn, data_utils
    cfg = {"N": 2, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 200, "ft_sizes": [2048, 128],
           "auto_encoder_ft": "query", "diff_encoder": True}
    sd = O.init_state_dict(cfg, 3)
    inp = O.synth_inputs(cfg, B=4, Q=16, C=24, H=70, T=12, Lv=[140, 40], seed=5)
    model = mtn.make_model(200, 200, N=2, d_model=512, d_ff=2048, h=8, dropout=p, ft_sizes=[2048, 128], diff_encoder=True,
                           auto_encoder_ft="query")
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = p                                   # also the attention modules (make_model leaves them at 0.1)
    g = lambda t: t.cuda()
    b = data_utils.Batch(g(inp["query"]), g(inp["his"]), None, [g(f).permute(1, 0, 2).contiguous() for f in inp["fts"]],
                         g(inp["cap"]), g(inp["trg"]), g(inp["trg_y"]), 1)
    return model, b, data_utils


def _loss(model, b, du, seed, backward=False):
    from mtn_b200 import autograd as AG, label_smoothing
    AG.manual_seed(seed)
    out, ae = model.forward(b)
    lc = du.SimpleLossCompute(model.generator, None, label_smoothing.LabelSmoothing(200, 1, 0.1), opt=None)
    loss = lc.loss(out, b.trg_y, int(b.ntokens), ae, b.query, int((b.query != 1).sum()))
    if backward:
        for q in model.parameters():
            q.grad = None
        loss.backward()
    return float(loss.detach())


def test_model_dropout_determinism_and_statistics():
    model, b, du = _model_and_batch(0.1)
    l1, l2, l3 = _loss(model, b, du, 11), _loss(model, b, du, 11), _loss(model, b, du, 12)
    assert l1 == l2 and l1 != l3                       # same seed -> bit-identical, new seed -> new masks
    model.eval()
    with torch.no_grad():
        le = _loss(model, b, du, 11)
    model.train()
    ls = [_loss(model, b, du, 100 + i) for i in range(8)]
    print("loss eval %.4f  train-mode dropout mean %.4f (min %.4f max %.4f)" % (le, np.mean(ls), min(ls), max(ls)))
    assert abs(np.mean(ls) - le) < 0.15 * le and max(ls) - min(ls) > 1e-4


def test_training_step_forward_backward_consistency_under_dropout():
    """Directional derivative: with the seed fixed the loss is a deterministic function of the parameters, so a step
    of size eps along -grad must lower it by eps * ||grad||^2 (first order).  A backward that regenerated different
    dropout decisions than its forward (or mis-scaled them) fails this by a wide margin."""
    for p in (0.0, 0.3):
        model, b, du = _model_and_batch(p)
        l0 = _loss(model, b, du, 5, backward=True)
        g2 = sum(float((q.grad.double() ** 2).sum()) for q in model.parameters() if q.grad is not None)
        eps = 0.02 * l0 / g2                             # aim at a 2 % decrease
        with torch.no_grad():
            for q in model.parameters():
                if q.grad is not None:
                    q.add_(q.grad, alpha=-eps)
        l1 = _loss(model, b, du, 5)
        ratio = (l0 - l1) / (eps * g2)
        print("p=%.1f: loss %.5f -> %.5f, predicted decrease %.5f, ratio %.3f" % (p, l0, l1, eps * g2, ratio))
        assert 0.8 < ratio < 1.1, ratio
