"""GPU tier: mtn_ln_linear_fwd (LayerNorm fused into the projection's operand staging, csrc/ln_gemm.cu)
against (a) the CPU oracle's LayerNorm (mtn.py:111-114 restated) followed by the f16-operand product and
(b) the two-launch form mtn_layernorm_fwd + mtn_linear_fwd it replaces: same LayerNorm formula, same f16
rounding of the operand, same k-order of the accumulation, so the two agree to accumulation-order level
(2e-5 normwise; the only freedom is the compiler's FMA contraction inside the LayerNorm, which can move a
single f16 operand element by one ulp -- measured: 0 to 21 differing output elements per case).

Edge cases: ragged row counts (M % 128 != 0, M < 128, M == 1), ragged N (N % 128 != 0), every supported d,
strided W / output views (a column block of a packed [Q|K|V] buffer), ReLU, no bias, more n-tiles than CTAs
per row block (tile loop + both TMEM accumulators), one CTA per row block with several n-tiles.
"""
import pytest
import torch

import golden_util as G
import mtn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from mtn_b200 import _lib
    _lib.lib()
    return _lib


CASES = [  # M, N, d, act, bias, strided
    (8192, 512, 512, 0, True, False),      # target-stream Q projection (cfg2, T=256): 64 row blocks x 2 CTAs
    (8192, 1536, 512, 0, True, True),      # packed Q|K|V of the self-attention site
    (2048, 2048, 512, 1, True, False),     # QAE FFN w_1 (ReLU)
    (1280, 512, 512, 0, True, False),      # decode step t=20, batch 64
    (64, 512, 512, 0, True, False),        # decode step t=1: one partial row block
    (1, 512, 512, 0, False, False),
    (300, 200, 128, 0, True, False),       # ragged M and N, d=128 (cfg1 family)
    (129, 128, 256, 1, True, True),
    (20000, 512, 512, 0, True, False),     # more row blocks than SMs: one CTA per row block walks 4 n-tiles
    (16, 3000 // 8 * 8, 512, 0, True, False),  # generator-like: 24 n-tiles over the machine
]


@pytest.mark.parametrize("M,N,d,act,use_bias,strided", CASES)
def test_ln_linear(L, M, N, d, act, use_bias, strided):
    assert L.ln_linear_supported(d)
    g = torch.Generator().manual_seed(M * 13 + N * 5 + d)
    x = torch.randn(M, d, generator=g) * 2 + 0.5
    a = 1 + 0.1 * torch.randn(d, generator=g)
    b = 0.1 * torch.randn(d, generator=g)
    ldw = d + 64 if strided else d
    Wfull = (torch.randn(N, ldw, generator=g) / d ** 0.5).half()
    bias = torch.randn(N, generator=g) if use_bias else None
    # oracle arithmetic: f32 LayerNorm, operand rounded to f16, exact product
    xn = O.layer_norm(x, a, b, 1e-6).half()
    ref = xn.double() @ Wfull[:, :d].double().t()
    if use_bias:
        ref = ref + bias.double()
    if act:
        ref = ref.clamp_min(0)
    ref = ref.float()

    xd, ad, bd = x.cuda(), a.cuda(), b.cuda()
    Wd = Wfull.cuda()[:, :d]
    biasd = bias.cuda() if use_bias else None
    ld16 = N + 8 if strided else N
    out_full = torch.zeros(M, ld16, device="cuda", dtype=torch.float16)
    out = out_full[:, :N]
    L.ln_linear(xd, ad, bd, 1e-6, Wd, bias=biasd, act=act, out_f16=out)
    # the two-launch form
    xn16 = torch.empty(M, d, device="cuda", dtype=torch.float16)
    two = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    L.layernorm(xd, ad, bd, 1e-6, out_f16=xn16)
    L.linear(xn16, Wd, bias=biasd, act=act, out_f16=two)
    torch.cuda.synchronize()
    e_ref = G.rel_err(out.float().cpu(), ref)
    e_two = G.rel_err(out.float().cpu(), two.float().cpu())
    nbad = int((out != two).sum())
    print("ln_linear M=%d N=%d d=%d: e_ref=%.3e e_two=%.3e differing elements=%d of %d" % (M, N, d, e_ref, e_two, nbad, M * N))
    assert e_ref < 6e-4, (e_ref, e_two, nbad)
    assert e_two < 2e-5 and nbad <= max(8, M * N // 1000), (e_two, nbad)
    if strided:
        assert float(out_full[:, N:].abs().sum()) == 0.0      # padding columns untouched


def test_ln_linear_rejects_unsupported(L):
    x = torch.randn(8, 1024, device="cuda")
    W = torch.randn(64, 1024, device="cuda").half()
    out = torch.empty(8, 64, device="cuda", dtype=torch.float16)
    assert not L.ln_linear_supported(1024)
    with pytest.raises(L.MtnError):
        L.ln_linear(x, torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda"), 1e-6, W, out_f16=out)


def test_ln_linear_inside_cuda_graph_with_pdl_neighbours(L):
    """The fused kernel between two of its consumers / producers on one stream, captured and replayed:
    programmatic dependent launch must not let it read the residual stream before the previous kernel's
    epilogue has written it."""
    M, d = 2048, 512
    g = torch.Generator().manual_seed(7)
    x0 = torch.randn(M, d, generator=g).cuda()
    a, b = (1 + 0.1 * torch.randn(d, generator=g)).cuda(), (0.1 * torch.randn(d, generator=g)).cuda()
    W1 = (torch.randn(d, d, generator=g) / d ** 0.5).half().cuda()
    W2 = (torch.randn(d, d, generator=g) / d ** 0.5).half().cuda()
    bias = torch.randn(d, generator=g).cuda()

    def chain(fused):
        x = x0.clone()
        q = torch.empty(M, d, device="cuda", dtype=torch.float16)
        xn = torch.empty(M, d, device="cuda", dtype=torch.float16)
        for _ in range(3):
            if fused:
                L.ln_linear(x, a, b, 1e-6, W1, bias=bias, out_f16=q)
            else:
                L.layernorm(x, a, b, 1e-6, out_f16=xn)
                L.linear(xn, W1, bias=bias, out_f16=q)
            L.linear(q, W2, bias=bias, addend=x, out_f32=x)     # residual update feeding the next round
        return x

    ref = chain(False)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain(True)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = chain(True)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    assert G.rel_err(out.cpu(), ref.cpu()) < 1e-5
