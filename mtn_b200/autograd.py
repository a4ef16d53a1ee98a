"""torch.autograd bindings of the backward kernels (ABI v3): one ``autograd.Function`` per reference module on
the training path -- embeddings (mtn.py:282-309), Encoder stream LayerNorm (mtn.py:83-101), video encoder
(mtn.py:377-379), the decoder cascade (mtn.py:158-218, ``train_engine.DecoderTrainer``), Generator
(mtn.py:62-69) and the label-smoothed criterion (label_smoothing.py:20-32).  Autograd only ORDERS these
Functions and sums gradients of shared parameters; every gradient is computed by a kernel of
``libmtn_b200.so``.  f16 tensor-core operands of a backward pass are scaled by a device-side power of two
(``_lib.grad_scale``) and results un-scaled in the producing kernels.
"""
import torch
from torch.autograd import Function

from . import _lib


def recording(module, *tensors):
    """True when autograd would record this call: grad mode on and an input or a parameter requires grad."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        return True
    return module is not None and any(p.requires_grad for p in module.parameters())


# ---- dropout seed: one 64-bit counter per device, in device memory so that a captured training step advances it
_SEEDS = {}


def manual_seed(seed, device=None):
    """Seed of the fused dropout (every training-mode forward draws a fresh value from this counter)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _SEEDS[dev] = torch.tensor([int(seed)], dtype=torch.int64, device=dev)


def next_seed(device):
    """Snapshot of the device's dropout counter for ONE forward pass (its backward regenerates the same decisions),
    then advance the counter.  Two tiny stream-ordered ops: CUDA-graph capturable, no host synchronisation."""
    dev = torch.device(device)
    if dev not in _SEEDS:
        manual_seed(0x6d746e, dev)
    snap = _SEEDS[dev].clone()
    _lib.seed_bump(_SEEDS[dev])
    return snap


def _linear_bwd(dy32, x16, w16, want_dx=True):
    """Gradients of y = x W^T + b given dy (f32 [rows, N]).  Returns (dx f32 or None, dW f32 [N, K], db f32 [N])."""
    rows, N = dy32.shape
    K = w16.shape[1]
    dev = dy32.device
    S2 = _lib.grad_scale([dy32])
    S, invS = S2[0:1], S2[1:2]
    dy16 = torch.empty(rows, N, dtype=torch.float16, device=dev)
    db = torch.zeros(N, dtype=torch.float32, device=dev)
    _lib.cast_colsum(dy32, dst_f16=dy16, colsum=db, scale=S, alpha=invS)
    dW = torch.zeros(N, K, dtype=torch.float32, device=dev)
    _lib.linear_wgrad(dy16, x16, dW, alpha=invS)
    dx = None
    if want_dx:
        dx = torch.empty(rows, K, dtype=torch.float32, device=dev)
        _lib.linear_dgrad(dy16, w16, alpha=invS, out_f32=dx)
    return dx, dW, db


class LayerNormFn(Function):
    """LayerNorm.forward (mtn.py:111-114) and its gradient."""

    @staticmethod
    def forward(ctx, x, a_2, b_2, eps):
        xc = x.contiguous().float()
        y = torch.empty_like(xc)
        _lib.layernorm(xc, a_2, b_2, eps, out_f32=y)
        ctx.save_for_backward(xc, a_2)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, a_2 = ctx.saved_tensors
        dyc = dy.contiguous().float()
        dx = torch.empty_like(x)
        da, db = torch.zeros_like(a_2), torch.zeros_like(a_2)
        _lib.layernorm_bwd(x, a_2, ctx.eps, dyc, dx, da_2=da, db_2=db)
        return dx, da, db, None


class EmbedFn(Function):
    """Embeddings * sqrt(d) + positional encoding (+ the Encoder's stream LayerNorm), mtn_embed_fwd / _bwd."""

    @staticmethod
    def forward(ctx, ids, lut, pe, scale, a_2, b_2, eps, p_drop=0.0):
        B, L = ids.shape
        out = torch.empty(B, L, lut.shape[1], dtype=torch.float32, device=ids.device)
        ln = None if a_2 is None else (a_2, b_2, eps)
        ctx.drop = _lib.drop_cfg(next_seed(ids.device), 0, p_drop) if p_drop > 0 else None   # mtn.py:309
        _lib.embed(ids, lut, pe, scale, ln=ln, out_f32=out, drop=ctx.drop)
        ctx.save_for_backward(ids, lut, pe, a_2)
        ctx.scale, ctx.eps = scale, eps
        return out

    @staticmethod
    def backward(ctx, dy):
        ids, lut, pe, a_2 = ctx.saved_tensors
        dyc = dy.contiguous().float()
        dlut = torch.zeros_like(lut)
        da = db = None
        ln = None
        if a_2 is not None:
            da, db = torch.zeros_like(a_2), torch.zeros_like(a_2)
            ln = (a_2, None, ctx.eps)
        _lib.embed_bwd(ids, lut, pe, ctx.scale, dyc, dlut, ln=ln, da_2=da, db_2=db, drop=ctx.drop)
        return None, dlut, None, None, da, db, None, None


class VideoEncoderFn(Function):
    """relu(ft W^T + b) + pe[t] (mtn.py:377-379) as one fused GEMM; the f16 ReLU output is kept as the mask."""

    @staticmethod
    def forward(ctx, ft16, weight, bias, w16, pe, Lv, p_drop=0.0):
        rows = ft16.shape[0]
        d = weight.shape[0]
        out = torch.empty(rows, d, dtype=torch.float32, device=ft16.device)
        relu16 = torch.empty(rows, d, dtype=torch.float16, device=ft16.device)
        ctx.drop = _lib.drop_cfg(next_seed(ft16.device), 0, p_drop) if p_drop > 0 else None   # mtn.py:309, after + pe
        _lib.linear(ft16, w16, bias, act=_lib.ACT_RELU, addend=pe, add_period=Lv, out_f32=out, out_f16=relu16,
                    out16_pre_add=True, drop=ctx.drop, drop_after_add=True)
        ctx.save_for_backward(ft16, relu16)
        ctx.wshape = weight.shape
        return out

    @staticmethod
    def backward(ctx, dy):
        ft16, relu16 = ctx.saved_tensors
        dyc = dy.contiguous().float()
        dev = dyc.device
        S2 = _lib.grad_scale([dyc])
        S, invS = S2[0:1], S2[1:2]
        dpre16 = torch.empty_like(relu16)
        db = torch.zeros(ctx.wshape[0], dtype=torch.float32, device=dev)
        _lib.cast_colsum(dyc, dst_f16=dpre16, colsum=db, scale=S, alpha=invS, relu_mask=relu16, drop=ctx.drop)
        dW = torch.zeros(ctx.wshape, dtype=torch.float32, device=dev)
        _lib.linear_wgrad(dpre16, ft16, dW, alpha=invS)
        return None, dW, db, None, None, None, None


class ProjectFn(Function):
    """Generator projection (mtn.py:68): logits[:, :V8] = x W^T + b with the vocabulary padded to a multiple of 8;
    ``log_probs`` additionally applies log_softmax over the first V columns (mtn.py:69)."""

    @staticmethod
    def forward(ctx, x, weight, bias, w16, b_pad, V, log_probs):
        d = weight.shape[1]
        x16 = _lib.cast_f16(x.contiguous().float().view(-1, d))
        logits = torch.empty(x16.shape[0], w16.shape[0], dtype=torch.float32, device=x.device)
        _lib.linear(x16, w16, b_pad, out_f32=logits)
        ctx.V, ctx.log_probs, ctx.xshape = V, log_probs, x.shape
        if log_probs:
            y = torch.empty(x16.shape[0], V, dtype=torch.float32, device=x.device)
            _lib.log_softmax(logits, V, out=y)
            ctx.save_for_backward(x16, w16, y)
            return y
        ctx.save_for_backward(x16, w16)
        return logits

    @staticmethod
    def backward(ctx, dy):
        V = ctx.V
        dyc = dy.contiguous().float()
        if ctx.log_probs:
            x16, w16, y = ctx.saved_tensors
            dz = torch.empty(x16.shape[0], w16.shape[0], dtype=torch.float32, device=dyc.device)
            _lib.log_softmax_bwd(y, dyc, V, dz)
        else:
            x16, w16 = ctx.saved_tensors
            dz = dyc          # columns [V, V8) of an upstream logits gradient are zero by construction
        dx, dW, db = _linear_bwd(dz, x16, w16)
        return dx.view(ctx.xshape), dW[:V], db[:V], None, None, None, None


class LabelSmoothingFn(Function):
    """Label-smoothed KL criterion (label_smoothing.py:20-32) from logits or log-probabilities."""

    @staticmethod
    def forward(ctx, z, target, V, padding_idx, smoothing, scale):
        loss = torch.zeros(1, dtype=torch.float32, device=z.device)
        tgt = target.reshape(-1).contiguous()
        _lib.label_smoothing_loss(z, V, tgt, padding_idx, smoothing, loss, scale=scale)
        ctx.save_for_backward(z, tgt)
        ctx.args = (V, padding_idx, smoothing, scale)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        z, tgt = ctx.saved_tensors
        V, padding_idx, smoothing, scale = ctx.args
        dz = torch.empty_like(z)
        _lib.label_smoothing_loss_bwd(z, V, tgt, padding_idx, smoothing, dz, gscale=scale,
                                      gout=g.reshape(1).contiguous().float())
        return dz, None, None, None, None, None


class DecoderFn(Function):
    """The N-layer decoder cascade (mtn.py:158-164) with a hand-written backward (train_engine.py)."""

    @staticmethod
    def forward(ctx, trainer, meta, n_vid, n_ae, *tensors):
        x, his, cap, qm = tensors[:4]
        vid = list(tensors[4:4 + n_vid])
        ae = list(tensors[4 + n_vid:4 + n_vid + n_ae])
        if meta["ae_list"]:
            ae_ft = ae
        else:
            ae_ft = ae[0] if n_ae == 1 else None
        f = lambda t: t.contiguous().float()
        seed = next_seed(x.device) if trainer.dec.training else None
        out, ae_outs, tape = trainer.forward([f(v) for v in vid], meta["vid_mask"], f(x), f(his), meta["his_mask"],
                                             f(cap), meta["cap_mask"], f(qm), meta["q_mask"], meta["tgt_mask"],
                                             [f(a) for a in ae_ft] if isinstance(ae_ft, list) else
                                             (f(ae_ft) if ae_ft is not None else None), meta["ae_features"], seed=seed)
        ctx.trainer, ctx.tape, ctx.n_vid, ctx.n_ae, ctx.ae_list = trainer, tape, n_vid, n_ae, meta["ae_list"]
        ctx.n_params = len(tensors) - 4 - n_vid - n_ae
        return (out,) + tuple(ae_outs)

    @staticmethod
    def backward(ctx, g_out, *g_ae):
        tape, ctx.tape = ctx.tape, None
        if tape is None:
            raise RuntimeError("mtn_b200: the decoder's backward can run only once per forward")
        grads, per = ctx.trainer.backward(tape, g_out, list(g_ae))
        g_qm, g_cap = grads["src"], grads["cap"]
        if ctx.ae_list:
            g_ae_in = list(grads["ae"])
        else:
            # one tensor feeds every modality's auto-encoder stream: the query / caption memory itself (ae_ft is
            # None, mtn.py:200-201, 205-208) or a single given tensor -- sum the per-modality gradients
            if ctx.n_ae == 1:
                tot = grads["ae"][0]
            else:
                tot = g_qm if tape["ae_shared"] == "src" else g_cap
                _lib.scale_f32(grads["ae"][0].view(-1), None, tot.view(-1), accumulate=True)
            for g in grads["ae"][1:]:
                _lib.scale_f32(g.view(-1), None, tot.view(-1), accumulate=True)
            g_ae_in = [tot] if ctx.n_ae == 1 else []
        params = ctx.trainer.param_list()
        assert len(params) == ctx.n_params
        # per is None: the gradients were accumulated straight into the attached .grad arena
        pg = tuple(per[id(p)] for p in params) if per is not None else (None,) * len(params)
        return (None, None, None, None, grads["x"], grads["his"], g_cap, g_qm) + tuple(grads["vid"]) + \
            tuple(g_ae_in) + pg
