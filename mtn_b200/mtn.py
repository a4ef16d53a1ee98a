"""Host-side mirror of the reference's ``mtn.py`` operator interface.

Same class names, constructor signatures, attribute names (hence identical
``state_dict`` keys -- SURVEY.md 8b) and call protocol as henryhungle/MTN's
``mtn.py``, so ``train.py`` / ``generate.py`` / ``data_utils.py`` of the reference
drive it unchanged (see ``compat/`` and INTEGRATION.md).  The arithmetic of the hot
path -- LayerNorm, MultiHeadedAttention, PositionwiseFeedForward, SublayerConnection,
DecoderLayer, Decoder (mtn.py:103-280 of the reference) -- runs in hand-written
sm_100a kernels behind the C ABI of ``include/mtn_b200.h``; PyTorch only owns the
memory.  There is no CPU / eager fallback: on a machine without the CUDA library
the hot-path modules raise.

Three levels of entry, all the same kernels:
  * module level  -- ``LayerNorm``, ``MultiHeadedAttention``, ``PositionwiseFeedForward``
    are drop-ins for the reference modules of the same name (f32 in, f32 out);
  * sublayer level -- ``DecoderLayer.forward`` issues one ``mtn_attn_site_fwd`` /
    ``mtn_ffn_fwd`` per SublayerConnection (LN + projections + core + residual);
  * model level   -- ``Decoder.forward`` runs the whole N-layer cascade through
    ``engine.DecoderEngine`` (hoisted memory K/V projections for all layers, the
    target-independent Query-Aware Auto-Encoder branch computed once per memory set).
"""
import copy
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from . import autograd as AG
from .engine import DecoderEngine, PackedWeights, ensure_inference
from .train_engine import DecoderTrainer


def clones(module, N):
    "N deep copies (reference mtn.py:71-73)."
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


# ----------------------------------------------------------------------------------
# hot-path modules
# ----------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """Reference mtn.py:103-114: a_2 * (x - mean) / (std_unbiased + eps) + b_2."""

    def __init__(self, features, eps=1e-6):
        super(LayerNorm, self).__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        if AG.recording(self, x):            # training: mtn_layernorm_fwd + mtn_layernorm_bwd
            if not x.is_cuda:
                raise _lib.MtnError("mtn_b200 hot path needs CUDA tensors (no CPU implementation)")
            return AG.LayerNormFn.apply(x, self.a_2, self.b_2, self.eps)
        ensure_inference(self, x)
        xc = x.contiguous().float()
        y = torch.empty_like(xc)
        _lib.layernorm(xc, self.a_2.data, self.b_2.data, self.eps, out_f32=y)
        return y


class SublayerConnection(nn.Module):
    """Reference mtn.py:116-127: x + dropout(sublayer(norm(x))) for an ARBITRARY callable.
    (The fused, structure-aware path is DecoderLayer / Decoder.)"""

    def __init__(self, size, dropout):
        super(SublayerConnection, self).__init__()
        self.norm = LayerNorm(size)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, sublayer):
        return x + self.dropout(sublayer(self.norm(x)))


def _mask_to_bits(mask, B, Lq, Lk):
    """Any reference-style mask -- (B,1,Lk), (B,Lq,Lk), (1,Lq,Lk) [subsequent_mask in decode,
    data_utils.py:205] -- to packed bit words."""
    if mask is None:
        return None
    if mask.dim() == 4:           # already unsqueezed for heads (mtn.py:252): same mask for all heads
        mask = mask[:, 0]
    assert mask.dim() == 3 and mask.shape[-1] == Lk and mask.shape[1] in (1, Lq), tuple(mask.shape)
    if mask.shape[0] != B:
        mask = mask.expand(B, -1, -1)
    return _lib.mask_pack(mask)


def attention(query, key, value, mask=None, dropout=None):
    """Reference mtn.py:221-231 on (B, h, L, d_k) tensors.  Returns (output, None): the
    probability matrix is never materialised (the reference only stores it in
    ``self.attn`` and nothing reads it)."""
    ensure_inference(None, query)
    B, h, Lq, dk = query.shape
    Lk = key.shape[2]

    def rows(t, L):   # (B,h,L,dk) -> [B*L, h*dk] f16
        return _lib.cast_f16(t.transpose(1, 2).reshape(B * L, h * dk).float())

    out = torch.empty(B * Lq, h * dk, dtype=torch.float16, device=query.device)
    _lib.attn_core(rows(query, Lq), rows(key, Lk), rows(value, Lk), B, h, Lq, Lk, dk, out,
                   mask_bits=_mask_to_bits(mask, B, Lq, Lk))
    return out.float().view(B, Lq, h, dk).transpose(1, 2), None


class MultiHeadedAttention(nn.Module):
    """Reference mtn.py:233-267."""

    def __init__(self, h, d_model, d_in=-1, dropout=0.1):
        super(MultiHeadedAttention, self).__init__()
        assert d_model % h == 0
        self.d_k = d_model // h
        self.h = h
        if d_in < 0:
            d_in = d_model
        self.linears = clones(nn.Linear(d_in, d_model), 3)
        self.linears.append(nn.Linear(d_model, d_in))
        self.attn = None
        self.dropout = nn.Dropout(p=dropout)
        self._packed = PackedWeights()

    def _weights(self):
        """f16 copies of the four nn.Linear weights, repacked when a parameter changes."""
        def build():
            w = [l.weight.data for l in self.linears]
            b = [l.bias.data for l in self.linears]
            return {
                "w_qkv": _lib.cast_f16(torch.cat(w[:3], 0).contiguous()),
                "b_qkv": torch.cat(b[:3], 0).contiguous(),
                "w_o": _lib.cast_f16(w[3].contiguous()), "b_o": b[3].contiguous(),
            }
        return self._packed.get(list(self.linears.parameters()), build)

    def forward(self, query, key, value, mask=None):
        ensure_inference(self, query)
        B, Lq, d_in = query.shape
        Lk = key.shape[1]
        d = self.h * self.d_k
        W = self._weights()
        dev = query.device

        def f16(t):
            return _lib.cast_f16(t.contiguous().float().view(-1, t.shape[-1]))

        if query is key and key is value:                       # self-attention: one [3d] GEMM
            qkv = torch.empty(B * Lq, 3 * d, dtype=torch.float16, device=dev)
            _lib.linear(f16(query), W["w_qkv"], W["b_qkv"], out_f16=qkv)
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        else:
            q = torch.empty(B * Lq, d, dtype=torch.float16, device=dev)
            _lib.linear(f16(query), W["w_qkv"][:d], W["b_qkv"][:d], out_f16=q)
            if key is value:                                     # memory: one [2d] GEMM
                kv = torch.empty(B * Lk, 2 * d, dtype=torch.float16, device=dev)
                _lib.linear(f16(key), W["w_qkv"][d:], W["b_qkv"][d:], out_f16=kv)
                k, v = kv[:, :d], kv[:, d:]
            else:
                k = torch.empty(B * Lk, d, dtype=torch.float16, device=dev)
                v = torch.empty(B * Lk, d, dtype=torch.float16, device=dev)
                _lib.linear(f16(key), W["w_qkv"][d:2 * d], W["b_qkv"][d:2 * d], out_f16=k)
                _lib.linear(f16(value), W["w_qkv"][2 * d:], W["b_qkv"][2 * d:], out_f16=v)
        o = torch.empty(B * Lq, d, dtype=torch.float16, device=dev)
        _lib.attn_core(q, k, v, B, self.h, Lq, Lk, self.d_k, o, mask_bits=_mask_to_bits(mask, B, Lq, Lk))
        self.attn = None     # never materialised (reference: mtn.py:261 stores it, nothing reads it)
        out = torch.empty(B * Lq, d_in, dtype=torch.float32, device=dev)
        _lib.linear(o, W["w_o"], W["b_o"], out_f32=out)
        return out.view(B, Lq, d_in)


class PositionwiseFeedForward(nn.Module):
    """Reference mtn.py:269-280 (ReLU)."""

    def __init__(self, d_model, d_ff, dropout=0.1, d_out=-1):
        super(PositionwiseFeedForward, self).__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        if d_out < 0:
            d_out = d_model
        self.w_2 = nn.Linear(d_ff, d_out)
        self.dropout = nn.Dropout(dropout)
        self._packed = PackedWeights()

    def _weights(self):
        def build():
            return {"w_1": _lib.cast_f16(self.w_1.weight.data.contiguous()), "b_1": self.w_1.bias.data,
                    "w_2": _lib.cast_f16(self.w_2.weight.data.contiguous()), "b_2": self.w_2.bias.data}
        return self._packed.get(list(self.parameters()), build)

    def forward(self, x):
        ensure_inference(self, x)
        W = self._weights()
        shp = x.shape
        x16 = _lib.cast_f16(x.contiguous().float().view(-1, shp[-1]))
        hid = torch.empty(x16.shape[0], W["w_1"].shape[0], dtype=torch.float16, device=x.device)
        _lib.linear(x16, W["w_1"], W["b_1"], act=_lib.ACT_RELU, out_f16=hid)
        out = torch.empty(x16.shape[0], W["w_2"].shape[0], dtype=torch.float32, device=x.device)
        _lib.linear(hid, W["w_2"], W["b_2"], out_f32=out)
        return out.view(*shp[:-1], out.shape[-1])


class DecoderLayer(nn.Module):
    """Reference mtn.py:166-218: the 5 + 4*M pre-norm residual sublayers."""

    def __init__(self, size, self_attn, cap_attn, his_attn, q_attn, auto_encoder_self_attn,
                 auto_encoder_vid_attn, auto_encoder_attn, feed_forward, auto_encoder_feed_forward,
                 dropout):
        super(DecoderLayer, self).__init__()
        self.size = size
        self.self_attn = self_attn
        self.src_attn = q_attn
        self.feed_forward = feed_forward
        self.his_attn = his_attn
        self.cap_attn = cap_attn
        self.auto_encoder_attn = auto_encoder_attn
        self.auto_encoder_self_attn = auto_encoder_self_attn
        self.auto_encoder_vid_attn = auto_encoder_vid_attn
        self.auto_encoder_feed_forward = auto_encoder_feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 5 + 4 * len(auto_encoder_vid_attn))

    # -- one fused SublayerConnection(attention) through the C ABI site entry point
    def _attn_site(self, c, att, x, mem, mask):
        B, Lq, d = x.shape
        self_att = mem is None
        Lk = Lq if self_att else mem.shape[1]
        W = att._weights()
        nrm = self.sublayer[c].norm
        a = _lib.AttnSiteArgs()
        a.B, a.Lq, a.Lk, a.d, a.h = B, Lq, Lk, d, att.h
        xc = x.contiguous()
        out = torch.empty_like(xc)
        a.x, a.x_out = xc.data_ptr(), out.data_ptr()
        a.ln_a, a.ln_b, a.ln_eps = nrm.a_2.data_ptr(), nrm.b_2.data_ptr(), nrm.eps
        keep = [xc, out, W]
        if self_att:
            a.w_q, a.b_q = W["w_qkv"].data_ptr(), W["b_qkv"].data_ptr()
        else:
            wq, bq, wkv, bkv = W["w_qkv"][:d], W["b_qkv"][:d], W["w_qkv"][d:], W["b_qkv"][d:]
            mem16 = _lib.cast_f16(mem.contiguous().float().view(-1, d))
            a.w_q, a.b_q, a.w_kv, a.b_kv = wq.data_ptr(), bq.data_ptr(), wkv.data_ptr(), bkv.data_ptr()
            a.mem_f16 = mem16.data_ptr()
            keep += [mem16]
        a.w_o, a.b_o = W["w_o"].data_ptr(), W["b_o"].data_ptr()
        bits = _mask_to_bits(mask, B, Lq, Lk)
        if bits is not None:
            a.mask_bits, a.mask_rows_q = bits.data_ptr(), bits.shape[1]
        nbytes = _lib.lib().mtn_attn_site_workspace_bytes(B, Lq, Lk, d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
        _lib.check(_lib.lib().mtn_attn_site_fwd(C.byref(a), _lib.stream_ptr()))
        return out

    def _ffn_site(self, c, ff, x):
        shp = x.shape
        d = shp[-1]
        W = ff._weights()
        nrm = self.sublayer[c].norm
        xc = x.contiguous()
        out = torch.empty_like(xc)
        a = _lib.FfnArgs()
        a.rows, a.d, a.d_ff = xc.numel() // d, d, W["w_1"].shape[0]
        a.x, a.x_out = xc.data_ptr(), out.data_ptr()
        a.ln_a, a.ln_b, a.ln_eps = nrm.a_2.data_ptr(), nrm.b_2.data_ptr(), nrm.eps
        a.w_1, a.b_1, a.w_2, a.b_2 = (W["w_1"].data_ptr(), W["b_1"].data_ptr(), W["w_2"].data_ptr(),
                                      W["b_2"].data_ptr())
        nbytes = _lib.lib().mtn_ffn_workspace_bytes(a.rows, a.d, a.d_ff)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
        _lib.check(_lib.lib().mtn_ffn_fwd(C.byref(a), _lib.stream_ptr()))
        return out

    def forward(self, x, cap_memory, cap_mask, his_memory, his_mask, q_memory, q_mask, tgt_mask,
                vid_fts, vid_mask, ae_fts, ae_features):
        ensure_inference(self, x)
        c = 0
        x = self._attn_site(c, self.self_attn, x, None, tgt_mask); c += 1
        x = self._attn_site(c, self.his_attn, x, his_memory, his_mask); c += 1
        if ae_features == 'caption' or ae_features == 'summary':
            x = self._attn_site(c, self.src_attn, x, q_memory, q_mask); c += 1
            x = self._attn_site(c, self.cap_attn, x, cap_memory, cap_mask); c += 1
            if ae_fts is None:
                ae_fts = cap_memory
            ae_mask = cap_mask
        elif ae_features == 'query':
            x = self._attn_site(c, self.cap_attn, x, cap_memory, cap_mask); c += 1
            x = self._attn_site(c, self.src_attn, x, q_memory, q_mask); c += 1
            if ae_fts is None:
                ae_fts = q_memory
            ae_mask = q_mask
        else:
            raise ValueError("auto_encoder_ft must be 'query', 'caption' or 'summary' "
                             "(the reference leaves ae_mask unbound otherwise, mtn.py:187-202)")
        out_ae_fts = []
        for i, vid_ft in enumerate(vid_fts):
            ae_ft = ae_fts[i] if type(ae_fts) == list else ae_fts
            ae_ft = self._attn_site(c, self.auto_encoder_self_attn[i], ae_ft, None, ae_mask); c += 1
            ae_ft = self._attn_site(c, self.auto_encoder_vid_attn[i], ae_ft, vid_ft, vid_mask[i]); c += 1
            ae_ft = self._ffn_site(c, self.auto_encoder_feed_forward[i], ae_ft); c += 1
            x = self._attn_site(c, self.auto_encoder_attn[i], x, ae_ft, ae_mask); c += 1
            out_ae_fts.append(ae_ft)
        return self._ffn_site(c, self.feed_forward, x), out_ae_fts


class Decoder(nn.Module):
    """Reference mtn.py:149-164."""

    def __init__(self, layer, N, ft_sizes=None):
        super(Decoder, self).__init__()
        self.layers = clones(layer, N)
        self.norm = LayerNorm(layer.size)
        self.ae_norm = nn.ModuleList()
        for ft_size in ft_sizes:
            self.ae_norm.append(LayerNorm(layer.size))
        self._engine = None
        self._trainer = None

    @property
    def engine(self):
        if self._engine is None:
            self._engine = DecoderEngine(self)
        return self._engine

    @property
    def trainer(self):
        if getattr(self, "_trainer", None) is None:
            self._trainer = DecoderTrainer(self)
        return self._trainer

    def __getstate__(self):          # torch.save(model) (train.py:217): engines are rebuilt lazily
        st = self.__dict__.copy()
        st["_engine"] = None
        st["_trainer"] = None
        return st

    def forward(self, vid_ft, vid_mask, x, his_memory, his_mask, cap_memory, cap_mask, query_memory,
                query_mask, tgt_mask, auto_encoded_ft, auto_encoded_features):
        ae_in = list(auto_encoded_ft) if isinstance(auto_encoded_ft, (list, tuple)) else \
            ([auto_encoded_ft] if auto_encoded_ft is not None else [])
        if AG.recording(self, x, his_memory, cap_memory, query_memory, *(list(vid_ft) + ae_in)):
            # training (train.py:33): one autograd.Function for the whole cascade, hand-written backward
            if not x.is_cuda:
                raise _lib.MtnError("mtn_b200 hot path needs CUDA tensors (no CPU implementation)")
            meta = {"vid_mask": list(vid_mask), "his_mask": his_mask, "cap_mask": cap_mask, "q_mask": query_mask,
                    "tgt_mask": tgt_mask, "ae_features": auto_encoded_features,
                    "ae_list": isinstance(auto_encoded_ft, (list, tuple))}
            res = AG.DecoderFn.apply(self.trainer, meta, len(vid_ft), len(ae_in), x, his_memory, cap_memory,
                                     query_memory, *(list(vid_ft) + ae_in + self.trainer.param_list()))
            return res[0], list(res[1:])
        ensure_inference(self, x)
        return self.engine.forward(vid_ft, vid_mask, x, his_memory, his_mask, cap_memory, cap_mask,
                                   query_memory, query_mask, tgt_mask, auto_encoded_ft,
                                   auto_encoded_features)


# ----------------------------------------------------------------------------------
# feeders and API shell (kept in PyTorch except LayerNorm and the video-encoder GEMM)
# ----------------------------------------------------------------------------------
class Encoder(nn.Module):
    """Reference mtn.py:75-101: one distinct LayerNorm per input stream."""

    def __init__(self, size, nb_layers):
        super(Encoder, self).__init__()
        self.norm = nn.ModuleList()
        self.nb_layers = nb_layers
        for n in range(nb_layers):
            self.norm.append(LayerNorm(size))

    def forward(self, *seqs):
        output, i = [], 0
        for seq in seqs:
            if isinstance(seq, list):
                group = []
                for s in seq:
                    group.append(self.norm[i](s)); i += 1
                output.append(group)
            else:
                output.append(self.norm[i](seq)); i += 1
            if i == self.nb_layers:
                break
        return output


class Generator(nn.Module):
    """Reference mtn.py:62-69: log_softmax(proj(x)).  SURVEY 8f row f2: the projection runs on the
    tcgen05 linear kernel (vocabulary padded to a multiple of 8 rows), followed by a row
    log-softmax kernel; ``argmax`` skips the log-softmax for greedy decoding (data_utils.py:183)."""

    def __init__(self, d_model, vocab):
        super(Generator, self).__init__()
        self.proj = nn.Linear(d_model, vocab)
        self._packed = PackedWeights()

    def packed(self):
        """{"w": f16 [V8, d] (vocabulary padded to a multiple of 8 rows, padding rows zero), "b": f32 [V8]}, rebuilt when
        the projection's parameters change."""
        V, d = self.proj.weight.shape
        V8 = (V + 7) // 8 * 8

        def build():
            w = torch.zeros(V8, d, dtype=torch.float32, device=self.proj.weight.device)
            w[:V] = self.proj.weight.data
            b = torch.zeros(V8, dtype=torch.float32, device=w.device)
            b[:V] = self.proj.bias.data
            return {"w": _lib.cast_f16(w), "b": b}
        return self._packed.get(list(self.proj.parameters()), build)

    def _logits(self, x, log_probs=False):
        V, d = self.proj.weight.shape
        V8 = (V + 7) // 8 * 8
        W = self.packed()
        if AG.recording(self, x):       # training: projection (+ log-softmax) with its backward kernels
            return AG.ProjectFn.apply(x, self.proj.weight, self.proj.bias, W["w"], W["b"], V, log_probs), V
        x16 = _lib.cast_f16(x.contiguous().float().view(-1, d))
        logits = torch.empty(x16.shape[0], V8, dtype=torch.float32, device=x.device)
        _lib.linear(x16, W["w"], W["b"], out_f32=logits)
        return logits, V

    def forward(self, x):
        if AG.recording(self, x):
            y, V = self._logits(x, log_probs=True)
            return y.view(*x.shape[:-1], V)
        ensure_inference(self, x)
        logits, V = self._logits(x)
        out = torch.empty(logits.shape[0], V, dtype=torch.float32, device=x.device)
        _lib.log_softmax(logits, V, out=out)
        return out.view(*x.shape[:-1], V)

    def argmax(self, x):
        """argmax_v log_softmax(proj(x))[..., v] without materialising the log-probabilities."""
        ensure_inference(self, x)
        logits, V = self._logits(x)
        idx = torch.empty(logits.shape[0], dtype=torch.int64, device=x.device)
        _lib.log_softmax(logits, V, argmax=idx)
        return idx.view(*x.shape[:-1])


class Embeddings(nn.Module):
    """Reference mtn.py:282-289."""

    def __init__(self, d_model, vocab):
        super(Embeddings, self).__init__()
        self.lut = nn.Embedding(vocab, d_model)
        self.d_model = d_model

    def forward(self, x):
        return self.lut(x) * math.sqrt(self.d_model)


class PositionalEncoding(nn.Module):
    """Reference mtn.py:291-309."""

    def __init__(self, d_model, dropout, max_len=5000):
        super(PositionalEncoding, self).__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0., max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0))

    def forward(self, x):
        return self.dropout(x + self.pe[:, :x.size(1)])


class VideoEncoder(nn.Sequential):
    """``Linear(F_i -> d) + ReLU + PositionalEncoding`` (reference mtn.py:377-379; an
    nn.Sequential there, so the state_dict keys stay ``{0.weight, 0.bias, 2.pe}``).
    SURVEY 8f row f1: the GEMM, ReLU and positional add are one launch of the fused
    linear kernel (ReLU first, then + pe[t], exactly the reference's order)."""

    def __init__(self, ft_size, d_model, position):
        super(VideoEncoder, self).__init__(nn.Linear(ft_size, d_model), nn.ReLU(), position)
        self._packed = PackedWeights()

    def forward(self, ft):
        lin, pos = self[0], self[2]
        B, Lv, Fdim = ft.shape
        W = self._packed.get(list(lin.parameters()),
                             lambda: {"w": _lib.cast_f16(lin.weight.data.contiguous()), "b": lin.bias.data})
        if ft.dtype == torch.float16:      # Batch already produced the masked f16 operand (feature_prep kernel)
            x16 = ft.contiguous().view(B * Lv, Fdim)
        else:
            x16 = _lib.cast_f16(ft.detach().contiguous().float().view(B * Lv, Fdim))
        if AG.recording(self):             # training: the features are data, only W / b receive gradients
            out = AG.VideoEncoderFn.apply(x16, lin.weight, lin.bias, W["w"], pos.pe[0, :Lv], Lv,
                                          pos.dropout.p if self.training else 0.0)
            return out.view(B, Lv, -1)
        ensure_inference(self, ft)
        out = torch.empty(B * Lv, lin.out_features, dtype=torch.float32, device=ft.device)
        _lib.linear(x16, W["w"], W["b"], act=_lib.ACT_RELU, addend=pos.pe[0, :Lv], add_period=Lv,
                    out_f32=out)
        return out.view(B, Lv, -1)


class EncoderDecoder(nn.Module):
    """Reference mtn.py:10-60."""

    def __init__(self, query_encoder, his_encoder, cap_encoder, vid_encoder, decoder, query_embed,
                 his_embed, cap_embed, tgt_embed, generator, diff_encoder=False,
                 auto_encoder_embed=None, auto_encoder_ft=None, auto_encoder_generator=None):
        super(EncoderDecoder, self).__init__()
        self.query_encoder = query_encoder
        self.his_encoder = his_encoder
        self.cap_encoder = cap_encoder
        self.vid_encoder = vid_encoder
        self.decoder = decoder
        self.query_embed = query_embed
        self.his_embed = his_embed
        self.cap_embed = cap_embed
        self.tgt_embed = tgt_embed
        self.generator = generator
        self.diff_encoder = diff_encoder
        self.auto_encoder_embed = auto_encoder_embed
        self.auto_encoder_ft = auto_encoder_ft
        self.auto_encoder_generator = auto_encoder_generator

    def forward(self, b):
        q, vid, cap, his, ae = self.encode(b.query, b.query_mask, b.his, b.his_mask, b.cap, b.cap_mask,
                                           b.fts, b.fts_mask)
        return self.decode(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask, b.trg,
                           b.trg_mask, ae)

    def vid_encode(self, video_features, video_features_mask, encoded_query=None):
        return [self.vid_encoder[i](ft) for i, ft in enumerate(video_features)]

    def _embed(self, seq, ids, norm=None):
        """``seq`` = nn.Sequential(Embeddings, PositionalEncoding) applied to ids, optionally followed by
        the Encoder's stream LayerNorm ``norm`` -- one fused kernel."""
        emb, pos = seq[0], seq[1]
        if AG.recording(seq) or (norm is not None and AG.recording(norm)):
            return AG.EmbedFn.apply(ids, emb.lut.weight, pos.pe[0], math.sqrt(emb.d_model),
                                    None if norm is None else norm.a_2, None if norm is None else norm.b_2,
                                    0.0 if norm is None else norm.eps, pos.dropout.p if seq.training else 0.0)
        ensure_inference(self, emb.lut.weight.data)
        B, L = ids.shape
        out = torch.empty(B, L, emb.d_model, dtype=torch.float32, device=ids.device)
        ln = None if norm is None else (norm.a_2.data, norm.b_2.data, norm.eps)
        _lib.embed(ids, emb.lut.weight.data, pos.pe[0], math.sqrt(emb.d_model), ln=ln, out_f32=out)
        return out

    def _fused_embed_ok(self):
        return (isinstance(self.query_embed, nn.Sequential) and len(self.query_embed) == 2 and
                isinstance(self.query_embed[0], Embeddings) and self.query_embed[0].d_model in (128, 256, 512, 1024)
                and self.query_embed[0].lut.weight.is_cuda)

    def encode(self, query, query_mask, his=None, his_mask=None, cap=None, cap_mask=None, vid=None,
               vid_mask=None):
        if self._fused_embed_ok():
            # same streams / LayerNorm order as Encoder.forward (mtn.py:83-101): query, vid_0.., cap, his, ae_0..
            nrm, M = self.query_encoder.norm, len(vid)
            q_mem = self._embed(self.query_embed, query, nrm[0])
            vid_mem = [nrm[1 + i](v) for i, v in enumerate(self.vid_encode(vid, vid_mask))]
            cap_mem = self._embed(self.query_embed, cap, nrm[1 + M])
            his_mem = self._embed(self.query_embed, his, nrm[2 + M])
            if not self.diff_encoder:
                return [q_mem, vid_mem, cap_mem, his_mem, None]
            if self.auto_encoder_ft in ('caption', 'summary'):
                ft = cap
            elif self.auto_encoder_ft == 'query':
                ft = query
            else:
                raise ValueError("auto_encoder_ft must be 'query', 'caption' or 'summary'")
            ae_mem = [self._embed(self.auto_encoder_embed[i] if self.auto_encoder_embed is not None
                                  else self.query_embed, ft, nrm[3 + M + i]) for i in range(M)]
            return [q_mem, vid_mem, cap_mem, his_mem, ae_mem]
        if self.diff_encoder:
            if self.auto_encoder_ft in ('caption', 'summary'):
                ft = cap
            elif self.auto_encoder_ft == 'query':
                ft = query
            else:
                raise ValueError("auto_encoder_ft must be 'query', 'caption' or 'summary'")
            if self.auto_encoder_embed is not None:
                ae_encoded = [self.auto_encoder_embed[i](ft) for i in range(len(vid))]
            else:
                ae_encoded = [self.query_embed(ft) for i in range(len(vid))]
            return self.query_encoder(self.query_embed(query), self.vid_encode(vid, vid_mask),
                                      self.query_embed(cap), self.query_embed(his), ae_encoded)
        output = self.query_encoder(self.query_embed(query), self.vid_encode(vid, vid_mask),
                                    self.query_embed(cap), self.query_embed(his))
        output.append(None)
        return output

    def decode(self, encoded_vid_features, his_memory, cap_memory, query_memory, vid_features_mask,
               his_mask, cap_mask, query_mask, tgt, tgt_mask, auto_encoded_ft):
        x = self._embed(self.tgt_embed, tgt) if self._fused_embed_ok() else self.tgt_embed(tgt)
        return self.decoder(encoded_vid_features, vid_features_mask, x, his_memory,
                            his_mask, cap_memory, cap_mask, query_memory, query_mask, tgt_mask,
                            auto_encoded_ft, self.auto_encoder_ft)


    # ---- KV-cached, last-token-only decoding (extension of the reference API) ----------------
    def decode_begin(self, encoded_vid_features, his_memory, cap_memory, query_memory, vid_features_mask,
                     his_mask, cap_mask, query_mask, auto_encoded_ft, max_len, rows_per_dialogue=1):
        """Same memories / masks as ``decode`` (data_utils.py:202-210), no target yet: returns the incremental
        decoding state (memory stage + per-layer self-attention caches for ``max_len`` positions).
        rows_per_dialogue > 1: that many hypotheses per dialogue (beam search), rows ordered dialogue-major."""
        ensure_inference(self, query_memory)
        return self.decoder.engine.decode_begin(encoded_vid_features, vid_features_mask, his_memory, his_mask,
                                                cap_memory, cap_mask, query_memory, query_mask, auto_encoded_ft,
                                                self.auto_encoder_ft, max_len, rows_per_dialogue)

    def decode_reorder(self, state, parents):
        """Beam search bookkeeping: row i of the next step continues the hypothesis of row parents[i]."""
        self.decoder.engine.decode_reorder(state, parents)

    def decode_step(self, state, tokens, t=None):
        """tokens: [B] int64 = the target token at position t (default: the next position of ``state``).  Returns the
        decoder output row of that position, [B, d] -- equal to ``decode(..., ys[:, :t+1], ...)[0][:, -1]`` of the
        full-prefix form (the causal mask makes row t independent of later rows; SURVEY 8a invariant ii)."""
        t = state["t"] if t is None else int(t)
        emb, pos = self.tgt_embed[0], self.tgt_embed[1]
        B = tokens.shape[0]
        x = torch.empty(B, 1, emb.d_model, dtype=torch.float32, device=tokens.device)
        _lib.embed(tokens.reshape(B, 1), emb.lut.weight.data, pos.pe[0][t:], math.sqrt(emb.d_model), out_f32=x)
        return self.decoder.engine.decode_step(state, x, t)

    def decode_step_argmax(self, state, tokens, t=None, out=None):
        """Greedy decoding's step (data_utils.py:180-184): ``generator.argmax(decode_step(state, tokens, t))`` -> int64 [B]
        (written into ``out`` when given; any stride).  Where the step runs as the cluster kernel (csrc/decode_cluster.cu)
        the generator's projection and the arg-max are its last stage: the whole position is embedding + ONE kernel."""
        t = state["t"] if t is None else int(t)
        emb, pos = self.tgt_embed[0], self.tgt_embed[1]
        B = tokens.shape[0]
        x = torch.empty(B, 1, emb.d_model, dtype=torch.float32, device=tokens.device)
        _lib.embed(tokens.reshape(B, 1), emb.lut.weight.data, pos.pe[0][t:], math.sqrt(emb.d_model), out_f32=x)
        if out is None:
            out = torch.empty(B, dtype=torch.int64, device=tokens.device)
        res = self.decoder.engine.decode_step(state, x, t, generator=self.generator, tokens_out=out)
        if res is not out:                       # (launch-sequence step: the generator runs on its own kernels)
            out.copy_(self.generator.argmax(res))
        return out


def make_model(src_vocab, tgt_vocab, N=6, d_model=512, d_ff=2048, h=8, dropout=0.1,
               separate_his_embed=False, separate_cap_embed=False, ft_sizes=None, diff_encoder=False,
               diff_embed=False, diff_gen=False, auto_encoder_ft=None, auto_encoder_attn=False):
    """Reference mtn.py:332-414.  Sub-modules are constructed in the reference's order so
    that, under the same ``torch.manual_seed``, the initial weights are identical."""
    c = copy.deepcopy
    attn = MultiHeadedAttention(h, d_model)
    ff = PositionwiseFeedForward(d_model, d_ff, dropout)
    position = PositionalEncoding(d_model, dropout)
    generator = Generator(d_model, tgt_vocab)
    query_embed = nn.Sequential(Embeddings(d_model, src_vocab), c(position))
    tgt_embed = nn.Sequential(Embeddings(d_model, tgt_vocab), c(position))
    his_embed = nn.Sequential(Embeddings(d_model, src_vocab), c(position)) if separate_his_embed else None
    cap_embed = nn.Sequential(Embeddings(d_model, src_vocab), c(position)) if separate_cap_embed else None
    auto_encoder_embed = None
    if diff_embed:
        auto_encoder_embed = nn.ModuleList(
            [nn.Sequential(Embeddings(d_model, src_vocab), c(position)) for _ in ft_sizes])
    query_encoder = Encoder(d_model, nb_layers=3 + (2 if diff_encoder else 1) * len(ft_sizes))
    self_attn, vid_attn, ae_ff = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
    vid_encoder, ae_attn = nn.ModuleList(), nn.ModuleList()
    for ft_size in ft_sizes:
        vid_encoder.append(VideoEncoder(ft_size, d_model, c(position)))
        self_attn.append(c(attn))
        vid_attn.append(c(attn))
        ae_ff.append(c(ff))
        ae_attn.append(c(attn))
    auto_encoder_generator = nn.ModuleList([c(generator) for _ in ft_sizes]) if diff_gen else None
    decoder = Decoder(DecoderLayer(d_model, c(attn), c(attn), c(attn), c(attn), self_attn, vid_attn,
                                   ae_attn, c(ff), ae_ff, dropout), N, ft_sizes)
    model = EncoderDecoder(query_encoder=query_encoder, his_encoder=None, cap_encoder=None,
                           vid_encoder=vid_encoder, decoder=decoder, query_embed=query_embed,
                           his_embed=his_embed, cap_embed=cap_embed, tgt_embed=tgt_embed,
                           generator=generator, auto_encoder_generator=auto_encoder_generator,
                           auto_encoder_embed=auto_encoder_embed, diff_encoder=diff_encoder,
                           auto_encoder_ft=auto_encoder_ft)
    for p in model.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
    return model
