"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A functional, fp32, CPU restatement of the MTN attention/FFN hot path of the
reference (``/root/reference/mtn.py`` @ 5105934).  It is written against a flat
``state_dict`` (the reference's own keys) instead of the
reference's nn.Module tree, so it shares no code with it; every function cites
the reference lines it restates.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this file against
(1) the three known-answer vectors extracted from the reference,
(2) golden tensors produced by importing the unmodified reference in the build
container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and
(3) -- when ``/root/reference`` is present -- the live reference, bit for bit.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module, and only as
the checker / CPU baseline.  The product package ``mtn_b200`` never does.
"""
import math

import torch


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def layer_norm(x, a_2, b_2, eps=1e-6):
    """mtn.py:111-114.  NOT nn.LayerNorm: unbiased std (divide by d-1) and eps is
    added to the *std*, not to the variance."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)           # unbiased by default, as in the reference
    return a_2 * (x - mean) / (std + eps) + b_2


def linear(x, w, b):
    """nn.Linear with [out, in] row-major weights (mtn.py:243-244, 273-276)."""
    return torch.nn.functional.linear(x, w, b)


def attention(q, k, v, mask):
    """mtn.py:221-231 with dropout disabled (eval).  ``mask`` broadcasts against
    (B, h, Lq, Lk); masked scores are set to the FINITE value -1e9, so a row whose
    keys are all masked becomes a uniform average over every key."""
    d_k = q.shape[-1]
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d_k)
    if mask is not None:
        s = s.masked_fill(mask == 0, -1e9)
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v), p


def mha(sd, pfx, h, query, key, value, mask):
    """mtn.py:248-267.  ``pfx`` is the state_dict prefix of a MultiHeadedAttention
    (its four nn.Linear live under ``linears.{0..3}``)."""
    if mask is not None:
        mask = mask.unsqueeze(1)                                   # mtn.py:250-252
    nb = query.shape[0]
    d_model = sd[pfx + "linears.0.weight"].shape[0]
    d_k = d_model // h

    def split(x, i):
        y = linear(x, sd[pfx + "linears.%d.weight" % i], sd[pfx + "linears.%d.bias" % i])
        return y.view(nb, -1, h, d_k).transpose(1, 2)              # mtn.py:256-258

    o, _ = attention(split(query, 0), split(key, 1), split(value, 2), mask)
    o = o.transpose(1, 2).contiguous().view(nb, -1, h * d_k)      # mtn.py:265-266
    return linear(o, sd[pfx + "linears.3.weight"], sd[pfx + "linears.3.bias"])


def ffn(sd, pfx, x):
    """mtn.py:279-280 (ReLU, not GELU; dropout off)."""
    hid = torch.relu(linear(x, sd[pfx + "w_1.weight"], sd[pfx + "w_1.bias"]))
    return linear(hid, sd[pfx + "w_2.weight"], sd[pfx + "w_2.bias"])


def sublayer(sd, pfx, x, fn):
    """mtn.py:125-127: pre-norm residual  x + f(norm(x))  (dropout off)."""
    return x + fn(layer_norm(x, sd[pfx + "norm.a_2"], sd[pfx + "norm.b_2"]))


# --------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------
def decoder_layer(sd, pfx, h, x, cap_mem, cap_mask, his_mem, his_mask, q_mem, q_mask,
                  tgt_mask, vid_fts, vid_mask, ae_fts, ae_features):
    """mtn.py:181-218.  ``pfx`` = 'decoder.layers.{l}.'"""
    c = 0

    def sub(xin, fn):
        nonlocal c
        y = sublayer(sd, pfx + "sublayer.%d." % c, xin, fn)
        c += 1
        return y

    x = sub(x, lambda t: mha(sd, pfx + "self_attn.", h, t, t, t, tgt_mask))          # :183
    x = sub(x, lambda t: mha(sd, pfx + "his_attn.", h, t, his_mem, his_mem, his_mask))  # :185
    if ae_features in ("caption", "summary"):                                            # :187-194
        x = sub(x, lambda t: mha(sd, pfx + "src_attn.", h, t, q_mem, q_mem, q_mask))
        x = sub(x, lambda t: mha(sd, pfx + "cap_attn.", h, t, cap_mem, cap_mem, cap_mask))
        if ae_fts is None:
            ae_fts = cap_mem
        ae_mask = cap_mask
    elif ae_features == "query":                                                         # :195-202
        x = sub(x, lambda t: mha(sd, pfx + "cap_attn.", h, t, cap_mem, cap_mem, cap_mask))
        x = sub(x, lambda t: mha(sd, pfx + "src_attn.", h, t, q_mem, q_mem, q_mask))
        if ae_fts is None:
            ae_fts = q_mem
        ae_mask = q_mask
    else:
        raise ValueError("auto_encoder_ft must be query|caption|summary (mtn.py:187-202)")
    out_ae = []
    for i, vid in enumerate(vid_fts):                                                    # :204-217
        ae = ae_fts[i] if isinstance(ae_fts, (list, tuple)) else ae_fts
        ae = sub(ae, lambda t: mha(sd, pfx + "auto_encoder_self_attn.%d." % i, h, t, t, t, ae_mask))
        ae = sub(ae, lambda t: mha(sd, pfx + "auto_encoder_vid_attn.%d." % i, h, t, vid, vid, vid_mask[i]))
        ae = sub(ae, lambda t: ffn(sd, pfx + "auto_encoder_feed_forward.%d." % i, t))
        x = sub(x, lambda t: mha(sd, pfx + "auto_encoder_attn.%d." % i, h, t, ae, ae, ae_mask))
        out_ae.append(ae)
    x = sub(x, lambda t: ffn(sd, pfx + "feed_forward.", t))                              # :218
    return x, out_ae


def decoder(sd, cfg, vid_ft, vid_mask, x, his_mem, his_mask, cap_mem, cap_mask, q_mem, q_mask,
            tgt_mask, ae_ft):
    """mtn.py:158-164."""
    for l in range(cfg["N"]):
        x, ae_ft = decoder_layer(sd, "decoder.layers.%d." % l, cfg["h"], x, cap_mem, cap_mask,
                                 his_mem, his_mask, q_mem, q_mask, tgt_mask, vid_ft, vid_mask,
                                 ae_ft, cfg["auto_encoder_ft"])
    out_ae = [layer_norm(a, sd["decoder.ae_norm.%d.a_2" % i], sd["decoder.ae_norm.%d.b_2" % i])
              for i, a in enumerate(ae_ft)]
    return layer_norm(x, sd["decoder.norm.a_2"], sd["decoder.norm.b_2"]), out_ae


# --------------------------------------------------------------------------
# feeders either side of the hot path (embeddings, video encoder, stream norms)
# --------------------------------------------------------------------------
def embed(sd, pfx, ids, d_model):
    """mtn.py:288-289 (lut * sqrt(d)) followed by mtn.py:307-309 (add sinusoid PE)."""
    x = torch.nn.functional.embedding(ids, sd[pfx + "0.lut.weight"]) * math.sqrt(d_model)
    return x + sd[pfx + "1.pe"][:, :x.shape[1]]


def vid_encode(sd, i, ft):
    """mtn.py:32-36, 377-379: Linear(F_i -> d) + ReLU + PE."""
    y = torch.relu(linear(ft, sd["vid_encoder.%d.0.weight" % i], sd["vid_encoder.%d.0.bias" % i]))
    return y + sd["vid_encoder.%d.2.pe" % i][:, :y.shape[1]]


def encode(sd, cfg, query, his, cap, fts):
    """mtn.py:38-56 + Encoder.forward mtn.py:83-101 for diff_encoder=True and no
    separate his/cap/ae embeddings (run.sh defaults): every text stream goes
    through ``query_embed``; one distinct LayerNorm per stream in the order
    query, vid_0..vid_{M-1}, cap, his, ae_0..ae_{M-1}."""
    d = cfg["d_model"]
    M = len(fts)
    ae_src = cap if cfg["auto_encoder_ft"] in ("caption", "summary") else query   # :40-43
    k = 0

    def norm(x):
        nonlocal k
        y = layer_norm(x, sd["query_encoder.norm.%d.a_2" % k], sd["query_encoder.norm.%d.b_2" % k])
        k += 1
        return y

    q_mem = norm(embed(sd, "query_embed.", query, d))
    vid_mem = [norm(vid_encode(sd, i, ft)) for i, ft in enumerate(fts)]
    cap_mem = norm(embed(sd, "query_embed.", cap, d))
    his_mem = norm(embed(sd, "query_embed.", his, d))
    if cfg.get("diff_encoder", True):
        ae_mem = [norm(embed(sd, "query_embed.", ae_src, d)) for _ in range(M)]
    else:
        ae_mem = None                                                                 # :54-56
    return q_mem, vid_mem, cap_mem, his_mem, ae_mem


# --------------------------------------------------------------------------
# Batch masks (data_utils.py:10-54)
# --------------------------------------------------------------------------
def subsequent_mask(size):
    """data_utils.py:10-14: (1, size, size) bool, True on and below the diagonal."""
    return torch.tril(torch.ones(1, size, size, dtype=torch.bool))


def make_masks(query, his, cap, trg, fts, pad):
    """data_utils.py:23-46.  Feature frames whose elements are ALL exactly 1.0 are
    padding (:29) and are zeroed (:30)."""
    m = {
        "query_mask": (query != pad).unsqueeze(-2),
        "his_mask": (his != pad).unsqueeze(-2),
        "cap_mask": (cap != pad).unsqueeze(-2),
    }
    if trg is not None:
        m["trg_mask"] = (trg != pad).unsqueeze(-2) & subsequent_mask(trg.shape[-1]).to(trg.device)
    m["fts_mask"] = [(torch.sum(ft != 1, dim=2) != 0).unsqueeze(-2) for ft in fts]
    m["fts"] = [ft * m["fts_mask"][i].squeeze(1).unsqueeze(-1).float() for i, ft in enumerate(fts)]
    return m


def forward(sd, cfg, query, his, cap, trg, fts, pad=1):
    """EncoderDecoder.forward, mtn.py:28-30.  Returns (out, [ae_i])."""
    with torch.no_grad():
        m = make_masks(query, his, cap, trg, fts, pad)
        q_mem, vid_mem, cap_mem, his_mem, ae_mem = encode(sd, cfg, query, his, cap, m["fts"])
        x = embed(sd, "tgt_embed.", trg, cfg["d_model"])                         # mtn.py:59
        return decoder(sd, cfg, vid_mem, m["fts_mask"], x, his_mem, m["his_mask"], cap_mem,
                       m["cap_mask"], q_mem, m["query_mask"], m["trg_mask"], ae_mem)


def generator(sd, x):
    """mtn.py:68-69."""
    return torch.log_softmax(linear(x, sd["generator.proj.weight"], sd["generator.proj.bias"]), dim=-1)


def label_smoothing_loss(logp, target, size, padding_idx, smoothing):
    """label_smoothing.py:20-32 + nn.KLDivLoss(size_average=False): sum over rows of KL(true_dist || exp(logp)).
    true_dist: smoothing/(size-2) everywhere, 1-smoothing at the target, 0 in the padding column; rows whose
    target is padding are zeroed ONLY IF the sum of their row indices is > 0 (the reference tests
    ``mask.sum() > 0`` on the index tensor, label_smoothing.py:26-30)."""
    assert logp.shape[1] == size
    t = torch.full_like(logp, smoothing / (size - 2))
    t.scatter_(1, target.unsqueeze(1), 1.0 - smoothing)
    t[:, padding_idx] = 0
    idx = torch.nonzero(target == padding_idx)
    if idx.sum() > 0 and len(idx) > 0:
        t.index_fill_(0, idx.squeeze(1), 0.0)
    return torch.nn.functional.kl_div(logp, t, reduction="sum")


def simple_loss(sd, cfg, out, trg_y, ae_out, ae_y, pad=1, smoothing=0.1, lam=1.0):
    """data_utils.py:132-156 (evaluation, opt=None): returns loss * norm like the reference."""
    V = sd["generator.proj.weight"].shape[0]
    norm = (trg_y != pad).sum().float()
    loss = label_smoothing_loss(generator(sd, out).reshape(-1, V), trg_y.reshape(-1), V, pad, smoothing) / norm
    ae_norm = (ae_y != pad).sum().float()
    for a in ae_out:
        loss = loss + lam * label_smoothing_loss(generator(sd, a).reshape(-1, V), ae_y.reshape(-1), V, pad,
                                                 smoothing) / ae_norm
    return float(loss) * float(norm)


def loss_and_grads(sd, cfg, query, his, cap, trg, trg_y, fts, pad=1, smoothing=0.1, lam=1.0, norm=None, ae_norm=None):
    """One training step's loss and parameter gradients with dropout disabled: train.py:33-39 (forward, loss on
    the decoder output AND on every auto-encoder stream against the un-shifted query ids, normalised by
    ntokens / ntokens_query) + data_utils.py:132-152 (loss.backward()).  Plain torch autograd through the
    functional restatement above.  Returns (loss / 1, {name: grad}) -- the loss is the NORMALISED one that is
    differentiated (the reference returns loss * norm)."""
    p = {k: (v.clone().requires_grad_(True) if not k.endswith(".pe") else v) for k, v in sd.items()}
    m = make_masks(query, his, cap, trg, fts, pad)
    q_mem, vid_mem, cap_mem, his_mem, ae_mem = encode(p, cfg, query, his, cap, m["fts"])
    x = embed(p, "tgt_embed.", trg, cfg["d_model"])
    out, ae_out = decoder(p, cfg, vid_mem, m["fts_mask"], x, his_mem, m["his_mask"], cap_mem, m["cap_mask"], q_mem,
                          m["query_mask"], m["trg_mask"], ae_mem)
    V = p["generator.proj.weight"].shape[0]
    ae_y = query if cfg.get("auto_encoder_ft", "query") == "query" else cap          # train.py:34-39
    # norm / ae_norm: token counts of THIS batch (the reference) unless the caller passes the global counts of a
    # sharded batch (data parallelism)
    norm = (trg_y != pad).sum().float() if norm is None else torch.tensor(float(norm))
    ae_norm = (ae_y != pad).sum().float() if ae_norm is None else torch.tensor(float(ae_norm))
    loss = label_smoothing_loss(generator(p, out).reshape(-1, V), trg_y.reshape(-1), V, pad, smoothing) / norm
    for a in ae_out:
        loss = loss + lam * label_smoothing_loss(generator(p, a).reshape(-1, V), ae_y.reshape(-1), V, pad,
                                                 smoothing) / ae_norm
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items() if not k.endswith(".pe")}
    return float(loss.detach()), grads


class _Linear16(torch.autograd.Function):
    """nn.Linear with the arithmetic CONTRACT of the sm_100a GEMM kernels and none of their code: operands rounded
    to f16 (forward: x, W; backward: dy scaled by a power of two, x, W), f32 accumulation."""

    @staticmethod
    def forward(ctx, x, w, b):
        x16, w16 = x.half().float(), w.half().float()
        ctx.save_for_backward(x16, w16)
        return x16 @ w16.t() + b

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        s = 2.0 ** (8 - math.frexp(float(dy.abs().max()) + 1e-300)[1])
        dy16 = (dy * s).half().float() / s
        dw = dy16.reshape(-1, dy16.shape[-1]).t() @ x16.reshape(-1, x16.shape[-1])
        return dy16 @ w16, dw, dy.reshape(-1, dy.shape[-1]).sum(0)


class f16_operand_linears(object):
    """Context manager: inside it every ``linear`` of this module rounds its operands like the tensor-core kernels
    do.  Used by the tests to separate "precision of the operand format" from "bug" in gradient comparisons."""

    def __enter__(self):
        global linear
        self._orig = linear
        linear = lambda x, w, b: _Linear16.apply(x, w, b)

    def __exit__(self, *exc):
        global linear
        linear = self._orig


def greedy_decode(sd, cfg, query, his, cap, fts, max_len, sos=2, pad=1):
    """The *intended* semantics of data_utils.py:162-186, using the working call
    form of data_utils.py:202-210 (the reference's greedy_decode raises TypeError;
    SURVEY 8a row G): no EOS stop, full-prefix recompute, argmax of the last row.
    Batched over dim 0 (the reference is batch-1; rows are independent)."""
    with torch.no_grad():
        m = make_masks(query, his, cap, None, fts, pad)
        q_mem, vid_mem, cap_mem, his_mem, ae_mem = encode(sd, cfg, query, his, cap, m["fts"])
        B = query.shape[0]
        ys = torch.full((B, 1), sos, dtype=torch.long)
        for _ in range(max_len - 1):
            x = embed(sd, "tgt_embed.", ys, cfg["d_model"])
            out, _ = decoder(sd, cfg, vid_mem, m["fts_mask"], x, his_mem, m["his_mask"], cap_mem,
                             m["cap_mask"], q_mem, m["query_mask"], subsequent_mask(ys.shape[1]),
                             ae_mem)
            nxt = generator(sd, out[:, -1]).argmax(-1)
            ys = torch.cat([ys, nxt.unsqueeze(1)], dim=1)
        return ys


# --------------------------------------------------------------------------
# model construction without the reference (for the GPU box)
# --------------------------------------------------------------------------
def sinusoid_pe(d_model, max_len=5000):
    """mtn.py:298-304."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0., max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def init_state_dict(cfg, seed):
    """A state_dict with the reference's keys and init *distribution*
    (Xavier-uniform on every >1-D tensor, mtn.py:410-412; nn.Linear default bias;
    LayerNorm a_2=1, b_2=0).  The RNG stream is this function's own -- it is not
    meant to reproduce ``make_model`` draw for draw (golden fixtures carry the
    reference's actual draws)."""
    g = torch.Generator().manual_seed(seed)
    d, dff, V, M = cfg["d_model"], cfg["d_ff"], cfg["vocab"], len(cfg["ft_sizes"])
    sd = {}

    def xavier(out_f, in_f):
        a = math.sqrt(6.0 / (in_f + out_f))
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = xavier(out_f, in_f)
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    def ln(name):
        sd[name + ".a_2"] = torch.ones(d)
        sd[name + ".b_2"] = torch.zeros(d)

    def attn(name):
        for i in range(4):
            lin(name + ".linears.%d" % i, d, d)

    def ff(name):
        lin(name + ".w_1", dff, d)
        lin(name + ".w_2", d, dff)

    n_norm = 3 + 2 * M if cfg.get("diff_encoder", True) else 3 + M
    for k in range(n_norm):
        ln("query_encoder.norm.%d" % k)
    pe = sinusoid_pe(d)
    for i, f in enumerate(cfg["ft_sizes"]):
        lin("vid_encoder.%d.0" % i, d, f)
        sd["vid_encoder.%d.2.pe" % i] = pe.clone()
    for l in range(cfg["N"]):
        p = "decoder.layers.%d." % l
        for nm in ("self_attn", "src_attn", "his_attn", "cap_attn"):
            attn(p + nm)
        for i in range(M):
            attn(p + "auto_encoder_attn.%d" % i)
            attn(p + "auto_encoder_self_attn.%d" % i)
            attn(p + "auto_encoder_vid_attn.%d" % i)
            ff(p + "auto_encoder_feed_forward.%d" % i)
        ff(p + "feed_forward")
        for s in range(5 + 4 * M):
            ln(p + "sublayer.%d.norm" % s)
    ln("decoder.norm")
    for i in range(M):
        ln("decoder.ae_norm.%d" % i)
    a = math.sqrt(6.0 / (V + d))
    sd["query_embed.0.lut.weight"] = (torch.rand(V, d, generator=g) * 2 - 1) * a
    sd["query_embed.1.pe"] = pe.clone()
    sd["tgt_embed.0.lut.weight"] = (torch.rand(V, d, generator=g) * 2 - 1) * a
    sd["tgt_embed.1.pe"] = pe.clone()
    lin("generator.proj", V, d)
    return sd


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def synth_inputs(cfg, B, Q, C, H, T, Lv, seed, ragged=True, pad=1):
    """Seeded synthetic batch.  ids ~ randint(4, V); per-sample valid lengths
    ~ U[ceil(L/2), L] with sample 0 full length; text tails set to ``pad``,
    feature-frame tails set to all-1.0 (the reference's padding sentinel,
    data_handler.py:236).  If B > 2, sample 2 gets an all-pad history to exercise
    the uniform-softmax path (first-turn history, data_handler.py:113-114)."""
    g = torch.Generator().manual_seed(seed)
    V = cfg["vocab"]

    def ids(L):
        x = torch.randint(4, V, (B, L), generator=g)
        if ragged:
            for b in range(1, B):
                n = int(torch.randint((L + 1) // 2, L + 1, (1,), generator=g))
                x[b, n:] = pad
        return x

    query, his, cap, trg, trg_y = ids(Q), ids(H), ids(C), ids(T), None
    trg_y = torch.randint(4, V, (B, T), generator=g)
    trg_y[trg == pad] = pad
    if ragged and B > 2:
        his[2, :] = pad
    fts = []
    for L, F in zip(Lv, cfg["ft_sizes"]):
        f = torch.randn(B, L, F, generator=g)
        if ragged:
            for b in range(1, B):
                n = int(torch.randint((L + 1) // 2, L + 1, (1,), generator=g))
                f[b, n:] = 1.0
        fts.append(f)
    return {"query": query, "his": his, "cap": cap, "trg": trg, "trg_y": trg_y, "fts": fts}
