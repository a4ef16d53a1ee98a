"""Training execution of the MTN decoder cascade: forward with a tape, hand-written backward.

``autograd.DecoderFn`` (forward -> ``DecoderTrainer.forward``, backward -> ``DecoderTrainer.backward``) is what
``mtn.Decoder.forward`` (reference mtn.py:158-164) runs when autograd is recording.  The whole N-layer cascade
(reference mtn.py:181-218 per layer) is ONE ``torch.autograd.Function``:
its forward launches the same sm_100a kernels as the inference engine and keeps, per SublayerConnection, the
f32 residual input, the f16 LayerNorm output, the f16 Q/K/V and attention outputs and the softmax statistics;
its backward walks that tape in reverse and launches the backward kernels of ``include/mtn_b200.h`` (ABI v3):

  out-projection / w_2 : dY as f16 (handed over by the previous LayerNorm backward)  ->  dgrad GEMM, split-K wgrad GEMM
  attention core       : delta = rowsum(dO * O);  tcgen05 backward kernel (dQ f16 or f32-accumulated, dK / dV f16)
  Q / QKV / w_1        : dgrad + wgrad GEMMs (ReLU mask in the dgrad epilogue, bias gradient inside the wgrad GEMM)
  LayerNorm            : row kernel, residual-stream gradient updated in place, f16 copy + bias gradient for the next sublayer
  hoisted memory K/V   : every layer writes its dK / dV columns into ONE [rows, N*2d] buffer per memory;
                         one wide dgrad + one wgrad per memory at the end.

All intermediate gradients are multiplied by a power of two S chosen on the device from the incoming gradient
(``_lib.grad_scale``) so that the f16 tensor-core operands do not underflow; every result leaving the Function
(parameter and input gradients) is multiplied by 1/S in the kernel that produces it.  No host synchronisation,
so forward + backward can be captured in a CUDA graph.

Streams.  Forward: each modality's Query-Aware Auto-Encoder chain (target-independent, mtn.py:209-213) runs on its
own side stream next to the target path, as in the inference engine.  Backward: the target chain is the critical
path on the caller's stream; every weight-gradient GEMM (nothing downstream reads it) goes to a dedicated stream,
and each modality's QAE chain runs on its side stream as soon as the target layer that feeds it is done.  Side
streams only LAUNCH: every buffer is allocated stream-ordered on the caller's stream and kept alive until the
streams are joined at the end of the pass.

PyTorch autograd only connects this Function with the embedding / encoder / generator Functions around it
(``mtn_b200/autograd.py``); no arithmetic of the path runs in PyTorch.
"""
import torch

from . import _lib
from .engine import PackedWeights


class _Arena(object):
    """A flat buffer carved into views (every view starts 16-byte aligned)."""

    def __init__(self, buf):
        self.buf = buf
        self.off = 0

    def take(self, *shape):
        n = 1
        for s in shape:
            n *= s
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 7) // 8 * 8
        return v


class DecoderTrainer(object):
    def __init__(self, decoder):
        self.dec = decoder
        self._packed = PackedWeights(follow_generation=False)
        self.on_target_grads_ready = None   # callable(event on the wgrad stream) or None, see _carve
        self._prefix_numel = 0
        self._arena = None          # flat f32 parameter arena + f16 operand arena (same layout)
        self._grad = None           # optional persistent gradient arena (same layout) backing every .grad
        self._mc = 0                # byte offset local -> NVLS multicast mapping of that arena (0: local accumulation)

    # ------------------------------------------------------------------ streams
    def _streams(self, M, dev):
        if getattr(self, "_side", None) is None or len(self._side) != M or self._side_dev != dev:
            self._side = [torch.cuda.Stream(device=dev) for _ in range(M)]
            self._wstream = torch.cuda.Stream(device=dev)
            self._side_dev = dev
        return self._side, self._wstream

    # ------------------------------------------------------------------ parameters
    def param_list(self):
        return list(self.dec.parameters())

    def arena_numel(self):
        return sum((p.numel() + 7) // 8 * 8 for p in self.dec.parameters())

    def _ensure_arena(self):
        """Flat parameter arena: every decoder parameter's storage becomes a view of ONE f32 buffer laid out like
        the gradient arena ([Wq;Wk;Wv] contiguous per self-attention-like module, every layer's [Wk_l;Wv_l] of a
        hoisted memory contiguous).  The f16 tensor-core operands of ALL parameters are then one cast kernel over
        the arena, and every packed operand (w_qkv, hoisted K/V weights, ...) is a view -- no torch.cat, no
        per-tensor casts.  state_dict keys / values are unchanged (parameters stay separate nn.Parameters)."""
        params = self.param_list()
        dev = params[0].device
        ar = self._arena
        if ar is not None and ar["dev"] == dev and all(p.data_ptr() == v.data_ptr() for p, v in zip(params, ar["views"])):
            return ar
        n = self.arena_numel()
        flat32 = torch.zeros(n, dtype=torch.float32, device=dev)
        P, per = self._carve(flat32)
        views = []
        with torch.no_grad():
            for p in params:
                v = per[id(p)]
                v.copy_(p.data)
                p.data = v
                views.append(v)
        flat16 = torch.empty(n, dtype=torch.float16, device=dev)
        H, _ = self._carve(flat16)
        self._arena = {"dev": dev, "flat32": flat32, "flat16": flat16, "P": P, "H": H, "views": views, "W": self._build_W(P, H)}
        self._packed._key = None
        return self._arena

    def _build_W(self, P, H):
        """The engine's weight dictionary (same structure as DecoderEngine.weights()) out of arena views: f16
        operands from H, f32 biases / LayerNorm parameters from P."""
        dec = self.dec
        layers = dec.layers
        M = len(layers[0].auto_encoder_vid_attn)
        d = layers[0].size

        def packed(key, l, m):
            return {"w_qkv": H[(key, l, "wqkv")], "b_qkv": P[(key, l, "bqkv")], "w_o": H[(key, l, "wo")],
                    "b_o": P[(key, l, "bo")], "h": m.h, "d_k": m.d_k, "mod": m}

        def hoisted_site(key, l, m):     # only the query projection is used per site ([:d] of a [d, d] view is itself)
            return {"w_qkv": H[(key, l, "wq")], "b_qkv": P[(key, l, "bq")], "w_o": H[(key, l, "wo")],
                    "b_o": P[(key, l, "bo")], "h": m.h, "d_k": m.d_k, "mod": m}

        def ffn(key, l, m):
            return {"w_1": H[(key, l, "w1")], "b_1": P[(key, l, "b1")], "w_2": H[(key, l, "w2")], "b_2": P[(key, l, "b2")],
                    "mod": m}

        W = {"M": M, "d": d, "N": len(layers), "layers": []}
        for name, key in (("kv_his", ("his", 0)), ("kv_cap", ("cap", 0)), ("kv_q", ("src", 0))):
            W[name] = (H[(key, "wkv")], P[(key, "bkv")])
        W["kv_vid"] = [(H[(("ae_vid", i), "wkv")], P[(("ae_vid", i), "bkv")]) for i in range(M)]
        for l, L in enumerate(layers):
            W["layers"].append({
                "self": packed(("self", 0), l, L.self_attn),
                "his": hoisted_site(("his", 0), l, L.his_attn), "cap": hoisted_site(("cap", 0), l, L.cap_attn),
                "src": hoisted_site(("src", 0), l, L.src_attn),
                "ae_self": [packed(("ae_self", i), l, L.auto_encoder_self_attn[i]) for i in range(M)],
                "ae_vid": [hoisted_site(("ae_vid", i), l, L.auto_encoder_vid_attn[i]) for i in range(M)],
                "ae_attn": [packed(("ae_attn", i), l, L.auto_encoder_attn[i]) for i in range(M)],
                "ae_ffn": [ffn(("ae_ffn", i), l, L.auto_encoder_feed_forward[i]) for i in range(M)],
                "ffn": ffn(("ffn", 0), l, L.feed_forward),
                "sub": L.sublayer,
                "ln": [(P[("ln", l, c)][0], P[("ln", l, c)][1], s.norm.eps) for c, s in enumerate(L.sublayer)],
            })
        W["norm"] = (P[("norm",)][0], P[("norm",)][1], dec.norm.eps)
        W["ae_norm"] = [(P[("ae_norm", i)][0], P[("ae_norm", i)][1], m.eps) for i, m in enumerate(dec.ae_norm)]
        return W

    def weights(self):
        """Tensor-core (f16) operands of the decoder's parameters: ONE cast kernel over the parameter arena,
        re-run when a parameter changed (optimizer step, load_state_dict)."""
        ar = self._ensure_arena()

        def build():
            n = ar["flat32"].numel()
            _lib.cast_f16(ar["flat32"].view(-1, 8), ar["flat16"].view(-1, 8))
            return ar["W"]
        return self._packed.get(self.param_list(), build)

    def attach_grads(self, buf, multicast_offset=0):
        """Back every decoder parameter's ``.grad`` by a view of `buf` (f32, arena_numel() elements, arena layout):
        the backward then accumulates straight into it -- no per-parameter autograd accumulation kernels.
        multicast_offset: `buf` is symmetric memory with an NVLS multicast mapping at data_ptr() + multicast_offset:
        every gradient reduction of the backward then goes to the multicast address (multimem.red), i.e. into the
        gradient buffers of ALL ranks -- the data-parallel all-reduce fused into the weight-gradient epilogues."""
        self._mc = int(multicast_offset)
        G, per = self._carve(buf)
        for p in self.param_list():
            p.grad = per[id(p)]
        self._grad = {"buf": buf, "G": G, "per": per}

    # ------------------------------------------------------------------ forward building blocks (taped)
    @staticmethod
    def _drop(dr, p):
        """Next dropout site of this forward pass: (seed, site, thresh) or None (eval mode / p == 0)."""
        if dr is None or p <= 0:
            return None
        dr["site"] += 1
        return _lib.drop_cfg(dr["seed"], dr["site"], p)

    @classmethod
    def _attn_fwd(cls, tape, x_in, ln, A, B, Lq, Lk, kv, k_col, v_col, bits, names, dr=None, p_sub=0.0):
        """One pre-norm residual attention site (mtn.py:125-127 around :248-267).  kv=None: self-attention.
        Training: dropout of the probabilities (mtn.py:230) and of the sublayer output (mtn.py:127)."""
        drop_p = cls._drop(dr, A["mod"].dropout.p)
        drop_o = cls._drop(dr, p_sub)
        rows, d = x_in.shape
        dev = x_in.device
        f16 = torch.float16
        xn16 = torch.empty(rows, d, dtype=f16, device=dev)
        _lib.layernorm(x_in, ln[0], ln[1], ln[2], out_f16=xn16)
        if kv is None:
            qbuf = torch.empty(rows, 3 * d, dtype=f16, device=dev)
            _lib.linear(xn16, A["w_qkv"], A["b_qkv"], out_f16=qbuf)
            q, k, v = qbuf[:, :d], qbuf[:, d:2 * d], qbuf[:, 2 * d:]
        else:
            qbuf = torch.empty(rows, d, dtype=f16, device=dev)
            _lib.linear(xn16, A["w_qkv"][:d], A["b_qkv"][:d], out_f16=qbuf)
            q, k, v = qbuf, kv[:, k_col:k_col + d], kv[:, v_col:v_col + d]
        o16 = torch.empty(rows, d, dtype=f16, device=dev)
        stats = torch.empty(B, A["h"], Lq, 2, dtype=torch.float32, device=dev)
        _lib.attn_core(q, k, v, B, A["h"], Lq, Lk, A["d_k"], o16, mask_bits=bits, stats=stats, drop=drop_p)
        x_out = torch.empty_like(x_in)
        _lib.linear(o16, A["w_o"], A["b_o"], addend=x_in, out_f32=x_out, drop=drop_o)
        tape.append(("attn", dict(x_in=x_in, ln=ln, A=A, B=B, Lq=Lq, Lk=Lk, self_attn=kv is None, xn16=xn16, q=q, k=k,
                                  v=v, o16=o16, stats=stats, bits=bits, names=names, drop_p=drop_p, drop_o=drop_o)))
        return x_out

    @classmethod
    def _ffn_fwd(cls, tape, x_in, ln, Fw, names, want16=False, dr=None, p_sub=0.0):
        drop_h = cls._drop(dr, Fw["mod"].dropout.p)          # dropout of relu(w_1 x) (mtn.py:280)
        drop_o = cls._drop(dr, p_sub)                        # dropout of the sublayer output (mtn.py:127)
        rows, d = x_in.shape
        dev = x_in.device
        f16 = torch.float16
        xn16 = torch.empty(rows, d, dtype=f16, device=dev)
        _lib.layernorm(x_in, ln[0], ln[1], ln[2], out_f16=xn16)
        hid = torch.empty(rows, Fw["w_1"].shape[0], dtype=f16, device=dev)
        _lib.linear(xn16, Fw["w_1"], Fw["b_1"], act=_lib.ACT_RELU, out_f16=hid, drop=drop_h)
        x_out = torch.empty_like(x_in)
        out16 = torch.empty(rows, d, dtype=f16, device=dev) if want16 else None
        _lib.linear(hid, Fw["w_2"], Fw["b_2"], addend=x_in, out_f32=x_out, out_f16=out16, drop=drop_o)
        tape.append(("ffn", dict(x_in=x_in, ln=ln, Fw=Fw, xn16=xn16, hid=hid, names=names, drop_h=drop_h, drop_o=drop_o)))
        return x_out, out16

    # ------------------------------------------------------------------ forward
    def forward(self, vid_ft, vid_mask, x, his, his_mask, cap, cap_mask, qm, q_mask, tgt_mask, ae_ft, ae_features,
                seed=None):
        """Returns (out [B,T,d], [ae_out_i [B,La,d]], ctx) with ctx the tape for ``backward``.
        seed: 1-element int64 device tensor -- the dropout seed of THIS pass (training mode); None: no dropout."""
        W = self.weights()
        dr = {"seed": seed, "site": 0} if (seed is not None and self.dec.training) else None
        psub = lambda l, c: (W["layers"][l]["sub"][c].dropout.p if dr is not None else 0.0)
        d, N, M = W["d"], W["N"], W["M"]
        B, T, _ = x.shape
        dev = x.device
        f16 = torch.float16
        if ae_features in ("caption", "summary"):
            ae_default = cap
        elif ae_features == "query":
            ae_default = qm
        else:
            raise ValueError("auto_encoder_ft must be 'query', 'caption' or 'summary' "
                             "(reference mtn.py:187-202 leaves ae_mask unbound otherwise)")
        La = ae_default.shape[1]

        def bits(mask, Bm=B):
            if mask is None:
                return None
            if mask.dim() == 4:
                mask = mask[:, 0]
            if mask.shape[0] != Bm:
                mask = mask.expand(Bm, -1, -1)
            return _lib.mask_pack(mask)

        def hoisted(mem, wb):
            m32 = mem.contiguous().view(-1, d)
            m16 = _lib.cast_f16(m32)
            out = torch.empty(m16.shape[0], N * 2 * d, dtype=f16, device=dev)
            _lib.linear(m16, wb[0], wb[1], out_f16=out)
            return m16, out

        ctx = {"W": W, "B": B, "T": T, "La": La, "ae_features": ae_features, "ae_shared": None}
        main = torch.cuda.current_stream()
        side, _ = self._streams(M, dev)
        b_his, b_cap, b_q = bits(his_mask), bits(cap_mask), bits(q_mask)
        b_ae = b_q if ae_features == "query" else b_cap
        b_vids = [bits(vid_mask[i]) for i in range(M)]
        vid_c = [v.contiguous() for v in vid_ft]
        ae_src = []
        for i in range(M):
            src = ae_ft[i] if isinstance(ae_ft, (list, tuple)) else (ae_ft if ae_ft is not None else ae_default)
            ae_src.append(src.contiguous().view(B * La, d))
        for st in side:                     # fork: everything above (mask packing, input copies) is visible
            st.wait_stream(main)

        # ---- Query-Aware Auto-Encoder branch (mtn.py:209-213), all layers, per modality
        qae_tapes, kv_ae, ae16s, ae_outs, vid16s = [], [], [], [], []
        if isinstance(ae_ft, (list, tuple)):
            ctx["ae_shared"] = None
        else:
            ctx["ae_shared"] = "given" if ae_ft is not None else ("src" if ae_features == "query" else "cap")
        ev_ae = [[torch.cuda.Event() for _ in range(M)] for _ in range(N)]
        for i in range(M):
            with _lib.on_stream(side[i]):
                Lv = vid_ft[i].shape[1]
                v16, kv_vid = hoisted(vid_c[i], W["kv_vid"][i])
                vid16s.append(v16)
                b_vid = b_vids[i]
                ae = ae_src[i]
                tape, kvs, a16s = [], [], []
                for l in range(N):
                    Lw = W["layers"][l]
                    c0 = 4 + 4 * i
                    ae = self._attn_fwd(tape, ae, Lw["ln"][c0], Lw["ae_self"][i], B, La, La, None, 0, 0, b_ae,
                                        ("ae_self", l, i, c0), dr, psub(l, c0))
                    ae = self._attn_fwd(tape, ae, Lw["ln"][c0 + 1], Lw["ae_vid"][i], B, La, Lv, kv_vid, l * 2 * d,
                                        l * 2 * d + d, b_vid, ("ae_vid", l, i, c0 + 1), dr, psub(l, c0 + 1))
                    ae, a16 = self._ffn_fwd(tape, ae, Lw["ln"][c0 + 2], Lw["ae_ffn"][i], ("ae_ffn", l, i, c0 + 2), want16=True,
                                            dr=dr, p_sub=psub(l, c0 + 2))
                    A2 = Lw["ae_attn"][i]
                    kv = torch.empty(B * La, 2 * d, dtype=f16, device=dev)
                    _lib.linear(a16, A2["w_qkv"][d:], A2["b_qkv"][d:], out_f16=kv)
                    kvs.append(kv)
                    a16s.append(a16)
                    ev_ae[l][i].record(side[i])      # layer l's K/V of ae_i is ready for the target path
                out = torch.empty(B * La, d, dtype=torch.float32, device=dev)
                nrm = W["ae_norm"][i]
                _lib.layernorm(ae, nrm[0], nrm[1], nrm[2], out_f32=out)
                tape.append(("norm", dict(x_in=ae, ln=nrm, names=("ae_norm", i))))
                qae_tapes.append(tape)
                kv_ae.append(kvs)
                ae16s.append(a16s)
                ae_outs.append(out.view(B, La, d))
        ctx["qae_tapes"], ctx["ae16"], ctx["vid16"] = qae_tapes, ae16s, vid16s

        # ---- text memories' hoisted K/V and the target path on the caller's stream
        his16, kv_his = hoisted(his, W["kv_his"])
        cap16, kv_cap = hoisted(cap, W["kv_cap"])
        q16, kv_q = hoisted(qm, W["kv_q"])
        ctx["mem16"] = {"his": his16, "cap": cap16, "src": q16}
        ctx["mem_shape"] = {"his": his.shape, "cap": cap.shape, "src": qm.shape}
        tm = tgt_mask
        if tm is not None and tm.dim() == 4:
            tm = tm[:, 0]
        if tm is not None and tm.shape[0] != B:
            tm = tm.expand(B, -1, -1)
        bits_t = _lib.mask_pack(tm) if tm is not None else None
        order = (("src", kv_q, b_q, qm.shape[1]), ("cap", kv_cap, b_cap, cap.shape[1])) \
            if ae_features in ("caption", "summary") else \
            (("cap", kv_cap, b_cap, cap.shape[1]), ("src", kv_q, b_q, qm.shape[1]))
        tape = []
        xs = x.contiguous().view(B * T, d)
        for l in range(N):
            Lw = W["layers"][l]
            kc, vc = l * 2 * d, l * 2 * d + d
            xs = self._attn_fwd(tape, xs, Lw["ln"][0], Lw["self"], B, T, T, None, 0, 0, bits_t, ("self", l, 0, 0), dr,
                                psub(l, 0))
            xs = self._attn_fwd(tape, xs, Lw["ln"][1], Lw["his"], B, T, his.shape[1], kv_his, kc, vc, b_his,
                                ("his", l, 0, 1), dr, psub(l, 1))
            for c, (name, kvm, bm, Lm) in enumerate(order):
                xs = self._attn_fwd(tape, xs, Lw["ln"][2 + c], Lw[name], B, T, Lm, kvm, kc, vc, bm, (name, l, 0, 2 + c), dr,
                                    psub(l, 2 + c))
            for i in range(M):
                main.wait_event(ev_ae[l][i])
                xs = self._attn_fwd(tape, xs, Lw["ln"][7 + 4 * i], Lw["ae_attn"][i], B, T, La, kv_ae[i][l], 0, d, b_ae,
                                    ("ae_attn", l, i, 7 + 4 * i), dr, psub(l, 7 + 4 * i))
            xs, _ = self._ffn_fwd(tape, xs, Lw["ln"][4 + 4 * M], Lw["ffn"], ("ffn", l, 0, 4 + 4 * M), dr=dr,
                                  p_sub=psub(l, 4 + 4 * M))
        out = torch.empty(B * T, d, dtype=torch.float32, device=dev)
        _lib.layernorm(xs, W["norm"][0], W["norm"][1], W["norm"][2], out_f32=out)
        tape.append(("norm", dict(x_in=xs, ln=W["norm"], names=("norm",))))
        ctx["tape"] = tape
        for st in side:                     # join
            main.wait_stream(st)
        return out.view(B, T, d), ae_outs, ctx

    # ------------------------------------------------------------------ gradient buffers
    def _grad_views(self, dev):
        """(G, per_param, direct): the gradient arena for one backward.  With ``attach_grads`` and every .grad
        still being its view, gradients accumulate directly (direct=True, the Function returns None for the
        parameters); otherwise a fresh zeroed arena whose views are handed to autograd."""
        g = self._grad
        if g is not None and g["buf"].device == dev and all(
                p.grad is not None and p.grad.data_ptr() == g["per"][id(p)].data_ptr() for p in self.param_list()):
            return g["G"], g["per"], True
        G, per = self._carve(torch.zeros(self.arena_numel(), dtype=torch.float32, device=dev))
        return G, per, False

    def _carve(self, buf):
        """Carve a flat buffer in the arena layout, so that every GEMM of the backward writes one contiguous block:
        [Wq;Wk;Wv] per self-attention-like module, all layers' [Wk_l;Wv_l] per hoisted memory.
        Returns (G, per_param) with per_param: id(parameter) -> view."""
        dec = self.dec
        layers = dec.layers
        N = len(layers)
        M = len(layers[0].auto_encoder_vid_attn)
        d = layers[0].size
        ar = _Arena(buf)
        G, per = {}, {}

        def lin_views(m, idx, w, b):
            per[id(m.linears[idx].weight)], per[id(m.linears[idx].bias)] = w, b

        def packed_qkv(key, mods):      # modules whose q, k, v projections are used together: [3d, d]
            for l, m in mods:
                w, b = ar.take(3 * d, d), ar.take(3 * d)
                G[(key, l, "wqkv")], G[(key, l, "bqkv")] = w, b
                for j in range(3):
                    lin_views(m, j, w[j * d:(j + 1) * d], b[j * d:(j + 1) * d])
                wo, bo = ar.take(d, d), ar.take(d)
                G[(key, l, "wo")], G[(key, l, "bo")] = wo, bo
                lin_views(m, 3, wo, bo)

        def hoisted(key, mods):         # [N*2d, d]: layer l -> rows [l*2d, l*2d+d) = Wk_l, next d = Wv_l
            w, b = ar.take(N * 2 * d, d), ar.take(N * 2 * d)
            G[(key, "wkv")], G[(key, "bkv")] = w, b
            for l, m in mods:
                lin_views(m, 1, w[l * 2 * d:l * 2 * d + d], b[l * 2 * d:l * 2 * d + d])
                lin_views(m, 2, w[l * 2 * d + d:(l + 1) * 2 * d], b[l * 2 * d + d:(l + 1) * 2 * d])
                wq, bq, wo, bo = ar.take(d, d), ar.take(d), ar.take(d, d), ar.take(d)
                G[(key, l, "wq")], G[(key, l, "bq")], G[(key, l, "wo")], G[(key, l, "bo")] = wq, bq, wo, bo
                lin_views(m, 0, wq, bq)
                lin_views(m, 3, wo, bo)

        def ffn(key, mods):
            for l, m in mods:
                dff = m.w_1.weight.shape[0]
                w1, b1, w2, b2 = ar.take(dff, d), ar.take(dff), ar.take(d, dff), ar.take(d)
                G[(key, l, "w1")], G[(key, l, "b1")], G[(key, l, "w2")], G[(key, l, "b2")] = w1, b1, w2, b2
                per[id(m.w_1.weight)], per[id(m.w_1.bias)] = w1, b1
                per[id(m.w_2.weight)], per[id(m.w_2.bias)] = w2, b2

        packed_qkv(("self", 0), [(l, L.self_attn) for l, L in enumerate(layers)])
        hoisted(("his", 0), [(l, L.his_attn) for l, L in enumerate(layers)])
        hoisted(("cap", 0), [(l, L.cap_attn) for l, L in enumerate(layers)])
        hoisted(("src", 0), [(l, L.src_attn) for l, L in enumerate(layers)])
        ffn(("ffn", 0), [(l, L.feed_forward) for l, L in enumerate(layers)])
        # Everything carved so far belongs to the TARGET path's own modules (self / his / cap / src sites, FFN): these
        # gradients are final once the target path and the text memories' K/V projections have been differentiated --
        # while the Query-Aware Auto-Encoder chains still run on their side streams.  A data-parallel trainer starts
        # the all-reduce of this prefix then (trainer.TrainStep, `on_target_grads_ready`).
        self._prefix_numel = ar.off
        for i in range(M):
            packed_qkv(("ae_self", i), [(l, L.auto_encoder_self_attn[i]) for l, L in enumerate(layers)])
            hoisted(("ae_vid", i), [(l, L.auto_encoder_vid_attn[i]) for l, L in enumerate(layers)])
            packed_qkv(("ae_attn", i), [(l, L.auto_encoder_attn[i]) for l, L in enumerate(layers)])
            ffn(("ae_ffn", i), [(l, L.auto_encoder_feed_forward[i]) for l, L in enumerate(layers)])
        for l, L in enumerate(layers):
            for c, s in enumerate(L.sublayer):
                a, b = ar.take(d), ar.take(d)
                G[("ln", l, c)] = (a, b)
                per[id(s.norm.a_2)], per[id(s.norm.b_2)] = a, b
        a, b = ar.take(d), ar.take(d)
        G[("norm",)] = (a, b)
        per[id(dec.norm.a_2)], per[id(dec.norm.b_2)] = a, b
        for i, m in enumerate(dec.ae_norm):
            a, b = ar.take(d), ar.take(d)
            G[("ae_norm", i)] = (a, b)
            per[id(m.a_2)], per[id(m.b_2)] = a, b
        return G, per

    # ------------------------------------------------------------------ backward building blocks
    # `wq` collects the weight-gradient GEMMs of a site as closures: nothing downstream reads them, so `_flush`
    # launches them on the dedicated wgrad stream once the site's operands are complete.
    @staticmethod
    def _ln_bwd(t, dy, dres, dx, gab, invS, dy_scale=None, nxt=None, mc=0):
        """nxt = (dx16, bias_grad) of the sublayer processed next: the LayerNorm kernel also emits the f16 copy of
        the updated residual gradient (that sublayer's GEMM operand) and its column sums (its output-bias gradient)."""
        ln = t["ln"]
        _lib.layernorm_bwd(t["x_in"], ln[0], ln[2], dy, dx, dres=dres, da_2=gab[0], db_2=gab[1], dy_scale=dy_scale,
                           param_alpha=invS, dx_f16=None if nxt is None else nxt[0],
                           dx_colsum=None if nxt is None else nxt[1], drop=None if nxt is None else nxt[2], mc=mc)

    def _flush(self, bk, wq):
        """Launch the collected weight-gradient GEMMs on the wgrad stream, after everything issued so far on the
        launch stream (their operands).  The closures keep their operand tensors alive until the final join."""
        if not wq:
            return
        ev = torch.cuda.Event()
        ev.record(_lib.launch_stream())
        bk["ws"].wait_event(ev)
        with _lib.on_stream(bk["ws"]):
            for fn in wq:
                fn()
        bk["keep"].extend(wq)
        del wq[:]

    def _attn_bwd(self, bk, t, dx, dkv, dx16=None, nxt=None):
        """dx: [rows, d] f32 scaled residual-stream gradient at the site's OUTPUT; updated in place to the
        gradient at its input.  dkv: (dk view, dv view) f16 destination for cross sites (hoisted columns).
        dx16: f16 copy of dx if the previous LayerNorm backward already produced it (and the b_o gradient)."""
        G, invS, wq, mc = bk["G"], bk["invS"], [], bk["mc"]
        A = t["A"]
        B, Lq, Lk, h, dk_ = t["B"], t["Lq"], t["Lk"], A["h"], A["d_k"]
        rows, d = dx.shape
        dev = dx.device
        f16 = torch.float16
        key, l, i, c = t["names"]
        gk = (key, i)
        # ---- output projection (mtn.py:267) + residual (mtn.py:127)
        if dx16 is None:
            dx16 = torch.empty(rows, d, dtype=f16, device=dev)
            _lib.cast_colsum(dx, dst_f16=dx16, colsum=G[(gk, l, "bo")], alpha=invS, drop=t["drop_o"], mc=mc)
        do16 = torch.empty(rows, d, dtype=f16, device=dev)
        _lib.linear_dgrad(dx16, A["w_o"], out_f16=do16)
        wq.append(lambda: _lib.linear_wgrad(dx16, t["o16"], G[(gk, l, "wo")], alpha=invS, mc=mc))
        # ---- attention core
        delta = torch.empty(B, h, Lq, dtype=torch.float32, device=dev)
        _lib.attn_delta(do16, t["o16"], B, Lq, h, dk_, delta)
        one_tile = Lk <= 128             # dQ complete inside one CTA: stored as f16 directly, no f32 accumulation buffer
        if t["self_attn"]:
            dqkv = torch.empty(rows, 3 * d, dtype=f16, device=dev)
            dq16, dkd, dvd = dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:]
        else:
            dq16 = torch.empty(rows, d, dtype=f16, device=dev)
            dkd, dvd = dkv
        dq32 = None if one_tile else _lib.zero(torch.empty(rows, d, dtype=torch.float32, device=dev))
        _lib.attn_core_bwd(t["q"], t["k"], t["v"], do16, t["stats"], delta, B, h, Lq, Lk, dk_, dq16 if one_tile else dq32,
                           dkd, dvd, mask_bits=t["bits"], drop=t["drop_p"])
        # ---- Q (or QKV) projection
        dxn = torch.empty(rows, d, dtype=torch.float32, device=dev)
        # (bias gradients = column sums of the f16 gradient: computed inside the weight-gradient GEMM by one extra MMA
        #  against a tile of ones, off the critical path)
        if t["self_attn"]:
            if not one_tile:
                _lib.cast_colsum(dq32, dst_f16=dq16)
            _lib.linear_dgrad(dqkv, A["w_qkv"], out_f32=dxn)
            wq.append(lambda: _lib.linear_wgrad(dqkv, t["xn16"], G[(gk, l, "wqkv")], alpha=invS, dbias=G[(gk, l, "bqkv")], mc=mc))
        else:
            wq_key, bq_key = ("wq", "bq") if (gk, l, "wq") in G else ("wqkv", "bqkv")
            gw, gb = G[(gk, l, wq_key)], G[(gk, l, bq_key)]
            if not one_tile:
                _lib.cast_colsum(dq32, dst_f16=dq16)
            _lib.linear_dgrad(dq16, A["w_qkv"][:d], out_f32=dxn)
            wq.append(lambda: _lib.linear_wgrad(dq16, t["xn16"], gw[:d], alpha=invS, dbias=gb[:d], mc=mc))
        self._flush(bk, wq)
        self._ln_bwd(t, dxn, dx, dx, G[("ln", l, c)], invS, nxt=nxt, mc=mc)
        bk["keep"].append((dq32, delta, do16, dxn))

    def _ffn_bwd(self, bk, t, dx, dx16=None, nxt=None):
        G, invS, wq, mc = bk["G"], bk["invS"], [], bk["mc"]
        Fw = t["Fw"]
        rows, d = dx.shape
        dev = dx.device
        f16 = torch.float16
        key, l, i, c = t["names"]
        gk = (key, i)
        if dx16 is None:
            dx16 = torch.empty(rows, d, dtype=f16, device=dev)
            _lib.cast_colsum(dx, dst_f16=dx16, colsum=G[(gk, l, "b2")], alpha=invS, drop=t["drop_o"], mc=mc)
        dhid = torch.empty(rows, Fw["w_1"].shape[0], dtype=f16, device=dev)
        # through the ReLU and the hidden dropout (mtn.py:280): the saved hidden activation is > 0 exactly where the
        # unit was active AND kept; kept gradients are scaled by 1/(1-p)
        ms = 1.0 / (1.0 - t["drop_h"][2] / 65536.0) if t["drop_h"] is not None else 0.0
        _lib.linear_dgrad(dx16, Fw["w_2"], relu_mask=t["hid"], out_f16=dhid, mask_scale=ms)
        wq.append(lambda: _lib.linear_wgrad(dx16, t["hid"], G[(gk, l, "w2")], alpha=invS, mc=mc))
        wq.append(lambda: _lib.linear_wgrad(dhid, t["xn16"], G[(gk, l, "w1")], alpha=invS, dbias=G[(gk, l, "b1")], mc=mc))
        self._flush(bk, wq)
        dxn = torch.empty(rows, d, dtype=torch.float32, device=dev)
        _lib.linear_dgrad(dhid, Fw["w_1"], out_f32=dxn)
        self._ln_bwd(t, dxn, dx, dx, G[("ln", l, c)], invS, nxt=nxt, mc=mc)
        bk["keep"].append(dxn)

    def _mem_bwd(self, bk, dkv, mem16, w_kv, gw, gb):
        """Backward of a hoisted memory K/V projection: bias gradient, weight gradient, memory gradient."""
        invS = bk["invS"]
        wq = [lambda: _lib.linear_wgrad(dkv, mem16, gw, alpha=invS, dbias=gb, mc=bk["mc"])]
        self._flush(bk, wq)
        dmem = torch.empty(mem16.shape[0], mem16.shape[1], dtype=torch.float32, device=dkv.device)
        _lib.linear_dgrad(dkv, w_kv, alpha=invS, out_f32=dmem)
        return dmem

    @staticmethod
    def _unscale(dx, invS):
        """An input gradient leaves the Function: multiply by 1/S."""
        out = torch.empty_like(dx)
        _lib.scale_f32(dx, invS, out)
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, ctx, g_out, g_ae):
        """g_out: [B,T,d] f32 or None; g_ae: list of [B,La,d] f32 or None.  Returns (input_grads, per_param) with
        input_grads = dict(x, his, cap, src, vid=[...], ae=[...])."""
        W = ctx["W"]
        d, N, M = W["d"], W["N"], W["M"]
        B, T, La = ctx["B"], ctx["T"], ctx["La"]
        tape = ctx["tape"]
        dev = tape[0][1]["x_in"].device
        f16 = torch.float16
        main = torch.cuda.current_stream()
        side, ws = self._streams(M, dev)
        G, per, direct = self._grad_views(dev)
        if g_out is None:
            g_out = torch.zeros(B, T, d, dtype=torch.float32, device=dev)
        g_ae = [g if g is not None else torch.zeros(B, La, d, dtype=torch.float32, device=dev) for g in g_ae]
        g_out = g_out.contiguous().float()
        g_ae = [g.contiguous().float() for g in g_ae]
        S2 = _lib.grad_scale([g_out] + g_ae)
        S, invS = S2[0:1], S2[1:2]
        bk = {"G": G, "invS": invS, "ws": ws, "keep": [], "mc": self._mc if direct else 0}

        # hoisted dK/dV destinations, one per memory: every layer fills its own columns
        rows_mem = {k: v.shape[0] for k, v in ctx["mem16"].items()}
        dkv_mem = {k: torch.empty(r, N * 2 * d, dtype=f16, device=dev) for k, r in rows_mem.items()}
        dkv_vid = [torch.empty(v.shape[0], N * 2 * d, dtype=f16, device=dev) for v in ctx["vid16"]]
        dkv_ae = [[torch.empty(B * La, 2 * d, dtype=f16, device=dev) for _ in range(N)] for _ in range(M)]
        ev_kv = [[torch.cuda.Event() for _ in range(M)] for _ in range(N)]
        for st in side + [ws]:              # fork: gradient arena zeroed, scale computed
            st.wait_stream(main)

        # ---- target path (critical path, caller's stream), last sublayer first
        def bias_of(entry):                 # output-bias gradient of a taped sublayer
            knd, tt = entry
            key, l, i, c = tt["names"]
            return G[((key, i), l, "bo" if knd == "attn" else "b2")]

        def handoff(seq, k, rows, fused=lambda entry: True):
            """(dx16, bias gradient) the LayerNorm backward of seq[k] should produce for seq[k+1], or None."""
            if k + 1 >= len(seq) or not fused(seq[k + 1]) or d not in (128, 256, 512, 1024):
                return None        # (the vectorised LayerNorm backward that can emit the hand-off exists for these d)
            return (torch.empty(rows, d, dtype=f16, device=dev), bias_of(seq[k + 1]), seq[k + 1][1]["drop_o"])

        kind, t = tape[-1]
        seq = list(reversed(tape[:-1]))
        dx = torch.empty(B * T, d, dtype=torch.float32, device=dev)
        nxt = handoff(seq, -1, B * T)
        self._ln_bwd(t, g_out.view(B * T, d), None, dx, G[("norm",)], invS, dy_scale=S, nxt=nxt, mc=bk["mc"])   # mtn.py:164
        for k, (kind, t) in enumerate(seq):
            dx16, nxt = (nxt[0] if nxt is not None else None), handoff(seq, k, B * T)
            if kind == "ffn":
                self._ffn_bwd(bk, t, dx, dx16, nxt)
                continue
            key, l, i, c = t["names"]
            if key == "self":
                dkv = None
            elif key == "ae_attn":
                buf = dkv_ae[i][l]
                dkv = (buf[:, :d], buf[:, d:])
            else:
                buf = dkv_mem[key]
                dkv = (buf[:, l * 2 * d:l * 2 * d + d], buf[:, l * 2 * d + d:(l + 1) * 2 * d])
            self._attn_bwd(bk, t, dx, dkv, dx16, nxt)
            if key == "ae_attn":
                ev_kv[l][i].record(main)    # dK/dV of layer l's ae_i memory are complete: its QAE chain may proceed
        grads = {"x": self._unscale(dx, invS).view(B, T, d)}
        for name in ("his", "cap", "src"):
            gk = (name, 0)
            grads[name] = self._mem_bwd(bk, dkv_mem[name], ctx["mem16"][name],
                                        W["kv_" + ("q" if name == "src" else name)][0], G[(gk, "wkv")],
                                        G[(gk, "bkv")]).view(ctx["mem_shape"][name])

        hook = getattr(self, "on_target_grads_ready", None)
        if hook is not None and direct:
            ev_w = torch.cuda.Event()
            ev_w.record(ws)                 # the weight-gradient GEMMs issued so far (all of the prefix's)
            hook(ev_w)

        # ---- Query-Aware Auto-Encoder branch: one side stream per modality
        grads["ae"], grads["vid"] = [], []
        for i in range(M):
            with _lib.on_stream(side[i]):
                qt = ctx["qae_tapes"][i]
                kind, t = qt[-1]
                seq = list(reversed(qt[:-1]))
                not_ffn = lambda entry: entry[0] != "ffn"    # an FFN's input gradient gets the K/V term added first
                dae = torch.empty(B * La, d, dtype=torch.float32, device=dev)
                nxt = None
                self._ln_bwd(t, g_ae[i].view(B * La, d), None, dae, G[("ae_norm", i)], invS, dy_scale=S, mc=bk["mc"])   # mtn.py:162-163
                for k, (kind, t) in enumerate(seq):
                    dx16, nxt = (nxt[0] if nxt is not None else None), handoff(seq, k, B * La, not_ffn)
                    key, l, _, c = t["names"]
                    if kind == "ffn":
                        # the layer's output ae_i^l is also the memory of the target's auto_encoder_attn[i]
                        # (mtn.py:215): add the gradient that came back through its K/V projection
                        side[i].wait_event(ev_kv[l][i])
                        A2 = W["layers"][l]["ae_attn"][i]
                        gk = ("ae_attn", i)
                        buf, a16 = dkv_ae[i][l], ctx["ae16"][i][l]
                        gw, gb = G[(gk, l, "wqkv")][d:], G[(gk, l, "bqkv")][d:]
                        _lib.linear_dgrad(buf, A2["w_qkv"][d:], addend=dae, out_f32=dae)
                        self._flush(bk, [lambda buf=buf, a16=a16, gw=gw, gb=gb: _lib.linear_wgrad(buf, a16, gw, alpha=invS, dbias=gb,
                                                                                             mc=bk["mc"])])
                        self._ffn_bwd(bk, t, dae, None, nxt)
                    elif key == "ae_vid":
                        buf = dkv_vid[i]
                        self._attn_bwd(bk, t, dae, (buf[:, l * 2 * d:l * 2 * d + d], buf[:, l * 2 * d + d:(l + 1) * 2 * d]),
                                       dx16, nxt)
                    else:
                        self._attn_bwd(bk, t, dae, None, dx16, nxt)
                grads["ae"].append(self._unscale(dae, invS).view(B, La, d))
                bk["keep"].append(dae)
                gk = ("ae_vid", i)              # hoisted video K/V projection of modality i
                grads["vid"].append(self._mem_bwd(bk, dkv_vid[i], ctx["vid16"][i], W["kv_vid"][i][0], G[(gk, "wkv")],
                                                  G[(gk, "bkv")]).view(B, -1, d))
        for st in side + [ws]:              # join
            main.wait_stream(st)
        bk["keep"].append((dkv_mem, dkv_vid, dkv_ae, dx, g_out, g_ae, S2))
        ctx["_bwd_keep"] = bk["keep"]       # released with the tape by the caller, after the join has been enqueued
        return grads, (None if direct else per)
