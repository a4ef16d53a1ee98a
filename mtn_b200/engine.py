"""Model-level execution of the MTN decoder cascade on the sm_100a kernels.

``DecoderEngine.forward`` is what ``mtn.Decoder.forward`` (reference mtn.py:158-164)
runs.  It exploits two structural facts of the reference's DecoderLayer
(mtn.py:181-218) that the reference itself does not:

1. **Static memories.**  his / cap / query / video memories are the same tensors in
   every layer, so their K and V projections for ALL N layers are one wide GEMM per
   memory ([B*L, d] x [d, N*2d]) instead of 2N small ones.
2. **The Query-Aware Auto-Encoder branch never reads the target stream** (SURVEY 8a):
   ae-self -> ae->video -> ae-FFN of every layer, and the K/V projections of its
   outputs, depend only on the memories.  They are computed once per memory set
   ("memory stage") and cached, so repeated ``model.decode`` calls of beam / greedy
   search (data_utils.py:197-206) only pay for the target path.

The target path per layer is then: LN -> [Q|K|V] GEMM -> core -> out-proj(+residual) for
self-attention, and LN -> Q GEMM -> core (hoisted K/V) -> out-proj(+residual) for each
memory, then the fused FFN; the residual stream stays f32, every tensor-core operand
is f16 (see DESIGN.md "Precision").
"""
import os

import torch

from . import _lib


def ensure_inference(module, x):
    """The CUDA path implements the forward (eval) arithmetic.  Fail loudly instead of
    silently producing tensors that autograd cannot differentiate or that lack dropout."""
    if not x.is_cuda:
        raise _lib.MtnError("mtn_b200 hot path needs CUDA tensors (no CPU implementation)")
    if torch.is_grad_enabled() and (x.requires_grad or
                                    (module is not None and module.training and
                                     any(p.requires_grad for p in module.parameters()))):
        raise NotImplementedError(
            "mtn_b200: this entry point is the inference form (no autograd graph, no dropout); training goes "
            "through train_engine.DecoderTrainer / trainer.TrainStep -- or call model.eval() under torch.no_grad()")


def _ln_fused_max_rows():
    """Largest row count that takes the fused LayerNorm + projection kernel (csrc/ln_gemm.cu).  Default 0 = never:
    measured 3.5-5 us slower per sublayer than the two launches at every shape of the decoder
    (profiles/r01d_ln_linear.txt).  MTN_B200_LN_FUSED=<rows> opts in (results are the same either way)."""
    e = os.environ.get("MTN_B200_LN_FUSED")
    return LN_FUSED_MAX_ROWS if e is None else int(e)


LN_FUSED_MAX_ROWS = 0


# Debug taps (tools/bisect_decode.py): when TAP is a list, the target path appends (label, clone) of every
# intermediate -- also inside CUDA-graph capture, where the clones live in the graph's pool and are refreshed by
# every replay.  None in production: no extra work.
TAP = None
TAP_PREFIX = [""]


def _tap(label, t):
    if TAP is not None:
        TAP.append((TAP_PREFIX[0] + label, t.clone()))


def _ln_linear(x, ln, w, b, act, xn16, out16):
    """out16 = act(LN(x) w^T + b): one fused launch when the shape qualifies, else LayerNorm -> xn16 -> linear
    (same results either way, tests/test_gpu_ln_linear.py)."""
    rows, d = x.shape
    if _lib.ROWS_KERNELS and _lib.rows_ln_linear_ok(rows, w.shape[0], d):   # KV-cached decoding: few rows, one mma.sync launch
        _lib.rows_ln_linear(x, ln[0], ln[1], ln[2], w, bias=b, act=act, out_f16=out16)
        return
    if rows <= _ln_fused_max_rows() and _lib.ln_linear_supported(d):
        _lib.ln_linear(x, ln[0], ln[1], ln[2], w, bias=b, act=act, out_f16=out16)
        return
    _lib.layernorm(x, ln[0], ln[1], ln[2], out_f16=xn16)
    _tap("xn16", xn16)
    _lib.linear(xn16, w, b, act=act, out_f16=out16)


_GENERATION = [0]


def invalidate_weight_caches():
    """Public hook: call after modifying parameters in a way autograd's version counters do not see --
    writes through ``p.data`` (``p.data.mul_(2)`` leaves ``p._version`` unchanged), raw-pointer kernel updates
    (the fused arena Adam), ``p.data = new_tensor`` with a recycled address.  Every f16 weight pack and every cached
    memory stage (their keys carry this generation) is rebuilt on its next use."""
    _GENERATION[0] += 1


def _site_fused_ok(d, h, B, Lq, Lk):
    """Does this hoisted-K/V site take the ONE-kernel path (csrc/site_fused.cu)?  MTN_B200_SITE_FUSED = 1: whenever the
    shape is supported; 0: never; default "auto": where it measured faster than the launch sequence Q GEMM -> core ->
    out-proj GEMM (profiles/r02c_site_bench.txt: 1.23-1.33x at Lq = 256, 1.06-1.12x at Lq = 64, 1.09-1.16x for few query
    rows over a short memory, 0.97x for few rows over a long one)."""
    mode = os.environ.get("MTN_B200_SITE_FUSED", "auto")
    if mode == "0" or not _lib.attn_site_fused_supported(d, h):
        return False
    if mode == "1":
        return True
    if _lib.ROWS_KERNELS and Lq <= 8 and B * Lq <= 128:
        return False                 # KV-cached decoding: the few-row kernels (csrc/decode_rows.cu) are faster still
    return Lq >= 64 or Lk <= 64


# parallel.SmPartition or None: see DecoderEngine._streams
SM_PARTITION = None


# Pre-allocated _lib.StepProgram objects for decode_step calls made during CUDA-graph capture (graph.GraphedGreedyDecoder
# fills it before capturing: a program's pinned / device buffers cannot be allocated inside a capture)
PROGRAM_POOL = []


def _ffn_fused_ok(rows, d, d_ff):
    """Does this feed-forward sublayer take the ONE-kernel path (csrc/ffn_fused.cu)?  MTN_B200_FFN_FUSED = 1: whenever the
    shape is supported; 0 (default): never -- see DESIGN.md section 4 for the measurement behind the default."""
    mode = os.environ.get("MTN_B200_FFN_FUSED", FFN_FUSED_DEFAULT)
    if mode == "0" or not _lib.ffn_fused_supported(rows, d, d_ff):
        return False
    if mode == "1":
        return True
    return rows >= 4096          # "auto": full machine


FFN_FUSED_DEFAULT = "0"
DECODE_CLUSTER_DEFAULT = "1"


class PackedWeights(object):
    """Cache of tensor-core-ready (f16, concatenated) copies of nn.Parameters, rebuilt
    when any source parameter is modified in place (``_version``), re-assigned, or after
    ``invalidate_weight_caches()`` (writes the version counters cannot see)."""

    def __init__(self, follow_generation=True):
        # follow_generation=False: the owner refreshes the pack itself when the raw-pointer writer runs (the training
        # arena, whose f16 copy the fused Adam kernel rewrites in the same pass)
        self._key, self._val, self._gen = None, None, follow_generation

    def get(self, params, build):
        key = (_GENERATION[0] if self._gen else 0,) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            self._val, self._key = build(), key
        return self._val

    def __deepcopy__(self, memo):
        return PackedWeights()

    def __getstate__(self):
        return {}

    def __setstate__(self, st):
        self._key, self._val, self._gen = None, None, True


class _MemoryKey(object):
    """Identity of a memory set.  Holds STRONG references to the tensors: while they are
    alive the caching allocator cannot hand their storage to a different tensor, so
    ``is`` + ``_version`` equality really means "same contents" (data_ptr alone would
    alias across training-loop iterations)."""

    def __init__(self, wkey, ae_features, tensors):
        self.wkey, self.ae_features = wkey, ae_features
        self.tensors = list(tensors)
        self.versions = [None if t is None else t._version for t in self.tensors]

    def matches(self, other):
        return (other is not None and self.wkey == other.wkey and self.ae_features == other.ae_features
                and len(self.tensors) == len(other.tensors)
                and all(a is b for a, b in zip(self.tensors, other.tensors))
                and self.versions == other.versions)


class DecoderEngine(object):
    # run the modalities' QAE chains as strided-batch launches (MTN_B200_QAE_BATCHED=0: one chain per stream)
    qae_batched = __import__("os").environ.get("MTN_B200_QAE_BATCHED", "1") != "0"

    def __init__(self, decoder):
        self.dec = decoder
        self._packed = PackedWeights()
        self._mem_key, self._mem = None, None

    # ------------------------------------------------------------------ weights
    def weights(self):
        dec = self.dec
        params = list(dec.parameters())

        def build():
            f16 = lambda w: _lib.cast_f16(w.contiguous())
            layers = dec.layers
            M = len(layers[0].auto_encoder_vid_attn)
            d = layers[0].size

            def att(m):      # full per-site pack
                w = [l.weight.data for l in m.linears]
                b = [l.bias.data for l in m.linears]
                return {"w_qkv": f16(torch.cat(w[:3], 0)), "b_qkv": torch.cat(b[:3], 0).contiguous(),
                        "w_o": f16(w[3]), "b_o": b[3].contiguous(), "h": m.h, "d_k": m.d_k}

            def hoist(mods):  # [N*2d, d]: layer l -> rows [l*2d, l*2d+d) = Wk_l, next d = Wv_l
                w = torch.cat([torch.cat([m.linears[1].weight.data, m.linears[2].weight.data], 0)
                               for m in mods], 0)
                b = torch.cat([torch.cat([m.linears[1].bias.data, m.linears[2].bias.data], 0)
                               for m in mods], 0)
                return f16(w), b.contiguous()

            def ffn(m):
                return {"w_1": f16(m.w_1.weight.data), "b_1": m.w_1.bias.data.contiguous(),
                        "w_2": f16(m.w_2.weight.data), "b_2": m.w_2.bias.data.contiguous()}

            def ln(m):
                return (m.a_2.data, m.b_2.data, m.eps)

            W = {"M": M, "d": d, "N": len(layers), "layers": []}
            W["kv_his"] = hoist([l.his_attn for l in layers])
            W["kv_cap"] = hoist([l.cap_attn for l in layers])
            W["kv_q"] = hoist([l.src_attn for l in layers])
            W["kv_vid"] = [hoist([l.auto_encoder_vid_attn[i] for l in layers]) for i in range(M)]
            for l in layers:
                W["layers"].append({
                    "self": att(l.self_attn), "his": att(l.his_attn), "cap": att(l.cap_attn),
                    "src": att(l.src_attn),
                    "ae_self": [att(m) for m in l.auto_encoder_self_attn],
                    "ae_vid": [att(m) for m in l.auto_encoder_vid_attn],
                    "ae_attn": [att(m) for m in l.auto_encoder_attn],
                    "ae_ffn": [ffn(m) for m in l.auto_encoder_feed_forward],
                    "ffn": ffn(l.feed_forward),
                    "ln": [ln(s.norm) for s in l.sublayer],
                })
            W["norm"] = ln(dec.norm)
            W["ae_norm"] = [ln(m) for m in dec.ae_norm]
            if M > 1:
                # modality-stacked copies for the batched Query-Aware Auto-Encoder chain: problem i of a
                # strided-batch GEMM / parameter set i of a grouped LayerNorm is modality i
                st = lambda ts: torch.stack(list(ts), 0).contiguous()
                lnst = lambda ls: (st(t[0] for t in ls), st(t[1] for t in ls), ls[0][2])
                for l, Lw in enumerate(W["layers"]):
                    Lw["qae"] = {
                        "ln_self": lnst([Lw["ln"][4 + 4 * i] for i in range(M)]),
                        "ln_vid": lnst([Lw["ln"][5 + 4 * i] for i in range(M)]),
                        "ln_ffn": lnst([Lw["ln"][6 + 4 * i] for i in range(M)]),
                        "self_w_qkv": st(a["w_qkv"] for a in Lw["ae_self"]), "self_b_qkv": st(a["b_qkv"] for a in Lw["ae_self"]),
                        "self_w_o": st(a["w_o"] for a in Lw["ae_self"]), "self_b_o": st(a["b_o"] for a in Lw["ae_self"]),
                        "vid_w_q": st(a["w_qkv"][:d] for a in Lw["ae_vid"]), "vid_b_q": st(a["b_qkv"][:d] for a in Lw["ae_vid"]),
                        "vid_w_o": st(a["w_o"] for a in Lw["ae_vid"]), "vid_b_o": st(a["b_o"] for a in Lw["ae_vid"]),
                        "ffn_w_1": st(f["w_1"] for f in Lw["ae_ffn"]), "ffn_b_1": st(f["b_1"] for f in Lw["ae_ffn"]),
                        "ffn_w_2": st(f["w_2"] for f in Lw["ae_ffn"]), "ffn_b_2": st(f["b_2"] for f in Lw["ae_ffn"]),
                        "kv_w": st(a["w_qkv"][d:] for a in Lw["ae_attn"]), "kv_b": st(a["b_qkv"][d:] for a in Lw["ae_attn"]),
                    }
                W["ae_norm_stacked"] = lnst(W["ae_norm"])
            return W

        return self._packed.get(params, build)

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _attn_block(x, ln, A, B, Lq, Lk, q_w, q_b, kv, k_col, v_col, bits, xn16, qbuf, obuf):
        """One pre-norm residual attention site, in place on the f32 stream ``x`` [B*Lq, d].
        kv=None -> self-attention (q_w is the [3d, d] pack, K/V come out of the same GEMM)."""
        d = x.shape[1]
        if kv is not None and _site_fused_ok(d, A["h"], B, Lq, Lk):
            _lib.layernorm(x, ln[0], ln[1], ln[2], out_f16=xn16)
            _tap("xn16", xn16)
            _lib.attn_site_fused(xn16, x, q_w, q_b, A["w_o"], A["b_o"], kv, k_col, v_col, B, A["h"], Lq, Lk, mask_bits=bits)
            _tap("x", x)
            return
        _ln_linear(x, ln, q_w, q_b, _lib.ACT_NONE, xn16, qbuf)
        if kv is None:
            q, k, v = qbuf[:, :d], qbuf[:, d:2 * d], qbuf[:, 2 * d:]
        else:
            q, k, v = qbuf, kv[:, k_col:k_col + d], kv[:, v_col:v_col + d]
        _tap("q", qbuf)
        _lib.attn_core(q, k, v, B, A["h"], Lq, Lk, A["d_k"], obuf, mask_bits=bits)
        _tap("o", obuf)
        _lib.linear(obuf, A["w_o"], A["b_o"], addend=x, out_f32=x)
        _tap("x", x)

    @staticmethod
    def _ffn_block(x, ln, Fw, xn16, hid, out16=None):
        if out16 is None and _ffn_fused_ok(x.shape[0], x.shape[1], Fw["w_1"].shape[0]):
            # LayerNorm -> ONE kernel: both projections, the [rows, d_ff] hidden activation stays on the SM (csrc/ffn_fused.cu)
            _lib.layernorm(x, ln[0], ln[1], ln[2], out_f16=xn16)
            _lib.ffn_fused(xn16, x, Fw["w_1"], Fw["b_1"], Fw["w_2"], Fw["b_2"])
            _tap("x", x)
            return
        _ln_linear(x, ln, Fw["w_1"], Fw["b_1"], _lib.ACT_RELU, xn16, hid)
        _tap("hid", hid)
        _lib.linear(hid, Fw["w_2"], Fw["b_2"], addend=x, out_f32=x, out_f16=out16)
        _tap("x", x)

    # ------------------------------------------------------------------ memory stage
    def _streams(self, M, dev):
        if SM_PARTITION is not None:
            # SM partitioning (parallel.SmPartition): the side chain (side[0]: the batched Query-Aware Auto-Encoder branch)
            # runs on the small SM group, concurrently with the target path on the big group (the caller runs the forward
            # on SM_PARTITION.main); the other side streams (hoisted video K/V: wide GEMMs) stay on the big group
            return [SM_PARTITION.side] + [SM_PARTITION.extra(i) for i in range(M - 1)]
        if getattr(self, "_side", None) is None or len(self._side) != M or self._side_dev != dev:
            self._side = [torch.cuda.Stream(device=dev) for _ in range(M)]
            self._side_dev = dev
        return self._side

    def _qae_batched(self, W, S, mods, side, main, B, La, rows, ae_bits):
        """Both (all) modalities' Query-Aware Auto-Encoder chains in lockstep on one side stream: every
        projection is ONE strided-batch GEMM over the modalities (same shapes, different weights), every
        LayerNorm one grouped launch, the ae-self attention one launch over M*B "dialogues"; only the
        ae->video attention stays per modality (different video lengths).  13 launches per layer instead of 26,
        and each fills the machine twice as well."""
        d, N, M = W["d"], W["N"], W["M"]
        dev = mods[0]["ae"].device
        f16 = torch.float16
        dff = W["layers"][0]["ffn"]["w_1"].shape[0]
        A0 = W["layers"][0]["ae_self"][0]
        h, dk = A0["h"], A0["d_k"]
        R = M * rows
        ae = torch.stack([m["ae"] for m in mods], 0)                         # [M, rows, d] f32 residual streams
        xn16 = torch.empty(M, rows, d, dtype=f16, device=dev)
        qkv = torch.empty(M, rows, 3 * d, dtype=f16, device=dev)
        obuf = torch.empty(M, rows, d, dtype=f16, device=dev)
        hid = torch.empty(M, rows, dff, dtype=f16, device=dev)
        ae16 = torch.empty(M, rows, d, dtype=f16, device=dev)
        kv_ae = [torch.empty(M, rows, 2 * d, dtype=f16, device=dev) for _ in range(N)]
        out = torch.empty(M, rows, d, dtype=torch.float32, device=dev)
        bits_rep = ae_bits.repeat(M, 1, 1).contiguous() if ae_bits is not None else None
        S["kv_ae"] = [[kv_ae[l][i] for i in range(M)] for l in range(N)]
        evs = [torch.cuda.Event() for _ in range(N)]
        S["ev"] = [[evs[l]] * M for l in range(N)]
        S["ae_out"] = [out[i].view(B, La, d) for i in range(M)]
        S["side"] = side
        S["_keep_b"] = (ae, xn16, qkv, obuf, hid, ae16, kv_ae, out, bits_rep)
        for i in range(M):                                                   # hoisted video K/V, one stream each
            side[i].wait_stream(main)
            with torch.cuda.stream(side[i]):
                m = mods[i]
                _lib.cast_f16(m["vid"], m["vid16"])
                _lib.linear(m["vid16"], W["kv_vid"][i][0], W["kv_vid"][i][1], out_f16=m["kv_vid"])
        for i in range(1, M):
            side[0].wait_stream(side[i])
        with torch.cuda.stream(side[0]):
            flat = lambda t: t.view(R, t.shape[-1])
            for l in range(N):
                Q = W["layers"][l]["qae"]
                kc, vc = l * 2 * d, l * 2 * d + d
                ln = Q["ln_self"]
                _lib.layernorm(flat(ae), ln[0], ln[1], ln[2], out_f16=flat(xn16), rows_per_group=rows)
                _lib.linear_batched(xn16, Q["self_w_qkv"], Q["self_b_qkv"], out_f16=qkv)
                q2 = flat(qkv)
                _lib.attn_core(q2[:, :d], q2[:, d:2 * d], q2[:, 2 * d:], M * B, h, La, La, dk, flat(obuf), mask_bits=bits_rep)
                _lib.linear_batched(obuf, Q["self_w_o"], Q["self_b_o"], addend=ae, out_f32=ae)
                ln = Q["ln_vid"]
                _lib.layernorm(flat(ae), ln[0], ln[1], ln[2], out_f16=flat(xn16), rows_per_group=rows)
                _lib.linear_batched(xn16, Q["vid_w_q"], Q["vid_b_q"], out_f16=qkv[:, :, :d])
                for i in range(M):
                    m = mods[i]
                    _lib.attn_core(qkv[i][:, :d], m["kv_vid"][:, kc:kc + d], m["kv_vid"][:, vc:vc + d], B, h, La, m["Lv"],
                                   dk, obuf[i], mask_bits=m["bits_vid"])
                _lib.linear_batched(obuf, Q["vid_w_o"], Q["vid_b_o"], addend=ae, out_f32=ae)
                ln = Q["ln_ffn"]
                _lib.layernorm(flat(ae), ln[0], ln[1], ln[2], out_f16=flat(xn16), rows_per_group=rows)
                _lib.linear_batched(xn16, Q["ffn_w_1"], Q["ffn_b_1"], act=_lib.ACT_RELU, out_f16=hid)
                _lib.linear_batched(hid, Q["ffn_w_2"], Q["ffn_b_2"], addend=ae, out_f32=ae, out_f16=ae16)
                # K/V of this layer's ae_i for the target stream's auto_encoder_attn[i] (mtn.py:215)
                _lib.linear_batched(ae16, Q["kv_w"], Q["kv_b"], out_f16=kv_ae[l])
                evs[l].record(side[0])
            ln = W["ae_norm_stacked"]
            _lib.layernorm(flat(ae), ln[0], ln[1], ln[2], out_f32=flat(out), rows_per_group=rows)   # mtn.py:162-163

    def _memory_stage(self, W, vid_ft, vid_mask, his, his_mask, cap, cap_mask, qm, q_mask, ae_ft, ae_features):
        """Everything that does not depend on the target stream.  The text memories' hoisted K/V run
        on the caller's stream; each modality's Query-Aware Auto-Encoder chain (all N layers) runs on
        its own side stream, forked here and joined by ``forward`` -- the two chains and the target
        path are independent latency-bound sequences of small kernels, so they overlap.  Every
        buffer is allocated on the caller's stream; side streams only launch kernels.
        ``S["ev"][l][i]`` fires when layer l's K/V of ae_i is ready for the target path."""
        dev = qm.device
        d, N, M = W["d"], W["N"], W["M"]
        B = qm.shape[0]
        f16 = torch.float16
        main = torch.cuda.current_stream()
        side = self._streams(M, dev)
        dff = W["layers"][0]["ffn"]["w_1"].shape[0]

        def hoisted(mem, wb):
            m16 = _lib.cast_f16(mem.contiguous().view(-1, d))
            out = torch.empty(m16.shape[0], N * 2 * d, dtype=f16, device=dev)
            _lib.linear(m16, wb[0], wb[1], out_f16=out)
            return out

        def bits(mask):
            if mask is None:
                return None
            if mask.shape[0] != B:
                mask = mask.expand(B, -1, -1)
            return _lib.mask_pack(mask)

        S = {"H": his.shape[1], "C": cap.shape[1], "Q": qm.shape[1]}
        if ae_features in ("caption", "summary"):
            ae_default = cap
        elif ae_features == "query":
            ae_default = qm
        else:
            raise ValueError("auto_encoder_ft must be 'query', 'caption' or 'summary' "
                             "(reference mtn.py:187-202 leaves ae_mask unbound otherwise)")
        La = ae_default.shape[1]
        S["La"] = La
        rows = B * La
        # ---- allocations + tiny prep on the caller's stream
        S["bits_his"], S["bits_cap"], S["bits_q"] = bits(his_mask), bits(cap_mask), bits(q_mask)
        ae_bits = S["bits_q"] if ae_features == "query" else S["bits_cap"]
        S["bits_ae"] = ae_bits
        mods = []
        for i in range(M):
            Lv = vid_ft[i].shape[1]
            src = ae_ft[i] if isinstance(ae_ft, (list, tuple)) else (ae_ft if ae_ft is not None else ae_default)
            mods.append({
                "Lv": Lv, "bits_vid": bits(vid_mask[i]),
                "vid": vid_ft[i].contiguous().view(-1, d),
                "vid16": torch.empty(B * Lv, d, dtype=f16, device=dev),
                "kv_vid": torch.empty(B * Lv, N * 2 * d, dtype=f16, device=dev),
                "ae": src.contiguous().view(rows, d).clone(),        # f32 residual stream of the QAE branch
                "ae16": torch.empty(rows, d, dtype=f16, device=dev),
                "xn16": torch.empty(rows, d, dtype=f16, device=dev),
                "qkv": torch.empty(rows, 3 * d, dtype=f16, device=dev),
                "obuf": torch.empty(rows, d, dtype=f16, device=dev),
                "hid": torch.empty(rows, dff, dtype=f16, device=dev),
                "kv_ae": [torch.empty(rows, 2 * d, dtype=f16, device=dev) for _ in range(N)],
                "out": torch.empty(rows, d, dtype=torch.float32, device=dev),
            })
        if M > 1 and self.qae_batched and d in (128, 256, 512, 1024):
            self._qae_batched(W, S, mods, side, main, B, La, rows, ae_bits)
            S["kv_his"], S["kv_cap"], S["kv_q"] = hoisted(his, W["kv_his"]), hoisted(cap, W["kv_cap"]), hoisted(qm, W["kv_q"])
            S["_keep"] = mods
            return S
        S["kv_ae"] = [[mods[i]["kv_ae"][l] for i in range(M)] for l in range(N)]
        S["ev"] = [[torch.cuda.Event() for _ in range(M)] for _ in range(N)]
        S["ae_out"] = [m["out"].view(B, La, d) for m in mods]
        S["side"] = side
        # ---- fork: one side stream per modality
        for i in range(M):
            side[i].wait_stream(main)
            m = mods[i]
            with torch.cuda.stream(side[i]):
                _lib.cast_f16(m["vid"], m["vid16"])
                _lib.linear(m["vid16"], W["kv_vid"][i][0], W["kv_vid"][i][1], out_f16=m["kv_vid"])
                ae = m["ae"]
                for l in range(N):
                    Lw = W["layers"][l]
                    c0 = 4 + 4 * i
                    A = Lw["ae_self"][i]
                    self._attn_block(ae, Lw["ln"][c0], A, B, La, La, A["w_qkv"], A["b_qkv"], None, 0, 0, ae_bits,
                                     m["xn16"], m["qkv"], m["obuf"])
                    A = Lw["ae_vid"][i]
                    self._attn_block(ae, Lw["ln"][c0 + 1], A, B, La, m["Lv"], A["w_qkv"][:d], A["b_qkv"][:d],
                                     m["kv_vid"], l * 2 * d, l * 2 * d + d, m["bits_vid"], m["xn16"],
                                     m["qkv"][:, :d], m["obuf"])
                    self._ffn_block(ae, Lw["ln"][c0 + 2], Lw["ae_ffn"][i], m["xn16"], m["hid"], out16=m["ae16"])
                    # K/V of this layer's ae_i for the target stream's auto_encoder_attn[i] (mtn.py:215):
                    # the memory is the un-normed ae_i itself
                    A2 = Lw["ae_attn"][i]
                    _lib.linear(m["ae16"], A2["w_qkv"][d:], A2["b_qkv"][d:], out_f16=m["kv_ae"][l])
                    S["ev"][l][i].record(side[i])
                nrm = W["ae_norm"][i]
                _lib.layernorm(ae, nrm[0], nrm[1], nrm[2], out_f32=m["out"])          # mtn.py:162-163
        # ---- text memories on the caller's stream (overlaps with the side streams)
        S["kv_his"], S["kv_cap"], S["kv_q"] = hoisted(his, W["kv_his"]), hoisted(cap, W["kv_cap"]), hoisted(qm, W["kv_q"])
        S["_keep"] = mods
        return S

    # ------------------------------------------------------------------ memory stage, cached by tensor identity
    def memory(self, W, vid_ft, vid_mask, his, his_mask, cap, cap_mask, qm, q_mask, ae_ft, ae_features):
        """(S, fresh): the memory stage of this memory set -- computed now (fresh: its side streams are still running;
        the caller waits on S["ev"] / joins S["side"]) or found cached from an earlier call with the same tensors."""
        key = _MemoryKey(self._packed._key, ae_features,
                         list(vid_ft) + list(vid_mask) + [his, his_mask, cap, cap_mask, qm, q_mask] +
                         (list(ae_ft) if isinstance(ae_ft, (list, tuple)) else [ae_ft]))
        fresh = not key.matches(self._mem_key)
        if fresh:
            self._mem_key, self._mem = None, None
            self._mem = self._memory_stage(W, vid_ft, vid_mask, his, his_mask, cap, cap_mask, qm, q_mask,
                                           ae_ft, ae_features)
            self._mem_key = key
        return self._mem, fresh

    @staticmethod
    def _site_order(ae_features):
        return (("src", "kv_q", "bits_q", "Q"), ("cap", "kv_cap", "bits_cap", "C")) \
            if ae_features in ("caption", "summary") else \
            (("cap", "kv_cap", "bits_cap", "C"), ("src", "kv_q", "bits_q", "Q"))

    # ------------------------------------------------------------------ KV-cached, last-token-only decoding
    def decode_begin(self, vid_ft, vid_mask, his, his_mask, cap, cap_mask, qm, q_mask, ae_ft, ae_features, max_len,
                     rows_per_dialogue=1):
        """Start incremental decoding of one dialogue batch (reference call form data_utils.py:202-210
        without its full-prefix recompute): runs (or finds cached) the memory stage and allocates, per layer, the
        self-attention cache [B, max_len, 3d] f16 that holds the [Q|K|V] projection of every target position seen so
        far.  Exact by SURVEY 8a invariant (ii): row t of a full-prefix decode only depends on rows <= t.
        rows_per_dialogue = R > 1 (beam search): R hypotheses per dialogue decode in lockstep; target rows are ordered
        dialogue-major (row = dialogue * R + hypothesis).  Self-attention caches are per hypothesis; the cross sites
        treat the R hypotheses of a dialogue as R query rows of ONE batch element, so the memory stage (hoisted K/V,
        QAE branch) is computed and stored once per dialogue, not per hypothesis."""
        W = self.weights()
        d, N = W["d"], W["N"]
        R = int(rows_per_dialogue)
        B = qm.shape[0] * R
        S, fresh = self.memory(W, vid_ft, vid_mask, his, his_mask, cap, cap_mask, qm, q_mask, ae_ft, ae_features)
        main = torch.cuda.current_stream()
        if fresh:                                            # the per-step path has no event waits: join here
            for st in S["side"]:
                main.wait_stream(st)
        dev = qm.device
        f16 = torch.float16
        dff = W["layers"][0]["ffn"]["w_1"].shape[0]
        return {"S": S, "W": W, "B": B, "R": R, "t": 0, "max_len": int(max_len), "ae_features": ae_features,
                "cache": [torch.zeros(B, int(max_len), 3 * d, dtype=f16, device=dev) for _ in range(N)],
                "xs": torch.empty(B, d, dtype=torch.float32, device=dev), "xn16": torch.empty(B, d, dtype=f16, device=dev),
                "q": torch.empty(B, d, dtype=f16, device=dev), "obuf": torch.empty(B, d, dtype=f16, device=dev),
                "hid": torch.empty(B, dff, dtype=f16, device=dev), "out": torch.empty(B, d, dtype=torch.float32, device=dev)}

    def decode_step(self, st, x_t, t=None, generator=None, tokens_out=None):
        """One target position for every dialogue of the batch: x_t [B, d] f32 = embedding (+ positional encoding) of the
        token at position t (default: the next one).  Returns the decoder output row [B, d] (final LayerNorm applied;
        a buffer of `st`, overwritten by the next step) -- or, when ``generator`` (mtn.Generator) and ``tokens_out`` (int64
        [B]) are given AND the step runs as the cluster kernel, ``tokens_out`` filled with the arg-max tokens (the caller
        checks ``result is tokens_out``).  Per layer: LayerNorm -> [Q|K|V] projection of the ONE new row
        straight into the cache -> attention of that row over cache rows 0..t (no mask needed: every cached key is in
        the causal past) -> output projection + residual; then the cross sites with a 1-row query over the hoisted
        K/V of the memory stage; then the FFN.  M = B rows per GEMM instead of B*(t+1)."""
        prev_rows = _lib.ROWS_KERNELS
        _lib.ROWS_KERNELS = os.environ.get("MTN_B200_DECODE_ROWS", "1") != "0"      # few-row kernels (csrc/decode_rows.cu)
        try:
            return self._decode_step(st, x_t, t, generator, tokens_out)
        finally:
            _lib.ROWS_KERNELS = prev_rows

    def _decode_step(self, st, x_t, t, generator=None, tokens_out=None):
        B, d = st["B"], st["W"]["d"]
        t = st["t"] if t is None else int(t)
        assert 0 <= t < st["max_len"], "decode_step: position %d outside the cache (max_len %d)" % (t, st["max_len"])
        plan = self._cluster_plan(st)
        if plan is not None:
            # the whole step: ONE kernel, one thread-block cluster per dialogue group (csrc/decode_cluster.cu)
            x = x_t.reshape(B, d)
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = st["xs"].copy_(x)
            gen = None
            if generator is not None and tokens_out is not None and os.environ.get("MTN_B200_DECODE_ARGMAX_FUSED", "1") != "0":
                G = generator.packed()           # greedy decoding: generator projection + arg-max as the kernel's last stage
                gen = (G["w"], G["b"], generator.proj.weight.shape[0])
            plan.step(t, x, st["out"], taps=st.get("cluster_taps"), stamps=st.get("cluster_stamps"), gen=gen,
                      tokens=tokens_out if gen is not None else None)
            st["t"] = t + 1
            return tokens_out if gen is not None else st["out"]
        st["xs"].copy_(x_t.reshape(B, d))
        prog = self._step_program(st, t)
        if prog is not None:
            prog.launch()          # the whole step: ONE persistent kernel (csrc/decode_rows.cu, decode_prog_kernel)
        else:
            self._decode_step_body(st, t)
        st["t"] = t + 1
        return st["out"]

    def _cluster_plan(self, st):
        """The _lib.DecodeClusterPlan of this decoding state, or None when the step takes the launch sequence.  Default on
        (MTN_B200_DECODE_CLUSTER=0: off) for greedy decoding and beam search (R hypotheses per dialogue share its
        memories) at d = 512, h = 8, d_ff = 2048, at most 128 target rows: a decoding step's dependency chain is per dialogue, so a cluster of 8 CTAs (one per head) takes a group
        of dialogues through all sublayers of the step with cluster barriers only, streaming its weights and K / V ahead
        of the chain -- ONE launch instead of ~135."""
        if (os.environ.get("MTN_B200_DECODE_CLUSTER", DECODE_CLUSTER_DEFAULT) == "0" or not _lib.ROWS_KERNELS or
                TAP is not None or os.environ.get("MTN_B200_DECODE_PROG", "0") == "1"):
            return None
        plan = st.get("cluster_plan", False)
        if plan is False:
            plan = st["cluster_plan"] = self._build_cluster_plan(st)
        return plan

    def _build_cluster_plan(self, st):
        S, W, B = st["S"], st["W"], st["B"]
        d, N, M = W["d"], W["N"], W["M"]
        L0 = W["layers"][0]
        h, dff = L0["self"]["h"], L0["ffn"]["w_1"].shape[0]
        if not _lib.decode_cluster_supported(B, d, h, dff, N * (5 + M)):
            return None
        bits = [S["bits_his"], S["bits_cap"], S["bits_q"], S["bits_ae"]]
        if any(b is not None and b.shape[1] != 1 for b in bits) or max(S["H"], S["C"], S["Q"], S["La"]) > 1024:
            return None
        plan = _lib.DecodeClusterPlan(B, d, h, dff, rows_per_dialogue=st["R"])
        order = self._site_order(st["ae_features"])
        for l in range(N):
            Lw = W["layers"][l]
            A = Lw["self"]
            plan.self_attention(Lw["ln"][0], A["w_qkv"], A["b_qkv"], A["w_o"], A["b_o"], st["cache"][l])
            kc, vc = l * 2 * d, l * 2 * d + d
            A = Lw["his"]
            plan.cross_attention(Lw["ln"][1], A["w_qkv"][:d], A["b_qkv"][:d], A["w_o"], A["b_o"], S["kv_his"], kc, vc, S["H"],
                                 S["bits_his"])
            for c, (name, kvn, bn, Ln) in enumerate(order):
                A = Lw[name]
                plan.cross_attention(Lw["ln"][2 + c], A["w_qkv"][:d], A["b_qkv"][:d], A["w_o"], A["b_o"], S[kvn], kc, vc, S[Ln],
                                     S[bn])
            for i in range(M):
                A = Lw["ae_attn"][i]
                plan.cross_attention(Lw["ln"][7 + 4 * i], A["w_qkv"][:d], A["b_qkv"][:d], A["w_o"], A["b_o"], S["kv_ae"][l][i],
                                     0, d, S["La"], S["bits_ae"])
            F = Lw["ffn"]
            plan.feed_forward(Lw["ln"][4 + 4 * M], F["w_1"], F["b_1"], F["w_2"], F["b_2"])
        return plan.finish(W["norm"])

    def _step_program(self, st, t):
        """The recorded program of position t of this state (greedy decoding, few-row kernels, d = 512), or None.  The stage
        list depends on t (cache row, number of cached keys) and on the state's buffers only, so it is recorded once per
        (state, t).  OPT-IN (MTN_B200_DECODE_PROG=1): measured SLOWER than the launch sequence (907 vs 650 us per step at batch
        64): a stage costs >= 2.3 us (three dependent L2 round trips: operands, release-arrive, acquire-poll), no less than
        a kernel boundary under programmatic dependent launch, and the persistent grid holds 2 CTAs per SM (registers of
        the widest stage) where the stand-alone kernels run 8 -- the stages with 512-1024 virtual blocks take several
        rounds."""
        if (not _lib.ROWS_KERNELS or st["R"] != 1 or st["W"]["d"] != 512 or st["B"] > 128 or
                os.environ.get("MTN_B200_DECODE_PROG", "0") != "1" or TAP is not None):
            return None
        progs = st.setdefault("progs", {})
        prog = progs.get(t)
        if prog is None:
            if PROGRAM_POOL:
                prog = PROGRAM_POOL.pop()
            elif torch.cuda.is_current_stream_capturing():
                return None                     # (buffers cannot be allocated during capture: fill PROGRAM_POOL before)
            else:
                prog = _lib.StepProgram()
            with prog.record():
                self._decode_step_body(st, t)
            progs[t] = prog
        return prog

    def _decode_step_body(self, st, t):
        S, W, B, R = st["S"], st["W"], st["B"], st["R"]
        D = B // R                                           # dialogues; the cross sites see [D, R] query rows
        d, N, M = W["d"], W["N"], W["M"]
        xs, xn16, qb, obuf, hid = st["xs"], st["xn16"], st["q"], st["obuf"], st["hid"]
        order = self._site_order(st["ae_features"])
        for l in range(N):
            Lw = W["layers"][l]
            A = Lw["self"]
            cache = st["cache"][l]
            row = cache[:, t]                                            # [B, 3d], row stride max_len * 3d
            _ln_linear(xs, Lw["ln"][0], A["w_qkv"], A["b_qkv"], _lib.ACT_NONE, xn16, row)
            _lib.attn_core(cache[:, t:t + 1, :d], cache[:, :, d:2 * d], cache[:, :, 2 * d:], B, A["h"], 1, t + 1, A["d_k"],
                           obuf, mask_bits=None)
            _lib.linear(obuf, A["w_o"], A["b_o"], addend=xs, out_f32=xs)
            _tap("x", xs)
            kc, vc = l * 2 * d, l * 2 * d + d
            A = Lw["his"]
            self._attn_block(xs, Lw["ln"][1], A, D, R, S["H"], A["w_qkv"][:d], A["b_qkv"][:d], S["kv_his"], kc, vc,
                             S["bits_his"], xn16, qb, obuf)
            for c, (name, kvn, bn, Ln) in enumerate(order):
                A = Lw[name]
                self._attn_block(xs, Lw["ln"][2 + c], A, D, R, S[Ln], A["w_qkv"][:d], A["b_qkv"][:d], S[kvn], kc, vc,
                                 S[bn], xn16, qb, obuf)
            for i in range(M):
                A = Lw["ae_attn"][i]
                self._attn_block(xs, Lw["ln"][7 + 4 * i], A, D, R, S["La"], A["w_qkv"][:d], A["b_qkv"][:d],
                                 S["kv_ae"][l][i], 0, d, S["bits_ae"], xn16, qb, obuf)
            self._ffn_block(xs, Lw["ln"][4 + 4 * M], Lw["ffn"], xn16, hid)
        _lib.layernorm(xs, W["norm"][0], W["norm"][1], W["norm"][2], out_f32=st["out"])   # mtn.py:164

    @staticmethod
    def decode_reorder(st, parents):
        """Beam search: target row i continues the hypothesis that was row parents[i] (int64 [B], device) -- gathers
        the self-attention caches accordingly (the cross-attention side has no per-hypothesis state)."""
        st["cache"] = [c.index_select(0, parents) for c in st["cache"]]
        st.pop("cluster_plan", None)                           # (its sites point into the old caches)

    # ------------------------------------------------------------------ forward
    def forward(self, vid_ft, vid_mask, x, his, his_mask, cap, cap_mask, qm, q_mask, tgt_mask, ae_ft,
                ae_features):
        W = self.weights()
        d, N, M = W["d"], W["N"], W["M"]
        B, T, _ = x.shape
        dev = x.device
        S, fresh = self.memory(W, vid_ft, vid_mask, his, his_mask, cap, cap_mask, qm, q_mask, ae_ft, ae_features)
        main = torch.cuda.current_stream()

        rows = B * T
        f16 = torch.float16
        xs = x.contiguous().view(rows, d).clone()               # f32 residual stream (never aliases the input)
        xn16 = torch.empty(rows, d, dtype=f16, device=dev)
        qkv = torch.empty(rows, 3 * d, dtype=f16, device=dev)
        obuf = torch.empty(rows, d, dtype=f16, device=dev)
        dff = W["layers"][0]["ffn"]["w_1"].shape[0]
        hid = torch.empty(rows, dff, dtype=f16, device=dev)
        tm = tgt_mask
        if tm is not None and tm.shape[0] != B:
            tm = tm.expand(B, -1, -1)
        bits_t = _lib.mask_pack(tm) if tm is not None else None
        if bits_t is not None:
            _tap("bits_t", bits_t)
        order = self._site_order(ae_features)
        _tap("embed.x", xs)
        for l in range(N):
            Lw = W["layers"][l]
            A = Lw["self"]
            TAP_PREFIX[0] = "L%d.self." % l
            self._attn_block(xs, Lw["ln"][0], A, B, T, T, A["w_qkv"], A["b_qkv"], None, 0, 0, bits_t, xn16, qkv, obuf)
            kc, vc = l * 2 * d, l * 2 * d + d
            A = Lw["his"]
            TAP_PREFIX[0] = "L%d.his." % l
            self._attn_block(xs, Lw["ln"][1], A, B, T, S["H"], A["w_qkv"][:d], A["b_qkv"][:d], S["kv_his"], kc, vc,
                             S["bits_his"], xn16, qkv[:, :d], obuf)
            for c, (name, kvn, bn, Ln) in enumerate(order):
                A = Lw[name]
                TAP_PREFIX[0] = "L%d.%s." % (l, name)
                self._attn_block(xs, Lw["ln"][2 + c], A, B, T, S[Ln], A["w_qkv"][:d], A["b_qkv"][:d], S[kvn], kc, vc,
                                 S[bn], xn16, qkv[:, :d], obuf)
            for i in range(M):
                if fresh:
                    main.wait_event(S["ev"][l][i])          # layer l's K/V of ae_i (side stream i)
                A = Lw["ae_attn"][i]
                TAP_PREFIX[0] = "L%d.ae%d." % (l, i)
                self._attn_block(xs, Lw["ln"][7 + 4 * i], A, B, T, S["La"], A["w_qkv"][:d], A["b_qkv"][:d],
                                 S["kv_ae"][l][i], 0, d, S["bits_ae"], xn16, qkv[:, :d], obuf)
            TAP_PREFIX[0] = "L%d.ffn." % l
            self._ffn_block(xs, Lw["ln"][4 + 4 * M], Lw["ffn"], xn16, hid)
        out = torch.empty(rows, d, dtype=torch.float32, device=dev)
        _lib.layernorm(xs, W["norm"][0], W["norm"][1], W["norm"][2], out_f32=out)   # mtn.py:164
        TAP_PREFIX[0] = ""
        _tap("final.out", out)
        if fresh:
            for st in S["side"]:                                 # join (also required for graph capture)
                main.wait_stream(st)
        return out.view(B, T, d), list(S["ae_out"])
