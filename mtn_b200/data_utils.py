"""Host-side mirror of the reference's ``data_utils.py`` pieces that sit on the hot
path's boundary: ``Batch`` (masks, data_utils.py:21-54), ``subsequent_mask``
(:10-14), the ``encode`` helper (:158-160), ``beam_search_decode`` (:188-242) and a
WORKING ``greedy_decode`` (the reference's, :162-186, raises TypeError -- SURVEY 8a
row G; this one implements its intended semantics with the call form of :202-210).

Torchtext leftovers of the reference file are out of scope and are not mirrored.  ``Batch``'s field
block and ``NoamOpt`` are interface mirrors: they follow the reference's ``data_utils.py:21-46`` and ``:92-117`` almost
line for line on purpose (``train.py`` / ``generate.py`` read these attributes and drive this schedule); everything that
computes -- masks on the device, the fused loss, greedy / beam decoding -- is this repo's own code.
"""
import numpy as np
import torch


def subsequent_mask(size, device=None):
    """Mask out subsequent positions: (1, size, size) bool, True on/below the diagonal.
    ``device`` (an extension of the reference signature) builds it in place on the GPU so
    that Batch construction is CUDA-graph capturable (no host->device copy)."""
    return torch.tril(torch.ones(1, size, size, dtype=torch.bool, device=device))


def _default_device(t):
    if t.is_cuda:
        return t.device
    return torch.device("cuda") if torch.cuda.is_available() else t.device


class Batch:
    """Same constructor and fields as the reference ``Batch`` (data_utils.py:21-46).
    ``fts``: list of (L, B, F) numpy arrays or tensors (the reference's layout); they are
    moved to the GPU (the reference hard-codes ``.cuda()``, :28), permuted to (B, L, F),
    frames whose elements are all exactly 1.0 are padding (:29) and are zeroed (:30)."""

    def __init__(self, query, his, his_st, fts=None, cap=None, trg=None, trg_y=None, pad=0):
        self.query = query
        self.his = his
        self.his_st = his_st
        if fts is not None:
            dev = _default_device(query)
            def load(ft):      # f16 features (a loader that stores them as f16: half the upload) stay f16
                t = torch.from_numpy(ft) if isinstance(ft, np.ndarray) else ft
                return (t if t.dtype == torch.float16 and dev.type == "cuda" else t.float()).to(dev).permute(1, 0, 2)
            permuted = [load(ft) for ft in fts]
            if dev.type == "cuda" and all(ft.shape[2] % 8 == 0 for ft in permuted):
                # one pass per modality: padding mask + zeroing + f16 operand of the video encoder
                from . import _lib
                prepped = [_lib.feature_prep(ft) for ft in permuted]
                self.fts_mask = [m for m, _ in prepped]
                self.fts = [f for _, f in prepped]          # f16 on the CUDA path (f32 in the reference)
            else:
                self.fts_mask = [(torch.sum(ft != 1, dim=2) != 0).unsqueeze(-2) for ft in permuted]
                self.fts = [ft * self.fts_mask[i].squeeze(1).unsqueeze(-1).float()
                            for i, ft in enumerate(permuted)]
        else:
            self.fts = None
            self.fts_mask = None
        self.query_mask = (query != pad).unsqueeze(-2)
        self.his_mask = (his != pad).unsqueeze(-2)
        if cap is not None:
            self.cap = cap
            self.cap_mask = (cap != pad).unsqueeze(-2)
        else:
            self.cap = None
            self.cap_mask = None
        if trg is not None:
            self.trg = trg
            self.trg_y = trg_y
            self.trg_mask = self.make_std_mask(self.trg, pad)
            self.ntokens = (self.trg_y != pad).data.sum()

    @staticmethod
    def make_std_mask(tgt, pad):
        "Hide padding and future words: (B, T, T)."
        return (tgt != pad).unsqueeze(-2) & subsequent_mask(tgt.size(-1), tgt.device)


class NoamOpt:
    """Learning-rate schedule wrapper of the reference (data_utils.py:92-117; train.py:190-191 builds it around
    torch.optim.Adam): rate = factor * d_model^-0.5 * min(step^-0.5, step * warmup^-1.5).  Host logic only."""

    def __init__(self, model_size, factor, warmup, optimizer):
        self.optimizer = optimizer
        self._step = 0
        self.warmup = warmup
        self.factor = factor
        self.model_size = model_size
        self._rate = 0

    def step(self):
        self._step += 1
        rate = self.rate()
        for p in self.optimizer.param_groups:
            p['lr'] = rate
        self._rate = rate
        self.optimizer.step()

    def rate(self, step=None):
        if step is None:
            step = self._step
        return self.factor * (self.model_size ** (-0.5) * min(step ** (-0.5), step * self.warmup ** (-1.5)))


class SimpleLossCompute:
    """The reference's SimpleLossCompute (data_utils.py:123-156): main loss / norm plus
    l * sum_i auto-encoder loss_i / ae_norm (every stream through ``generator`` unless ``ae_generator`` is
    given); with ``opt`` it also runs ``loss.backward()``, ``opt.step()`` and ``opt.optimizer.zero_grad()``
    (data_utils.py:152-155).  Returns ``loss * norm`` like the reference.  The generator's logits feed the
    fused label-smoothing kernel directly (no (rows, vocab) log-prob tensor, no dense target distribution);
    in training the same two steps are autograd Functions backed by the backward kernels."""

    def __init__(self, generator, ae_generator, criterion, opt=None, l=1.0):
        self.generator, self.ae_generator, self.criterion, self.opt, self.l = generator, ae_generator, criterion, opt, l

    def loss(self, x, y, norm, ae_x=None, ae_y=None, ae_norm=None):
        """The normalised loss as a 0-d tensor (differentiable when autograd is recording).  No host sync: ``norm`` /
        ``ae_norm`` may be Python numbers (folded into the kernels' scale) or 0-d DEVICE tensors (the counts of the batch
        in flight, e.g. inside a captured step: the division then happens on the device, per replay)."""
        def scaled(logits, V, target, weight, n):
            if torch.is_tensor(n) and n.is_cuda:
                return self.criterion.from_logits(logits, V, target, scale=float(weight)) / n.to(torch.float32)
            return self.criterion.from_logits(logits, V, target, scale=float(weight) / float(n))

        logits, V = self.generator._logits(x)
        total = scaled(logits, V, y, 1.0, norm)
        if ae_x is not None:
            streams = ae_x if isinstance(ae_x, (list, tuple)) else [ae_x]
            for i, ae_in in enumerate(streams):
                gen = self.generator if self.ae_generator is None else (
                    self.ae_generator[i] if isinstance(ae_x, (list, tuple)) else self.ae_generator)
                lg, Vg = gen._logits(ae_in)
                total = total + scaled(lg, Vg, ae_y, self.l, ae_norm)
        return total

    def __call__(self, x, y, norm, ae_x=None, ae_y=None, ae_norm=None):
        total = self.loss(x, y, norm, ae_x, ae_y, ae_norm)
        if self.opt is not None:
            total.backward()
            self.opt.step()
            self.opt.optimizer.zero_grad()
        return total.item() * float(norm)


def encode(model, his, his_st, his_mask, cap, cap_mask, query, query_mask, video_features,
           video_features_mask):
    q_mem, vid_mem, cap_mem, his_mem, ae_ft = model.encode(query, query_mask, his, his_mask, cap,
                                                           cap_mask, video_features, video_features_mask)
    return his_mem, cap_mem, q_mem, vid_mem, ae_ft


def _encode_batch(model, batch):
    return encode(model, batch.his, batch.his_st, batch.his_mask, batch.cap, batch.cap_mask, batch.query,
                  batch.query_mask, batch.fts, batch.fts_mask)


def greedy_decode(model, batch, max_len, start_symbol, pad_symbol=None, cached=None):
    """ys = [sos]; repeat max_len-1 times: decode, take the argmax of the last position (no EOS stop, as in the
    reference).  Works for any batch size (the reference is batch-1; rows are independent).  The memory stage
    (hoisted K/V, QAE branch) is computed once per dialogue batch.

    cached=True (default where the model offers ``decode_begin`` / ``decode_step``): KV-cached decoding -- every
    step computes ONLY the new position (self-attention over the per-layer cache, 1-row cross-attention queries);
    cached=False: the reference's call form, the whole prefix is decoded again at every step (data_utils.py:202-210).
    Both produce the same tokens (tests/test_gpu_model.py)."""
    his_mem, cap_mem, q_mem, vid_mem, ae_ft = _encode_batch(model, batch)
    B = batch.query.shape[0]
    if cached is None:
        cached = hasattr(model, "decode_begin") and getattr(model, "_fused_embed_ok", lambda: False)()
    if cached:
        st = model.decode_begin(vid_mem, his_mem, cap_mem, q_mem, batch.fts_mask, batch.his_mask, batch.cap_mask,
                                batch.query_mask, ae_ft, max_len)
        ys = torch.full((B, max_len), start_symbol, dtype=batch.query.dtype, device=batch.query.device)
        for t in range(max_len - 1):
            if hasattr(model, "decode_step_argmax"):
                model.decode_step_argmax(st, ys[:, t], t, out=ys[:, t + 1])
            else:
                ys[:, t + 1] = model.generator.argmax(model.decode_step(st, ys[:, t], t))
        return ys
    ys = torch.full((B, 1), start_symbol, dtype=batch.query.dtype, device=batch.query.device)
    for _ in range(max_len - 1):
        out = model.decode(vid_mem, his_mem, cap_mem, q_mem, batch.fts_mask, batch.his_mask,
                           batch.cap_mask, batch.query_mask, ys,
                           subsequent_mask(ys.size(1), ys.device), ae_ft)
        last = out[0][:, -1]
        if hasattr(model.generator, "argmax"):
            nxt = model.generator.argmax(last).unsqueeze(1)
        else:
            nxt = model.generator(last).argmax(dim=1, keepdim=True)
        ys = torch.cat([ys, nxt.to(ys.dtype)], dim=1)
    return ys


class _BeamPool(object):
    """The reference's per-step hypothesis pool (data_utils.py:223-234): grows to ``width`` entries, then a
    candidate only enters by REPLACING the current worst entry (first minimal score), and the first
    candidate that does not beat the worst ends the expansion of that hypothesis."""

    def __init__(self, width):
        self.width, self.entries, self.worst = width, [], 0

    def _refresh_worst(self):
        self.worst = min(range(len(self.entries)), key=lambda i: self.entries[i][1])

    def offer(self, entry):
        """entry = (tokens, score, prefix).  False <=> the pool is full and the entry was rejected."""
        if len(self.entries) < self.width:
            self.entries.append(entry)
            if len(self.entries) == self.width:
                self._refresh_worst()
            return True
        if self.entries[self.worst][1] < entry[1]:
            self.entries[self.worst] = entry
            self._refresh_worst()
            return True
        return False


def beam_search_decode_batched(model, batch, max_len, start_symbol, unk_symbol, end_symbol, pad_symbol, beam=5,
                               penalty=1.0, nbest=5, min_len=1):
    """The reference's beam search (data_utils.py:188-242, the path generate.py:56 calls) for a batch of D dialogues
    at once: returns a list of D results ``(nbest hypotheses [(tokens, score)], best finished score)``, each identical
    to what the reference's batch-of-one search returns for that dialogue.

    All live hypotheses of all dialogues advance in ONE KV-cached ``model.decode_step`` per position (rows ordered
    dialogue-major, ``beam`` rows per dialogue), the generator's log-probabilities stay on the device, and ONE
    device-to-host copy per step brings back what the selection rules can possibly use: per hypothesis the best
    ``beam + 2`` continuations (at most ``beam`` can enter a pool of ``beam`` entries; <unk> and <eos> are skipped)
    and the <eos> log-probability.  The pool rules themselves (grow to ``beam``, then replace the current worst,
    first rejected candidate ends a hypothesis' expansion, data_utils.py:219-234) run on the host per dialogue."""
    his_mem, cap_mem, q_mem, vid_mem, ae_ft = _encode_batch(model, batch)
    ids = batch.query
    D, dev = ids.shape[0], ids.device
    R = int(beam)
    st = model.decode_begin(vid_mem, his_mem, cap_mem, q_mem, batch.fts_mask, batch.his_mask, batch.cap_mask,
                            batch.query_mask, ae_ft, max_len, rows_per_dialogue=R)
    tokens = torch.full((D * R,), start_symbol, dtype=ids.dtype, device=dev)
    # per dialogue: live hypotheses [(tokens, score, row of this step's batch)]
    live = [[([], 0., d * R)] for d in range(D)]
    finished = [[] for _ in range(D)]
    best = [None] * D
    K = R + 2
    for step in range(max_len):
        logp = model.generator(model.decode_step(st, tokens, step))             # [D*R, V] log-probabilities
        top_v, top_i = torch.topk(logp, min(K, logp.shape[1]), dim=1)           # sorted, best first
        pack = torch.cat([top_v, top_i.to(top_v.dtype), logp[:, end_symbol:end_symbol + 1]], dim=1).cpu().numpy()
        kk = top_v.shape[1]
        parents = np.arange(D * R, dtype=np.int64)
        nxt = np.full(D * R, int(pad_symbol if pad_symbol is not None else start_symbol), dtype=np.int64)
        for d in range(D):
            pool = _BeamPool(R)
            for toks, score, row in live[d]:
                vals, idxs, lp_eos = pack[row, :kk] + score, pack[row, kk:2 * kk].astype(np.int64), pack[row, 2 * kk] + score
                if step >= min_len:
                    done = lp_eos + penalty * (len(toks) + 1)
                    finished[d].append((toks, done))
                    if best[d] is None or best[d] < done:
                        best[d] = done
                for v, tok in zip(vals, idxs):
                    if tok == unk_symbol or tok == end_symbol:
                        continue
                    if not pool.offer((toks + [int(tok)], float(v), row)):
                        break
            live[d] = []
            for j, (toks, score, parent) in enumerate(pool.entries):
                row = d * R + j
                parents[row], nxt[row] = parent, toks[-1]
                live[d].append((toks, score, row))
        if step + 1 < max_len:
            model.decode_reorder(st, torch.from_numpy(parents).to(dev))
            tokens = torch.from_numpy(nxt).to(dev).to(ids.dtype)
    res = []
    for d in range(D):
        if finished[d]:
            res.append((sorted(finished[d], key=lambda h: -h[1])[:nbest], best[d]))
        else:
            res.append(([([], 0)], None))
    return res


def beam_search_decode(model, batch, max_len, start_symbol, unk_symbol, end_symbol, pad_symbol, beam=5,
                       penalty=1.0, nbest=5, min_len=1):
    """Beam search with the reference's rules (data_utils.py:188-242): batch of one dialogue; each live
    hypothesis is expanded best-token-first, never with <unk> or <eos>; from step ``min_len`` on every live
    hypothesis also contributes a finished candidate scored ``logp[eos] + penalty * (len + 1)``; the n-best
    finished candidates and the best finished score are returned.

    On the CUDA model this is ``beam_search_decode_batched`` for D = 1 (all hypotheses in one KV-cached step, one
    D2H per position).  A model without ``decode_begin`` (or MTN_B200_BEAM_SERIAL=1) takes the reference's own call
    form below: one ``model.decode`` of the whole prefix per hypothesis per step."""
    import os
    if (hasattr(model, "decode_begin") and getattr(model, "_fused_embed_ok", lambda: False)() and
            os.environ.get("MTN_B200_BEAM_SERIAL") != "1" and batch.query.shape[0] == 1):
        return beam_search_decode_batched(model, batch, max_len, start_symbol, unk_symbol, end_symbol, pad_symbol, beam,
                                          penalty, nbest, min_len)[0]
    his_mem, cap_mem, q_mem, vid_mem, ae_ft = _encode_batch(model, batch)
    ids = batch.query

    def next_logp(prefix):
        dec = model.decode(vid_mem, his_mem, cap_mem, q_mem, batch.fts_mask, batch.his_mask, batch.cap_mask,
                           batch.query_mask, prefix, subsequent_mask(prefix.size(1), prefix.device), ae_ft)
        dec = dec[0] if isinstance(dec, (tuple, list)) else dec
        return np.squeeze(model.generator(dec[:, -1]).cpu().data.numpy())

    def extend(prefix, token):
        return torch.cat([prefix, torch.full((1, 1), int(token), dtype=ids.dtype, device=ids.device)], dim=1)

    live = [([], 0., torch.full((1, 1), start_symbol, dtype=ids.dtype, device=ids.device))]
    finished, best_finished = [], None
    for step in range(max_len):
        pool = _BeamPool(beam)
        for tokens, score, prefix in live:
            scores = next_logp(prefix) + score
            if step >= min_len:
                done = scores[end_symbol] + penalty * (len(tokens) + 1)
                finished.append((tokens, done))
                if best_finished is None or best_finished < done:
                    best_finished = done
            for tok in np.argsort(scores)[::-1]:
                if tok == unk_symbol or tok == end_symbol:
                    continue
                if not pool.offer((tokens + [tok], scores[tok], extend(prefix, tok))):
                    break
        live = pool.entries
    if finished:
        return sorted(finished, key=lambda h: -h[1])[:nbest], best_finished
    return [([], 0)], None
