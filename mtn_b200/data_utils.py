"""Host-side mirror of the reference's ``data_utils.py`` pieces that sit on the hot
path's boundary: ``Batch`` (masks, data_utils.py:21-54), ``subsequent_mask``
(:10-14), the ``encode`` helper (:158-160), ``beam_search_decode`` (:188-242) and a
WORKING ``greedy_decode`` (the reference's, :162-186, raises TypeError -- SURVEY 8a
row G; this one implements its intended semantics with the call form of :202-210).

Optimiser / loss / torchtext leftovers of the reference file are out of scope
 and are not mirrored.
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
            permuted = [(torch.from_numpy(ft) if isinstance(ft, np.ndarray) else ft).float().to(dev)
                        .permute(1, 0, 2) for ft in fts]
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


def encode(model, his, his_st, his_mask, cap, cap_mask, query, query_mask, video_features,
           video_features_mask):
    q_mem, vid_mem, cap_mem, his_mem, ae_ft = model.encode(query, query_mask, his, his_mask, cap,
                                                           cap_mask, video_features, video_features_mask)
    return his_mem, cap_mem, q_mem, vid_mem, ae_ft


def _encode_batch(model, batch):
    return encode(model, batch.his, batch.his_st, batch.his_mask, batch.cap, batch.cap_mask, batch.query,
                  batch.query_mask, batch.fts, batch.fts_mask)


def greedy_decode(model, batch, max_len, start_symbol, pad_symbol=None):
    """ys = [sos]; repeat max_len-1 times: decode the whole prefix, take the argmax of the
    last position (no EOS stop, as in the reference).  Works for any batch size (the
    reference is batch-1; rows are independent).  The memory stage (hoisted K/V, QAE
    branch) is computed once and reused by every step through the engine's cache."""
    his_mem, cap_mem, q_mem, vid_mem, ae_ft = _encode_batch(model, batch)
    B = batch.query.shape[0]
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


def beam_search_decode(model, batch, max_len, start_symbol, unk_symbol, end_symbol, pad_symbol, beam=5,
                       penalty=1.0, nbest=5, min_len=1):
    """Same search rules as the reference (data_utils.py:188-242): batch-1, hypotheses are
    expanded best-first skipping <unk>/<eos>, an argmin-replacement beam, EOS hypotheses
    scored lp + penalty*(len+1) once l >= min_len, n-best sorted by score."""
    his_mem, cap_mem, q_mem, vid_mem, ae_ft = _encode_batch(model, batch)
    q = batch.query
    ds = torch.full((1, 1), start_symbol, dtype=q.dtype, device=q.device)
    hyplist = [([], 0., ds)]
    best_state = None
    comp_hyplist = []
    for l in range(max_len):
        new_hyplist = []
        argmin = 0
        for out, lp, st in hyplist:
            output = model.decode(vid_mem, his_mem, cap_mem, q_mem, batch.fts_mask, batch.his_mask,
                                  batch.cap_mask, batch.query_mask, st,
                                  subsequent_mask(st.size(1), st.device), ae_ft)
            logp = model.generator(output[0][:, -1])
            lp_vec = np.squeeze(logp.cpu().data.numpy() + lp)
            if l >= min_len:
                new_lp = lp_vec[end_symbol] + penalty * (len(out) + 1)
                comp_hyplist.append((out, new_lp))
                if best_state is None or best_state < new_lp:
                    best_state = new_lp
            for o in np.argsort(lp_vec)[::-1]:
                if o == unk_symbol or o == end_symbol:
                    continue
                new_lp = lp_vec[o]
                if len(new_hyplist) == beam:
                    if new_hyplist[argmin][1] < new_lp:
                        new_st = torch.cat([st, torch.full((1, 1), int(o), dtype=q.dtype, device=q.device)], dim=1)
                        new_hyplist[argmin] = (out + [o], new_lp, new_st)
                        argmin = min(enumerate(new_hyplist), key=lambda h: h[1][1])[0]
                    else:
                        break
                else:
                    new_st = torch.cat([st, torch.full((1, 1), int(o), dtype=q.dtype, device=q.device)], dim=1)
                    new_hyplist.append((out + [o], new_lp, new_st))
                    if len(new_hyplist) == beam:
                        argmin = min(enumerate(new_hyplist), key=lambda h: h[1][1])[0]
        hyplist = new_hyplist
    if len(comp_hyplist) > 0:
        maxhyps = sorted(comp_hyplist, key=lambda h: -h[1])[:nbest]
        return maxhyps, best_state
    return [([], 0)], None
