"""One data-parallel training step of MTN on the sm_100a kernels.

Mirrors the reference's step -- ``model.forward(b)`` (train.py:33), ``SimpleLossCompute`` on the decoder output and
on every auto-encoder stream against the query ids (train.py:37-39, data_utils.py:132-151), ``loss.backward()``,
``opt.step()``, ``zero_grad()`` (data_utils.py:152-155) -- plus what the reference does not have:

* independent dialogue shards per rank, the loss normalised by the GLOBAL token counts, and ONE NCCL all-reduce
  (SUM) over a flat f32 buffer that backs every parameter's ``.grad`` (so the sum of the shard gradients equals the
  single-process gradient of the whole batch);
* the whole step (forward, loss, backward, all-reduce, optimizer) captured once into a CUDA graph over static input
  buffers and replayed with one launch -- ~1700 kernels per step would otherwise be host-bound.

Dropout: the fused kernels implement p = 0 only; ``TrainStep`` sets every nn.Dropout.p to 0 and says so.
"""
import torch
import torch.distributed as dist

from .data_utils import Batch, SimpleLossCompute
from .label_smoothing import LabelSmoothing


def flatten_grads(model):
    """Back every parameter's .grad by a view of one flat f32 buffer (zeroed) -- the buffer of the single gradient
    all-reduce.  The decoder's segment uses the training engine's arena layout, so its backward kernels accumulate
    straight into it.  Returns the buffer."""
    dec = getattr(model, "decoder", None)
    tr = dec.trainer if dec is not None and hasattr(dec, "trainer") else None
    dec_ids = set(id(p) for p in dec.parameters()) if tr is not None else set()
    rest = [p for p in model.parameters() if p.requires_grad and id(p) not in dec_ids]
    n_dec = tr.arena_numel() if tr is not None else 0
    total = n_dec + sum((p.numel() + 7) // 8 * 8 for p in rest)
    flat = torch.zeros(total, dtype=torch.float32, device=next(model.parameters()).device)
    if tr is not None:
        tr.attach_grads(flat[:n_dec])
    off = n_dec
    for p in rest:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += (p.numel() + 7) // 8 * 8
    return flat


def invalidate_packed_weights(model):
    """Force every cached f16 weight pack to be rebuilt on the next forward (used before graph capture so that the
    re-packing of the freshly updated parameters is part of the captured step)."""
    for m in model.modules():
        pw = getattr(m, "_packed", None)
        if pw is not None:
            pw._key = None
    dec = getattr(model, "decoder", None)
    if dec is not None and getattr(dec, "_engine", None) is not None:
        dec._engine._packed._key = None


class TrainStep(object):
    def __init__(self, model, vocab, pad=1, smoothing=0.1, lam=1.0, lr=1e-4, optimizer=None, graph=True):
        self.model, self.pad, self.graph_enabled = model, pad, graph
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        model.train()
        self.flat = flatten_grads(model)
        self.opt = optimizer if optimizer is not None else torch.optim.Adam(
            model.parameters(), lr=lr, betas=(0.9, 0.98), eps=1e-9, fused=True, capturable=graph)   # train.py:190-191
        self.loss_compute = SimpleLossCompute(model.generator, None, LabelSmoothing(vocab, pad, smoothing), opt=None, l=lam)
        self.static, self.graph, self.loss = None, None, None

    # one step on the batch held in `st` (dict of device tensors); ntokens are GLOBAL (summed over ranks) host numbers
    def _step(self, st, ntokens, ntokens_query):
        b = Batch(st["query"], st["his"], None, [f.permute(1, 0, 2) for f in st["fts"]], st["cap"], st["trg"], st["trg_y"],
                  self.pad)
        out, ae = self.model.forward(b)
        loss = self.loss_compute.loss(out, b.trg_y, ntokens, ae, b.query, ntokens_query)
        loss.backward()                                   # accumulates into the flat buffer's views
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)          # the ONE gradient collective
        self.opt.step()
        self.flat.zero_()
        return loss.detach()

    def eager(self, batch, ntokens, ntokens_query):
        return self._step(batch, ntokens, ntokens_query)

    def capture(self, batch, ntokens, ntokens_query, warmup=3):
        """Capture the step for this batch SHAPE and these normalisers; afterwards ``replay(batch)``."""
        self.static = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v]) for k, v in batch.items()}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step(self.static, ntokens, ntokens_query)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        invalidate_packed_weights(self.model)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step(self.static, ntokens, ntokens_query)
        return self

    def replay(self, batch=None):
        if batch is not None:
            for k, v in batch.items():
                if torch.is_tensor(v):
                    self.static[k].copy_(v, non_blocking=True)
                else:
                    for dst, src in zip(self.static[k], v):
                        dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
