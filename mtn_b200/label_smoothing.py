"""Host-side mirror of the reference's ``label_smoothing.py`` (forward / evaluation): the label-smoothed
KL criterion of train.py:198-209, computed by ``mtn_label_smoothing_loss_fwd`` straight from the generator's
logits (or from log-probabilities) -- no dense target distribution, no (rows, vocab) log-prob tensor."""
import torch
import torch.nn as nn

from . import _lib
from .engine import ensure_inference


class LabelSmoothing(nn.Module):
    """Same constructor and call as the reference (label_smoothing.py:9-32): ``forward(x, target)`` with
    x = log-probabilities [rows, size], target [rows]; returns the summed KL divergence (a 0-d tensor).
    ``self.true_dist`` is not materialised (the reference stores it; nothing reads it)."""

    def __init__(self, size, padding_idx, smoothing=0.0):
        super(LabelSmoothing, self).__init__()
        self.padding_idx = padding_idx
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing
        self.size = size
        self.true_dist = None

    def forward(self, x, target):
        assert x.size(1) == self.size
        return self.from_logits(x, x.size(1), target)

    def from_logits(self, logits, V, target, scale=1.0, out=None, accumulate=False):
        """logits: [rows, ld >= V] (raw generator logits or log-probs -- the loss formula is the same)."""
        if torch.is_grad_enabled() and logits.requires_grad:      # training: criterion + its backward kernel
            from . import autograd as AG
            z = logits.float() if logits.dtype != torch.float32 else logits
            val = AG.LabelSmoothingFn.apply(z, target, V, self.padding_idx, self.smoothing, float(scale))
            if out is None:
                return val
            return out + val if accumulate else val
        ensure_inference(None, logits)
        loss = out if out is not None else torch.zeros(1, dtype=torch.float32, device=logits.device)
        _lib.label_smoothing_loss(logits.float() if logits.dtype != torch.float32 else logits, V,
                                  target.reshape(-1), self.padding_idx, self.smoothing, loss, scale=scale,
                                  accumulate=accumulate)
        return loss[0] if out is None else out
