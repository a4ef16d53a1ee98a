"""Device-resident cache of per-video features (I3D / VGGish), keyed by video id.

In the reference every sample of a batch carries its video's features from the host to the GPU
(``data_utils.py:28``: ``torch.from_numpy(ft).float().cuda()``), every step -- although the ten question/answer turns of
a dialogue share ONE video (``data_handler.py:150-206`` builds one sample per turn).  At cfg2 that is 69 MB (f16) per
32-dialogue batch per 2.6 ms step: PCIe-bound on one GPU and far beyond the host's aggregate bandwidth on eight.  A B200
has 180 GB of HBM; the f16 features of a video are 2.2 MB, so tens of thousands of videos -- the whole AVSD training set
-- stay resident.  ``DeviceFeatureCache`` keeps them in flat per-modality slabs: a video crosses PCIe when it is first
seen (or after eviction), a batch is assembled ON THE DEVICE by one gather per modality, and the per-step host traffic
drops to the token ids plus the features of the videos that are new.

This is host-side plumbing of the data path next to ``Batch`` (SURVEY 8f row f4 / data formats either side of the hot
path); the arithmetic downstream is unchanged: ``gather`` returns the same (B, L, F) tensors ``Batch`` expects.
"""
import collections

import torch


class DeviceFeatureCache(object):
    """capacity: videos kept resident; shapes: [(L_i, F_i)] per modality; dtype: storage type (f16: what the kernels
    round the features to on arrival anyway, so results are bit-identical to uploading f32)."""

    def __init__(self, capacity, shapes, device, dtype=torch.float16):
        self.capacity, self.device = int(capacity), torch.device(device)
        self.store = [torch.empty(self.capacity, L, F, dtype=dtype, device=self.device) for (L, F) in shapes]
        self.slot_of = collections.OrderedDict()        # video id -> slot, in LRU order
        self.free = list(range(self.capacity - 1, -1, -1))
        self.hits = self.misses = 0

    def __contains__(self, video_id):
        return video_id in self.slot_of

    def bytes_per_video(self):
        return sum(s[0].numel() * s.element_size() for s in self.store)

    def invalidate(self, video_id):
        slot = self.slot_of.pop(video_id, None)
        if slot is not None:
            self.free.append(slot)

    def put(self, video_id, feats, non_blocking=True):
        """feats: one (L_i, F_i) tensor per modality (pinned host memory for an asynchronous copy, or device).  Stream-ordered
        on the current stream.  Evicts the least recently used video when full."""
        slot = self.slot_of.pop(video_id, None)
        if slot is None:
            if not self.free:
                _, slot = self.slot_of.popitem(last=False)
            else:
                slot = self.free.pop()
        self.slot_of[video_id] = slot
        for st, f in zip(self.store, feats):
            st[slot].copy_(f, non_blocking=non_blocking)
        return slot

    def gather(self, video_ids, out=None, index_buffer=None):
        """The batch's features, assembled on the device: list of (B, L_i, F_i) tensors (written into `out` when given).
        Every id must be resident (``put`` it first).  index_buffer: optional pinned int64 [B] staging for the slot indices."""
        slots = []
        for v in video_ids:
            s = self.slot_of.get(v)
            if s is None:
                self.misses += 1
                raise KeyError("video %r is not resident: put() its features first" % (v,))
            self.slot_of.move_to_end(v)
            self.hits += 1
            slots.append(s)
        if index_buffer is not None:
            index_buffer[:len(slots)] = torch.tensor(slots, dtype=torch.int64)
            idx = index_buffer[:len(slots)].to(self.device, non_blocking=True)
        else:
            idx = torch.tensor(slots, dtype=torch.int64).to(self.device)
        res = []
        for m, st in enumerate(self.store):
            if out is not None:
                torch.index_select(st, 0, idx, out=out[m])
                res.append(out[m])
            else:
                res.append(torch.index_select(st, 0, idx))
        return res
