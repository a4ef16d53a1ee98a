"""CUDA-graph execution of a whole ``EncoderDecoder.forward`` for a fixed shape.

One forward is ~450 short kernels; issued one by one from Python the host is the
bottleneck (ctypes + launch latency >> kernel time).  ``GraphedForward`` captures
``Batch(...)`` construction + ``model.forward(batch)`` once into a CUDA graph over
static input buffers and replays it with a single launch per step -- the B200-native
replacement for a tracing compiler: streams + graphs, no code generation.
"""
import os

import torch

from .data_utils import Batch


class GraphedForward(object):
    """inputs: dict with int64 ``query, his, cap, trg, trg_y`` of shape (B, L) and ``fts``: list of
    f32 (B, Lv, F) tensors -- all on the GPU.  After construction, ``copy_inputs`` (or writing
    into ``self.static`` directly) + ``replay()`` run one forward; results are in ``self.out`` /
    ``self.ae`` (static output buffers, overwritten by the next replay)."""

    def __init__(self, model, inputs, pad=1, warmup=2, partition=None):
        """partition: a parallel.SmPartition -- the forward is captured on the big SM group's stream and the engine's side
        chain on the small group's (engine.SM_PARTITION), so the two run concurrently on disjoint SMs."""
        from . import engine as _engine
        self.model, self.pad = model, pad
        self.static = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v])
                       for k, v in inputs.items()}
        prev, _engine.SM_PARTITION = _engine.SM_PARTITION, partition
        try:
            s = torch.cuda.Stream() if partition is None else partition.main
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(warmup):                 # first-launch work (func attributes, weight packing)
                    self._run()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            kw = {} if partition is None else {"stream": partition.main}
            with torch.no_grad(), torch.cuda.graph(self.graph, **kw):
                self.out, self.ae, self.ntokens = self._run()
        finally:
            _engine.SM_PARTITION = prev

    def _run(self):
        st = self.static
        # Batch takes features in the reference's (L, B, F) layout (data_utils.py:28)
        b = Batch(st["query"], st["his"], None, [f.permute(1, 0, 2) for f in st["fts"]], st["cap"],
                  st["trg"], st["trg_y"], self.pad)
        out, ae = self.model.forward(b)
        return out, ae, b.ntokens

    def copy_inputs(self, inputs, non_blocking=True):
        for k, v in inputs.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=non_blocking)
            else:
                for dst, src in zip(self.static[k], v):
                    dst.copy_(src, non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.out, self.ae


class GraphedGreedyDecoder(object):
    """Batched greedy decoding (BASELINE configs[3]; reference semantics data_utils.py:162-186 with the
    working call form of :202-210: no EOS stop, arg-max of the last position) as CUDA graphs.

    * prefill graph: Batch(...) + ``model.encode`` + the first decode step.  The first step makes the
      engine run its memory stage (hoisted K/V of every memory for all layers, the whole
      Query-Aware Auto-Encoder branch) -- once per dialogue batch.
    * one graph per later step t: embed the t generated tokens, run the target path only (the engine
      finds the memory stage cached: same memory tensors), arg-max of the last row into ``ys[:, t]``.

    Inputs: dict with int64 ``query, his, cap`` (B, L) and ``fts`` [(B, Lv, F) f32] on the GPU.
    """

    def __init__(self, model, inputs, max_len, sos=2, pad=1, cached=True):
        """cached=True: KV-cached steps (only the new position is computed; ``model.decode_begin`` / ``decode_step``);
        cached=False: every step decodes the whole prefix again (the reference's call form)."""
        from .data_utils import subsequent_mask
        self.model, self.max_len, self.cached = model, max_len, cached
        dev = inputs["query"].device
        self.static = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v]) for k, v in inputs.items()}
        B = inputs["query"].shape[0]
        self.ys = torch.full((B, max_len), pad, dtype=torch.int64, device=dev)
        self.ys[:, 0] = sos
        self._sos = sos
        # Built outside capture and read by every replay: they MUST stay referenced for the lifetime of the graphs.
        # (Round 1 kept them in a local only: once __init__ returned the caching allocator recycled their memory and the
        # replayed steps read whatever the next small allocation wrote there -- the "batch-64 decoder instances
        # disagree" item: row 0 of the causal mask gained a stray key.  tools/bisect_decode.py found it.)
        masks = self.masks = [subsequent_mask(t, dev) for t in range(max_len)]

        def prefill():
            st = self.static
            b = Batch(st["query"], st["his"], None, [f.permute(1, 0, 2) for f in st["fts"]], st["cap"], None, None, pad)
            self.b = b
            self.mem = model.encode(b.query, b.query_mask, b.his, b.his_mask, b.cap, b.cap_mask, b.fts, b.fts_mask)
            if cached:
                q, vid, cap, his, ae = self.mem
                self.state = model.decode_begin(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask, ae,
                                                max_len)
            step(1)

        def step(t):      # produces ys[:, t] from the prefix ys[:, :t]
            q, vid, cap, his, ae = self.mem
            b = self.b
            if cached:
                model.decode_step_argmax(self.state, self.ys[:, t - 1], t - 1, out=self.ys[:, t])
                return
            out = model.decode(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask,
                               self.ys[:, :t], masks[t], ae)
            self.ys[:, t] = model.generator.argmax(out[0][:, -1])

        if cached:
            # KV-cached steps run as ONE persistent kernel each (engine._step_program); the captured prefill creates a
            # fresh decoding state, whose step programs are recorded during capture -- their buffers must exist already
            from . import engine as _engine
            from . import _lib
            _lib.lib()
            if os.environ.get("MTN_B200_DECODE_PROG", "0") == "1":
                _engine.PROGRAM_POOL.extend(_lib.StepProgram() for _ in range(max_len))
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            prefill()                                   # warm-up (weight packing, function attributes)
            for t in range(2, max_len):
                step(t)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if cached:                                      # the warm-up state's step programs go back to the pool
            _engine.PROGRAM_POOL.extend(self.state.pop("progs", {}).values())
        self.graphs = []
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            prefill()
        self.graphs.append(g)
        pool = g.pool()
        # cached decoding: ALL later positions in ONE graph (a step is a handful of launches -- embedding, the cluster
        # kernel, generator + arg-max: a graph per step would pay a graph launch for each); full-prefix mode keeps one
        # graph per position.  ``steps_in_graph[i]`` = positions graph i decodes.
        self.steps_in_graph = [1]
        if cached and max_len > 2 and os.environ.get("MTN_B200_DECODE_ONE_GRAPH", "1") != "0":
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g, pool=pool):
                for t in range(2, max_len):
                    step(t)
            self.graphs.append(g)
            self.steps_in_graph.append(max_len - 2)
        else:
            for t in range(2, max_len):
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, pool=pool):
                    step(t)
                self.graphs.append(g)
                self.steps_in_graph.append(1)

    def copy_inputs(self, inputs, non_blocking=True):
        for k, v in inputs.items():
            if k not in self.static:
                continue
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=non_blocking)
            else:
                for dst, src in zip(self.static[k], v):
                    dst.copy_(src, non_blocking=non_blocking)

    def upload(self, inputs, cache=None, video_ids=None, new_videos=(), index_buffer=None):
        """Asynchronous host -> device copy of the NEXT dialogue batch into staging buffers on a copy stream, so the
        PCIe transfer (277 MB of f32 features at BASELINE configs[3]) overlaps the decoding of the current batch.
        ``decode(staged=True)`` then starts from the staged batch.

        With ``cache`` (feature_cache.DeviceFeatureCache): ``inputs`` carries only the token ids; the features of the
        batch's ``video_ids`` are gathered ON THE DEVICE from the cache after the ``new_videos`` -- [(video id, [host
        feature tensors])], the videos not yet resident -- have been uploaded.  generate.py decodes the ten turns of a
        dialogue one after the other: nine of ten samples find their video resident."""
        if getattr(self, "_stage", None) is None:
            self._stage = {k: (torch.empty_like(v) if torch.is_tensor(v) else [torch.empty_like(t) for t in v])
                           for k, v in self.static.items()}
            self._copy = torch.cuda.Stream()
            self._ev_up, self._ev_used = torch.cuda.Event(), torch.cuda.Event()
            self._ev_used.record()
        self._copy.wait_event(self._ev_used)            # the previous staged batch has been moved into the static buffers
        with torch.cuda.stream(self._copy):
            for k, v in inputs.items():
                if k not in self._stage or (cache is not None and k == "fts"):
                    continue
                if torch.is_tensor(v):
                    self._stage[k].copy_(v, non_blocking=True)
                else:
                    for dst, src in zip(self._stage[k], v):
                        dst.copy_(src, non_blocking=True)
            if cache is not None:
                for vid, feats in new_videos:
                    cache.put(vid, feats)
                cache.gather(video_ids, out=self._stage["fts"], index_buffer=index_buffer)
            self._ev_up.record(self._copy)

    def decode(self, staged=False):
        """Runs all graphs; returns the (B, max_len) token buffer (static, overwritten by the next call)."""
        if staged:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._ev_up)
            for k, v in self._stage.items():            # device -> device, ~0.1 ms
                if torch.is_tensor(v):
                    self.static[k].copy_(v, non_blocking=True)
                else:
                    for dst, src in zip(self.static[k], v):
                        dst.copy_(src, non_blocking=True)
            self._ev_used.record(cur)
        for g in self.graphs:
            g.replay()
        return self.ys
