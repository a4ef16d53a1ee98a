#!/usr/bin/env python
"""bench.py -- decoder tokens/sec of the MTN hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one ``EncoderDecoder.forward`` (encode + the 6-layer decoder cascade,
reference mtn.py:28-30) over one synthetic batch of BASELINE.json configs[1]:
N=6, d_model=512, h=8, d_ff=2048, batch 32 per GPU, query/caption 64, history 256,
I3D (2048-d, 512 frames) + VGGish (128-d, 256 frames), target length 256 (the
north_star's "seq=256"; ``--tgt-len 20`` gives the dialogue-realistic variant), ragged
padding.  tokens = non-pad target tokens (train.py:41-48, data_utils.py:46).

Prints ONE JSON line (rank 0):
  value        whole-job tokens/s with inputs resident in HBM (CUDA-graph replay per step,
               CUDA events, max over ranks; batches rotate so consecutive steps read
               different inputs and the working set exceeds L2)
  e2e          same metric through the public API with HOST inputs: pinned-host -> device
               copy of ids + features and device -> host copy of the decoder output inside
               the timed region, every step
  roofline     tensor-pipe roofline of the dominant kernel (the tcgen05 linear kernel),
               algorithmic FLOPs / CUDA-event time of its launches in one traced step
  cpu_baseline the CPU oracle (port of the reference forward, oracle/mtn_oracle.py) timed on
               this box's host cores on a bounded sample of the same workload
``--impl reference`` times that CPU implementation as the reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = {"N": 6, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 3000, "ft_sizes": [2048, 128],
       "auto_encoder_ft": "query", "diff_encoder": True}
SHAPE = {"B": 32, "Q": 64, "C": 64, "H": 256, "Lv": [512, 256]}
METRIC = "decoder tokens/sec at d_model=512 h=8 L=6 (forward)"


def oracle():
    """The CPU oracle is test / baseline infrastructure: imported only by the cpu_baseline and
    --impl reference legs (and by tests), never by the measured GPU path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mtn_oracle
    return mtn_oracle


def flops_forward(B, T, cfg=CFG, shp=SHAPE):
    """Algorithmic FLOPs of one forward (SURVEY 8d formula; causal self-attention counted dense)."""
    d, N = cfg["d_model"], cfg["N"]
    Q, C, H, Lv = shp["Q"], shp["C"], shp["H"], shp["Lv"]
    site = lambda lq, lk: 4 * B * d * d * (lq + lk) + 4 * B * lq * lk * d
    ffn = lambda l: 16 * B * l * d * d
    layer = site(T, T) + site(T, H) + site(T, C) + site(T, Q) + ffn(T)
    for lv in Lv:
        layer += site(Q, Q) + site(Q, lv) + ffn(Q) + site(T, Q)
    vid = sum(2 * B * lv * f * d for lv, f in zip(Lv, cfg["ft_sizes"]))
    return N * layer + vid


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_from_profiles(n_launches, train=False):
    """(DRAM bytes per launch of the linear kernel, provenance dict) from the committed ncu capture of the same step
    (profiles/rNN_gemm_traffic.json for the forward, rNN_train_gemm_traffic.json for the training step, written by
    tools/summarize_profiles.py); (None, None) if absent."""
    import glob
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json"))
                   if ("_train_" in os.path.basename(f)) == train)
    if not files:
        return None, None
    t = json.load(open(files[-1]))
    per_launch = (t["dram_bytes_read"] + t["dram_bytes_write"]) / max(1, t["launches"])
    return per_launch, {"launches_in_capture": t["launches"], "launches_in_this_run": n_launches,
                        "source": t["source"], "unit": "bytes per launch (average over the capture)"}


def synth(O, B, T, seed):
    return O.synth_inputs(CFG, B=B, Q=SHAPE["Q"], C=SHAPE["C"], H=SHAPE["H"], T=T, Lv=SHAPE["Lv"], seed=seed)


def cpu_forward_seconds(O, sd, inp, reps):
    """Median wall time of the oracle forward (all host threads) on `inp`."""
    ts = []
    O.forward(sd, CFG, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["fts"])      # warm-up
    for _ in range(reps):
        t0 = time.perf_counter()
        O.forward(sd, CFG, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["fts"])
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) if ts else None


def reference_forward_fn(sd_seed):
    """The reference arm's step function.  Lookup: $MTN_REF_DIR, then baseline/_ref/ -- a copy of the
    UNMODIFIED reference that travels with the repo snapshot; when one is there its own ``mtn.make_model`` /
    ``EncoderDecoder.forward`` run (kind "reference").  Otherwise (the reference is six Python files without package
    metadata, nothing is installed by default, and /root/reference does not exist on the GPU box) the CPU oracle port
    runs (kind "port").  Returns (kind, fn(inp) -> None)."""
    O = oracle()
    sd = O.init_state_dict(CFG, sd_seed)
    try:
        import ref_loader
        d = ref_loader.ref_dir(allow_build_container_path=False)
        if d is not None:
            ref_mtn, _ = ref_loader.load(d)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                model = ref_mtn.make_model(CFG["vocab"], CFG["vocab"], N=CFG["N"], d_model=CFG["d_model"], d_ff=CFG["d_ff"],
                                           h=CFG["h"], dropout=0.1, ft_sizes=CFG["ft_sizes"], diff_encoder=True,
                                           auto_encoder_ft="query").eval()
            model.load_state_dict(sd, strict=True)

            def fn(inp):
                b = ref_loader.make_cpu_batch(inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
                with torch.no_grad():
                    model.forward(b)
            return "reference", fn
    except Exception as e:      # fall back to the port, say why on stderr
        print("bench: reference import failed (%r); timing the oracle port" % (e,), file=sys.stderr)
    return "port", (lambda inp: O.forward(sd, CFG, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["fts"]))


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on this box's host cores, all threads, the same
    workload as our arm.  Each step is the FULL batch unless K steps of it would not finish within ~3 minutes; then
    each step is a bounded sample (the first n dialogues) and both `config.global_batch` and `cpu_baseline.sample`
    say so -- the line never claims a batch it did not run."""
    if rank != 0:
        return
    O = oracle()
    torch.set_num_threads(os.cpu_count() or 1)
    kind, fwd = reference_forward_fn(7)
    full = synth(O, args.batch, args.tgt_len, 1000)
    take = lambda n: {k: (v[:n] if torch.is_tensor(v) else [f[:n] for f in v]) for k, v in full.items()}
    probe_n = min(args.batch, max(1, args.cpu_batch))
    fwd(take(probe_n))                                                  # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    fwd(take(probe_n))
    per_dialogue = (time.perf_counter() - t0) / probe_n
    budget = float(os.environ.get("MTN_B200_REF_BUDGET_S", "170"))
    n = args.batch if per_dialogue * args.batch * (args.steps + 1) <= budget else \
        max(1, min(args.batch, int(budget / (per_dialogue * (args.steps + 1)))))
    inp = take(n)
    ntok = int((inp["trg_y"] != 1).sum())
    for _ in range(max(1, min(args.warmup, 1))):
        fwd(inp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd(inp)
    dt = (time.perf_counter() - t0) / args.steps
    v = ntok / dt
    cfg = workload_config(args, n)
    cfg["reference_sample"] = ("full batch of %d dialogues per step" % n) if n == args.batch else \
        ("first %d of the %d dialogues per step (bounded sample: the full batch would take %.0f s for %d steps)"
         % (n, args.batch, per_dialogue * args.batch * args.steps, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": "%d dialogues per step (%d tokens), %d steps" % (n, ntok, args.steps)},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, B):
    return {"workload": "MTN EncoderDecoder.forward, BASELINE configs[%s]: N=%d d_model=%d h=%d d_ff=%d, "
                        "batch=%d/GPU, query=caption=64, history=256, I3D 2048-d x%d + VGGish 128-d x%d, "
                        "target len %d, ragged padding" % ("1" if args.preset == "cfg2" else "4", CFG["N"], CFG["d_model"],
                                                           CFG["h"], CFG["d_ff"], B, SHAPE["Lv"][0], SHAPE["Lv"][1],
                                                           args.tgt_len),
            "global_batch": B * args.gpus, "tgt_len": args.tgt_len, "parallelism": "dp%d (independent dialogue "
            "batches per GPU, no data-path collective in forward)" % args.gpus,
            "l2": "inputs rotate over %d distinct batches per GPU (> L2 working set: features alone are 138 MB "
                  "per batch)" % args.rot}


def site_roofline(dev, peak_tf):
    """BASELINE metric, second half: the fused decoder cross-attention SITE at d_model=512, h=8, seq=256, video_len=512
    (north_star; SURVEY 8d: projections 25.77 GF + core 8.59 GF = 34.36 GF at B=32) through the C-ABI site entry point
    mtn_attn_site_fwd (LayerNorm + Q projection + K/V projection of the memory + attention core + output projection +
    residual), replayed back to back in a CUDA graph; algorithmic FLOPs / CUDA-event time / measured tensor peak."""
    import ctypes as C
    from mtn_b200 import _lib
    B, Lq, Lk, d, h = 32, 256, 512, 512, 8
    g = torch.Generator(device="cpu").manual_seed(3)
    r = lambda *shape: torch.randn(*shape, generator=g).to(dev)
    x, out = r(B * Lq, d), torch.empty(B * Lq, d, device=dev)
    mem16 = r(B * Lk, d).half()
    w_q, w_kv, w_o = (r(d, d) * 0.04).half(), (r(2 * d, d) * 0.04).half(), (r(d, d) * 0.04).half()
    b_q, b_kv, b_o, ln_a, ln_b = r(d) * 0.1, r(2 * d) * 0.1, r(d) * 0.1, 1 + 0.1 * r(d), 0.1 * r(d)
    lens = torch.randint(Lk // 2, Lk + 1, (B,), generator=g)
    mask = (torch.arange(Lk)[None, :] < lens[:, None]).view(B, 1, Lk).to(dev)
    bits = _lib.mask_pack(mask)
    a = _lib.AttnSiteArgs()
    a.B, a.Lq, a.Lk, a.d, a.h = B, Lq, Lk, d, h
    a.x, a.x_out = x.data_ptr(), out.data_ptr()
    a.ln_a, a.ln_b, a.ln_eps = ln_a.data_ptr(), ln_b.data_ptr(), 1e-6
    a.w_q, a.b_q, a.w_kv, a.b_kv, a.w_o, a.b_o = (w_q.data_ptr(), b_q.data_ptr(), w_kv.data_ptr(), b_kv.data_ptr(),
                                                  w_o.data_ptr(), b_o.data_ptr())
    a.mem_f16 = mem16.data_ptr()
    a.mask_bits, a.mask_rows_q = bits.data_ptr(), 1
    n = _lib.lib().mtn_attn_site_workspace_bytes(B, Lq, Lk, d)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), n
    call = lambda: _lib.check(_lib.lib().mtn_attn_site_fwd(C.byref(a), _lib.stream_ptr()))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        call(); call()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    reps = 20
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            call()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    proj, core = 4 * B * d * d * (Lq + Lk), 4 * B * Lq * Lk * d
    tf = (proj + core) / (us * 1e-6) / 1e12
    return {"workload": "mtn_attn_site_fwd: B=32 Lq=256 Lk=512 d=512 h=8, key-padding mask, memory K/V projected in the call "
                        "(3 kernels: LayerNorm || K/V GEMM of the memory on an internal side stream, then ONE fused kernel: "
                        "Q projection + attention + out-projection + residual, csrc/site_fused.cu)",
            "gflop": (proj + core) / 1e9, "us_per_site": us, "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": tf / peak_tf, "bound": "tensor",
            "note": "north_star target 0.70; round 1: five dependent launches, 79.1 us (0.31).  Half of the FLOPs are the "
                    "K/V projection (a plain GEMM near the GEMM ceiling); the fused kernel's attention phase is bound by the "
                    "per-row softmax latency chain (two engines per CTA), see DESIGN.md section 4"}


def bind_to_gpu_numa_node(local):
    """Multi-GPU runs: pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) BEFORE the
    pinned staging buffers are allocated, so that they are first-touched on the GPU's own NUMA node -- with all ranks
    uploading features every step, cross-socket traffic is what caps the end-to-end leg.  Best effort: returns the
    number of CPUs bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after or (len(after) < 4 and len(before) >= 4):     # never starve the rank
            os.sched_setaffinity(0, before)
            return None
        return len(after)
    except Exception:
        return None


def train_leg(args, model, devb, ntok, host, dev, rank, world, barrier, max_over_ranks, sum_over_ranks):
    """Tokens/s of one TRAINING step on the same workload: forward (train.py:33), label-smoothed loss on the decoder
    and both auto-encoder streams normalised by the global token counts (train.py:37-39), backward through the
    hand-written kernels, one NCCL all-reduce of the flat gradient buffer (N > 1), fused Adam.  Dropout p = 0.1 (the
    make_model default) runs inside the kernels."""
    import torch.distributed as dist
    from mtn_b200 import _lib
    from mtn_b200.trainer import TrainStep
    torch.cuda.empty_cache()
    res = {"workload": "train step = forward + label-smoothed loss (decoder + 2 auto-encoder streams) + backward + "
                       "%s + fused Adam; dropout p=0.1 inside the kernels (Philox, regenerated by the backward)"
                       % ("one NCCL all-reduce of the flat f32 gradient (%d ranks)" % world if world > 1 else "no collective (1 GPU)")}
    try:
        ts = TrainStep(model, CFG["vocab"], graph=not args.train_eager)
        n_params = sum(p.numel() for p in model.parameters())
        # loss normalisers: counted ON THE DEVICE from the batch in flight (global over the ranks), per step / replay
        if args.train_eager:
            run = lambda i: ts.eager(devb[i % args.rot])
        else:
            ts.capture(devb[0])
            run = lambda i: ts.replay(devb[i % args.rot])
        # launches of one step (eager trace, untimed)
        _lib.RECORD = []
        ts.eager(devb[0])
        torch.cuda.synchronize()
        rec, _lib.RECORD = _lib.RECORD, None
        per = {}
        for r in rec:
            per[r[0]] = per.get(r[0], 0) + 1
        # per-kernel-type time: the step's launches of one type replayed back to back in a CUDA graph
        kt = {}
        for name in sorted(per):
            mine = [r for r in rec if r[0] == name]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for r in mine:
                    r[3]()
            g.replay(); torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(3):
                g.replay()
            ev1.record(); torch.cuda.synchronize()
            msk = ev0.elapsed_time(ev1) / 3
            fl = sum(r[1] for r in mine)
            kt[name] = {"launches": len(mine), "ms": round(msk, 4), "tflops": round(fl / (msk * 1e-3) / 1e12, 1),
                        "gbs": round(sum(r[2] for r in mine) / (msk * 1e-3) / 1e9, 1)}
            del g
        res["kernel_breakdown_one_step"] = kt
        gemm_ms = sum(kt[k]["ms"] for k in ("linear", "linear_dgrad", "linear_wgrad") if k in kt)
        gemm_fl = sum(r[1] for r in rec if r[0] in ("linear", "linear_dgrad", "linear_wgrad"))
        pk = 1400.0
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", pk)
        except Exception:
            pass
        res["roofline"] = {"bound": "tensor", "kernel": "gemm_f16_tc_kernel, all forward / dgrad / wgrad launches of one training step",
                           "achieved": gemm_fl / (gemm_ms * 1e-3) / 1e12, "peak": pk, "unit": "TFLOP/s",
                           "frac": gemm_fl / (gemm_ms * 1e-3) / 1e12 / pk, "launches": sum(kt[k]["launches"] for k in
                                                                                          ("linear", "linear_dgrad", "linear_wgrad") if k in kt)}
        res["roofline"]["traffic"], res["roofline"]["traffic_source"] = \
            traffic_from_profiles(res["roofline"]["launches"], train=True)
        n_launches = len(rec)
        del rec
        for i in range(3):
            run(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.train_steps):
            loss = run(i)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.train_steps
        tokens = sum_over_ranks(sum(ntok[i % args.rot] for i in range(args.train_steps)) / args.train_steps)
        res.update({"tokens_per_s": tokens / (ms * 1e-3), "ms_per_step": ms, "steps": args.train_steps,
                    "mode": "eager launches" if args.train_eager else "CUDA graph replay (one launch per step)",
                    "launches_per_step": n_launches, "launches_by_kernel": per, "params": n_params,
                    "allreduce_bytes_per_step": (0 if world == 1 else (ts.flat.numel() - ts.n_dec) * 4 if ts.nvls is not None
                                                 else ts.flat.numel() * 4),
                    "grad_reduction": ("none (1 GPU)" if world == 1 else
                                       "NVLS multimem.red fused into the weight-gradient / LayerNorm / bias epilogues "
                                       "(%d MB decoder segment) + NCCL all-reduce of the rest" % (ts.n_dec * 4 >> 20)
                                       if ts.nvls is not None else "one NCCL all-reduce of the flat gradient buffer"),
                    "loss": float(loss), "model_tflops": 3 * flops_forward(args.batch, args.tgt_len) * world / (ms * 1e-3) / 1e12,
                    "normaliser_note": "ntokens / ntokens_query are counted on the device inside the captured step and "
                                       "summed over the ranks (one 2-element all-reduce); `loss` is summed over the ranks",
                    "normalisers_last_step": [float(x) for x in ts.norms] if ts.norms is not None else None})
        # how much of the gradient exchange is exposed: the same captured step WITHOUT its collectives (measurement only)
        if world > 1 and not args.train_eager:
            try:
                ts.allreduce = False
                ts.capture(devb[0])
                for i in range(3):
                    ts.replay(devb[i % args.rot])
                barrier()
                e0.record()
                for i in range(args.train_steps):
                    ts.replay(devb[i % args.rot])
                e1.record()
                barrier()
                ms0 = max_over_ranks(e0.elapsed_time(e1)) / args.train_steps
                res["ms_per_step_without_allreduce"] = ms0
                res["allreduce_exposed_ms"] = ms - ms0
                res["allreduce_overlap"] = ("prefix of %d MB (target path's modules) all-reduced on a communication stream during "
                                            "the auto-encoder chains' backward, then the remaining %d MB"
                                            % (ts.n_prefix * 4 >> 20, (ts.flat.numel() - ts.n_prefix) * 4 >> 20)) \
                    if ts.n_prefix else "none (one all-reduce after the backward)"
            except Exception as e:
                res["allreduce_exposed_error"] = repr(e)[:300]
        del ts
    except Exception as e:           # the forward headline must survive a failing auxiliary leg
        import traceback
        traceback.print_exc()
        res["error"] = repr(e)[:400]
    for p in model.parameters():
        p.grad = None
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tgt-len", type=int, default=256)
    ap.add_argument("--batch", type=int, default=SHAPE["B"])
    ap.add_argument("--rot", type=int, default=4, help="distinct input batches rotated through")
    ap.add_argument("--cpu-batch", type=int, default=4, help="dialogues in the CPU-baseline sample")
    ap.add_argument("--preset", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 = BASELINE configs[1] (the metric's config, default); cfg5 = configs[4] stress config "
                         "(N=12 d=1024 h=16 d_ff=4096, video_len=1024, batch 4/GPU) for roofline captures")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the auxiliary greedy-decode (configs[3]) leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--train-steps", type=int, default=30)
    ap.add_argument("--train-eager", action="store_true", help="time the training step without CUDA-graph capture")
    ap.add_argument("--decode-batch", type=int, default=64)
    ap.add_argument("--decode-in-flight", type=int, default=4,
                    help="dialogue batches decoded concurrently (own stream + graphs each) in the decode leg's second measurement")
    ap.add_argument("--decode-len", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.preset == "cfg5":
        CFG.update({"N": 12, "d_model": 1024, "d_ff": 4096, "h": 16})
        SHAPE.update({"B": 4, "Lv": [1024, 256]})
        if args.batch == 32:
            args.batch = 4
        global METRIC
        METRIC = "decoder tokens/sec at d_model=1024 h=16 L=12 (forward, stress config)"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("MTN_B200_NO_NUMA_BIND") != "1" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mtn_b200 import _lib, mtn
    from mtn_b200.graph import GraphedForward
    _lib.lib()          # fail loudly if the CUDA library is missing
    O = oracle()        # synthetic-input generator + weight generator (host side, untimed)

    B, T = args.batch, args.tgt_len
    torch.manual_seed(7)
    model = mtn.make_model(CFG["vocab"], CFG["vocab"], N=CFG["N"], d_model=CFG["d_model"], d_ff=CFG["d_ff"],
                           h=CFG["h"], ft_sizes=CFG["ft_sizes"], diff_encoder=True,
                           auto_encoder_ft="query").to(dev).eval()
    # distinct batches per rank (independent dialogue shards) and per rotation slot
    host = [synth(O, B, T, 1000 + 100 * rank + r) for r in range(args.rot)]
    pin = lambda t: t.pin_memory()
    host = [{k: (pin(v) if torch.is_tensor(v) else [pin(f) for f in v]) for k, v in h.items()} for h in host]
    devb = [{k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v]) for k, v in h.items()} for h in host]
    ntok = [int((h["trg_y"] != 1).sum()) for h in host]
    graphs = [GraphedForward(model, d) for d in devb]       # static buffers already hold batch r
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ------------------------------------------------------------- device-resident throughput
    for i in range(args.warmup):
        graphs[i % args.rot].replay()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # EXACTLY K steps per timed region; a region shorter than ~100 ms (K = 20 is 54 ms) is repeated -- the same K steps,
    # each repeat bracketed by barrier + synchronize -- and the reported time is the mean over the repeats.
    region_ms, repeats = [], 0
    while True:
        e0.record()
        for i in range(args.steps):
            graphs[i % args.rot].replay()
        e1.record()
        barrier()
        region_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        repeats += 1
        if sum(region_ms) >= 100.0 or repeats >= 50:
            break
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(region_ms) / repeats
    tokens = sum_over_ranks(sum(ntok[i % args.rot] for i in range(args.steps)))
    value = tokens / (ms * 1e-3)

    # Context: the same K steps with TWO independent batches in flight (graphs of even / odd rotation slots on two
    # streams).  The headline stays the one-batch-at-a-time number above; this shows how much of the step is
    # dependent-launch latency that a second batch can hide.
    in_flight = None
    rot2 = args.rot - args.rot % 2
    if rot2 >= 2 and world == 1:      # context only; single-process so that a failure cannot desynchronise ranks
        try:
            sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
            cur = torch.cuda.current_stream()

            def conc(n):
                for st in (sa, sb):
                    st.wait_stream(cur)
                for i in range(n):
                    with torch.cuda.stream(sa if i % 2 == 0 else sb):
                        graphs[i % rot2].replay()
                for st in (sa, sb):
                    cur.wait_stream(st)

            conc(4)
            barrier()
            e0.record()
            conc(args.steps)
            e1.record()
            barrier()
            ms2 = max_over_ranks(e0.elapsed_time(e1))
            tok2 = sum_over_ranks(sum(ntok[i % rot2] for i in range(args.steps)))
            in_flight = {"batches_in_flight": 2, "value": tok2 / (ms2 * 1e-3), "unit": "tokens/s",
                         "ms_per_step_amortised": ms2 / args.steps}
        except Exception as e:
            in_flight = {"error": repr(e)[:300]}

    # ------------------------------------------------------------- end to end (host buffers)
    # Public API with HOST inputs, every step: pinned-host -> device copy of ids + features, the
    # forward, device -> host copy of the decoder output.  Two input/graph slots are pipelined over
    # three streams (copy-in, compute, copy-out) so the PCIe transfer of step i+1 overlaps the
    # forward of step i; all copies and forwards of the K steps are inside the timed region.
    h2d = sum(v.numel() * v.element_size() if torch.is_tensor(v) else sum(f.numel() * f.element_size() for f in v)
              for v in host[0].values())
    slots = graphs[:2] if len(graphs) >= 2 else [graphs[0], graphs[0]]
    outs_of = lambda g: [g.out] + list(g.ae)         # forward() returns the decoder output AND both auto-encoder outputs
    out_host = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs_of(g)] for g in slots]
    d2h = sum(t.numel() * t.element_size() for t in out_host[0])
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in slots]      # inputs of slot landed
    ev_cmp = [torch.cuda.Event() for _ in slots]     # forward of slot finished
    ev_out = [torch.cuda.Event() for _ in slots]     # output of slot copied out (slot reusable)

    def e2e_steps(n, slots=slots, host=host, out_host=out_host):
        for i in range(n):
            k = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_cmp[k])                # previous forward on this slot has consumed its inputs
                slots[k].copy_inputs(host[i % args.rot])
                ev_in[k].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[k])
                s_cmp.wait_event(ev_out[k])               # previous output of this slot has left
                slots[k].replay()
                ev_cmp[k].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[k])
                for dst, src in zip(out_host[k], outs_of(slots[k])):
                    dst.copy_(src, non_blocking=True)
                ev_out[k].record(s_out)

    def time_e2e(**kw):
        e2e_steps(4, **kw)
        barrier()
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for st in (s_in, s_cmp, s_out):
            st.wait_stream(cur)
        e2e_steps(args.steps, **kw)
        for st in (s_in, s_cmp, s_out):
            cur.wait_stream(st)
        e1.record(cur)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # Reference data format first: f32 features in pinned host memory (what the reference's loader hands to Batch).
    ms_e2e32 = time_e2e()
    e2e32 = {"value": tokens / (ms_e2e32 * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e32 / args.steps,
             "h2d_bytes_per_step": h2d, "h2d_gbs_per_gpu": h2d / (ms_e2e32 / args.steps * 1e-3) / 1e9,
             "note": "features uploaded as f32 every step (the reference's storage format): PCIe-bound"}
    # HEADLINE e2e: the features are STORED as f16 in pinned host memory -- a one-time dataset-preparation choice of the
    # loader (the kernels round features to f16 on arrival anyway, so the outputs are bit-identical to the f32 upload);
    # every step still uploads its ids + features and downloads all three outputs inside the timed region.
    host16 = [{k: (v if torch.is_tensor(v) else [pin(f.half()) for f in v]) for k, v in h.items()} for h in host]
    d16 = {k: (v if torch.is_tensor(v) else [f.half() for f in v]) for k, v in devb[0].items()}
    slots16 = [GraphedForward(model, d16), GraphedForward(model, d16)]
    oh16 = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs_of(g)] for g in slots16]
    ms_e2e16 = time_e2e(slots=slots16, host=host16, out_host=oh16)
    torch.cuda.synchronize()
    last = (args.steps - 1) % 2                      # both runs end with the same batch in slot (steps-1) % 2
    e2e_identical = bool(torch.equal(slots16[last].out, slots[last].out))
    h2d16 = sum(v.numel() * v.element_size() if torch.is_tensor(v) else sum(f.numel() * f.element_size() for f in v)
                for v in host16[0].values())
    e2e16 = {"value": tokens / (ms_e2e16 * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e16 / args.steps,
             "h2d_bytes_per_step": h2d16, "h2d_gbs_aggregate": h2d16 * world / (ms_e2e16 / args.steps * 1e-3) / 1e9,
             "outputs_bit_identical_to_f32_upload": e2e_identical,
             "note": "every sample's features uploaded every step, stored as f16 on the host"}
    # HEADLINE e2e: the ten turns of a dialogue share one video (data_handler.py:150-206 builds one sample per turn), so a
    # video's features cross PCIe ONCE and stay in a device-resident cache (mtn_b200/feature_cache.py: 2.2 MB per video in
    # f16, the 180 GB of HBM hold the whole AVSD feature set); every step uploads the token ids of all its samples, the
    # slot indices, and the features of the videos that are NEW -- one tenth of the batch -- assembles the batch on the
    # device with one gather per modality, and downloads all three outputs.  Everything inside the timed region.
    from mtn_b200.feature_cache import DeviceFeatureCache
    shapes = [(f.shape[1], f.shape[2]) for f in host16[0]["fts"]]
    cache = DeviceFeatureCache(args.rot * B + 8, shapes, dev)
    for r in range(args.rot):                        # (a first epoch's uploads: untimed warm state)
        for j in range(B):
            cache.put(r * B + j, [f[j] for f in host16[r]["fts"]])
    n_new = max(1, (B + 9) // 10)
    idx_ring = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(8)]
    id_keys = [k for k, v in host16[0].items() if torch.is_tensor(v)]
    h2d_c = sum(host16[0][k].numel() * host16[0][k].element_size() for k in id_keys) + n_new * cache.bytes_per_video() + B * 8

    def e2e_cached(n):
        for i in range(n):
            k, r = i % 2, i % args.rot
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_cmp[k])
                for key in id_keys:
                    slots16[k].static[key].copy_(host16[r][key], non_blocking=True)
                for t in range(n_new):               # this step's new videos (rotating through the batch)
                    j = (i * n_new + t) % B
                    cache.put(r * B + j, [f[j] for f in host16[r]["fts"]])
                cache.gather([r * B + j for j in range(B)], out=slots16[k].static["fts"], index_buffer=idx_ring[i % 8])
                ev_in[k].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[k])
                s_cmp.wait_event(ev_out[k])
                slots16[k].replay()
                ev_cmp[k].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[k])
                for dst, src in zip(oh16[k], outs_of(slots16[k])):
                    dst.copy_(src, non_blocking=True)
                ev_out[k].record(s_out)

    e2e_cached(4)
    barrier()
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for st_ in (s_in, s_cmp, s_out):
        st_.wait_stream(cur)
    e2e_cached(args.steps)
    for st_ in (s_in, s_cmp, s_out):
        cur.wait_stream(st_)
    e1.record(cur)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    e2e_cache_identical = bool(torch.equal(slots16[last].out, slots[last].out))
    e2e = tokens / (ms_e2e * 1e-3)
    del slots16, oh16, cache

    # ------------------------------------------------------------- traced step: launches + roofline
    # One eager step with recording on: counts our kernel launches and keeps a re-launch closure per
    # launch.  Each kernel type is then replayed back to back inside its own CUDA graph (no host gaps,
    # same operands, same stream) and timed with CUDA events: that is the per-kernel duration the
    # roofline uses.  (Inside the real step kernels of the three streams overlap, so per-kernel time
    # cannot be read off the step time.)
    from mtn_b200.data_utils import Batch
    _lib.RECORD = []
    with torch.no_grad():
        d0 = devb[0]
        bt = Batch(d0["query"], d0["his"], None, [f.permute(1, 0, 2) for f in d0["fts"]], d0["cap"], d0["trg"],
                   d0["trg_y"], 1)
        keep = model.forward(bt)
    torch.cuda.synchronize()
    rec, _lib.RECORD = _lib.RECORD, None
    launches = len(rec)
    per = {}
    for name in sorted(set(r[0] for r in rec)):
        mine = [r for r in rec if r[0] == name]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for r in mine:
                r[3]()
        g.replay(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        ev0.record()
        for _ in range(reps):
            g.replay()
        ev1.record(); torch.cuda.synchronize()
        per[name] = {"n": len(mine), "ms": ev0.elapsed_time(ev1) / reps, "flops": sum(r[1] for r in mine),
                     "bytes": sum(r[2] for r in mine)}
    del keep
    lin = per.get("linear", {"n": 1, "ms": 1e-9, "flops": 0})
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)     # kernel timed inside a long step -> sustained figure
    ach = lin["flops"] / (lin["ms"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_f16_tc_kernel (tcgen05 linear; all %d launches of one step)" % lin["n"],
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (f16 and bf16 share the tensor rate)"
                if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                "traffic": traffic_from_profiles(lin["n"])[0], "traffic_source": traffic_from_profiles(lin["n"])[1],
                "flops_per_launch_avg": lin["flops"] / lin["n"], "us_per_launch_avg": lin["ms"] * 1e3 / lin["n"],
                "note": "algorithmic 2MNK of the step's linear launches / CUDA-event time of those launches "
                        "replayed back to back in one CUDA graph"}
    site = None
    if rank == 0:
        try:
            site = site_roofline(dev, peak_tf)
        except Exception as e:
            site = {"error": repr(e)[:300]}
    hbm = peaks.get("hbm_gbs", 6650.0)
    breakdown = {k: {"launches": v["n"], "ms": round(v["ms"], 4), "gflop": round(v["flops"] / 1e9, 2),
                     "mb": round(v["bytes"] / 1e6, 1),
                     "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1),
                     "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} for k, v in per.items()}
    ln = per.get("layernorm")
    if ln:
        breakdown["layernorm"]["hbm_frac"] = round(ln["bytes"] / (ln["ms"] * 1e-3) / 1e9 / hbm, 3)

    # ------------------------------------------------------------- auxiliary: greedy decode (configs[3])
    # batch 64 dialogues, 10-turn history (H=256), target length 20, CUDA-graph decoder; timed end to end
    # per dialogue batch: pinned-host inputs -> device, prefill (encode + memory stage) + 19 steps, tokens -> host.
    decode = None
    if not args.no_decode:
        from mtn_b200.graph import GraphedGreedyDecoder
        del graphs, slots
        torch.cuda.empty_cache()
        Bd, Ld = args.decode_batch, args.decode_len
        dh = [O.synth_inputs(CFG, B=Bd, Q=SHAPE["Q"], C=SHAPE["C"], H=SHAPE["H"], T=4, Lv=SHAPE["Lv"],
                             seed=5000 + 10 * rank + r) for r in range(2)]
        # features stored as f16 on the host (bit-identical results, see the e2e leg); a video's features cross PCIe once
        # per dialogue: generate.py decodes the ten turns of a dialogue one after the other, so per batch of 64 samples
        # ceil(64 / 10) videos are new and the rest are gathered from the device feature cache
        dh = [{k: (pin(v) if torch.is_tensor(v) else [pin(f.half()) for f in v]) for k, v in h.items()
               if k in ("query", "his", "cap", "fts")} for h in dh]
        dec = GraphedGreedyDecoder(model, {k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v])
                                           for k, v in dh[0].items()}, Ld)
        from mtn_b200.feature_cache import DeviceFeatureCache
        dcache = DeviceFeatureCache(2 * Bd + 8, [(f.shape[1], f.shape[2]) for f in dh[0]["fts"]], dev)
        for r in range(2):
            for j in range(Bd):
                dcache.put(r * Bd + j, [f[j] for f in dh[r]["fts"]])
        n_new_d = (Bd + 9) // 10
        d_ring = [torch.empty(Bd, dtype=torch.int64).pin_memory() for _ in range(64)]
        up_count = [0]

        def upload_next(dcd, r):
            i = up_count[0]
            up_count[0] += 1
            new = [(r * Bd + (i * n_new_d + t) % Bd, [f[(i * n_new_d + t) % Bd] for f in dh[r]["fts"]]) for t in range(n_new_d)]
            dcd.upload(dh[r], cache=dcache, video_ids=[r * Bd + j for j in range(Bd)], new_videos=new,
                       index_buffer=d_ring[i % len(d_ring)])

        toks_host = torch.empty(Bd, Ld, dtype=torch.int64).pin_memory()
        reps = 5
        # every batch's inputs come from pinned host memory inside the timed region; the upload of batch i+1 runs on a
        # copy stream while batch i decodes (staging buffers + one device-to-device copy)
        upload_next(dec, 0)
        for i in range(2):
            toks = dec.decode(staged=True); upload_next(dec, (i + 1) % 2); toks_host.copy_(toks, non_blocking=True)
        barrier()
        e0.record()
        for i in range(reps):
            toks = dec.decode(staged=True)
            upload_next(dec, (i + 1) % 2)
            toks_host.copy_(toks, non_blocking=True)
        e1.record()
        barrier()
        ms_dec = max_over_ranks(e0.elapsed_time(e1)) / reps
        decode = {"workload": "BASELINE configs[3]: greedy decode, batch %d/GPU, 10-turn history (H=256), target len %d, "
                              "N=6 d=512; memory stage once per batch, %d graph-replayed KV-cached steps (only the new "
                              "position is computed)" % (Bd, Ld, Ld - 1),
                  "generated_tokens_per_s": sum_over_ranks(Bd * (Ld - 1)) / (ms_dec * 1e-3), "ms_per_batch": ms_dec,
                  "includes": "H2D of the ids of all samples + the f16 features of the batch's NEW videos (%d of %d: ten turns per "
                              "dialogue share a video, device feature cache) pipelined with the previous batch's decoding, "
                              "device-side gather, encode, memory stage, all steps, D2H of tokens" % (n_new_d, Bd)}
        # HBM roofline of the cached steps (decode is HBM-bound): per step the target path reads its f16 weights
        # (22 d^2 per layer + generator) and the cached K/V of every memory it attends (his, cap, query, 2 x ae: f16
        # [B, L, 2d] per layer) plus the self-attention cache so far.
        try:
            d_, N_ = CFG["d_model"], CFG["N"]
            w_bytes = N_ * 22 * d_ * d_ * 2 + d_ * CFG["vocab"] * 2
            kv_mem = N_ * Bd * (SHAPE["H"] + SHAPE["C"] + SHAPE["Q"] + 2 * SHAPE["Q"]) * 2 * d_ * 2
            kv_self = sum(N_ * Bd * (t + 1) * 2 * d_ * 2 for t in range(1, Ld - 1)) / max(1, Ld - 2)
            step_bytes = w_bytes + kv_mem + kv_self
            for g_ in dec.graphs:
                g_.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                for g_ in dec.graphs[1:]:
                    g_.replay()
            e1.record(); torch.cuda.synchronize()
            us_step = e0.elapsed_time(e1) * 1e3 / (3 * sum(dec.steps_in_graph[1:]))
            hbm_pk = peaks.get("hbm_gbs", 6650.0)
            # launches of one cached step (eager, untimed): our kernels by name
            try:
                _lib.RECORD = []
                with torch.no_grad():
                    model.decode_step(dec.state, dec.ys[:, 3], 3)
                torch.cuda.synchronize()
                rec_d, _lib.RECORD = _lib.RECORD, None
                by = {}
                for r_ in rec_d:
                    by[r_[0]] = by.get(r_[0], 0) + 1
                decode["launches_per_step"] = len(rec_d)
                decode["launches_by_kernel"] = by
                del rec_d
            except Exception as e:
                _lib.RECORD = None
                decode["launches_per_step_error"] = repr(e)[:200]
            decode["roofline"] = {"bound": "hbm", "achieved": step_bytes / (us_step * 1e-6) / 1e9, "peak": hbm_pk, "unit": "GB/s",
                                  "frac": step_bytes / (us_step * 1e-6) / 1e9 / hbm_pk, "us_per_step": us_step,
                                  "bytes_per_step": step_bytes,
                                  "note": "algorithmic bytes of one cached step (f16 weights %.0f MB + memory K/V %.0f MB + "
                                          "self cache %.1f MB) / CUDA-event time of the step graph; %s"
                                          % (w_bytes / 1e6, kv_mem / 1e6, kv_self / 1e6,
                                             "the target path of a step is ONE kernel (csrc/decode_cluster.cu): a cluster of 8 CTAs "
                                             "per dialogue group, one head per CTA, operands streamed ahead of the per-dialogue "
                                             "dependency chain; bound by that chain's latency (LayerNorm -> projection -> attention "
                                             "-> exchange per sublayer), not by bandwidth"
                                             if "decode_cluster" in decode.get("launches_by_kernel", {}) else
                                             "the step is a chain of ~130 dependent few-row launches (csrc/decode_rows.cu), each one "
                                             "memory round trip deep: latency bound, not bandwidth bound")}
        except Exception as e:
            decode["roofline"] = {"error": repr(e)[:300]}
        # Several dialogue batches in flight: a decoding step is a chain of small dependent kernels (<= 40 CTAs each on
        # 148 SMs), so independent batches on their own streams fill the idle SMs -- same per-batch work, same tokens.
        if args.decode_in_flight > 1 and world == 1:
            try:
                K = args.decode_in_flight
                dev0 = {k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v]) for k, v in dh[0].items()}
                decs = [dec] + [GraphedGreedyDecoder(model, dev0, Ld) for _ in range(K - 1)]
                streams = [torch.cuda.Stream() for _ in range(K)]
                outs = [torch.empty(Bd, Ld, dtype=torch.int64).pin_memory() for _ in range(K)]
                main = torch.cuda.current_stream()

                def one(i):          # dialogue batch i on decoder / stream i % K (its round i // K decodes dh[round % 2])
                    j = i % K
                    with torch.cuda.stream(streams[j]):
                        t = decs[j].decode(staged=True)
                        upload_next(decs[j], (i // K + 1) % 2)
                        outs[j].copy_(t, non_blocking=True)

                for st_ in streams:
                    st_.wait_stream(main)
                for j in range(K):
                    with torch.cuda.stream(streams[j]):
                        upload_next(decs[j], 0)
                for i in range(2 * K):
                    one(i)
                torch.cuda.synchronize()
                # every decoder decoded dh[1] last, with the other batches in flight: compare with the same decoder
                # decoding dh[1] ALONE (tools/concurrency_check.py does the same for the forward graphs)
                same = True
                for j in range(K):
                    with torch.cuda.stream(streams[j]):
                        upload_next(decs[j], 1)
                        solo = decs[j].decode(staged=True).clone()
                        upload_next(decs[j], 0)
                    torch.cuda.synchronize()
                    same = same and bool((solo.cpu() == outs[j]).all())
                barrier()
                e0.record()
                for st_ in streams:
                    st_.wait_stream(main)
                nb = reps * K
                for i in range(nb):
                    one(i)
                for st_ in streams:
                    main.wait_stream(st_)
                e1.record()
                barrier()
                ms_k = max_over_ranks(e0.elapsed_time(e1)) / nb
                decode["in_flight"] = {"batches_in_flight": K, "generated_tokens_per_s": sum_over_ranks(Bd * (Ld - 1)) / (ms_k * 1e-3),
                                       "ms_per_batch_amortised": ms_k, "tokens_equal_solo_decoding": same,
                                       "note": "K independent dialogue batches of %d, each on its own stream with its own "
                                               "CUDA graphs and staging buffers; H2D / D2H inside the timed region" % Bd}
                del decs
            except Exception as e:
                decode["in_flight"] = {"error": repr(e)[:300]}
        del dec
        # ---- beam search (generate.py's actual path, data_utils.py:188-242): all hypotheses of all dialogues in one
        # KV-cached step and one D2H per position, next to the reference's call form (one full-prefix decode and one
        # D2H per hypothesis per step) on the same dialogues.  Context numbers, rank 0 of a 1-GPU run only.
        if world == 1:
            try:
                from mtn_b200 import data_utils as du
                Db, beam_w, blen = 8, 5, Ld
                dvb = {k: (v[:Db].to(dev) if torch.is_tensor(v) else [f[:Db].to(dev) for f in v]) for k, v in dh[0].items()}
                mkb = lambda sl: du.Batch(dvb["query"][sl], dvb["his"][sl], None, [f[sl].permute(1, 0, 2) for f in dvb["fts"]],
                                          dvb["cap"][sl], None, None, 1)
                with torch.no_grad():
                    du.beam_search_decode_batched(model, mkb(slice(0, Db)), blen, 2, 0, 3, 1, beam=beam_w)      # warm-up
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    got = du.beam_search_decode_batched(model, mkb(slice(0, Db)), blen, 2, 0, 3, 1, beam=beam_w)
                    torch.cuda.synchronize(); t_b = time.perf_counter() - t0
                    os.environ["MTN_B200_BEAM_SERIAL"] = "1"
                    du.beam_search_decode(model, mkb(slice(0, 1)), blen, 2, 0, 3, 1, beam=beam_w)               # warm-up
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    ref1 = du.beam_search_decode(model, mkb(slice(0, 1)), blen, 2, 0, 3, 1, beam=beam_w)
                    torch.cuda.synchronize(); t_s = time.perf_counter() - t0
                    os.environ.pop("MTN_B200_BEAM_SERIAL", None)
                decode["beam_search"] = {
                    "workload": "beam %d, %d positions, %d dialogues at once (eager launches, host-side pool rules)" % (beam_w, blen, Db),
                    "batched_ms_per_dialogue": t_b * 1e3 / Db, "serial_reference_call_form_ms_per_dialogue": t_s * 1e3,
                    "speedup": t_s / (t_b / Db),
                    "same_best_hypothesis_as_serial": [int(t) for t in got[0][0][0][0]] == [int(t) for t in ref1[0][0][0]]}
            except Exception as e:
                os.environ.pop("MTN_B200_BEAM_SERIAL", None)
                decode["beam_search"] = {"error": repr(e)[:300]}

        # ---- the cluster decoding step at the row capacity of the device (context, rank 0 of a 1-GPU run): a cluster
        # takes up to 8 dialogues; 75 % of a batch-64 step is the fixed per-cluster chain, so more rows per launch
        # amortise it
        if world == 1:
            try:
                cap_b = max(b_ for b_ in range(64, 129, 8) if _lib.decode_cluster_supported(b_, CFG["d_model"], CFG["h"], CFG["d_ff"], 7 * CFG["N"]))
                hb = O.synth_inputs(CFG, B=cap_b, Q=SHAPE["Q"], C=SHAPE["C"], H=SHAPE["H"], T=4, Lv=SHAPE["Lv"], seed=5100)
                decb = GraphedGreedyDecoder(model, {k: (v.to(dev) if torch.is_tensor(v) else [f.half().to(dev) for f in v])
                                                    for k, v in hb.items() if k in ("query", "his", "cap", "fts")}, Ld)
                decb.decode(); torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    decb.decode()
                e1.record(); torch.cuda.synchronize()
                ms_b = e0.elapsed_time(e1) / 3
                e0.record()
                for _ in range(3):
                    for g_ in decb.graphs[1:]:
                        g_.replay()
                e1.record(); torch.cuda.synchronize()
                decode["at_row_capacity"] = {"batch": cap_b, "generated_tokens_per_s": cap_b * (Ld - 1) / (ms_b * 1e-3), "ms_per_batch": ms_b,
                                             "us_per_step": e0.elapsed_time(e1) * 1e3 / (3 * sum(decb.steps_in_graph[1:])),
                                             "note": "device-resident inputs, memory stage + all steps; the largest batch (multiple "
                                                     "of 8) whose rows fit the co-resident clusters"}
                del decb, hb
            except Exception as e:
                decode["at_row_capacity"] = {"error": repr(e)[:300]}

    # ------------------------------------------------------------- training step (forward + loss + backward +
    # ONE NCCL gradient all-reduce + Adam), BASELINE configs[1] / [2]
    train = None
    if not args.no_train:
        train = train_leg(args, model, devb, ntok, host, dev, rank, world, barrier, max_over_ranks, sum_over_ranks)
        model.eval()

    # ------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        # bounded sample: whole batch 0 (the same 32 dialogues the GPU step processes), repeated for ~10 s of CPU work
        sub = {k: (v.clone() if torch.is_tensor(v) else [f.clone() for f in v]) for k, v in host[0].items()}
        t0 = time.perf_counter()
        cpu_forward_seconds(O, sd, sub, reps=0)                      # warm-up forward, also sizes the sample
        t_one = time.perf_counter() - t0
        reps = max(3, min(10, int(10.0 / max(t_one, 1e-3))))
        sec = cpu_forward_seconds(O, sd, sub, reps=reps)
        cpu = {"value": int((sub["trg_y"] != 1).sum()) / sec, "unit": "tokens/s", "cores": torch.get_num_threads(),
               "kind": "port", "sample": "oracle forward on all %d dialogues of batch 0, median of %d (%.2f s each)"
               % (sub["query"].shape[0], reps, sec)}
        # context only: the same port (stock PyTorch eager ops, f32, what the reference would run on a GPU) on
        # this B200, full batch -- the reference ships no GPU kernels of its own.
        try:
            sdg = {k: v.to(dev) for k, v in sd.items()}
            fullg = {k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v]) for k, v in host[0].items()}
            run = lambda: O.forward(sdg, CFG, fullg["query"], fullg["his"], fullg["cap"], fullg["trg"], fullg["fts"])
            run(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            cpu["gpu_eager_port_tokens_per_s"] = ntok[0] / ((time.perf_counter() - t0) / 3)
            cpu["gpu_eager_port_note"] = "oracle port in stock PyTorch eager f32 on the same GPU, full batch (context)"
            del sdg, fullg
        except Exception as e:       # context number only
            cpu["gpu_eager_port_note"] = "failed: %r" % (e,)

    if rank == 0:
        fl = flops_forward(B, T)
        scal = lambda d, k: (d.get(k) if isinstance(d, dict) else None)
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "timed_repeats": repeats, "timed_region_ms": ms,
                # scalars lifted to the top level so that a record that keeps only flat keys still has them
                "attn_site_frac": scal(site, "frac"), "attn_site_us": scal(site, "us_per_site"),
                "decode_tokens_per_s": scal(decode, "generated_tokens_per_s"), "decode_hbm_frac": scal(scal(decode, "roofline"), "frac"),
                "train_ms_per_step": scal(train, "ms_per_step"), "train_tokens_per_s": scal(train, "tokens_per_s"),
                "vs_baseline": None, "dtype": "f16",
                "precision": "f16 tensor-core operands (11-bit significand), f32 accumulate / softmax / LayerNorm / "
                             "residual stream; 4-6e-4 normwise vs the f32 reference (bar 1e-3)",
                "data": "synthetic", "config": workload_config(args, B),
                "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": h2d_c, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "h2d_gbs_aggregate": h2d_c * world / (ms_e2e / args.steps * 1e-3) / 1e9,
                        "input_format": "every step: token ids (int64) of all samples + slot indices + the f16 features of the "
                                        "NEW videos (%d of %d: the ten turns of a dialogue share one video, which crosses PCIe "
                                        "once and stays in the device feature cache, mtn_b200/feature_cache.py); the batch is "
                                        "gathered on the device; D2H = decoder output + both auto-encoder outputs; outputs "
                                        "bit-identical to the f32 upload: %s" % (n_new, B, e2e_cache_identical),
                        "every_sample_every_step_f16": e2e16, "every_sample_every_step_f32": e2e32,
                        "host_threads_bound_to_gpu_numa_cpus": numa},
                "gpu_launches": launches * args.steps, "launches_per_step": launches,
                "roofline": roofline, "attn_site_roofline": site, "cpu_baseline": cpu, "clocks": clocks,
                "two_batches_in_flight": in_flight,
                "tokens_per_step_per_gpu": sum(ntok) / len(ntok),
                "model_tflops": fl * world / (ms / args.steps * 1e-3) / 1e12, "gflop_per_step_per_gpu": fl / 1e9,
                "kernel_breakdown_one_step": breakdown, "decode": decode, "train": train}
        print(json.dumps(line), flush=True)
    if world > 1:
        # every rank has finished its legs (the line above is out): leave without NCCL's teardown -- with captured
        # graphs that hold collectives, destroy_process_group() was seen to hang a finished run until the caller's timeout
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _watchdog(limit_s):
    """A bench run that exceeds its wall-clock budget is killed from inside (exit code 3) instead of holding the GPU
    box until the caller's timeout: MTN_B200_BENCH_LIMIT_S, default 900 s."""
    def run():
        time.sleep(limit_s)
        sys.stderr.write("bench.py: wall-clock limit of %d s exceeded -- aborting\n" % limit_s)
        sys.stderr.flush()
        os._exit(3)
    threading.Thread(target=run, daemon=True).start()


if __name__ == "__main__":
    _watchdog(int(os.environ.get("MTN_B200_BENCH_LIMIT_S", "900")))
    # The contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL prints its version
    # banner there) are pointed at stderr for the duration of the run; print() keeps the real stdout.
    sys.stdout.flush()
    _real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real, "w")
    main()
    sys.stdout.flush()
