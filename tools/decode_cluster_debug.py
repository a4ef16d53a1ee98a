"""Bisecting aid for csrc/decode_cluster.cu (one KV-cached decoding step = one kernel, one cluster per dialogue group):
runs T positions of cached decoding with the cluster kernel and with the launch sequence (few-row kernels) on the same
inputs and prints, per position, the first sublayer whose residual rows disagree (the kernel's `taps` against
engine.TAP of the launch sequence) and the error of the step's output rows; then times both forms as CUDA graphs.

    python tools/decode_cluster_debug.py [--batch 64] [--steps 20] [--N 6] [--time]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--N", type=int, default=6)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--stamps", action="store_true", help="print the in-kernel phase timeline of the last position (CTA 0)")
    a = ap.parse_args()
    import mtn_oracle as O
    from mtn_b200 import mtn, data_utils as du, engine, _lib
    _lib.lib()
    cfg = {"N": a.N, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 3000, "ft_sizes": [2048, 128],
           "auto_encoder_ft": "query", "diff_encoder": True}
    torch.manual_seed(7)
    model = mtn.make_model(3000, 3000, N=a.N, d_model=512, d_ff=2048, h=8, ft_sizes=[2048, 128], diff_encoder=True,
                           auto_encoder_ft="query").cuda().eval()
    B, T = a.batch, a.steps
    inp = O.synth_inputs(cfg, B=B, Q=64, C=64, H=256, T=T, Lv=[512, 256], seed=17 + B)
    g = lambda t: t.cuda()
    b = du.Batch(g(inp["query"]), g(inp["his"]), None, [g(f).permute(1, 0, 2).contiguous() for f in inp["fts"]], g(inp["cap"]),
                 g(inp["trg"]), g(inp["trg_y"]), 1)
    rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30))
    nsite = a.N * 7
    with torch.no_grad():
        q, vid, cap, his, ae = model.encode(b.query, b.query_mask, b.his, b.his_mask, b.cap, b.cap_mask, b.fts, b.fts_mask)
        os.environ["MTN_B200_DECODE_CLUSTER"] = "0"
        st0 = model.decode_begin(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask, ae, T)
        os.environ["MTN_B200_DECODE_CLUSTER"] = "1"
        st1 = model.decode_begin(vid, his, cap, q, b.fts_mask, b.his_mask, b.cap_mask, b.query_mask, ae, T)
        st1["cluster_taps"] = torch.zeros(nsite, B, 512, device="cuda")
        worst = 0.0
        for t in range(T):
            os.environ["MTN_B200_DECODE_CLUSTER"] = "0"
            engine.TAP = []
            r0 = model.decode_step(st0, b.trg[:, t]).clone()
            taps0 = [x for (n, x) in engine.TAP if n.endswith("x")]
            engine.TAP = None
            os.environ["MTN_B200_DECODE_CLUSTER"] = "1"
            r1 = model.decode_step(st1, b.trg[:, t]).clone()
            torch.cuda.synchronize()
            assert "cluster_plan" in st1 and st1["cluster_plan"] is not None, "cluster path not taken"
            errs = [rel(st1["cluster_taps"][s], taps0[s]) for s in range(min(nsite, len(taps0)))]
            bad = [(s, e) for s, e in enumerate(errs) if not (e < 2e-3)]
            ce = max(rel(c1[:, t], c0[:, t]) for c0, c1 in zip(st0["cache"], st1["cache"]))
            e = rel(r1, r0)
            worst = max(worst, e if e == e else 1e9)
            print("t=%2d  out rel %.2e  cache row rel %.2e  max site err %.2e  finite %s%s" %
                  (t, e, ce, max(errs) if errs else -1, bool(torch.isfinite(r1).all()),
                   ("  FIRST BAD SITE %d (layer %d, sublayer %d): %.2e" % (bad[0][0], bad[0][0] // 7, bad[0][0] % 7, bad[0][1])) if bad else ""))
        print("worst output error over %d positions: %.2e" % (T, worst))
        if a.stamps:
            st1.pop("cluster_taps")
            st1["cluster_stamps"] = torch.zeros(nsite + 8, 8, dtype=torch.int64, device="cuda")
            for _ in range(3):
                model.decode_step(st1, b.trg[:, T - 1], T - 1)
            torch.cuda.synchronize()
            sm = st1.pop("cluster_stamps").cpu()
            names = ["LN", "in-proj/w1", "attention", "merge", "barrier1", "out-proj", "barrier2"]
            print("in-kernel timeline (cycles; CTA 0 thread 0), per sublayer of layer 0 and layer %d:" % (a.N - 1))
            print("site kind   " + "  ".join("%10s" % n for n in names) + "       total")
            kinds = ["self", "his", "cap", "src", "ae0", "ae1", "ffn"]
            for s_ in list(range(7)) + list(range(nsite - 7, nsite)):
                r = sm[s_].tolist()
                nxt = sm[s_ + 1, 0].item() if s_ + 1 < nsite else r[7]
                seg = []
                last = r[0]
                for k in range(1, 8):
                    if r[k] == 0:
                        seg.append(0); continue
                    seg.append(r[k] - last); last = r[k]
                print("%3d  %-5s  " % (s_, kinds[s_ % 7]) + "  ".join("%10d" % v for v in seg) + "  %10d" % (r[7] - r[0]))
            print("whole step: %d cycles" % (sm[nsite - 1, 7].item() - sm[0, 0].item()))
            f = sm[nsite:].reshape(-1).tolist()
            if f[15]:
                t0 = f[15]
                print("site 2 in-proj detail (cycles after entering dc_proj): chunk0 A-loaded %d acquired %d W-loaded+released %d mma done %d | "
                      "chunk1 %d %d %d %d | returned %d | before cbar %d | stamp1->entry %d, cbar done %d" %
                      (f[0] - t0, f[1] - t0, f[2] - t0, f[3] - t0, f[4] - t0, f[5] - t0, f[6] - t0, f[7] - t0, f[8] - t0, f[9] - t0,
                       t0 - sm[2, 1].item(), sm[2, 2].item() - t0))
    if a.time:
        from mtn_b200.graph import GraphedGreedyDecoder
        d = {k: (v.cuda() if torch.is_tensor(v) else [f.cuda() for f in v]) for k, v in inp.items() if k in ("query", "his", "cap", "fts")}
        ys = {}
        for mode in ("0", "1"):
            os.environ["MTN_B200_DECODE_CLUSTER"] = mode
            dec = GraphedGreedyDecoder(model, d, T, cached=True)
            ys[mode] = dec.decode().clone()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                dec.decode()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            # steps only (graphs 1..): the prefill graph holds the memory stage
            ev[0].record()
            for _ in range(10):
                for gph in dec.graphs[1:]:
                    gph.replay()
            ev[1].record()
            torch.cuda.synchronize()
            us = ev[0].elapsed_time(ev[1]) / 10 / max(1, sum(dec.steps_in_graph[1:])) * 1e3
            print("cluster=%s: %.3f ms per batch of %d dialogues x %d tokens = %.1f k tokens/s; %.1f us per cached step" %
                  (mode, ms, B, T, B * T / ms, us))
        print("tokens equal: %s (%d of %d sequences differ)" % (bool(torch.equal(ys["0"], ys["1"])),
                                                               int((ys["0"] != ys["1"]).any(1).sum()), B))


if __name__ == "__main__":
    main()
