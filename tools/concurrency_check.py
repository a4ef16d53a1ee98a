"""Are two CUDA-graph instances of the hot path independent?  Runs two GraphedForward / GraphedGreedyDecoder
instances of the bench model (a) one after the other and (b) at the same time on two streams, on the SAME input,
and compares the results bit for bit:
  run-to-run   : the same graph twice                      (determinism of the kernels)
  graph-to-graph: two graphs captured from the same model   (no capture-time state leaks into the results)
  concurrent   : both graphs in flight at once vs their solo results  (no shared scratch memory between graphs)
Usage: python tools/concurrency_check.py [out_file]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mtn_b200 import _lib, mtn  # noqa: E402
from mtn_b200.graph import GraphedForward, GraphedGreedyDecoder  # noqa: E402


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    _lib.lib()
    O = bench.oracle()
    CFG, SHAPE = bench.CFG, bench.SHAPE
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    model = mtn.make_model(CFG["vocab"], CFG["vocab"], N=CFG["N"], d_model=CFG["d_model"], d_ff=CFG["d_ff"], h=CFG["h"],
                           ft_sizes=CFG["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").to(dev).eval()
    to_dev = lambda h: {k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v]) for k, v in h.items()}
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def both(fa, fb, n=1):
        cur = torch.cuda.current_stream()
        for st in (sa, sb):
            st.wait_stream(cur)
        for _ in range(n):
            with torch.cuda.stream(sa):
                fa()
            with torch.cuda.stream(sb):
                fb()
        for st in (sa, sb):
            cur.wait_stream(st)
        torch.cuda.synchronize()

    if "--decode-only" not in sys.argv:
        forward_graphs(model, O, CFG, SHAPE, to_dev, both, out)
    decoders(model, O, CFG, SHAPE, to_dev, both, out)
    out.flush()


def forward_graphs(model, O, CFG, SHAPE, to_dev, both, out):
    # ---------------------------------------------------------------- forward graphs (B=32, T=256)
    inp = to_dev(O.synth_inputs(CFG, B=32, Q=SHAPE["Q"], C=SHAPE["C"], H=SHAPE["H"], T=256, Lv=SHAPE["Lv"], seed=1000))
    g0, g1 = GraphedForward(model, inp), GraphedForward(model, inp)
    g0.replay(); torch.cuda.synchronize(); a = g0.out.clone()
    g0.replay(); torch.cuda.synchronize(); a2 = g0.out.clone()
    g1.replay(); torch.cuda.synchronize(); b = g1.out.clone()
    print("forward  run-to-run    identical: %s" % torch.equal(a, a2), file=out)
    print("forward  graph-to-graph identical: %s (max abs diff %.3e)" % (torch.equal(a, b), float((a - b).abs().max())), file=out)
    bad = 0
    for _ in range(5):
        both(g0.replay, g1.replay, n=3)
        bad += int(not torch.equal(g0.out, a)) + int(not torch.equal(g1.out, b))
    print("forward  concurrent     mismatching results: %d of 10 (max abs diff %.3e / %.3e)" %
          (bad, float((g0.out - a).abs().max()), float((g1.out - b).abs().max())), file=out)
    del g0, g1
    torch.cuda.empty_cache()



def decoders(model, O, CFG, SHAPE, to_dev, both, out):
    # ---------------------------------------------------------------- greedy decoders (B=64, 20 tokens)
    dh = to_dev({k: v for k, v in O.synth_inputs(CFG, B=64, Q=SHAPE["Q"], C=SHAPE["C"], H=SHAPE["H"], T=4, Lv=SHAPE["Lv"],
                                               seed=5001).items() if k in ("query", "his", "cap", "fts")})
    d0, d1 = GraphedGreedyDecoder(model, dh, 20), GraphedGreedyDecoder(model, dh, 20)
    ta = d0.decode().clone(); torch.cuda.synchronize()
    ta2 = d0.decode().clone(); torch.cuda.synchronize()
    tb = d1.decode().clone(); torch.cuda.synchronize()
    print("decode   run-to-run    identical: %s" % torch.equal(ta, ta2), file=out)
    print("decode   graph-to-graph identical: %s (%d of %d sequences differ)" %
          (torch.equal(ta, tb), int((ta != tb).any(1).sum()), ta.shape[0]), file=out)
    bad = 0
    for _ in range(5):
        both(d0.decode, d1.decode, n=1)
        bad += int(not torch.equal(d0.ys, ta)) + int(not torch.equal(d1.ys, tb))
    print("decode   concurrent     mismatching results: %d of 10 (%d / %d sequences differ in the last run)" %
          (bad, int((d0.ys != ta).any(1).sum()), int((d1.ys != tb).any(1).sum())), file=out)
    # where do two decoder instances part?  first differing position, the encoder outputs, and the eager decoder
    d0.decode(); d1.decode(); torch.cuda.synchronize()
    diff = (d0.ys != d1.ys)
    first = [int(r.nonzero()[0]) for r in diff if r.any()]
    print("decode   first differing position per differing sequence: %s" % sorted(first), file=out)

    def flat(m):
        res = []
        for t in m:
            if isinstance(t, (list, tuple)):
                res += flat(t)
            elif t is not None:
                res.append(t)
        return res
    ma, mb = flat(d0.mem), flat(d1.mem)
    print("decode   encoder outputs identical: %s" % [bool(torch.equal(x, y)) for x, y in zip(ma, mb)], file=out)
    from mtn_b200.data_utils import Batch, greedy_decode
    with torch.no_grad():
        bt = Batch(dh["query"], dh["his"], None, [f.permute(1, 0, 2) for f in dh["fts"]], dh["cap"], None, None, 1)
        eager = greedy_decode(model, bt, 20, 2)
    torch.cuda.synchronize()
    print("decode   eager greedy_decode vs decoder 0: %d sequences differ; vs decoder 1: %d" %
          (int((eager != d0.ys).any(1).sum()), int((eager != d1.ys).any(1).sum())), file=out)
    with torch.no_grad():
        eager2 = greedy_decode(model, bt, 20, 2)
    torch.cuda.synchronize()
    print("decode   eager run-to-run: %d sequences differ" % int((eager != eager2).any(1).sum()), file=out)


if __name__ == "__main__":
    main()
