"""Bisect the batch-64 greedy-decode discrepancy (DESIGN.md section 7 of round 1): two GraphedGreedyDecoder instances
and the eager decoder disagree on some sequences although every kernel is meant to be a pure function of its operands.

Part A (graphs): capture two decoder instances with debug taps (engine.TAP: a clone of every target-path intermediate
inside the captured graphs), replay both, and report the FIRST tap (step, layer, site, tensor) that differs, the rows /
sequences affected and the size of the difference; also compares the cached memory stages of the two instances.
Part B (eager): run the eager decoder under two allocator layouts (a persistent dummy allocation shifts every later
buffer) with every launch followed by a synchronize + checksum of its operands, and report the first launch whose
outputs differ while its inputs agree.

Usage: python tools/bisect_decode.py [out_file] [--no-taps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mtn_b200 import _lib, engine, mtn  # noqa: E402
from mtn_b200.graph import GraphedGreedyDecoder  # noqa: E402
from mtn_b200.data_utils import Batch, greedy_decode  # noqa: E402


def flat_tensors(x, prefix=""):
    res = []
    if torch.is_tensor(x):
        res.append((prefix, x))
    elif isinstance(x, (list, tuple)):
        for i, t in enumerate(x):
            res += flat_tensors(t, "%s[%d]" % (prefix, i))
    elif isinstance(x, dict):
        for k in sorted(x, key=str):
            if str(k).startswith("_") or k in ("side", "ev"):
                continue
            res += flat_tensors(x[k], "%s.%s" % (prefix, k))
    return res


def describe_diff(a, b, rows_per_seq, out):
    a2, b2 = a.float().reshape(-1, a.shape[-1]), b.float().reshape(-1, b.shape[-1])
    bad = (a2 != b2)
    rows = bad.any(1).nonzero().flatten().tolist()
    cols = bad.any(0).nonzero().flatten().tolist()
    print("      shape %s: %d differing elements in %d rows x %d cols; max abs diff %.4e (max abs value %.3e)" %
          (tuple(a.shape), int(bad.sum()), len(rows), len(cols), float((a2 - b2).abs().max()), float(a2.abs().max())), file=out)
    print("      rows: %s%s" % (rows[:40], " ..." if len(rows) > 40 else ""), file=out)
    if rows_per_seq:
        print("      sequences: %s" % sorted(set(r // rows_per_seq for r in rows)), file=out)
    print("      cols: %s%s" % (cols[:24], " ..." if len(cols) > 24 else ""), file=out)


def explain_self_attn(la, lb, i, t, out):
    """The first differing tap is a self-attention output: compare both instances with what the operands imply."""
    names = [n for n, _ in la]
    qkv_a, qkv_b = la[i - 1][1], lb[i - 1][1]
    o_a, o_b = la[i][1], lb[i][1]
    bi = names.index("bits_t") if "bits_t" in names else None
    if bi is not None:
        ba, bb = la[bi][1], lb[bi][1]
        print("      mask bits identical between the instances: %s; bits[0, :, 0] = %s / %s" %
              (bool(torch.equal(ba, bb)), ba[0, :, 0].tolist(), bb[0, :, 0].tolist()), file=out)
        print("      bits[1, :, 0] = %s / %s ; bits[63, :, 0] = %s / %s" %
              (ba[1, :, 0].tolist(), bb[1, :, 0].tolist(), ba[63, :, 0].tolist(), bb[63, :, 0].tolist()), file=out)
    d = o_a.shape[1]
    B = o_a.shape[0] // t
    v = qkv_a[:, 2 * d:].float().view(B, t, d)
    for name, o in (("instance0", o_a), ("instance1", o_b)):
        o3 = o.float().view(B, t, d)
        r0 = o3[:, 0]
        cands = {"V[0]": v[:, 0], "mean V[0:t]": v.mean(1), "V[1]": v[:, min(1, t - 1)], "zeros": torch.zeros_like(r0)}
        print("      %s row 0 of every sequence vs candidates (max abs diff over all sequences/columns): %s" %
              (name, {k: "%.3e" % float((r0 - c).abs().max()) for k, c in cands.items()}), file=out)
        print("      %s seq 0 row 0 cols 0:6 = %s ; V[0] = %s" % (name, [round(float(x), 4) for x in o3[0, 0, :6]],
                                                                  [round(float(x), 4) for x in v[0, 0, :6]]), file=out)
    # recompute with the library's own check kernel and with the fast kernel, eagerly, from the tapped operands
    h = 8
    dk = d // h
    tm = torch.tril(torch.ones(1, t, t, dtype=torch.bool, device=o_a.device)).expand(B, -1, -1)
    bits = _lib.mask_pack(tm)
    for kind in (True, False):
        o = torch.full_like(o_a, float("nan"))
        _lib.attn_core(qkv_a[:, :d], qkv_a[:, d:2 * d], qkv_a[:, 2 * d:], B, h, t, t, dk, o, mask_bits=bits, _check_kernel=kind)
        torch.cuda.synchronize()
        print("      eager %s kernel on the tapped operands: equal to instance0 %s, instance1 %s" %
              ("check" if kind else "fast", bool(torch.equal(o, o_a)), bool(torch.equal(o, o_b))), file=out)
        if not kind:
            for name, oo in (("instance0", o_a), ("instance1", o_b)):
                bad = (o != oo).any(1).nonzero().flatten().tolist()
                print("         rows where %s differs from the eager fast kernel: %s" % (name, bad[:30]), file=out)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out = open(args[0], "w") if args else sys.stdout
    use_taps = "--no-taps" not in sys.argv
    _lib.lib()
    O = bench.oracle()
    CFG, SHAPE = bench.CFG, bench.SHAPE
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    model = mtn.make_model(CFG["vocab"], CFG["vocab"], N=CFG["N"], d_model=CFG["d_model"], d_ff=CFG["d_ff"], h=CFG["h"],
                           ft_sizes=CFG["ft_sizes"], diff_encoder=True, auto_encoder_ft="query").to(dev).eval()
    dh = {k: (v.to(dev) if torch.is_tensor(v) else [f.to(dev) for f in v])
          for k, v in O.synth_inputs(CFG, B=64, Q=SHAPE["Q"], C=SHAPE["C"], H=SHAPE["H"], T=4, Lv=SHAPE["Lv"], seed=5001).items()
          if k in ("query", "his", "cap", "fts")}
    STEPS = 20

    # ------------------------------------------------------------------ part A: graph instances with taps
    calls = []
    orig_decode = model.decode

    def tapped_decode(*a, **k):
        engine.TAP = [] if use_taps else None
        try:
            return orig_decode(*a, **k)
        finally:
            calls.append(engine.TAP)
            engine.TAP = None
    model.decode = tapped_decode
    insts, taps, mems = [], [], []
    for _ in range(2):
        del calls[:]
        d = GraphedGreedyDecoder(model, dh, STEPS)
        insts.append(d)
        taps.append(list(calls[-(STEPS - 1):]))          # the captured calls (the eager warm-up ones come first)
        mems.append(model.decoder.engine._mem)           # keeps this instance's memory stage referenced
    model.decode = orig_decode
    t0 = insts[0].decode().clone(); torch.cuda.synchronize()
    t1 = insts[1].decode().clone(); torch.cuda.synchronize()
    diff = (t0 != t1)
    first = {i: int(r.nonzero()[0]) for i, r in enumerate(diff) if r.any()}
    print("A: graph instances (taps %s): %d of 64 sequences differ; first differing position per sequence: %s" %
          ("on" if use_taps else "off", len(first), first), file=out)
    # memory stages
    insts[0].decode(); insts[1].decode(); torch.cuda.synchronize()
    fa, fb = flat_tensors(mems[0], "S"), flat_tensors(mems[1], "S")
    nbad = 0
    for (na, x), (nb, y) in zip(fa, fb):
        if x.shape == y.shape and x.dtype == y.dtype and not torch.equal(x, y):
            nbad += 1
            print("   memory stage tensor %s differs between the instances:" % na, file=out)
            describe_diff(x, y, 0, out)
    print("A: memory stage: %d tensors compared, %d differ" % (len(fa), nbad), file=out)
    ea = flat_tensors(insts[0].mem, "mem"); eb = flat_tensors(insts[1].mem, "mem")
    print("A: encoder outputs identical: %s" % [bool(torch.equal(x, y)) for (_, x), (_, y) in zip(ea, eb)], file=out)
    if use_taps:
        found = False
        for step, (la, lb) in enumerate(zip(taps[0], taps[1])):
            t = step + 1                                   # prefix length of this decode call
            assert len(la) == len(lb)
            for i, ((na, x), (nb, y)) in enumerate(zip(la, lb)):
                if not torch.equal(x, y):
                    print("A: FIRST differing tap: prefix length %d, tap %d %r (previous tap: %r)" %
                          (t, i, na, la[i - 1][0] if i else None), file=out)
                    describe_diff(x, y, t, out)
                    print("      data_ptr instance0 0x%x  instance1 0x%x" % (x.data_ptr(), y.data_ptr()), file=out)
                    # how many later taps of this step differ
                    later = [n for (n, u), (_, v) in zip(la[i:], lb[i:]) if not torch.equal(u, v)]
                    print("      taps differing from here on in this step: %d of %d; next ones: %s" %
                          (len(later), len(la) - i, later[:8]), file=out)
                    if na.endswith("self.o") and i >= 1:
                        explain_self_attn(la, lb, i, t, out)
                    found = True
                    break
            if found:
                break
        if not found:
            print("A: no tap differs between the instances", file=out)
    out.flush()
    del insts, taps, mems
    torch.cuda.empty_cache()

    if "--no-eager" in sys.argv:
        return
    # ------------------------------------------------------------------ part B: eager decoder under two allocator layouts
    log = []
    orig = _lib._launch
    keepalive = []

    def hooked(name, flops, nbytes, fn, keep=()):
        ins = []
        for x in keep:
            if torch.is_tensor(x):
                ins.append(int(x.contiguous().view(-1).view(torch.uint8).to(torch.int64).sum()))
        orig(name, flops, nbytes, fn, keep)
        torch.cuda.synchronize()
        outs, ptrs = [], []
        for x in keep:
            if torch.is_tensor(x):
                outs.append(int(x.contiguous().view(-1).view(torch.uint8).to(torch.int64).sum()))
                ptrs.append((tuple(x.shape), str(x.dtype).replace("torch.", ""), x.data_ptr()))
        log.append((name, ins, outs, ptrs))

    def eager(hook):
        del log[:]
        if hook:
            _lib._launch = hooked
        try:
            with torch.no_grad():
                bt = Batch(dh["query"], dh["his"], None, [f.permute(1, 0, 2) for f in dh["fts"]], dh["cap"], None, None, 1)
                ys = greedy_decode(model, bt, STEPS, 2)
        finally:
            _lib._launch = orig
        torch.cuda.synchronize()
        return ys.clone(), list(log)

    ya, _ = eager(False)
    ya2, _ = eager(False)
    print("B: eager run-to-run (same layout): %d sequences differ" % int((ya != ya2).any(1).sum()), file=out)
    results = []
    for shift_mb in (0, 3, 7, 33):
        model.decoder.engine._mem_key, model.decoder.engine._mem = None, None
        torch.cuda.empty_cache()
        if shift_mb:
            keepalive.append(torch.empty(shift_mb << 20, dtype=torch.uint8, device=dev))
            keepalive.append(torch.empty(300 << 10, dtype=torch.uint8, device=dev))
        y, _ = eager(False)
        results.append(y)
        print("B: eager, layout shift %2d MB: vs unshifted %d sequences differ; vs graph instance 0: %d" %
              (shift_mb, int((y != results[0]).any(1).sum()), int((y != t0).any(1).sum())), file=out)
    out.flush()
    # hooked runs (synchronize + checksum after every launch) under two layouts
    del keepalive[:]
    model.decoder.engine._mem_key, model.decoder.engine._mem = None, None
    torch.cuda.empty_cache()
    yh0, l0 = eager(True)
    model.decoder.engine._mem_key, model.decoder.engine._mem = None, None
    torch.cuda.empty_cache()
    keepalive.append(torch.empty(7 << 20, dtype=torch.uint8, device=dev))
    keepalive.append(torch.empty(300 << 10, dtype=torch.uint8, device=dev))
    yh1, l1 = eager(True)
    print("B: hooked eager runs: %d launches; sequences differing between the layouts: %d" %
          (len(l0), int((yh0 != yh1).any(1).sum())), file=out)
    shown = 0
    for i, (x, y) in enumerate(zip(l0, l1)):
        if x[0] != y[0]:
            print("B: launch %d: different kernels %s / %s" % (i, x[0], y[0]), file=out)
            break
        if x[2] != y[2]:
            print("B: launch %d (%s): inputs %s, outputs DIFFER" % (i, x[0], "agree" if x[1] == y[1] else "DIFFER"), file=out)
            for k, (px, py) in enumerate(zip(x[3], y[3])):
                print("      operand %d %s %s  ptr 0x%x / 0x%x  in %s out %s" %
                      (k, px[0], px[1], px[2], py[2], "same" if x[1][k] == y[1][k] else "DIFF",
                       "same" if x[2][k] == y[2][k] else "DIFF"), file=out)
            print("      previous launches: %s" % [n for n, _, _, _ in l0[max(0, i - 5):i]], file=out)
            shown += 1
            if shown >= 4:
                break
    if not shown:
        print("B: no launch differs between the two hooked layouts", file=out)
    out.flush()


if __name__ == "__main__":
    main()
