"""TEST INFRASTRUCTURE: golden GRADIENTS of one training step of the UNMODIFIED reference (CPU), for the
backward kernels.  Run in the build container:

    python oracle/make_golden_grads.py        ->  tests/golden/grads_cfg1.npz

The reference's own training code path is executed: ``model.train()`` with every ``nn.Dropout.p`` set to 0
(dropout noise cannot be reproduced across implementations), ``model.forward(batch)`` (train.py:33), the
reference ``SimpleLossCompute`` + ``LabelSmoothing`` (train.py:37-39, data_utils.py:132-155) with a recording
stand-in for the optimizer, so ``loss.backward()`` is the reference's call.  Weights come from
``mtn_oracle.init_state_dict(cfg, seed)`` (a pure generator; the fixture stores cfg/seed/checksum), inputs from
``mtn_oracle.synth_inputs``.  To keep the fixture small only a digest of every gradient tensor is stored:
its L2 norm, its sum and 48 entries at seeded positions.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mtn_oracle  # noqa: E402
import ref_loader  # noqa: E402
from make_golden import cfg_np, ref_model_from_seed  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CFG = {"N": 1, "d_model": 128, "d_ff": 512, "h": 4, "vocab": 100, "ft_sizes": [2048, 128],
       "auto_encoder_ft": "query", "diff_encoder": True}
SEED, INPUT_SEED = 11, 5
SHAPES = dict(B=3, Q=8, C=8, H=16, T=8, Lv=[16, 8])


def digest_positions(name, numel, n=48):
    g = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def hash_name(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000000007
    return h


def reference_step(model, du, ls_mod, inp, pad=1, smoothing=0.1):
    """The reference's training step; returns (loss * norm as the reference reports it, {name: grad})."""
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    b = ref_loader.make_cpu_batch(inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"], pad)
    grads = {}

    class RecordingOpt(object):          # stands in for NoamOpt: snapshots .grad where Adam would consume it
        class optimizer(object):
            @staticmethod
            def zero_grad():
                pass

        @staticmethod
        def step():
            for k, p in model.named_parameters():
                grads[k] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)

    V = model.generator.proj.weight.shape[0]
    criterion = ls_mod.LabelSmoothing(size=V, padding_idx=pad, smoothing=smoothing)
    lc = du.SimpleLossCompute(model.generator, None, criterion, opt=RecordingOpt, l=1.0)
    out, ae_out = model.forward(b)                                            # train.py:33
    ntokens_query = (b.query != pad).data.sum()                               # train.py:38
    loss = lc(out, b.trg_y, b.ntokens, ae_out, b.query, ntokens_query)        # train.py:39
    return float(loss), grads


def load_ref_label_smoothing():
    spec = importlib.util.spec_from_file_location("ref_label_smoothing",
                                                  os.path.join(ref_loader.ref_dir(), "label_smoothing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    warnings.simplefilter("ignore")
    mtn, du = ref_loader.load()
    ls_mod = load_ref_label_smoothing()
    torch.set_num_threads(1)
    model, sd, chk = ref_model_from_seed(mtn, CFG, SEED)
    inp = mtn_oracle.synth_inputs(CFG, seed=INPUT_SEED, **SHAPES)
    loss, grads = reference_step(model, du, ls_mod, inp)
    d = cfg_np(CFG, SEED, chk)
    d["input_seed"] = np.int64(INPUT_SEED)
    d["loss_times_norm"] = np.float64(loss)
    for k, g in grads.items():
        flat = g.reshape(-1).double()
        pos = digest_positions(k, flat.numel())
        d["g/%s/norm" % k] = np.float64(flat.norm())
        d["g/%s/sum" % k] = np.float64(flat.sum())
        d["g/%s/pos" % k] = pos.numpy()
        d["g/%s/val" % k] = g.reshape(-1)[pos].numpy()
    np.savez_compressed(os.path.join(OUT, "grads_cfg1.npz"), **d)
    # cross-check the restatement right here
    oloss, og = mtn_oracle.loss_and_grads(sd, CFG, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"],
                                          inp["fts"])
    norm = float((inp["trg_y"] != 1).sum())
    worst = max(float((og[k].double() - grads[k].double()).norm() / grads[k].double().norm().clamp_min(1e-30))
                for k in grads)
    print("reference loss*norm = %.6f   oracle = %.6f   worst grad rel err oracle vs reference = %.2e  (%d tensors)"
          % (loss, oloss * norm, worst, len(grads)))


if __name__ == "__main__":
    main()
