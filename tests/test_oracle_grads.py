"""CPU tier: the oracle's training-step restatement (loss + parameter gradients, dropout disabled) against the
golden digest produced by the UNMODIFIED reference (oracle/make_golden_grads.py -> tests/golden/grads_cfg1.npz)."""
import numpy as np
import torch

import golden_util as G
import mtn_oracle as O


def golden_grad_case():
    z = G.load("grads_cfg1.npz")
    cfg, sd = G.seeded_state_dict(z)
    inp = O.synth_inputs(cfg, seed=int(z["input_seed"]), B=3, Q=8, C=8, H=16, T=8, Lv=[16, 8])
    return z, cfg, sd, inp


def grad_rms_floor(norms, numels):
    """Global rms magnitude of a set of gradient tensors given their L2 norms and sizes."""
    return float(np.sqrt(sum(n * n for n in norms.values()) / sum(numels.values())))


def grad_errors(grads, ref):
    """{name: ||g - ref|| / max(||ref||, 1e-2 * global rms * sqrt(numel))} (see check_against_digest)."""
    floor = 1e-2 * grad_rms_floor({k: float(v.double().norm()) for k, v in ref.items()},
                                  {k: v.numel() for k, v in ref.items()})
    out = {}
    for k, r in ref.items():
        r = r.detach().double().cpu()
        g = grads[k].detach().double().cpu()
        den = max(float(r.norm()), floor * np.sqrt(r.numel()))
        if k.endswith("linears.1.bias"):
            # the key-projection bias has an analytically ZERO gradient (softmax is invariant to a per-query
            # constant); what every implementation returns is rounding noise -- measure it against its sibling,
            # the value-projection bias gradient
            den = max(den, float(ref[k.replace("linears.1.bias", "linears.2.bias")].double().norm()))
        out[k] = float((g - r).norm() / den)
    return out


def check_against_digest(z, grads, tol):
    """grads: {name: tensor}.  Returns the worst normwise error estimate over all tensors."""
    names = sorted(k[2:-5] for k in z if k.startswith("g/") and k.endswith("/norm"))
    assert sorted(grads) == names
    worst = 0.0
    floor = 1e-2 * grad_rms_floor({k: float(z["g/%s/norm" % k]) for k in names}, {k: grads[k].numel() for k in names})
    for k in names:
        g = grads[k].detach().double().cpu().reshape(-1)
        # gradients that are analytically zero (the key bias: softmax is invariant to it) are rounding noise in
        # every implementation; errors are measured against max(||ref||, floor * sqrt(numel))
        ref_norm = max(float(z["g/%s/norm" % k]), floor * np.sqrt(g.numel()))
        pos = torch.from_numpy(z["g/%s/pos" % k])
        val = torch.from_numpy(z["g/%s/val" % k]).double()
        scale = ref_norm / np.sqrt(g.numel())                       # rms magnitude of the tensor
        e_norm = abs(float(g.norm()) - float(z["g/%s/norm" % k])) / ref_norm
        e_val = float((g[pos] - val).abs().max()) / scale / 8.0     # sampled entries, in units of 8 rms
        e_sum = abs(float(g.sum()) - float(z["g/%s/sum" % k])) / (ref_norm * np.sqrt(g.numel()))
        err = max(e_norm, e_val, e_sum)
        assert err <= tol, (k, e_norm, e_val, e_sum)
        worst = max(worst, err)
    return worst


def test_oracle_training_step_matches_reference_golden():
    z, cfg, sd, inp = golden_grad_case()
    loss, grads = O.loss_and_grads(sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"], inp["fts"])
    norm = float((inp["trg_y"] != 1).sum())
    assert abs(loss * norm - float(z["loss_times_norm"])) <= 1e-5 * abs(float(z["loss_times_norm"]))
    assert check_against_digest(z, grads, 2e-5) < 2e-5
    # every parameter of the model receives a gradient in this configuration
    assert all(float(g.abs().max()) > 0 for g in grads.values())
