"""Helpers shared by the parity tests: load tests/golden/*.npz into torch."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def state_dict_from(z, d_model):
    """state_dict stored under 'sd/' keys; the sinusoid tables are regenerated."""
    import mtn_oracle
    sd = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd/")}
    pe = mtn_oracle.sinusoid_pe(d_model)
    for k in ("query_embed.1.pe", "tgt_embed.1.pe", "vid_encoder.0.2.pe", "vid_encoder.1.2.pe"):
        sd[k] = pe.clone()
    return sd


def cfg_from(z):
    cfg = {k: int(z["cfg/" + k]) for k in ("N", "d_model", "d_ff", "h", "vocab")}
    cfg["ft_sizes"] = [int(x) for x in z["cfg/ft_sizes"]]
    cfg["auto_encoder_ft"] = "query"
    cfg["diff_encoder"] = True
    return cfg


def seeded_state_dict(z):
    """Weights of a 'weights by seed' fixture, checked against the stored checksum."""
    import mtn_oracle
    cfg = cfg_from(z)
    sd = mtn_oracle.init_state_dict(cfg, int(z["cfg/seed"]))
    gs = float(z["cfg/gen_scale"])
    if gs != 1.0:
        sd["generator.proj.weight"] = sd["generator.proj.weight"] * gs
    chk = float(sum(v.double().sum() for k, v in sd.items() if not k.endswith(".pe")))
    assert abs(chk - float(z["cfg/weight_checksum"])) < 1e-6 * max(1.0, abs(chk)), \
        "seeded weights drifted from the fixture (torch CPU RNG changed?)"
    return cfg, sd


CFG1 = {"N": 1, "d_model": 128, "d_ff": 512, "h": 4, "vocab": 100, "ft_sizes": [2048, 128],
        "auto_encoder_ft": "query", "diff_encoder": True}


def rel_err(a, b):
    """normwise relative error ||a-b|| / ||b|| (SURVEY 8d parity metric)."""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
