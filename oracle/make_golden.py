"""TEST INFRASTRUCTURE: generate ``tests/golden/*.npz`` by running the UNMODIFIED
reference (``/root/reference``, imported through ``oracle/ref_loader.py``) on CPU.

Run in the build container (the reference is not present on the GPU box):

    python oracle/make_golden.py

Fixtures (all fp32 / int64, torch 2.11 CPU):
  kat.npz          -- the three known-answer vectors of SURVEY.md 8c
                      (LayerNorm(4), fully-masked attention row, last-key-masked row)
  cfg1.npz         -- BASELINE.json configs[0]: N=1 d=128 h=4 B=2, text-only
                      (features all zero), the model-level recipe of SURVEY 8c:
                      full state_dict + inputs + out + ae outputs + argmax
  cfg1b.npz        -- same model, random features with padded frames and an
                      all-pad history row (uniform-softmax path)
  site_*.npz       -- one SublayerConnection(MultiHeadedAttention) site and one
                      SublayerConnection(PositionwiseFeedForward) at d=128,h=4
                      with key-pad, causal and all-masked rows
  greedy.npz       -- 7 greedy steps (fixed call form) on an
                      N=2 d=128 model: tokens + per-step last-row log-probs
  mini512.npz      -- N=2 d=512 h=8 (the cfg2 family, d_k=64) forward, ragged pads
For greedy/mini512 the weights are NOT stored: the reference model is loaded with
``mtn_oracle.init_state_dict(cfg, seed)`` (a pure weight generator) and the
fixture records (cfg, seed, checksum); outputs are the reference's own.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


import mtn_oracle  # noqa: E402  (weight generator only; outputs come from the reference)


def sd_np(model):
    """state_dict without the (deterministic, 2.5 MB each) sinusoid tables."""
    return {"sd/" + k: v.detach().numpy() for k, v in model.state_dict().items()
            if not k.endswith(".pe")}


def ref_model_from_seed(mtn, cfg, seed, gen_scale=1.0):
    """Reference model whose weights are ``mtn_oracle.init_state_dict(cfg, seed)``:
    lets a fixture carry (cfg, seed, checksum) instead of megabytes of weights."""
    model = mtn.make_model(cfg["vocab"], cfg["vocab"], N=cfg["N"], d_model=cfg["d_model"],
                           d_ff=cfg["d_ff"], h=cfg["h"], dropout=0.1, ft_sizes=cfg["ft_sizes"],
                           diff_encoder=True, auto_encoder_ft=cfg["auto_encoder_ft"]).eval()
    sd = mtn_oracle.init_state_dict(cfg, seed)
    if gen_scale != 1.0:
        sd["generator.proj.weight"] = sd["generator.proj.weight"] * gen_scale
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    chk = float(sum(v.double().sum() for k, v in sd.items() if not k.endswith(".pe")))
    return model, sd, chk


def cfg_np(cfg, seed, chk, gen_scale=1.0):
    return {"cfg/N": np.int64(cfg["N"]), "cfg/d_model": np.int64(cfg["d_model"]),
            "cfg/d_ff": np.int64(cfg["d_ff"]), "cfg/h": np.int64(cfg["h"]),
            "cfg/vocab": np.int64(cfg["vocab"]), "cfg/ft_sizes": np.array(cfg["ft_sizes"], np.int64),
            "cfg/seed": np.int64(seed), "cfg/weight_checksum": np.float64(chk),
            "cfg/gen_scale": np.float64(gen_scale)}


def main():
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    mtn, du = ref_loader.load()
    torch.set_num_threads(1)

    # ---- KATs
    ln = mtn.LayerNorm(4)
    kat = {"ln_in": np.array([1., 2., 3., 4.], np.float32)}
    kat["ln_out"] = ln(torch.tensor(kat["ln_in"])).detach().numpy()
    q = torch.tensor([[1., 0.]]); k = torch.tensor([[1., 0.], [0., 1.], [5., 5.]])
    v = torch.tensor([[1., 2.], [3., 4.], [5., 6.]])
    for name, mask in (("allmasked", torch.tensor([[False, False, False]])),
                       ("lastmasked", torch.tensor([[True, True, False]]))):
        o, p = mtn.attention(q, k, v, mask=mask)
        kat["attn_%s_o" % name] = o.numpy(); kat["attn_%s_p" % name] = p.numpy()
        kat["attn_%s_mask" % name] = mask.numpy()
    kat["attn_q"], kat["attn_k"], kat["attn_v"] = q.numpy(), k.numpy(), v.numpy()
    np.savez(os.path.join(OUT, "kat.npz"), **kat)

    # ---- cfg1 (SURVEY 8c model-level recipe)
    torch.manual_seed(1234)
    model = mtn.make_model(100, 100, N=1, d_model=128, d_ff=512, h=4, dropout=0.1,
                           ft_sizes=[2048, 128], diff_encoder=True,
                           auto_encoder_ft='query').eval()
    g = torch.Generator().manual_seed(4321)
    qy = torch.randint(4, 100, (2, 8), generator=g); his = torch.randint(4, 100, (2, 16), generator=g)
    cap = torch.randint(4, 100, (2, 8), generator=g); trg = torch.randint(4, 100, (2, 8), generator=g)
    trg_y = torch.randint(4, 100, (2, 8), generator=g)
    qy[1, 6:] = 1; his[1, 10:] = 1; cap[1, 5:] = 1; trg[1, 5:] = 1; trg_y[1, 5:] = 1
    fts = [torch.zeros(2, 16, 2048), torch.zeros(2, 8, 128)]

    def run(fts_, his_, name, with_sd=True):
        b = ref_loader.make_cpu_batch(qy, his_, cap, trg, trg_y, fts_)
        with torch.no_grad():
            out, ae = model.forward(b)
            logp = model.generator(out)
        d = sd_np(model) if with_sd else {}
        d.update(query=qy.numpy(), his=his_.numpy(), cap=cap.numpy(), trg=trg.numpy(),
                 trg_y=trg_y.numpy(), ft0=fts_[0].numpy(), ft1=fts_[1].numpy(),
                 out=out.numpy(), ae0=ae[0].numpy(), ae1=ae[1].numpy(),
                 argmax=logp.argmax(-1).numpy(), ntokens=np.int64(b.ntokens.item()),
                 trg_mask=b.trg_mask.numpy(), fts_mask0=b.fts_mask[0].numpy(),
                 fts_mask1=b.fts_mask[1].numpy())
        np.savez_compressed(os.path.join(OUT, name), **d)
        return out

    out = run(fts, his, "cfg1.npz")
    print("cfg1  sum|out| = %.7f  (SURVEY: 1645.1214289)" % out.abs().sum().item())
    fts_b = [torch.randn(2, 16, 2048, generator=g), torch.randn(2, 8, 128, generator=g)]
    fts_b[0][1, 11:] = 1.0; fts_b[1][1, 5:] = 1.0
    his_b = his.clone(); his_b[1, :] = 1            # all-pad history -> uniform softmax
    run(fts_b, his_b, "cfg1b.npz", with_sd=False)   # weights: see cfg1.npz

    # ---- site-level: SublayerConnection(MHA) / SublayerConnection(FFN), d=128 h=4
    torch.manual_seed(77)
    d, h = 128, 4
    sub = mtn.SublayerConnection(d, 0.1).eval()
    att = mtn.MultiHeadedAttention(h, d).eval()
    ff = mtn.PositionwiseFeedForward(d, 4 * d, 0.1).eval()
    for m_ in (sub, att, ff):
        for p in m_.parameters():
            if p.dim() > 1:
                torch.nn.init.xavier_uniform_(p)
    with torch.no_grad():
        sub.norm.a_2.copy_(1 + 0.1 * torch.randn(d)); sub.norm.b_2.copy_(0.1 * torch.randn(d))
    B, Lq, Lk = 3, 10, 37
    x = torch.randn(B, Lq, d) * 3 + 0.5
    mem = torch.randn(B, Lk, d)
    kmask = torch.ones(B, 1, Lk, dtype=torch.bool); kmask[1, 0, 20:] = False; kmask[2, 0, :] = False
    cmask = (torch.ones(B, 1, Lq, dtype=torch.bool) & du.subsequent_mask(Lq))
    cmask = cmask.clone(); cmask[1, :, 7:] = False
    with torch.no_grad():
        y_cross = sub(x, lambda t: att(t, mem, mem, kmask))
        y_self = sub(x, lambda t: att(t, t, t, cmask))
        y_nomask = sub(x, lambda t: att(t, mem, mem, None))
        y_ffn = sub(x, ff)
    site = {"sub/" + k_: v_.detach().numpy() for k_, v_ in sub.state_dict().items()}
    site.update({"att/" + k_: v_.detach().numpy() for k_, v_ in att.state_dict().items()})
    site.update({"ff/" + k_: v_.detach().numpy() for k_, v_ in ff.state_dict().items()})
    site.update(x=x.numpy(), mem=mem.numpy(), kmask=kmask.numpy(), cmask=cmask.numpy(),
                y_cross=y_cross.numpy(), y_self=y_self.numpy(), y_nomask=y_nomask.numpy(),
                y_ffn=y_ffn.numpy(), h=np.int64(h))
    np.savez_compressed(os.path.join(OUT, "site_d128.npz"), **site)

    # ---- greedy (fixed call form), N=2 d=128, 7 steps, weights by seed
    gcfg = {"N": 2, "d_model": 128, "d_ff": 512, "h": 4, "vocab": 60, "ft_sizes": [64, 32],
            "auto_encoder_ft": "query"}
    model2, _, chk = ref_model_from_seed(mtn, gcfg, 99, gen_scale=8.0)  # x8: realistic margins
    g = torch.Generator().manual_seed(5)
    Bd = 3
    qy2 = torch.randint(4, 60, (Bd, 9), generator=g); his2 = torch.randint(4, 60, (Bd, 21), generator=g)
    cap2 = torch.randint(4, 60, (Bd, 11), generator=g)
    qy2[1, 6:] = 1; his2[1, 10:] = 1; cap2[2, 5:] = 1; his2[2, :] = 1
    f2 = [torch.randn(Bd, 13, 64, generator=g), torch.randn(Bd, 7, 32, generator=g)]
    f2[0][1, 9:] = 1.0
    toks, lps = [], []
    for bi in range(Bd):                                   # the reference decodes batch-1
        b = ref_loader.make_cpu_batch(qy2[bi:bi + 1], his2[bi:bi + 1], cap2[bi:bi + 1], None, None,
                                      [f[bi:bi + 1] for f in f2])
        with torch.no_grad():
            his_m, cap_m, q_m, vid_m, ae_m = du.encode(model2, b.his, None, b.his_mask, b.cap,
                                                       b.cap_mask, b.query, b.query_mask, b.fts,
                                                       b.fts_mask)
            ys = torch.full((1, 1), 2, dtype=torch.long)
            lp_steps = []
            for _ in range(7):
                o = model2.decode(vid_m, his_m, cap_m, q_m, b.fts_mask, b.his_mask, b.cap_mask,
                                  b.query_mask, ys, du.subsequent_mask(ys.size(1)), ae_m)
                lp = model2.generator(o[0][:, -1])
                lp_steps.append(lp[0].numpy())
                ys = torch.cat([ys, lp.argmax(-1, keepdim=True)], dim=1)
        toks.append(ys[0].numpy()); lps.append(np.stack(lp_steps))
    gd = cfg_np(gcfg, 99, chk, 8.0)
    gd.update(query=qy2.numpy(), his=his2.numpy(), cap=cap2.numpy(), ft0=f2[0].numpy(),
              ft1=f2[1].numpy(), tokens=np.stack(toks), logp=np.stack(lps))
    np.savez_compressed(os.path.join(OUT, "greedy.npz"), **gd)

    # ---- mini512: the cfg2 architecture family (d=512 h=8 d_k=64) at N=2 and small
    #      lengths, ragged padding + all-pad history; weights by seed
    mcfg = {"N": 2, "d_model": 512, "d_ff": 2048, "h": 8, "vocab": 200, "ft_sizes": [2048, 128],
            "auto_encoder_ft": "query"}
    model3, _, chk3 = ref_model_from_seed(mtn, mcfg, 512)
    inp = mtn_oracle.synth_inputs(mcfg, B=3, Q=16, C=16, H=40, T=12, Lv=[40, 24], seed=11)
    b = ref_loader.make_cpu_batch(inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"],
                                  inp["fts"])
    with torch.no_grad():
        out3, ae3 = model3.forward(b)
    md = cfg_np(mcfg, 512, chk3)
    md.update(query=inp["query"].numpy(), his=inp["his"].numpy(), cap=inp["cap"].numpy(),
              trg=inp["trg"].numpy(), trg_y=inp["trg_y"].numpy(), ft0=inp["fts"][0].numpy(),
              ft1=inp["fts"][1].numpy(), out=out3.numpy(), ae0=ae3[0].numpy(), ae1=ae3[1].numpy())
    np.savez_compressed(os.path.join(OUT, "mini512.npz"), **md)
    # ---- label smoothing (label_smoothing.py) incl. the padding-row index-sum quirk
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_label_smoothing", os.path.join(ref_loader.ref_dir(), "label_smoothing.py"))
    ref_ls = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_ls)
    g = torch.Generator().manual_seed(8)
    ls = {}
    for name, tgt in (("mixed", [5, 1, 7, 1, 2, 9]), ("quirk_pad_row0_only", [1, 4, 6, 3, 2, 8]), ("nopad", [5, 3, 7, 4, 2, 9])):
        logp = torch.log_softmax(torch.randn(6, 12, generator=g) * 2, -1)
        crit = ref_ls.LabelSmoothing(size=12, padding_idx=1, smoothing=0.1)
        ls[name + "/logp"] = logp.numpy(); ls[name + "/target"] = np.array(tgt, np.int64)
        ls[name + "/loss"] = np.float32(crit(logp, torch.tensor(tgt)).item())
    crit0 = ref_ls.LabelSmoothing(size=12, padding_idx=1, smoothing=0.0)
    ls["nosmooth/logp"] = ls["mixed/logp"]; ls["nosmooth/target"] = ls["mixed/target"]
    ls["nosmooth/loss"] = np.float32(crit0(torch.from_numpy(ls["mixed/logp"]), torch.from_numpy(ls["mixed/target"])).item())
    np.savez(os.path.join(OUT, "label_smoothing.npz"), **ls)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
