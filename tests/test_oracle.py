"""CPU tier: pin oracle/mtn_oracle.py against the reference's known answers, the
committed golden tensors (made by the unmodified reference, oracle/make_golden.py)
and -- when /root/reference is present -- the live reference."""
import numpy as np
import pytest
import torch

import mtn_oracle as O
import golden_util as G


@pytest.fixture(autouse=True)
def _single_thread():
    """The fixtures were generated single-threaded; MKL's multi-threaded GEMM
    changes the fp32 reduction order (~1e-6), which would hide a real 1-ulp
    restatement bug behind a tolerance.  One thread => bit-exact comparisons."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_kat_layernorm():
    z = G.load("kat.npz")
    y = O.layer_norm(G.t(z["ln_in"]), torch.ones(4), torch.zeros(4))
    assert torch.equal(y, G.t(z["ln_out"]))
    # SURVEY 8c known answer (NOT nn.LayerNorm's +-1.3416)
    assert np.allclose(y.numpy(), [-1.1618942, -0.3872980, 0.3872980, 1.1618942], atol=1e-6)


@pytest.mark.parametrize("name,p_expect,o_expect", [
    ("allmasked", [1 / 3, 1 / 3, 1 / 3], [3.0, 4.0]),
    ("lastmasked", [0.66976154, 0.33023846, 0.0], [1.6604769, 2.6604769]),
])
def test_kat_attention(name, p_expect, o_expect):
    z = G.load("kat.npz")
    o, p = O.attention(G.t(z["attn_q"]), G.t(z["attn_k"]), G.t(z["attn_v"]),
                       G.t(z["attn_%s_mask" % name]))
    assert torch.equal(o, G.t(z["attn_%s_o" % name]))
    assert torch.equal(p, G.t(z["attn_%s_p" % name]))
    assert np.allclose(p.numpy()[0], p_expect, atol=1e-6)
    assert np.allclose(o.numpy()[0], o_expect, atol=1e-6)


def _site_sd(z):
    sd = {}
    for k, v in z.items():
        if k.startswith("att/"):
            sd["att." + k[4:]] = G.t(v)
        elif k.startswith("ff/"):
            sd["ff." + k[3:]] = G.t(v)
        elif k.startswith("sub/"):
            sd["sub." + k[4:]] = G.t(v)
    return sd


def test_site_golden():
    z = G.load("site_d128.npz")
    sd, h = _site_sd(z), int(z["h"])
    x, mem = G.t(z["x"]), G.t(z["mem"])
    km, cm = G.t(z["kmask"]), G.t(z["cmask"])
    with torch.no_grad():
        y = O.sublayer(sd, "sub.", x, lambda t: O.mha(sd, "att.", h, t, mem, mem, km))
        assert torch.equal(y, G.t(z["y_cross"]))
        y = O.sublayer(sd, "sub.", x, lambda t: O.mha(sd, "att.", h, t, t, t, cm))
        assert torch.equal(y, G.t(z["y_self"]))
        y = O.sublayer(sd, "sub.", x, lambda t: O.mha(sd, "att.", h, t, mem, mem, None))
        assert torch.equal(y, G.t(z["y_nomask"]))
        y = O.sublayer(sd, "sub.", x, lambda t: O.ffn(sd, "ff.", t))
        assert torch.equal(y, G.t(z["y_ffn"]))


@pytest.mark.parametrize("name", ["cfg1.npz", "cfg1b.npz"])
def test_cfg1_golden(name):
    zm = G.load("cfg1.npz")
    z = G.load(name)
    sd = G.state_dict_from(zm, 128)
    out, ae = O.forward(sd, G.CFG1, G.t(z["query"]), G.t(z["his"]), G.t(z["cap"]), G.t(z["trg"]),
                        [G.t(z["ft0"]), G.t(z["ft1"])])
    assert torch.equal(out, G.t(z["out"]))
    assert torch.equal(ae[0], G.t(z["ae0"])) and torch.equal(ae[1], G.t(z["ae1"]))
    assert torch.equal(O.generator(sd, out).argmax(-1), G.t(z["argmax"]))
    m = O.make_masks(G.t(z["query"]), G.t(z["his"]), G.t(z["cap"]), G.t(z["trg"]),
                     [G.t(z["ft0"]), G.t(z["ft1"])], 1)
    assert torch.equal(m["trg_mask"], G.t(z["trg_mask"]))
    assert torch.equal(m["fts_mask"][0], G.t(z["fts_mask0"]))
    assert int((G.t(z["trg_y"]) != 1).sum()) == int(z["ntokens"])
    if name == "cfg1.npz":
        # SURVEY 8c model-level known answers
        assert abs(float(out.abs().sum()) - 1645.1214289) < 2e-3
        assert np.allclose(out[0, 0, :4].numpy(), [-0.9352665, 1.7873830, -0.7340517, -0.3389700],
                           atol=1e-5)
        assert z["argmax"].tolist() == [[4, 71, 40, 39, 39, 4, 34, 53],
                                        [53, 43, 53, 53, 14, 53, 53, 53]]
        assert int(z["ntokens"]) == 13


def test_mini512_golden():
    z = G.load("mini512.npz")
    cfg, sd = G.seeded_state_dict(z)
    out, ae = O.forward(sd, cfg, G.t(z["query"]), G.t(z["his"]), G.t(z["cap"]), G.t(z["trg"]),
                        [G.t(z["ft0"]), G.t(z["ft1"])])
    assert torch.equal(out, G.t(z["out"]))
    assert torch.equal(ae[0], G.t(z["ae0"])) and torch.equal(ae[1], G.t(z["ae1"]))


def test_greedy_golden():
    z = G.load("greedy.npz")
    cfg, sd = G.seeded_state_dict(z)
    fts = [G.t(z["ft0"]), G.t(z["ft1"])]
    ys = O.greedy_decode(sd, cfg, G.t(z["query"]), G.t(z["his"]), G.t(z["cap"]), fts, max_len=8)
    assert ys.tolist() == z["tokens"].tolist()          # token-exact vs the reference


def test_sinusoid_matches_reference_table():
    import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    mtn, _ = ref_loader.load()
    pe = mtn.PositionalEncoding(128, 0.0).pe
    assert torch.equal(pe, O.sinusoid_pe(128))


def test_live_reference_bitexact():
    """Oracle vs the unmodified reference on a fresh random model + ragged batch."""
    import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    import warnings
    warnings.simplefilter("ignore")
    mtn, du = ref_loader.load()
    cfg = {"N": 2, "d_model": 64, "d_ff": 256, "h": 2, "vocab": 50, "ft_sizes": [40, 24],
           "auto_encoder_ft": "query", "diff_encoder": True}
    sd = O.init_state_dict(cfg, 3)
    model = mtn.make_model(50, 50, N=2, d_model=64, d_ff=256, h=2, ft_sizes=[40, 24],
                           diff_encoder=True, auto_encoder_ft="query").eval()
    model.load_state_dict(sd, strict=True)
    inp = O.synth_inputs(cfg, B=4, Q=7, C=9, H=15, T=6, Lv=[11, 5], seed=2)
    b = ref_loader.make_cpu_batch(inp["query"], inp["his"], inp["cap"], inp["trg"], inp["trg_y"],
                                  inp["fts"])
    with torch.no_grad():
        out_r, ae_r = model.forward(b)
    out, ae = O.forward(sd, cfg, inp["query"], inp["his"], inp["cap"], inp["trg"], inp["fts"])
    assert torch.equal(out, out_r)
    assert all(torch.equal(a, b_) for a, b_ in zip(ae, ae_r))
    # state_dict key contract
    assert set(sd.keys()) == set(model.state_dict().keys())


@pytest.mark.parametrize("name,smoothing", [("mixed", 0.1), ("quirk_pad_row0_only", 0.1), ("nopad", 0.1), ("nosmooth", 0.0)])
def test_label_smoothing_golden(name, smoothing):
    """oracle.label_smoothing_loss vs the reference's LabelSmoothing (label_smoothing.py), including the quirk that a
    lone padding target in row 0 is NOT zeroed."""
    z = G.load("label_smoothing.npz")
    loss = O.label_smoothing_loss(G.t(z[name + "/logp"]), G.t(z[name + "/target"]), 12, 1, smoothing)
    assert abs(float(loss) - float(z[name + "/loss"])) <= 1e-6 * max(1.0, abs(float(z[name + "/loss"])))
