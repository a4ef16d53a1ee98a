"""Where does the fused site kernel differ from the launch sequence (rounding-level)?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib as L
L.lib()
torch.manual_seed(0)
for (B, Lq, Lk, d, h) in [(4, 64, 512, 512, 8), (2, 128, 64, 512, 8), (5, 1, 256, 512, 8)]:
    dk = d // h
    dev = "cuda"
    xn = torch.randn(B * Lq, d, device=dev).half()
    x = torch.randn(B * Lq, d, device=dev) * 2
    kv = (torch.randn(B * Lk, 2 * d, device=dev) * 1.2).half()
    wq, wo = (torch.randn(d, d, device=dev) * 0.05).half(), (torch.randn(d, d, device=dev) * 0.05).half()
    bq, bo = torch.randn(d, device=dev) * 0.1, torch.randn(d, device=dev) * 0.1
    qb = torch.empty(B * Lq, d, device=dev, dtype=torch.float16)
    ob = torch.empty(B * Lq, d, device=dev, dtype=torch.float16)
    x_seq = x.clone()
    L.linear(xn, wq, bq, out_f16=qb)
    L.attn_core(qb, kv[:, :d], kv[:, d:], B, h, Lq, Lk, dk, ob)
    L.linear(ob, wo, bo, addend=x_seq, out_f32=x_seq)
    x_f = x.clone()
    L.attn_site_fused(xn, x_f, wq, bq, wo, bo, kv, 0, d, B, h, Lq, Lk)
    # variant: zero Wo-bias path check -- identity Wo exposes O itself
    eye = torch.eye(d, device=dev).half()
    z = torch.zeros(d, device=dev)
    o_seq = torch.zeros(B * Lq, d, device=dev); L.linear(ob, eye, z, addend=o_seq, out_f32=o_seq)
    o_f = torch.zeros(B * Lq, d, device=dev); L.attn_site_fused(xn, o_f, wq, bq, eye, z, kv, 0, d, B, h, Lq, Lk)
    torch.cuda.synchronize()
    dd = (x_f - x_seq)
    do = (o_f - o_seq)
    print((B, Lq, Lk, d, h), "x: differing %d of %d, max %.3e; per column half: %s; rows differing %d of %d" %
          (int((dd != 0).sum()), dd.numel(), float(dd.abs().max()), [int((dd[:, :d // 2] != 0).sum()), int((dd[:, d // 2:] != 0).sum())],
           int((dd != 0).any(1).sum()), dd.shape[0]))
    print("   O (identity Wo): differing %d of %d, max %.3e; per head: %s" %
          (int((do != 0).sum()), do.numel(), float(do.abs().max()), [int((do[:, i * dk:(i + 1) * dk] != 0).sum()) for i in range(h)]))
