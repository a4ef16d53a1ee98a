"""Debug: the attention-core case (2, 2, 300, 300, 64, causal) of tests/test_gpu_kernels.py, per (batch, head, warp) error."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib as L
L.lib()
B, h, Lq, Lk, dk = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (2, 2, 300, 300, 64))]
kind = sys.argv[6] if len(sys.argv) > 6 else "causal"
g = torch.Generator().manual_seed(B * 1000 + Lq * 10 + Lk + dk)
d = h * dk
q = (torch.randn(B, Lq, d, generator=g) * 1.5).half()
k = (torch.randn(B, Lk, d, generator=g) * 1.5).half()
v = torch.randn(B, Lk, d, generator=g).half()
if kind == "causal":
    mask = torch.tril(torch.ones(Lq, Lk, dtype=torch.bool)).expand(B, Lq, Lk).clone()
    mask[B - 1, :, Lk - 3:] = False
elif kind == "holes":
    mask = torch.rand(B, Lq, Lk, generator=g) > 0.3
    mask[:, :, Lk // 2 + 5:] = False
else:
    mask = None
bits = L.mask_pack(mask.cuda()) if mask is not None else None
qd = q.cuda().view(-1, d).contiguous(); kd = k.cuda().view(-1, d).contiguous(); vd = v.cuda().view(-1, d).contiguous()
out = torch.zeros(B * Lq, d, device="cuda", dtype=torch.float16)
chk = torch.zeros(B * Lq, d, device="cuda", dtype=torch.float16)
L.attn_core(qd, kd, vd, B, h, Lq, Lk, dk, chk, mask_bits=bits, _check_kernel=True)
L.attn_core(qd, kd, vd, B, h, Lq, Lk, dk, out, mask_bits=bits)
torch.cuda.synchronize()
e = (out.float() - chk.float()).abs().view(B, Lq, h, dk).amax(3)   # [B, Lq, h]
bad = 0
for b in range(B):
    for hd in range(h):
        rows = (e[b, :, hd] > 5e-3).nonzero().flatten().tolist()
        if rows:
            bad += 1
            if bad <= 12:
                print("b=%d h=%d bad rows %d: %s ... max %.3f" % (b, hd, len(rows), rows[:12], float(e[b, :, hd].max())))
print("DBG=%s PT=%s: %d bad (b,h) of %d, max err %.4f" % (os.environ.get("MTN_B200_ATTN_DBG"), os.environ.get("MTN_B200_ATTN_PTMEM"),
                                                       bad, B * h, float(e.max())))
