"""In-kernel phase timeline of the fused attention-site kernel (csrc/site_fused.cu v2): clock64 stamps of the role leaders,
averaged over the CTAs, relative to the kernel body start.  Usage: python tools/site_phases.py [out_file]"""
import ctypes as C
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtn_b200 import _lib as L
lib = L.lib()
lib.mtn_debug_site_stamps.restype = C.c_int
lib.mtn_debug_site_stamps.argtypes = [C.c_void_p]
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
NAMES = {1: "phase-1 loads issued", 2: "phase-1 MMAs issued", 3: "Q accumulator complete", 4: "Q tiles in smem (drain done)",
         5: "first S tile ready", 6: "engine 0 heads done", 10: "engine 1 heads done", 11: "own O tiles complete",
         12: "  eng 0: P of head 0 stored", 13: "  eng 0: P V of head 0 complete", 14: "  eng 0: O of head 0 in smem",
         15: "  eng 0: S of head 1 observed", 7: "peer O tiles received", 8: "Y accumulator complete", 9: "epilogue done"}
d, h = 512, 8
torch.manual_seed(0)
for (B, Lq, Lk) in [(32, 256, 64), (32, 256, 256)]:
    x = torch.randn(B * Lq, d, device="cuda")
    kv = torch.randn(B * Lk, 2 * d, device="cuda").half()
    wq, wo = (torch.randn(d, d, device="cuda") * 0.05).half(), (torch.randn(d, d, device="cuda") * 0.05).half()
    bq, bo = torch.randn(d, device="cuda") * 0.1, torch.randn(d, device="cuda") * 0.1
    lens = torch.randint(Lk // 2, Lk + 1, (B,))
    bits = L.mask_pack((torch.arange(Lk)[None, :] < lens[:, None]).view(B, 1, Lk).cuda())
    xn = torch.randn(B * Lq, d, device="cuda").half()
    grid = 2 * B * ((Lq + 127) // 128)
    st = torch.zeros(grid * 16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        L.attn_site_fused(xn, x, wq, bq, wo, bo, kv, 0, d, B, h, Lq, Lk, mask_bits=bits)
    torch.cuda.synchronize()
    assert lib.mtn_debug_site_stamps(C.c_void_p(st.data_ptr())) == 0
    L.attn_site_fused(xn, x, wq, bq, wo, bo, kv, 0, d, B, h, Lq, Lk, mask_bits=bits)
    torch.cuda.synchronize()
    assert lib.mtn_debug_site_stamps(None) == 0
    t = st.view(grid, 16).cpu().double()
    rel = (t - t[:, :1])
    print("B=%d Lq=%d Lk=%d  (%d CTAs; cycles after the kernel body start: mean / max over CTAs)" % (B, Lq, Lk, grid), file=out)
    for i in (1, 2, 3, 4, 5, 12, 13, 14, 15, 6, 10, 11, 7, 8, 9):
        print("  %-32s %8.0f %8.0f" % (NAMES[i], rel[:, i].mean(), rel[:, i].max()), file=out)
out.flush()
