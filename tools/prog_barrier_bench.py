import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtn_b200 import _lib as L
L.lib()
x = torch.randn(64, 512, device="cuda"); a = torch.ones(512, device="cuda"); b = torch.zeros(512, device="cuda")
y = torch.empty(64, 512, device="cuda", dtype=torch.float16)
for rows in (8, 64):
    prog = L.StepProgram()
    with prog.record():
        for _ in range(128):
            L.layernorm(x[:rows], a, b, 1e-6, out_f16=y[:rows])
    prog.launch(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): prog.launch()
    e1.record(); torch.cuda.synchronize()
    print("rows=%d: %d stages, %.2f us per stage" % (rows, prog.n, e0.elapsed_time(e1) * 1e3 / 10 / prog.n))
