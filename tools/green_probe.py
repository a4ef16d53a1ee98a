"""Feasibility probe: SM partitioning with CUDA green contexts -- create two disjoint SM groups, run this library's
kernels on streams of either group (eager and inside ONE captured CUDA graph), and time a wide GEMM on the big group
with / without small GEMMs running concurrently on the small group."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuda.bindings import driver as cu
from mtn_b200 import _lib as L

def ck(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    assert err == cu.CUresult.CUDA_SUCCESS, err
    return rest[0] if len(rest) == 1 else rest

torch.cuda.init(); torch.zeros(1, device="cuda")
L.lib()
dev = ck(cu.cuDeviceGet(0))
res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs:", res.sm.smCount)
small = int(sys.argv[1]) if len(sys.argv) > 1 else 16
groups, nb, rem = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, small))
print("split: group of", groups[0].sm.smCount, "SMs, remaining", rem.sm.smCount)
streams = {}
for name, r in (("small", groups[0]), ("big", rem)):
    desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
    g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    s = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
    streams[name] = torch.cuda.ExternalStream(int(s))
    print(name, "stream", hex(int(s)))

def gemm_set(M, N, K):
    A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    o = torch.empty(M, N, device="cuda", dtype=torch.float16)
    return A, W, o
big = gemm_set(16384, 6144, 512)
sm_ = [gemm_set(4096, 512, 512) for _ in range(4)]
torch.cuda.synchronize()

def run_big():
    L.linear(big[0], big[1], None, out_f16=big[2])
def run_small():
    for a, w, o in sm_:
        L.linear(a, w, None, out_f16=o)

ref = torch.empty_like(big[2]); L.linear(big[0], big[1], None, out_f16=ref); torch.cuda.synchronize()
for name in ("big", "small"):
    with torch.cuda.stream(streams[name]):
        run_big()
    streams[name].synchronize()
    print("eager on", name, "group: equal to default-stream result:", bool(torch.equal(ref, big[2])))

def timed(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

# one graph: fork from the capturing stream into both groups, join
def body(use_small, use_big, reps=5):
    cur = torch.cuda.current_stream()
    for st in streams.values(): st.wait_stream(cur)
    if use_big:
        with torch.cuda.stream(streams["big"]):
            for _ in range(reps): run_big()
    if use_small:
        with torch.cuda.stream(streams["small"]):
            for _ in range(reps * 6): run_small()
    for st in streams.values(): cur.wait_stream(st)

for us, ub in ((False, True), (True, False), (True, True)):
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body(us, ub)
        t = timed(g.replay)
        print("graph small=%s big=%s: %.1f us per replay" % (us, ub, t))
    except Exception as e:
        print("graph capture failed:", repr(e)[:300])
        torch.cuda.synchronize()
        for rep in range(2):
            t = timed(lambda: body(us, ub), n=5)
        print("eager small=%s big=%s: %.1f us" % (us, ub, t))
# baseline: both on ordinary streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def body_plain(reps=5):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        for _ in range(reps): run_big()
    with torch.cuda.stream(s2):
        for _ in range(reps * 6): run_small()
    cur.wait_stream(s1); cur.wait_stream(s2)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    body_plain()
print("graph, ordinary streams, both: %.1f us per replay" % timed(g.replay))
