"""A/B of matcher configs, interleaved repetitions (kernel only, 1024 x 2k pairs)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops
ctx = ops.context(0)
n, batch = 2048, 1024
des = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (batch + 1, n, 32), dtype=np.uint8)).cuda()
pq = torch.arange(1, batch + 1, dtype=torch.int32).cuda(); pt = pq - 1
idx = torch.empty((batch, n, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
keep = torch.empty((batch, n), dtype=torch.uint8, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
def fn():
    ops.check(ctx.lib.tvg_knn2_hamming256(ctx.handle, P(des), P(des), batch, n, n, n, P(pq), P(pt), None, None, 0.75,
                                          P(idx), P(dist), P(keep), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "knn")
cfgs = [(4, 3), (4, 11), (2, 11), (2, 3), (4, 7)]
res = {c: [] for c in cfgs}
for rep in range(6):
    for c in cfgs:
        ctx.set_matcher_config(c[0], c[1], 6)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); b.synchronize()
        res[c].append(a.elapsed_time(b) / 5)
for c in cfgs:
    v = np.array(res[c]); print(c, "ms median %.3f min %.3f max %.3f -> %.0f Gcmp/s" % (np.median(v), v.min(), v.max(), batch*n*n/np.median(v)/1e6))
