"""The matcher alone over bench.py's batch (1024 pairs of 2048 x 2048): ms per tvg_knn2_hamming256 call.
Used for the DRAM-traffic captures in profiles/r1_tensor_matcher_dram_traffic.md (ncu --cache-control none)."""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch, ctypes as C
import bench
from openwork_b200 import ops, synth
P = 1024
des_h, kps_h = bench.make_batch(P)
des = torch.from_numpy(des_h).cuda()
pq = torch.arange(1, 2 * P, 2, dtype=torch.int32).cuda(); pt = pq - 1
idx = torch.empty((P, 2048, 2), dtype=torch.int32, device="cuda"); dst = torch.empty_like(idx)
keep = torch.empty((P, 2048), dtype=torch.uint8, device="cuda")
ctx = ops.context(0)
if os.environ.get("TVG_TC_L2_MB"):
    ctx.set_matcher_l2_budget(int(os.environ["TVG_TC_L2_MB"]) << 20)
def knn():
    ops.check(ctx.lib.tvg_knn2_hamming256(ctx.handle, C.c_void_p(des.data_ptr()), C.c_void_p(des.data_ptr()), P, 2048, 2048, 2048,
              C.c_void_p(pq.data_ptr()), C.c_void_p(pt.data_ptr()), None, None, 0.75, C.c_void_p(idx.data_ptr()),
              C.c_void_p(dst.data_ptr()), C.c_void_p(keep.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "knn")
for _ in range(3): knn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): knn()
e1.record(); torch.cuda.synchronize()
print("l2_budget_mb", os.environ.get("TVG_TC_L2_MB"), "knn ms", e0.elapsed_time(e1) / 10, flush=True)
