"""Front-end throughput: ops.extract_points on a batch of frames vs the reference's cv2 calls on the host."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops
from oracle import frontend as fe
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = [fe.synthetic_image(i) for i in range(8)]
imgs = np.stack([base[i % 8] for i in range(B)])
d = torch.from_numpy(imgs).cuda()
for _ in range(3):
    k, de, n = ops.extract_points(d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    k, de, n = ops.extract_points(d)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
t = time.perf_counter()
for i in range(8):
    fe.extract_points_cv2(base[i])
cpu_ms = (time.perf_counter() - t) / 8 * 1e3
print({"frames": B, "gpu_ms_per_batch": ms, "gpu_frames_per_s": B / ms * 1e3, "cv2_ms_per_frame": cpu_ms,
       "cv2_frames_per_s_one_process": 1e3 / cpu_ms, "kps_per_frame": float(n.float().mean())})
