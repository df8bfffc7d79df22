"""One C2 / C3 frame pair through the device pipeline a few times (for ncu launch lists)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth  # noqa: E402

n, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
K = synth.K_640x480
lk, ld, ck, cd = synth.make_pair_large(n, 1)
des = torch.from_numpy(np.stack([ld, cd])).cuda()
kps = torch.from_numpy(np.stack([lk, ck])).cuda()
words = torch.from_numpy(synth.sample_words(H, 2).view(np.int32)).cuda()
pq = torch.tensor([1], dtype=torch.int32).cuda()
pt = torch.tensor([0], dtype=torch.int32).cuda()
so = torch.tensor([0, 1], dtype=torch.int32).cuda()
out = None
for _ in range(6):
    out = ops.process_pairs_device(des, kps, K, words=words, pair_q=pq, pair_t=pt, seq_off=so, out=out)
torch.cuda.synchronize()
print("ok", int(out["n_inl"][0]))
