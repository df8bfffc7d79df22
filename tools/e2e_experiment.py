import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from openwork_b200 import ops, synth
P = 1024
K = synth.K_640x480
des_h, kps_h = bench.make_batch(P)
des_pin, kps_pin = torch.from_numpy(des_h).pin_memory(), torch.from_numpy(kps_h).pin_memory()
words_pin = torch.from_numpy(synth.sample_words(1024, 1).view(np.int32)).pin_memory()
des, kps, words = des_pin.cuda(), kps_pin.cuda(), words_pin.cuda()
def T(name, fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps * 1e3
    print("%-40s %7.3f ms" % (name, dt), flush=True)
def h2d():
    des.copy_(des_pin, non_blocking=True); kps.copy_(kps_pin, non_blocking=True)
T("H2D 201 MB pinned", h2d)
pq = torch.arange(1, 2 * P, 2, dtype=torch.int32); pt = pq - 1; so = torch.arange(0, P + 1, dtype=torch.int32)
pqd, ptd, sod = pq.cuda(), pt.cuda(), so.cuda()
outs = ("m_count", "E", "n_inl", "Rt", "poses", "status")
ctx = ops.context(0)
o = [None]
def dev():
    o[0] = ops.process_pairs_device(des, kps, K, words=words, pair_q=pqd, pair_t=ptd, seq_off=sod, outputs=outs, out=o[0])
T("device whole batch", dev)
for nchunk in (2, 4, 8):
    per = P // nchunk
    oc = [None] * nchunk
    sl = [(pqd[i*per:(i+1)*per].contiguous(), ptd[i*per:(i+1)*per].contiguous(), torch.arange(0, per+1, dtype=torch.int32).cuda()) for i in range(nchunk)]
    def devc():
        for i in range(nchunk):
            oc[i] = ops.process_pairs_device(des, kps, K, words=words, pair_q=sl[i][0], pair_t=sl[i][1], seq_off=sl[i][2], outputs=outs, out=oc[i])
    T("device in %d chunks" % nchunk, devc)
ho = [None]
for c in (128, 256, 512, 1024):
    ctx.set_host_chunk(c)
    def e2e():
        ho[0], a, b = ops.process_pairs_host(des_pin, kps_pin, K, words=words_pin, pair_q=pq, pair_t=pt, seq_off=so, outputs=outs, out=ho[0])
    T("e2e host entry chunk %d" % c, e2e)
