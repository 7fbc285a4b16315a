"""Timing of the hard-negative mining row at the config-2 shape: CUDA path (device resident, CUDA
events, rotating inputs > L2) next to the numpy oracle on one host core.  python tools/bench_hnm.py [B] [N]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import functional as F
from oracle import reference_np as R

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 34125
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
sets = []
for r in range(8):
    cls = rng.normal(0., 3., size=(B, N, 2)).astype(np.float32)
    tg = rng.choice([-1, 0, 1], p=[0.05, 0.948, 0.002], size=(B, N)).astype(np.int64)
    loc = rng.normal(size=(B * N, 4)).astype(np.float32)
    lt = rng.normal(size=(B, N, 4)).astype(np.float32)
    sets.append((cls, loc, tg, lt))
dsets = [tuple(torch.from_numpy(a).to(dev) for a in s) for s in sets]
outs = [None] * len(dsets)
def run(k):
    cls, loc, tg, lt = dsets[k % len(dsets)]
    return F.hard_negative_mining(cls, loc, tg, lt, B, 3., 2)
for k in range(16):
    run(k)
torch.cuda.synchronize()
# one CUDA graph per input set (the python wrapper + 6 launches cost more host time than the kernels take)
graphs = []
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for k in range(len(dsets)):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            outs[k] = run(k)
        graphs.append(g)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
def run(k):
    graphs[k % len(graphs)].replay()
for k in range(16):
    run(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 200
e0.record()
for k in range(K):
    run(k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
alg = B * N * (8 + 8 + 1) + 0   # logits + labels in, mask out (+ compacted rows, small)
print("GPU: %.1f us per batch of %d x %d  -> %.0f img/s, %.0f GB/s algorithmic (logits+labels+mask)" % (1e3 * ms, B, N, B / (ms * 1e-3), alg / (ms * 1e-3) / 1e9))
t0 = time.perf_counter()
n_img = 4
cls, loc, tg, lt = sets[0]
R.mining_hard_neg(n_img, cls[:n_img], loc[:n_img * N], tg[:n_img], None, lt[:n_img])
dt = time.perf_counter() - t0
print("CPU oracle (numpy, 1 core): %.1f ms per image -> %.0f img/s" % (1e3 * dt / n_img, n_img / dt))
