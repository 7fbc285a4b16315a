"""Timing of bbox_vote: CUDA path (batch of images, one CTA each, device resident) next to the numpy
restatement on one host core (the reference's own function is numpy too and runs at the same speed).
python tools/bench_vote.py [batch] [dets per image]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import functional as F, synthetic
from oracle import reference_np as R

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6750       # 6 scales / flips x 1125 (eval_sfd.py:112, multi_scale_test)
dev = torch.device("cuda", 0)
dets = [synthetic.gen_vote_dets(i, n, 120, 0.6) for i in range(B)]
cap = max(d.shape[0] for d in dets)
batch = np.zeros((B, cap, 5), np.float32)
for i, d in enumerate(dets):
    batch[i, :d.shape[0]] = d
counts = torch.tensor([d.shape[0] for d in dets], dtype=torch.int32, device=dev)
x = torch.from_numpy(batch).to(dev)
for _ in range(3):
    out, cnt = F.bbox_vote_batch(x, counts, 0.3, 750)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    out, cnt = F.bbox_vote_batch(x, counts, 0.3, 750)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("GPU: %.3f ms per batch of %d images x %d detections -> %.0f img/s (groups per image: %.0f)" %
      (ms, B, cap, B / (ms * 1e-3), float(cnt.float().mean())))
t0 = time.perf_counter()
for d in dets[:4]:
    R.bbox_vote(d)
dt = (time.perf_counter() - t0) / 4
print("CPU numpy restatement (1 core): %.1f ms per image -> %.1f img/s" % (1e3 * dt, 1 / dt))
