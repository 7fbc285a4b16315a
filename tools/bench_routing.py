"""Timing of the DynamicAnchorRouting evaluation row at the DAN 640x640 shape (6 layers, 34 125 anchors):
CUDA path (device resident, CUDA graph, rotating inputs) next to the reference's own compiled functor / the C++ port on
one host core.  python tools/bench_routing.py [batch]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import functional as F, synthetic
from oracle import native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = dict(feat_heights=[160, 80, 40, 20, 10, 5], feat_widths=[160, 80, 40, 20, 10, 5], depths=[1] * 6, strides=[4, 8, 16, 32, 64, 128])
dev = torch.device("cuda", 0)
imgs = [synthetic.gen_routing(i, cfg["feat_heights"], cfg["feat_widths"], cfg["depths"], cfg["strides"]) for i in range(B)]
N = imgs[0][0].shape[0]
layers = F.routing_layers(cfg["feat_heights"], cfg["feat_widths"], cfg["depths"], cfg["strides"])
sets = []
for r in range(8):
    order = [(i + r) % B for i in range(B)]
    sets.append(tuple(torch.from_numpy(np.stack([imgs[i][k] for i in order])).to(dev) for k in range(4)))
for s in sets:
    F.dynamic_anchor_routing_eval(layers, *s)
torch.cuda.synchronize()
graphs, keep = [], []
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep.append(F.dynamic_anchor_routing_eval(layers, *s))
        graphs.append(g)
torch.cuda.current_stream().wait_stream(side)
for k in range(16):
    graphs[k % 8].replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 400
e0.record()
for k in range(K):
    graphs[k % 8].replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
alg = B * N * (16 + 16 + 4 + 4 + 4 + 16)      # boxes, offsets, label, mask in; mask, boxes out
print("GPU: %.1f us per batch of %d images x %d anchors -> %.0f img/s, %.0f GB/s algorithmic (60 B/anchor)" %
      (1e3 * ms, B, N, B / (ms * 1e-3), alg / (ms * 1e-3) / 1e9))
impl = "reference" if native.have_reference_dar() else "port"
t0 = time.perf_counter()
n_img = 8
for i in range(n_img):
    a, t, lab, m = imgs[i]
    off = 0
    for H, W, D, S in zip(cfg["feat_heights"], cfg["feat_widths"], cfg["depths"], cfg["strides"]):
        n = H * W * D
        native.dynamic_anchor_routing(a[off:off + n], t[off:off + n], lab[off:off + n], m[off:off + n], H, W, D, S, 640, 640, impl=impl)
        off += n
dt = time.perf_counter() - t0
print("CPU %s functor (1 core): %.3f ms per image -> %.0f img/s" % (impl, 1e3 * dt / n_img, n_img / dt))
