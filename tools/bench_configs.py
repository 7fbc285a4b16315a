"""The other BASELINE.json configurations (parity-test cases, not bench lines), timed once so that DESIGN.md can state how
the kernels behave away from the headline workload: config 3 (DAN pyramid, dual matcher, batch 64, 200-1000 dense tiny GT)
and config 4 (postprocess at 640^2 ... 1600^2).  Device resident, CUDA graph replay, rotating inputs.
python tools/bench_configs.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import functional as F, synthetic
from dan_b200.utility import anchor_manipulator as am

dev = torch.device("cuda", 0)
ps = [0.1, 0.1, 0.2, 0.2]


def timed(fn, sets, K=100):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    graphs, keep = [], []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                keep.append(fn(s))
            graphs.append(g)
    torch.cuda.current_stream().wait_stream(side)
    for k in range(8):
        graphs[k % len(graphs)].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        graphs[k % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


# ---- config 3: DAN anchors, do_dual_max_match, thresholds 0.35/0.35, batch 64, G2 dense tiny GT -------------------
enc = am.AnchorEncoder(0.35, 0.35, ps)
anchors = synthetic.build_anchors(enc, synthetic.pyramid_config("dan", (640, 640)))
N = anchors[0].numel()
B = 64
for mining in (False, True):
    params = F.encode_params(0.35, 0.35, ps, match_mining=mining)
    sets = []
    for r in range(4):
        cat, offs = synthetic.to_csr([synthetic.gen_dense_tiny(r * B + i) for i in range(B)])
        sets.append((torch.from_numpy(cat).to(dev), torch.from_numpy(offs).to(dev), int(offs[-1])))
    ms = timed(lambda s: F.encode_batch(params, *anchors[:4], anchors[4], s[0], s[1]), sets)
    mean_gt = np.mean([s[2] for s in sets]) / B
    flops = B * N * (20. * mean_gt + 30.)
    print("config 3 (%s): batch %d, %.0f GT/image: %.3f ms/batch -> %.0f img/s; %.1f Tflop/s of the dense N(20M+30) figure "
          "(culling skips empty pairs)" % ("mining" if mining else "dual", B, mean_gt, ms, B / (ms * 1e-3), flops / (ms * 1e-3) / 1e12))

# ---- config 4: postprocess at several input sizes, batch 8 ------------------------------------------------------------
e2 = am.AnchorEncoder(0.4, 0.4, ps)
for size in (640, 1024, 1600):
    a_eval = synthetic.build_anchors(e2, synthetic.pyramid_config("s3fd", (size, size), border=0.))
    an = np.stack([a.cpu().numpy() for a in a_eval[:4]], -1)
    n = an.shape[0]
    Bp = 8
    pp = F.postprocess_params(2, (size, size), 0.01, 0, 5000, 750, 0.3, ps)
    sets = []
    for r in range(3):
        preds = [synthetic.gen_predictions(r * Bp + i, an, size=(size, size), max_faces=300) for i in range(Bp)]
        sets.append((torch.from_numpy(np.stack([p[0] for p in preds])).to(dev), torch.from_numpy(np.stack([p[1] for p in preds])).to(dev)))
    ms = timed(lambda s: F.postprocess_batch(pp, s[0], loc_pred=s[1], anchors=a_eval[:4]), sets, K=60)
    print("config 4 @%d^2: %d anchors, batch %d: %.3f ms/batch -> %.0f img/s (%.0f GB/s of the 24 B/anchor input)" %
          (size, n, Bp, ms, Bp / (ms * 1e-3), Bp * n * 24 / (ms * 1e-3) / 1e9))
