"""GPU micro-benchmark: duration of the fused sort+NMS kernel for ONE image as a function of the number of surviving
candidates K (and of the batch), via the *_profile entry point.  python tools/nms_scaling.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dan_b200 import functional as F, synthetic  # noqa: E402
from dan_b200.utility import anchor_manipulator as am  # noqa: E402

dev = torch.device("cuda", 0)
ps = [0.1, 0.1, 0.2, 0.2]
enc = am.AnchorEncoder(0.4, 0.4, ps)
a_eval = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", (640, 640), border=0.))
an = np.stack([a.cpu().numpy() for a in a_eval[:4]], -1)
pp = F.postprocess_params(2, (640, 640), 0.01, 0, 5000, 750, 0.3, ps)
for faces in (5, 20, 60, 120, 200, 300, 450):
    for batch in (1, 32):
        preds = [synthetic.gen_predictions(1000 + faces * 7 + i, an, max_faces=faces) for i in range(batch)]
        # force exactly `faces` planted faces by regenerating with min == max is not supported; report K instead
        cls = torch.from_numpy(np.stack([p[0] for p in preds])).to(dev)
        loc = torch.from_numpy(np.stack([p[1] for p in preds])).to(dev)
        ms_all = []
        for _ in range(6):
            det, ms = F.postprocess_batch(pp, cls, loc_pred=loc, anchors=a_eval[:4], profile=True)
            ms_all.append(ms)
        ws = F._ws._buf[:4 * batch].view(torch.int32).cpu().numpy()
        k = np.minimum(ws, 5000)
        ms = np.median(np.array(ms_all[2:]), axis=0)
        print("max_faces %4d batch %3d: K mean %6.0f max %5d  kept mean %5.0f max %4d  filter %.1f  sort %.1f  pairs %.1f  resolve %.1f us"
              % (faces, batch, k.mean(), k.max(), det.counts.float().mean().item(), det.counts.max().item(), 1e3 * ms[0], 1e3 * ms[1],
                 1e3 * ms[2], 1e3 * ms[3]))
