"""sort-only vs full NMS timing: nms_topk=1 makes the NMS loop stop after its first round."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dan_b200 import functional as F, synthetic
from dan_b200.utility import anchor_manipulator as am
dev = torch.device("cuda", 0)
ps = [0.1, 0.1, 0.2, 0.2]
enc = am.AnchorEncoder(0.4, 0.4, ps)
a_eval = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", (640, 640), border=0.))
an = np.stack([a.cpu().numpy() for a in a_eval[:4]], -1)
for faces in (20, 120, 300, 450):
    preds = [synthetic.gen_predictions(1000 + faces * 7 + i, an, max_faces=faces) for i in range(1)]
    cls = torch.from_numpy(np.stack([p[0] for p in preds])).to(dev)
    loc = torch.from_numpy(np.stack([p[1] for p in preds])).to(dev)
    for topk in (1, 64, 750):
        pp = F.postprocess_params(2, (640, 640), 0.01, 0, 5000, topk, 0.3, ps)
        ms_all = []
        for _ in range(6):
            det, ms = F.postprocess_batch(pp, cls, loc_pred=loc, anchors=a_eval[:4], profile=True)
            ms_all.append(ms)
        k = int(F._ws._buf[:4].view(torch.int32).cpu().numpy()[0])
        ms = np.median(np.array(ms_all[2:]), axis=0)
        print("K %5d nms_topk %4d: kept %4d  sort %.1f pairs %.1f resolve %.1f us" % (k, topk, int(det.counts[0, 0]), 1e3 * ms[1], 1e3 * ms[2], 1e3 * ms[3]))
