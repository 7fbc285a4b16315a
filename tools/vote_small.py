"""bbox_vote on a small input against the oracle (debugging aid for compute-sanitizer runs): python tools/vote_small.py [n]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import functional as F, synthetic
from oracle import reference_np as R
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
dev = torch.device("cuda", 0)
d = synthetic.gen_vote_dets(7, n, max(n // 25, 1), 0.5)[:n]
out, cnt = F.bbox_vote_batch(torch.from_numpy(d[None]).to(dev), None, 0.3, 750)
torch.cuda.synchronize()
ref = R.bbox_vote(d)
k = int(cnt[0])
print("groups", k, "ref", len(ref), "equal", k == len(ref) and np.array_equal(out[0, :k].cpu().numpy().view(np.uint32), ref.view(np.uint32)))
