#!/bin/bash
# Timings of the SURVEY 8(f) rows on one GPU (CUDA path next to the CPU reference / restatement), written to gpurun_out/.
mkdir -p gpurun_out
{
echo "== f3 hard-negative mining (tools/bench_hnm.py)"; timeout 300 python tools/bench_hnm.py
echo "== f1 DynamicAnchorRouting, evaluation branch (tools/bench_routing.py)"; timeout 300 python tools/bench_routing.py
echo "== f2 bbox_vote (tools/bench_vote.py)"; timeout 300 python tools/bench_vote.py
echo "== kernel durations, one launch each (ncu --metrics gpu__time_duration.sum, cold L2)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hnm_|route_|vote|face_|handoff" -c 40 --csv python - <<'PY' 2>/dev/null | grep -E "hnm_|route_|vote|face_|handoff" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq | cut -c1-140
import sys, numpy as np, torch
sys.path.insert(0, ".")
from dan_b200 import functional as F, synthetic
from dan_b200.utility import eval_merge, input_handoff
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
B, N = 32, 34125
cls = torch.from_numpy(rng.normal(0, 3, (B, N, 2)).astype(np.float32)).to(dev)
tg = torch.from_numpy(rng.choice([-1, 0, 1], p=[0.05, 0.948, 0.002], size=(B, N)).astype(np.int64)).to(dev)
loc = torch.from_numpy(rng.normal(size=(B * N, 4)).astype(np.float32)).to(dev)
lt = torch.from_numpy(rng.normal(size=(B, N, 4)).astype(np.float32)).to(dev)
F.hard_negative_mining(cls, loc, tg, lt, B, 3., 2)
cfg = dict(h=[160, 80, 40, 20, 10, 5], w=[160, 80, 40, 20, 10, 5], d=[1] * 6, s=[4, 8, 16, 32, 64, 128])
imgs = [synthetic.gen_routing(i, cfg["h"], cfg["w"], cfg["d"], cfg["s"]) for i in range(B)]
F.dynamic_anchor_routing_eval(F.routing_layers(cfg["h"], cfg["w"], cfg["d"], cfg["s"]),
                              *[torch.from_numpy(np.stack([im[k] for im in imgs])).to(dev) for k in range(4)])
det = np.stack([synthetic.gen_vote_dets(i, 6750, 120, 0.6)[:6000] for i in range(B)])
F.bbox_vote_batch(torch.from_numpy(det).to(dev), None, 0.3, 750)
eval_merge.detect_face_select(torch.from_numpy(rng.uniform(0, 640, (34125, 4)).astype(np.float32)).to(dev),
                              torch.from_numpy(rng.uniform(0, 1, 34125).astype(np.float32)).to(dev), 1.0)
gts = [synthetic.gen_faces(i, 50) for i in range(B)]
cat, offs = synthetic.to_csr(gts)
input_handoff.prepare_gt_batch(torch.from_numpy(cat).to(dev), torch.from_numpy(offs).to(dev),
                               torch.full((B, 2), 640., device=dev), (640, 640), trim=False)
torch.cuda.synchronize()
PY
} > gpurun_out/next_rows_timing.txt 2>&1
cat gpurun_out/next_rows_timing.txt
