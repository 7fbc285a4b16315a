"""Per-phase SM-cycle timestamps of the per-list NMS kernel (CTA of list 0).  Builds a -DDAN_PHASE_TIMING copy of the
library in /tmp, runs one image with many candidates, prints phase durations in microseconds at the SM clock."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import build
lib_dbg = "/tmp/libdan_b200_phase.so"
cmd = ["nvcc"] + build.NVCC_FLAGS + ["-shared", "-DDAN_PHASE_TIMING"] + os.environ.get("DAN_EXTRA_NVCC_FLAGS", "").split() + [os.path.join(build.CSRC, s) for s in build.SOURCES] + ["-o", lib_dbg]
subprocess.run(cmd, check=True)
from dan_b200 import _lib
if "--prod" not in sys.argv:
    _lib.LIB_PATH = lib_dbg
from dan_b200 import functional as F, synthetic
from dan_b200.utility import anchor_manipulator as am
L = _lib.lib()
if "--prod" not in sys.argv:
    L.dan_debug_phases.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
dev = torch.device("cuda", 0)
ps = [0.1, 0.1, 0.2, 0.2]
enc = am.AnchorEncoder(0.4, 0.4, ps)
a_eval = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", (640, 640), border=0.))
an = np.stack([a.cpu().numpy() for a in a_eval[:4]], -1)
pp = F.postprocess_params(2, (640, 640), 0.01, 0, 5000, 750, 0.3, ps)
for faces, seed in ((60, 1420), (300, 3100), (300, 7), (450, 4150)):
    cls, loc, _ = synthetic.gen_predictions(seed, an, max_faces=faces)
    cls = torch.from_numpy(cls[None]).to(dev); loc = torch.from_numpy(loc[None]).to(dev)
    for _ in range(3):
        det, ms = F.postprocess_batch(pp, cls, loc_pred=loc, anchors=a_eval[:4], profile=True)
    if "--prod" in sys.argv:
        print("production build: kept=%d kernel us: %s" % (int(det.counts[0, 0]), [round(1e3 * v, 1) for v in ms]))
        continue
    buf = (ctypes.c_longlong * 32)()
    L.dan_debug_phases(buf)
    t = np.array(list(buf), dtype=np.float64) / 1965.0   # us at 1965 MHz
    print("K=%d kept=%d  kernel us: %s" % (buf[18], int(det.counts[0, 0]), [round(1e3 * v, 1) for v in ms]))
    print("  load+select %.1f | bitonic %.1f | decode %.1f | greedy %.1f | outputs %.1f" %
          (t[8] - t[0], t[1] - t[8], t[2] - t[1], t[3] - t[2], t[4] - t[3]))
    print("  greedy: a (vs kept) %.1f | b (compact) %.1f | c (pair bits) %.1f | d (relax) %.1f | e (append) %.1f ; %d chunks, %d survivors, %d sweeps"
          % (t[10], t[11], t[12], t[13], t[14], buf[17], buf[16], buf[15]))
    print("  b: pre-barrier %.2f | named barrier %.2f | copy %.2f ;  d per sweep total: read masks %.2f | barrier %.2f | decide+atomic %.2f | barrier.or %.2f"
          % (t[20], t[21], t[22], t[23], t[24], t[25], t[26]))
