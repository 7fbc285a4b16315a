"""Per-phase SM-cycle timestamps of the per-list kernels (CTA of list 0).  Builds a -DDAN_PHASE_TIMING copy of the library
in /tmp, runs one image with many candidates, prints phase durations in microseconds at the SM clock."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dan_b200 import build
lib_dbg = "/tmp/libdan_b200_phase.so"
cmd = ["nvcc"] + build.NVCC_FLAGS + ["-DDAN_PHASE_TIMING"] + [os.path.join(build.CSRC, s) for s in build.SOURCES] + ["-o", lib_dbg]
subprocess.run(cmd, check=True)
from dan_b200 import _lib
_lib.LIB_PATH = lib_dbg
from dan_b200 import functional as F, synthetic
from dan_b200.utility import anchor_manipulator as am
L = _lib.lib()
L.dan_debug_phases.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
dev = torch.device("cuda", 0)
ps = [0.1, 0.1, 0.2, 0.2]
enc = am.AnchorEncoder(0.4, 0.4, ps)
a_eval = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", (640, 640), border=0.))
an = np.stack([a.cpu().numpy() for a in a_eval[:4]], -1)
pp = F.postprocess_params(2, (640, 640), 0.01, 0, 5000, 750, 0.3, ps)
for faces, seed in ((60, 1420), (300, 3100), (450, 4150)):
    cls, loc, _ = synthetic.gen_predictions(seed, an, max_faces=faces)
    cls = torch.from_numpy(cls[None]).to(dev); loc = torch.from_numpy(loc[None]).to(dev)
    for _ in range(3):
        det, ms = F.postprocess_batch(pp, cls, loc_pred=loc, anchors=a_eval[:4], profile=True)
    buf = (ctypes.c_longlong * 32)()
    L.dan_debug_phases(buf)
    t = np.array(list(buf), dtype=np.float64) / 1965.0   # us at 1965 MHz
    k = int(F._ws._buf[:4].view(torch.int32).cpu()[0])
    print("K=%d kept=%d  kernel ms: %s" % (k, int(det.counts[0, 0]), [round(1e3 * v, 1) for v in ms]))
    print("  sort kernel: load+select %.1f | bitonic %.1f | decode+minmax %.1f | cell count %.1f | scan+scatter %.1f" %
          (t[8] - t[0], t[1] - t[8], t[2] - t[1], t[3] - t[2], t[4] - t[3]))
    print("  resolve kernel: relaxation %.1f (%d edges, %d rounds) | compaction %.1f | outputs %.1f" %
          (t[17] - t[16], buf[20], buf[21], t[18] - t[17], t[19] - t[18]))
    print("  pairs kernel (CTA 0): staging %.1f | search %.1f | flush %.1f" % (t[25] - t[24], t[26] - t[25], t[27] - t[26]))
    print("     warp 0: window setup %.1f us, item loops %.1f us, %d candidates in %d (box, class) searches" % (t[28], t[29], buf[30], buf[31]))
