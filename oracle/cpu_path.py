"""CPU ORACLE (test infrastructure): the reference's CPU implementation of the whole hot path for ONE image,
timed by bench.py's ``cpu_baseline`` / ``--impl reference`` legs (and nothing else).

Per image, exactly what the reference does on the host (SURVEY.md call stacks A and D):
  training side   encode_anchors(match_mining=True): materialised [N,M] IoU (numpy, one op per TF op) +
                  SmallMiningMatch (the reference's own functor from oracle/_ref when it was built, else the port)
                  + gather/encode                                     (utility/anchor_manipulator.py:275-326)
  evaluation side decode_anchors + parse_by_class(softmax, select, clip, filter, top_k, NMS)
                                                                      (anchor_manipulator.py:409-424, bbox_util.py:103-119)
Image-level parallelism over worker processes mirrors the reference's queue threads (train_sfd.py:37-39).
TensorFlow itself is not installable here, so this is a restatement ("port"), not TF's kernels."""
from __future__ import annotations

import importlib.util
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STATE = {}


def _synthetic():
    """dan_b200/synthetic.py loaded by path (numpy only) so that worker processes never import torch."""
    if "syn" not in _STATE:
        path = os.path.join(os.path.dirname(_HERE), "dan_b200", "synthetic.py")
        spec = importlib.util.spec_from_file_location("_dan_synthetic", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _STATE["syn"] = mod
    return _STATE["syn"]


def setup(cfg):
    """cfg: dict(kind, size, pos, ign, mining, max_gt, max_faces, pp=(thr,min_size,keep_topk,nms_topk,nms_thr))."""
    from . import native, reference_np as R
    syn = _synthetic()
    enc = R.AnchorEncoder(cfg["pos"], cfg["ign"], [0.1, 0.1, 0.2, 0.2])
    _STATE["cfg"] = cfg
    _STATE["enc"] = enc
    _STATE["train_anchors"] = syn.build_anchors(enc, syn.pyramid_config(cfg["kind"], tuple(cfg["size"])))
    _STATE["eval_anchors"] = syn.build_anchors(enc, syn.pyramid_config(cfg["kind"], tuple(cfg["size"]), border=0.))
    _STATE["impl"] = "reference" if native.have_reference() else "port"
    _STATE["inputs"] = {}
    return _STATE["impl"]


def _gen_one(i):
    syn = _synthetic()
    cfg = _STATE["cfg"]
    an = _STATE["eval_an"]
    gt = syn.gen_faces(i, cfg["max_gt"], size=tuple(cfg["size"]))
    cls, loc, _ = syn.gen_predictions(i, an, size=tuple(cfg["size"]), max_faces=cfg["max_faces"])
    return i, (gt, cls, loc)


def generate_inputs(cfg, indices, procs=None):
    """Synthetic inputs of the given images (untimed), generated in parallel and kept in this (parent) process so
    that a pool forked afterwards sees all of them."""
    import multiprocessing as mp
    setup(cfg)
    _STATE["eval_an"] = np.stack(_STATE["eval_anchors"][:4], -1)
    todo = [i for i in indices if i not in _STATE["inputs"]]
    if todo:
        with mp.get_context("fork").Pool(min(procs or os.cpu_count() or 1, len(todo))) as pool:
            for i, v in pool.map(_gen_one, todo, chunksize=1):
                _STATE["inputs"][i] = v
    return _STATE["inputs"]


def run_image(i):
    from . import reference_np as R
    cfg = _STATE["cfg"]
    enc = _STATE["enc"]
    gt, cls, loc = _STATE["inputs"][i]
    t = enc.encode_anchors(gt, *_STATE["train_anchors"], match_mining=cfg["mining"], mining_impl=_STATE["impl"])
    ea = _STATE["eval_anchors"]
    boxes = enc.decode_anchors(loc, *ea[:4])
    thr, min_size, keep_topk, nms_topk, nms_thr = cfg["pp"]
    sb, ss = R.parse_by_class(list(cfg["size"]), cls, boxes, cls.shape[1], thr, min_size, keep_topk, nms_topk, nms_thr)
    return int((t[1] == 1).sum()), int((ss[1] > 0).sum())


def run_images(indices):
    t0 = time.perf_counter()
    out = [run_image(i) for i in indices]
    return time.perf_counter() - t0, out


class CpuPool(object):
    """P forked worker processes, image-level parallelism.  Create it AFTER generate_inputs() (the workers inherit the
    inputs through fork) and BEFORE any CUDA initialisation in this process."""

    def __init__(self, cfg, procs=None):
        import multiprocessing as mp
        self.procs = procs or os.cpu_count() or 1
        if "cfg" not in _STATE:
            setup(cfg)
        self.impl = _STATE["impl"]
        self.pool = mp.get_context("fork").Pool(self.procs)

    def run(self, indices):
        """-> (wall seconds, images, sum of per-worker busy seconds)."""
        per = [[] for _ in range(self.procs)]
        for k, i in enumerate(indices):
            per[k % self.procs].append(i)
        chunks = [c for c in per if c]
        t0 = time.perf_counter()
        res = self.pool.map(run_images, chunks, chunksize=1)
        wall = time.perf_counter() - t0
        return wall, len(indices), sum(r[0] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()
