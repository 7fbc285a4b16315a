"""CPU ORACLE (test infrastructure): ctypes bindings for oracle/liboracle_native.so
(our restatement) and oracle/_ref/libsmm_ref.so (the reference's own functor)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = None
_REF = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(quiet=True):
    """(Re)build the native oracle; `make ref` only acts when /root/reference exists."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _port():
    global _PORT
    if _PORT is None:
        path = os.path.join(_HERE, "liboracle_native.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.oracle_small_mining_match.restype = ctypes.c_int
        lib.oracle_small_mining_match.argtypes = [_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                                                  ctypes.c_float, ctypes.c_float, ctypes.c_int32,
                                                  ctypes.c_float, _i32p, _f32p]
        lib.oracle_tf_nms.restype = ctypes.c_int
        lib.oracle_tf_nms.argtypes = [_f32p, _f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                                      ctypes.c_int32, _i32p]
        _PORT = lib
    return _PORT


def have_reference():
    return os.path.exists(os.path.join(_HERE, "_ref", "libsmm_ref.so"))


def _ref():
    global _REF
    if _REF is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libsmm_ref.so"))
        lib.ref_small_mining_match.restype = ctypes.c_int
        lib.ref_small_mining_match.argtypes = [_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                                               ctypes.c_float, ctypes.c_float, ctypes.c_int32,
                                               ctypes.c_float, _i32p, _f32p]
        _REF = lib
    return _REF


def small_mining_match(overlap, neg_low, neg_high, pos, min_match, stop, impl="port"):
    overlap = np.ascontiguousarray(overlap, dtype=np.float32)
    assert overlap.ndim == 2, "inputs must be in 'num_anchors x num_ground_truth' format."
    n, m = overlap.shape
    match = np.empty(n, dtype=np.int32)
    scores = np.empty(n, dtype=np.float32)
    fn = _port().oracle_small_mining_match if impl == "port" else _ref().ref_small_mining_match
    rc = fn(overlap.ctypes.data_as(_f32p), n, m, neg_low, neg_high, pos, int(min_match), stop,
            match.ctypes.data_as(_i32p), scores.ctypes.data_as(_f32p))
    if rc != 0:
        raise ValueError("SmallMiningMatch: invalid attribute (small_mining_match.cc:292-305)")
    return match, scores


def tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold, tie="stable"):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    k = boxes.shape[0]
    out = np.empty(max(int(max_output_size), 1), dtype=np.int32)
    cnt = _port().oracle_tf_nms(boxes.ctypes.data_as(_f32p), scores.ctypes.data_as(_f32p), k,
                                int(max_output_size), float(iou_threshold),
                                1 if tie == "std_sort" else 0, out.ctypes.data_as(_i32p))
    return out[:cnt].copy()
