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
        lib.oracle_dynamic_anchor_routing_eval.restype = ctypes.c_int
        lib.oracle_dynamic_anchor_routing_eval.argtypes = [_f32p, _f32p, _f32p, _i32p, ctypes.c_int64] + [ctypes.c_int32] * 4 + \
                                                          [_i32p, _f32p]
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


_REF_DAR = None


def have_reference_dar():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdar_ref.so"))


def _ref_dar():
    global _REF_DAR
    if _REF_DAR is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libdar_ref.so"))
        lib.ref_dynamic_anchor_routing.restype = ctypes.c_int
        lib.ref_dynamic_anchor_routing.argtypes = [_f32p, _f32p, _f32p, _i32p, ctypes.c_int64] + [ctypes.c_int32] * 7 + \
                                                  [ctypes.c_float, ctypes.c_float, _i32p, _f32p]
        _REF_DAR = lib
    return _REF_DAR


def dynamic_anchor_routing(anchors, gt_targets, labels, mask_in, feat_height, feat_width, anchor_depth, feat_strides,
                           img_height, img_width, trainging=False, thres=0.03, ignore_thres=0.0, impl="port"):
    """One layer of the DynamicAnchorRouting op (dynamic_anchor_routing.cc:31-59) -> (mask_out int32 [n], decode_out [n,4]).
    impl="reference": the reference's own functor (both branches; the training branch draws from std::random_device);
    impl="port": our restatement of the EVALUATION branch only."""
    anchors = np.ascontiguousarray(anchors, dtype=np.float32).reshape(-1, 4)
    gt_targets = np.ascontiguousarray(gt_targets, dtype=np.float32).reshape(-1, 4)
    labels = np.ascontiguousarray(labels, dtype=np.float32).reshape(-1)
    mask_in = np.ascontiguousarray(mask_in, dtype=np.int32).reshape(-1)
    n = anchors.shape[0]
    assert gt_targets.shape[0] == n and labels.shape[0] == n and mask_in.shape[0] == n
    if not (0. <= thres < 1.) or not (0. <= ignore_thres < 1.):        # dynamic_anchor_routing.cc:527-529
        raise ValueError("Need Attr 1 > thres >= 0. and 1 > ignore_thres >= 0.")
    mask_out = np.zeros(n, dtype=np.int32)
    decode_out = np.zeros((n, 4), dtype=np.float32)
    a = (anchors.ctypes.data_as(_f32p), gt_targets.ctypes.data_as(_f32p), labels.ctypes.data_as(_f32p),
         mask_in.ctypes.data_as(_i32p), n, int(feat_height), int(feat_width), int(anchor_depth), int(feat_strides))
    if impl == "reference":
        _ref_dar().ref_dynamic_anchor_routing(*a, int(img_height), int(img_width), int(bool(trainging)), float(thres),
                                              float(ignore_thres), mask_out.ctypes.data_as(_i32p), decode_out.ctypes.data_as(_f32p))
    else:
        if trainging:
            raise NotImplementedError("the training branch draws from std::random_device (unseeded): not restated")
        _port().oracle_dynamic_anchor_routing_eval(*a, mask_out.ctypes.data_as(_i32p), decode_out.ctypes.data_as(_f32p))
    return mask_out, decode_out
