"""CPU ORACLE (test infrastructure -- NOT a product path).

numpy-fp32, one-numpy-op-per-TF-op restatement of the reference's anchor hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``dan_b200``) never imports it and has no CPU fallback.

What is restated (citations relative to /root/reference):

* ``utility/anchor_manipulator.py:24-52``   areas / intersection / iou_matrix
* ``utility/anchor_manipulator.py:54-105``  do_dual_max_match
* ``utility/anchor_manipulator.py:118-424`` AnchorEncoder (anchors, encode, decode)
* ``utility/bbox_util.py:24-119``           select/clip/filter/sort/nms/parse_by_class
* ``cpp/ExtraLib/small_mining_match.cc:67-284`` SmallMiningMatch (native, see
  ``oracle/native/oracle_native.cpp``; cross-checked against the reference's own
  functor compiled verbatim into ``oracle/_ref`` when /root/reference is present)

Third-party arithmetic that is NOT under /root/reference (TensorFlow 1.8, pinned
only by README.md:54): ``tf.argmax`` (first max), ``tf.nn.top_k`` (descending,
ties -> lower index), ``tf.image.non_max_suppression`` and ``tf.nn.softmax`` /
``tf.exp`` / ``tf.log`` are restated from their published algorithms
(see ``oracle/README.md``).  No reference test pins top-k / NMS results, so for
those rows parity is UNPINNED (stated in DESIGN.md as well).  Matching is pinned
by the verbatim-compiled functor and by the known-answer vector derived from
``cpp/ExtraLib/test_op.py:41,56``.

Numerics discipline: every array is float32, every +,-,*,/ is one separately
rounded IEEE op exactly like TF's unfused Eigen elementwise kernels, in the
association order of the reference source.  exp/log are the Cephes single
precision polynomials (what Eigen's pexp/plog implement) evaluated WITHOUT fma,
so that the CUDA kernels can reproduce them bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
i32 = np.int32
i64 = np.int64


# --------------------------------------------------------------------------
# elementary functions (Cephes expf / logf, as in Eigen pexp / plog, no fma)
# --------------------------------------------------------------------------
def expf(x):
    """Cephes expf, op order of Eigen's pexp<Packet4f> with pmadd = mul then add."""
    x = np.asarray(x, dtype=f32)
    x = np.minimum(np.maximum(x, f32(-88.3762626647949)), f32(88.3762626647950))
    fx = np.floor(x * f32(1.44269504088896341) + f32(0.5))
    tmp = fx * f32(0.693359375)
    z = fx * f32(-2.12194440e-4)
    x = x - tmp
    x = x - z
    z = x * x
    y = np.full_like(x, f32(1.9875691500e-4))
    y = y * x + f32(1.3981999507e-3)
    y = y * x + f32(8.3334519073e-3)
    y = y * x + f32(4.1665795894e-2)
    y = y * x + f32(1.6666665459e-1)
    y = y * x + f32(5.0000001201e-1)
    y = y * z + x
    y = y + f32(1.0)
    n = fx.astype(i32)
    # n in [-127, 128]; 2^n is built in ONE step, (n + 127) << 23, like Eigen's pexp: n = -127 gives the factor +0 (the
    # result flushes to 0 instead of a denormal), n = 128 gives inf
    p = ((n + 127).astype(np.uint32) << np.uint32(23)).view(f32)
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        return y * p


def logf(x):
    """Cephes logf, op order of Eigen's plog<Packet4f> with pmadd = mul then add."""
    x = np.asarray(x, dtype=f32)
    invalid = x < f32(0.0)
    iszero = x == f32(0.0)
    x = np.maximum(x, f32(1.17549435e-38))
    bits = x.view(np.uint32)
    e = (bits >> np.uint32(23)).astype(i32) - i32(126)      # (exp - 0x7f) + 1
    m = ((bits & np.uint32(0x807FFFFF)) | np.uint32(0x3F000000)).view(f32)   # [0.5, 1)
    e = e.astype(f32)
    small = m < f32(0.707106781186547524)
    tmp = np.where(small, m, f32(0.0))
    m = m - f32(1.0)
    e = e - np.where(small, f32(1.0), f32(0.0))
    m = m + tmp
    x2 = m * m
    x3 = x2 * m
    y = f32(7.0376836292e-2) * m + f32(-1.1514610310e-1)
    y1 = f32(-1.2420140846e-1) * m + f32(1.4249322787e-1)
    y2 = f32(2.0000714765e-1) * m + f32(-2.4999993993e-1)
    y = y * m + f32(1.1676998740e-1)
    y1 = y1 * m + f32(-1.6668057665e-1)
    y2 = y2 * m + f32(3.3333331174e-1)
    y = y * x3 + y1
    y = y * x3 + y2
    y = y * x3
    y1 = e * f32(-2.12194440e-4)
    tmp = x2 * f32(0.5)
    y = y + y1
    m = m - tmp
    y2 = e * f32(0.693359375)
    m = m + y
    m = m + y2
    out = np.where(iszero, f32(-np.inf), m)
    out = np.where(invalid, f32(np.nan), out)
    return out.astype(f32)


def softmax(logits):
    """tf.nn.softmax on CPU (Eigen functor): exp(x - max) * (1 / sum(exp(x - max))).

    Call site: utility/bbox_util.py:105, eval_sfd.py:279.  The class sum runs in
    index order."""
    logits = np.asarray(logits, dtype=f32)
    shifted = logits - logits.max(axis=-1, keepdims=True)
    e = expf(shifted)
    s = e[..., 0].copy()
    for c in range(1, e.shape[-1]):
        s = s + e[..., c]
    inv = f32(1.0) / s
    return e * inv[..., None]


# --------------------------------------------------------------------------
# utility/anchor_manipulator.py:24-52
# --------------------------------------------------------------------------
def areas(boxes):
    """anchor_manipulator.py:24-27."""
    ymin, xmin, ymax, xmax = [boxes[:, i:i + 1] for i in range(4)]
    return (xmax - xmin + f32(1.)) * (ymax - ymin + f32(1.))


def intersection(a_boxes, g_boxes):
    """anchor_manipulator.py:29-43 (first arg = anchors [N,4], second = GT [M,4])."""
    ymin, xmin, ymax, xmax = [a_boxes[:, i:i + 1] for i in range(4)]
    g_ymin, g_xmin, g_ymax, g_xmax = [g_boxes[:, i:i + 1].T for i in range(4)]
    int_ymin = np.maximum(ymin, g_ymin)
    int_xmin = np.maximum(xmin, g_xmin)
    int_ymax = np.minimum(ymax, g_ymax)
    int_xmax = np.minimum(xmax, g_xmax)
    h = np.maximum(int_ymax - int_ymin + f32(1.), f32(0.))
    w = np.maximum(int_xmax - int_xmin + f32(1.), f32(0.))
    return h * w


def iou_matrix(a_boxes, g_boxes):
    """anchor_manipulator.py:44-52 -> [N, M] fp32, materialised like TF does."""
    a_boxes = np.asarray(a_boxes, dtype=f32)
    g_boxes = np.asarray(g_boxes, dtype=f32)
    inter = intersection(a_boxes, g_boxes)
    union = areas(a_boxes) + areas(g_boxes).T - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        q = inter / union
    return np.where(union == f32(0.0), np.zeros_like(inter), q).astype(f32)


# --------------------------------------------------------------------------
# utility/anchor_manipulator.py:54-105
# --------------------------------------------------------------------------
def do_dual_max_match(overlap, low_thres, high_thres, ignore_between=True, gt_max_first=True):
    """anchor_manipulator.py:54-105.  overlap: [N, M] fp32 -> (int64 [N], fp32 [N])."""
    overlap = np.asarray(overlap, dtype=f32)
    n, m = overlap.shape
    low = f32(low_thres)
    high = f32(high_thres)
    anchors_to_gt = overlap.argmax(axis=1).astype(i64)          # :62 first max
    match_values = overlap.max(axis=1)                           # :64
    less_mask = match_values < low                               # :67
    between_mask = (match_values < high) & (match_values >= low)  # :68
    negative_mask = less_mask if ignore_between else between_mask
    ignore_mask = between_mask if ignore_between else less_mask
    match_indices = np.where(negative_mask, i64(-1), anchors_to_gt)   # :75
    match_indices = np.where(ignore_mask, i64(-2), match_indices)     # :76
    gt_to_anchors_overlap = overlap.max(axis=0, keepdims=True)   # :84
    left_mask = overlap == gt_to_anchors_overlap                 # :88
    if not gt_max_first:                                         # :89-92
        onehot = np.zeros((n, m), dtype=i32)
        pos = match_indices >= 0
        onehot[np.nonzero(pos)[0], match_indices[pos]] = 1
        left_mask = (onehot.max(axis=0, keepdims=True) < 1) & left_mask
    left_i64 = left_mask.astype(i64)                             # :94
    left_scores = overlap * left_mask.astype(f32)                # :95
    has_claim = left_i64.max(axis=1) > 0
    claim_idx = left_scores.argmax(axis=1).astype(i64)
    sel = np.where(has_claim, claim_idx, anchors_to_gt)          # :98-101
    selected_scores = overlap[np.arange(n), sel]
    return np.where(has_claim, claim_idx, match_indices), selected_scores


# --------------------------------------------------------------------------
# cpp/ExtraLib/small_mining_match.cc  (native restatement, loaded lazily)
# --------------------------------------------------------------------------
def small_mining_match(overlap, negative_low_thres, negative_high_thres, positive_thres,
                       min_match, stop_positive_thres, impl="port"):
    """SmallMiningMatch op (small_mining_match.cc:31-54 signature, :288-342 kernel).

    impl="port": oracle/native restatement; impl="reference": the reference's own
    functor compiled from /root/reference into oracle/_ref (if built)."""
    from . import native
    return native.small_mining_match(overlap, negative_low_thres, negative_high_thres,
                                     positive_thres, min_match, stop_positive_thres, impl=impl)


# --------------------------------------------------------------------------
# utility/anchor_manipulator.py:107-424
# --------------------------------------------------------------------------
class AnchorEncoder(object):
    """numpy mirror of utility/anchor_manipulator.py:107 AnchorEncoder."""

    def __init__(self, positive_threshold, ignore_threshold, prior_scaling):
        self._positive_threshold = positive_threshold
        self._ignore_threshold = ignore_threshold
        self._prior_scaling = prior_scaling

    # :125-127
    def center2point(self, center_y, center_x, height, width):
        return (center_y - (height - f32(1.)) / f32(2.), center_x - (width - f32(1.)) / f32(2.),
                center_y + (height - f32(1.)) / f32(2.), center_x + (width - f32(1.)) / f32(2.))

    # :129-132
    def point2center(self, ymin, xmin, ymax, xmax):
        height, width = (ymax - ymin + f32(1.)), (xmax - xmin + f32(1.))
        return (ymin + ymax) / f32(2.), (xmin + xmax) / f32(2.), height, width

    # :134-161
    def get_anchors_width_height(self, anchor_scale, extra_anchor_scale, anchor_ratio, name=None):
        depth = len(anchor_scale) * len(anchor_ratio) + len(extra_anchor_scale)
        hs, ws = [], []
        for scale in extra_anchor_scale:
            hs.append(scale)
            ws.append(scale)
        for scale in anchor_scale:
            for ratio in anchor_ratio:
                hs.append(scale / math.sqrt(ratio))      # python float64, rounded once below
                ws.append(scale * math.sqrt(ratio))
        return np.asarray(hs, dtype=f32), np.asarray(ws, dtype=f32), depth

    # :163-198
    def generate_anchors_by_offset(self, anchors_height, anchors_width, anchor_depth, image_shape,
                                   layer_shape, feat_stride, offset=0.5, name=None):
        feat_stride = f32(feat_stride)
        x_on_layer, y_on_layer = np.meshgrid(np.arange(layer_shape[1]), np.arange(layer_shape[0]))
        if isinstance(offset, (list, tuple)):
            offset_h, offset_w = offset[0], offset[1]
        else:
            offset_h = offset_w = offset
        y_on_image = (y_on_layer.astype(f32) + f32(offset_h)) * feat_stride
        x_on_image = (x_on_layer.astype(f32) + f32(offset_w)) * feat_stride
        ymin, xmin, ymax, xmax = self.center2point(y_on_image[..., None], x_on_image[..., None],
                                                   np.asarray(anchors_height, f32), np.asarray(anchors_width, f32))
        return (ymin.reshape(-1, anchor_depth), xmin.reshape(-1, anchor_depth),
                ymax.reshape(-1, anchor_depth), xmax.reshape(-1, anchor_depth))

    # :200-211
    def get_anchors_count(self, anchors_depth, layer_shape, name=None):
        spatial = layer_shape[0] * layer_shape[1]
        return spatial, spatial * anchors_depth

    # :213-273
    def get_all_anchors(self, image_shape, anchors_height, anchors_width, anchors_depth, anchors_offsets,
                        layer_shapes, feat_strides, allowed_borders, should_clips, name=None):
        image_height, image_width = f32(image_shape[0]), f32(image_shape[1])
        l_ymin, l_xmin, l_ymax, l_xmax, l_border = [], [], [], [], []
        for ind, anchor_depth in enumerate(anchors_depth):
            ymin, xmin, ymax, xmax = self.generate_anchors_by_offset(
                anchors_height[ind], anchors_width[ind], anchor_depth, image_shape,
                layer_shapes[ind], feat_strides[ind], offset=anchors_offsets[ind])
            if should_clips[ind]:
                ymin = np.clip(ymin, f32(0.), image_height - f32(1.))
                xmin = np.clip(xmin, f32(0.), image_width - f32(1.))
                ymax = np.clip(ymax, f32(0.), image_height - f32(1.))
                xmax = np.clip(xmax, f32(0.), image_width - f32(1.))
            ymin, xmin, ymax, xmax = [v.reshape(-1).astype(f32) for v in (ymin, xmin, ymax, xmax)]
            l_ymin.append(ymin)
            l_xmin.append(xmin)
            l_ymax.append(ymax)
            l_xmax.append(xmax)
            l_border.append(np.ones_like(ymin, dtype=f32) * f32(allowed_borders[ind]))
        ymin = np.concatenate(l_ymin)
        xmin = np.concatenate(l_xmin)
        ymax = np.concatenate(l_ymax)
        xmax = np.concatenate(l_xmax)
        border = np.concatenate(l_border)
        inside_mask = ((ymin > -border) & (xmin > -border)) & \
                      ((ymax < (image_height - f32(1.) + border)) & (xmax < (image_width - f32(1.) + border)))
        return ymin, xmin, ymax, xmax, inside_mask

    def _encode(self, bboxes, match_anchors, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax,
                inside_mask, ignore_threshold, positive_threshold, match_mining, scale, debug,
                mining_impl="port", return_match=False):
        bboxes = np.asarray(bboxes, dtype=f32).reshape(-1, 4)
        if bboxes.shape[0] < 1:                                  # :286 / :346
            bboxes = np.asarray([[0., 0., 1., 1.]], dtype=f32)
        overlap = iou_matrix(match_anchors, bboxes) * inside_mask.astype(f32)[:, None]   # :287
        if match_mining:                                         # :290-291
            matched_gt, gt_scores = small_mining_match(overlap, 0., ignore_threshold, positive_threshold,
                                                       6, 0.3, impl=mining_impl)
            matched_gt = matched_gt.astype(i64)
        else:                                                    # :293
            matched_gt, gt_scores = do_dual_max_match(overlap, ignore_threshold, positive_threshold)
        matched_gt_mask = matched_gt > -1                        # :296
        matched_indices = np.clip(matched_gt, 0, np.iinfo(i32).max)
        gt_labels = matched_gt_mask.astype(i64)
        gt_labels = gt_labels + (i64(-1) * (matched_gt < -1).astype(i64))    # :302
        matched_gt_bbox = bboxes[matched_indices]                # :306
        gt_ymin, gt_xmin, gt_ymax, gt_xmax = [matched_gt_bbox[:, i] for i in range(4)]
        gt_cy, gt_cx, gt_h, gt_w = self.point2center(gt_ymin, gt_xmin, gt_ymax, gt_xmax)
        anchor_cy, anchor_cx, anchor_h, anchor_w = self.point2center(anchors_ymin, anchors_xmin,
                                                                     anchors_ymax, anchors_xmax)
        ps = [f32(v) for v in self._prior_scaling]
        gt_cy = (gt_cy - anchor_cy) / anchor_h / ps[0]           # :314
        gt_cx = (gt_cx - anchor_cx) / anchor_w / ps[1]
        if scale is None:                                        # encode_anchors :316-317
            gt_h = logf(gt_h / anchor_h) / ps[2]
            gt_w = logf(gt_w / anchor_w) / ps[3]
        else:                                                    # encode_pa_anchors :377-378
            gt_h = logf(gt_h * f32(scale) / anchor_h) / ps[2]
            gt_w = logf(gt_w * f32(scale) / anchor_w) / ps[3]
        if debug:
            gt_targets = np.stack([anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax], axis=-1)
        else:
            gt_targets = np.stack([gt_cy, gt_cx, gt_h, gt_w], axis=-1)
        posf = matched_gt_mask.astype(f32)[:, None]
        gt_targets = posf * gt_targets                           # :324
        out = (gt_targets.astype(f32), gt_labels, gt_scores.astype(f32), (matched_gt_bbox * posf).astype(f32))
        if return_match:
            return out + (matched_gt,)
        return out

    # :275-326
    def encode_anchors(self, bboxes, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax, inside_mask,
                       match_mining=False, debug=False, mining_impl="port", return_match=False):
        all_anchors = np.stack([anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax], axis=-1)
        return self._encode(bboxes, all_anchors, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax,
                            inside_mask, self._ignore_threshold, self._positive_threshold, match_mining,
                            None, debug, mining_impl, return_match)

    # :328-387
    def encode_pa_anchors(self, bboxes, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax, inside_mask,
                          ignore_threshold, positive_threshold, match_mining=True, scale=1., debug=False,
                          mining_impl="port", return_match=False):
        anchor_cy, anchor_cx, anchor_h, anchor_w = self.point2center(anchors_ymin, anchors_xmin,
                                                                     anchors_ymax, anchors_xmax)
        all_anchors = np.stack(self.center2point(anchor_cy, anchor_cx, anchor_h / f32(scale),
                                                 anchor_w / f32(scale)), axis=-1)   # :340-342
        return self._encode(bboxes, all_anchors, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax,
                            inside_mask, ignore_threshold, positive_threshold, match_mining,
                            scale, debug, mining_impl, return_match)

    # :389-408
    def batch_decode_anchors(self, pred_location, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax):
        pred_location = np.asarray(pred_location, dtype=f32)
        a = [np.asarray(v, f32)[None, :] for v in (anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax)]
        anchor_cy, anchor_cx, anchor_h, anchor_w = self.point2center(*a)
        ps = [f32(v) for v in self._prior_scaling]
        pred_h = expf(pred_location[:, :, -2] * ps[2]) * anchor_h
        pred_w = expf(pred_location[:, :, -1] * ps[3]) * anchor_w
        pred_cy = pred_location[:, :, 0] * ps[0] * anchor_h + anchor_cy
        pred_cx = pred_location[:, :, 1] * ps[1] * anchor_w + anchor_cx
        return np.stack(self.center2point(pred_cy, pred_cx, pred_h, pred_w), axis=-1).astype(f32)

    # :409-424
    def decode_anchors(self, pred_location, anchors_ymin, anchors_xmin, anchors_ymax, anchors_xmax):
        pred_location = np.asarray(pred_location, dtype=f32)
        return self.batch_decode_anchors(pred_location[None], anchors_ymin, anchors_xmin,
                                         anchors_ymax, anchors_xmax)[0]


# --------------------------------------------------------------------------
# TF library ops used by utility/bbox_util.py
# --------------------------------------------------------------------------
def tf_top_k(values, k):
    """tf.nn.top_k(sorted=True): descending, equal elements -> lower index first."""
    values = np.asarray(values, dtype=f32)
    order = np.argsort(-values, kind="stable")[:k]
    return values[order], order.astype(i32)


def tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold, tie="stable"):
    """tf.image.non_max_suppression (TF r1.8 core/kernels/non_max_suppression_op.cc), native.

    tie="stable": equal scores keep input order (the tie policy this repo documents);
    tie="std_sort": libstdc++ std::sort with TF's comparator (what TF literally calls)."""
    from . import native
    return native.tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold, tie=tie)


# --------------------------------------------------------------------------
# utility/bbox_util.py:24-119
# --------------------------------------------------------------------------
def select_bboxes(scores_pred, bboxes_pred, num_classes, select_threshold):
    """bbox_util.py:24-36 (mask-multiply, no compaction)."""
    selected_bboxes, selected_scores = {}, {}
    for class_ind in range(1, num_classes):
        class_scores = scores_pred[:, class_ind]
        select_mask = (class_scores > f32(select_threshold)).astype(f32)
        selected_bboxes[class_ind] = bboxes_pred * select_mask[:, None]
        selected_scores[class_ind] = class_scores * select_mask
    return selected_bboxes, selected_scores


def clip_bboxes(ymin, xmin, ymax, xmax, height, width):
    """bbox_util.py:38-48."""
    ymin = np.maximum(ymin, f32(0.))
    xmin = np.maximum(xmin, f32(0.))
    ymax = np.minimum(ymax, f32(height) - f32(1.))
    xmax = np.minimum(xmax, f32(width) - f32(1.))
    ymin = np.minimum(ymin, ymax)
    xmin = np.minimum(xmin, xmax)
    return ymin, xmin, ymax, xmax


def filter_bboxes(scores_pred, ymin, xmin, ymax, xmax, min_size):
    """bbox_util.py:50-59."""
    width = xmax - xmin + f32(1.)
    height = ymax - ymin + f32(1.)
    thr = f32(min_size + 1.)
    filter_mask = ((width > thr) & (height > thr)).astype(f32)
    return (scores_pred * filter_mask, ymin * filter_mask, xmin * filter_mask,
            ymax * filter_mask, xmax * filter_mask)


def sort_bboxes(scores_pred, ymin, xmin, ymax, xmax, keep_topk):
    """bbox_util.py:61-72."""
    cur = scores_pred.shape[0]
    scores, idxes = tf_top_k(scores_pred, min(keep_topk, cur))
    ymin, xmin, ymax, xmax = ymin[idxes], xmin[idxes], ymax[idxes], xmax[idxes]
    pad = max(keep_topk - cur, 0)
    p = lambda v: np.pad(v, (0, pad)).astype(f32)
    return p(scores), p(ymin), p(xmin), p(ymax), p(xmax), idxes


def nms_bboxes(scores_pred, bboxes_pred, nms_topk, nms_threshold, tie="stable"):
    """bbox_util.py:75-78."""
    idxes = tf_non_max_suppression(bboxes_pred, scores_pred, nms_topk, nms_threshold, tie=tie)
    return scores_pred[idxes], bboxes_pred[idxes], idxes


def nms_bboxes_with_padding(scores_pred, bboxes_pred, nms_topk, nms_threshold, tie="stable"):
    """bbox_util.py:80-90 (paddings evaluate to [[0,pad]] and [[0,pad],[0,0]]: zeros at the END)."""
    idxes = tf_non_max_suppression(bboxes_pred, scores_pred, nms_topk, nms_threshold, tie=tie)
    scores = scores_pred[idxes]
    bboxes = bboxes_pred[idxes]
    pad = max(nms_topk - idxes.shape[0], 0)
    return (np.pad(scores, (0, pad)).astype(f32), np.pad(bboxes, ((0, pad), (0, 0))).astype(f32), idxes)


def bbox_point2center(bboxes):
    """bbox_util.py:92-96."""
    ymin, xmin, ymax, xmax = [bboxes[..., i] for i in range(4)]
    height, width = (ymax - ymin + f32(1.)), (xmax - xmin + f32(1.))
    return np.stack([(ymin + ymax) / f32(2.), (xmin + xmax) / f32(2.), height, width], axis=-1)


def bbox_center2point(bboxes):
    """bbox_util.py:98-101."""
    y, x, h, w = [bboxes[..., i] for i in range(4)]
    return np.stack([y - (h - f32(1.)) / f32(2.), x - (w - f32(1.)) / f32(2.),
                     y + (h - f32(1.)) / f32(2.), x + (w - f32(1.)) / f32(2.)], axis=-1)


def parse_by_class(image_shape, cls_pred, bboxes_pred, num_classes, select_threshold, min_size,
                   keep_topk, nms_topk, nms_threshold, tie="stable", return_indices=False):
    """bbox_util.py:103-119.  Returns ({c: [nms_topk,4]}, {c: [nms_topk]}) and, when
    ``return_indices`` is set, also {c: (topk anchor indices, nms keep positions)}."""
    cls_pred = np.asarray(cls_pred, dtype=f32)
    bboxes_pred = np.asarray(bboxes_pred, dtype=f32)
    scores_pred = softmax(cls_pred)
    selected_bboxes, selected_scores = select_bboxes(scores_pred, bboxes_pred, num_classes, select_threshold)
    indices = {}
    for class_ind in range(1, num_classes):
        ymin, xmin, ymax, xmax = [selected_bboxes[class_ind][:, i] for i in range(4)]
        ymin, xmin, ymax, xmax = clip_bboxes(ymin, xmin, ymax, xmax, image_shape[0], image_shape[1])
        sc, ymin, xmin, ymax, xmax = filter_bboxes(selected_scores[class_ind], ymin, xmin, ymax, xmax, min_size)
        sc, ymin, xmin, ymax, xmax, top_idx = sort_bboxes(sc, ymin, xmin, ymax, xmax, keep_topk)
        boxes = np.stack([ymin, xmin, ymax, xmax], axis=-1)
        sc, boxes, keep = nms_bboxes_with_padding(sc, boxes, nms_topk, nms_threshold, tie=tie)
        selected_scores[class_ind], selected_bboxes[class_ind] = sc, boxes
        indices[class_ind] = (top_idx, keep)
    if return_indices:
        return selected_bboxes, selected_scores, indices
    return selected_bboxes, selected_scores


# --------------------------------------------------------------------------
# per-image hard-negative mining, train_dan.py:286-324 (and the inline copy train_sfd.py:349-384)
# --------------------------------------------------------------------------
def mining_hard_neg(batch_size, cls_pred, location_pred, cls_targets, match_scores, loc_targets,
                    negative_ratio=3., num_classes=2, at_least_one=True):
    """train_dan.py:286-324 (``at_least_one=True``: the ``tf.maximum(..., 1)`` of :302) and train_sfd.py:349-384
    (``at_least_one=False``: no such clamp; an image without positives then indexes position -1 of the sorted row,
    which ``tf.gather_nd`` rejects on the CPU -> ValueError here).

    cls_pred [B*N, C] or [B, N, C] logits, location_pred [B*N, 4], cls_targets [B, N] int64, loc_targets [B, N, 4].
    Returns (cls_pred[final_mask], location_pred[positive_mask], clip(cls_targets)[final_mask], loc_targets[positive_mask])
    plus the dict of intermediates the tests compare (final_mask, n_neg_select, score_at_k)."""
    cls_targets = np.asarray(cls_targets, dtype=np.int64)
    B, N = cls_targets.shape
    assert B == batch_size
    cls_flat = np.asarray(cls_pred, dtype=f32).reshape(B * N, -1)
    loc_flat = np.asarray(location_pred, dtype=f32).reshape(B * N, 4)
    flat_targets = cls_targets.reshape(-1)                                   # :288
    flat_loc_targets = np.asarray(loc_targets, dtype=f32).reshape(-1, 4)     # :290
    positive_mask = flat_targets > 0                                         # :293
    batch_n_positives = np.count_nonzero(cls_targets > 0, axis=-1)           # :296
    batch_negative_mask = cls_targets == 0                                   # :298
    batch_n_negatives = np.count_nonzero(batch_negative_mask, axis=-1)       # :299
    n_sel = (f32(negative_ratio) * batch_n_positives.astype(f32)).astype(f32).astype(i32)   # :301 to_int32 truncates
    n_sel = np.minimum(n_sel, batch_n_negatives.astype(i32))
    if at_least_one:
        n_sel = np.maximum(n_sel, 1)                                         # :302
    elif (n_sel < 1).any():
        raise ValueError("indices[%d] = [%d, -1] does not index into param" % (int(np.argmax(n_sel < 1)), int(np.argmax(n_sel < 1))))
    bg = softmax(cls_flat.reshape(B, N, -1))[:, :, 0]                        # :305
    prob = np.where(batch_negative_mask, f32(0.) - bg, f32(0.) - np.ones_like(bg))   # :306-309
    score_at_k = np.empty(B, dtype=f32)
    for b in range(B):
        vals, _ = tf_top_k(prob[b], N)                                       # :310 full descending sort
        score_at_k[b] = vals[n_sel[b] - 1]                                   # :311
    selected = prob >= score_at_k[:, None]                                   # :313
    final_mask = np.logical_or(np.logical_and(batch_negative_mask, selected).reshape(-1), positive_mask)   # :316
    out = (cls_flat[final_mask], loc_flat[positive_mask],
           np.clip(flat_targets, 0, num_classes)[final_mask], flat_loc_targets[positive_mask])    # :319-322
    return out, {"final_mask": final_mask, "n_neg_select": n_sel.astype(i32), "score_at_k": score_at_k}


def mining_hard_neg_across_batch(batch_size, cls_pred, location_pred, cls_targets, match_scores, loc_targets,
                                 negative_ratio=3., num_classes=2):
    """train_dan.py:247-284: ONE selection over the flattened batch, strict ``>`` against the k-th value (:274)."""
    flat_targets = np.asarray(cls_targets, dtype=np.int64).reshape(-1)
    T = flat_targets.shape[0]
    cls_flat = np.asarray(cls_pred, dtype=f32).reshape(T, -1)
    loc_flat = np.asarray(location_pred, dtype=f32).reshape(T, 4)
    flat_loc_targets = np.asarray(loc_targets, dtype=f32).reshape(-1, 4)
    positive_mask = flat_targets > 0
    negative_mask = flat_targets == 0
    n_sel = int(np.minimum((f32(negative_ratio) * f32(np.count_nonzero(positive_mask))).astype(i32),
                           i32(np.count_nonzero(negative_mask))))
    if n_sel < 1:
        raise ValueError("slice index -1 of dimension 0 out of bounds")       # topk[-1] of an empty vector
    bg = softmax(cls_flat)[:, 0]
    prob = np.where(negative_mask, f32(0.) - bg, f32(0.) - np.ones_like(bg))
    vals, _ = tf_top_k(prob, n_sel)
    selected = prob > vals[-1]
    final_mask = np.logical_or(np.logical_and(negative_mask, selected), positive_mask)
    out = (cls_flat[final_mask], loc_flat[positive_mask], np.clip(flat_targets, 0, num_classes)[final_mask],
           flat_loc_targets[positive_mask])
    return out, {"final_mask": final_mask, "n_neg_select": np.array([n_sel], dtype=i32),
                 "score_at_k": np.array([vals[-1]], dtype=f32)}


# --------------------------------------------------------------------------
# DynamicAnchorRouting (C++ op) -- the oracle is native code (oracle/native.py); the only numpy piece
# is the restatement of glibc's expf that the CUDA kernel evaluates (std::exp(float) in dynamic_anchor_routing.cc:399-400)
# --------------------------------------------------------------------------
_EXP2F_N = 32


def _exp2f_table():
    """T[i] = bits(2^(i/32)) - (i << 47), computed (not copied) from correctly rounded decimal powers."""
    import struct
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    t = []
    for i in range(_EXP2F_N):
        d = float(Decimal(2) ** (Decimal(i) / Decimal(_EXP2F_N)))
        t.append((struct.unpack("<Q", struct.pack("<d", d))[0] - (i << 47)) & 0xFFFFFFFFFFFFFFFF)
    return np.array(t, dtype=np.uint64)


def libm_expf(x):
    """glibc >= 2.27 expf (the ARM optimized-routines algorithm, sysdeps/ieee754/flt-32/e_expf.c), restated: double
    arithmetic, k = round(32 x / ln 2) by the 0x1.8p52 shift, table of 2^(i/32), degree-3 polynomial, one final rounding."""
    x = np.asarray(x, dtype=f32)
    with np.errstate(all="ignore"):
        xd = x.astype(np.float64)
        z = float.fromhex("0x1.71547652b82fep+5") * xd
        kd = z + float.fromhex("0x1.8p52")
        ki = kd.view(np.uint64)
        kd = kd - float.fromhex("0x1.8p52")
        r = z - kd
        t = _exp2f_table()[(ki % np.uint64(_EXP2F_N)).astype(np.int64)] + (ki << np.uint64(47))
        s = t.view(np.float64)
        p = float.fromhex("0x1.c6af84b912394p-20") * r + float.fromhex("0x1.ebfce50fac4f3p-13")
        y = float.fromhex("0x1.62e42ff0c52d6p-6") * r + 1.0
        y = (p * (r * r) + y) * s
        out = y.astype(f32)
    out = np.where(x > f32(float.fromhex("0x1.62e42ep6")), f32(np.inf), out)
    out = np.where(x < f32(float.fromhex("-0x1.9fe368p6")), f32(0.), out)
    return np.where(np.isnan(x), x, out).astype(f32)


# --------------------------------------------------------------------------
# the evaluation merge the scripts actually use: detect_face's top-k (eval_sfd.py:95-114) and
# bbox_vote (eval_sfd.py:170-210; identical in eval_dan.py:201-241).  Boxes here are (xmin, ymin, xmax, ymax, score),
# float32, +1 pixel convention.
# --------------------------------------------------------------------------
def detect_face_select(bboxes, scores, shrink, max_per_image=750):
    """eval_sfd.py:101-112 (everything after net.run): boxes (ymin, xmin, ymax, xmax) / shrink -> columns
    (xmin, ymin, xmax, ymax, score); keep the top min(N - 1, int(1.5 * max_per_image)) by descending score.
    Tie policy: ``argsort()[::-1]`` with numpy's default (unstable, platform dependent) sort leaves equal scores
    unspecified; this restatement uses the stable sort, i.e. among equal scores the HIGHER index comes first."""
    bboxes = np.asarray(bboxes, dtype=f32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=f32).reshape(-1)
    sh = f32(shrink)
    det = np.column_stack((bboxes[:, 1] / sh, bboxes[:, 0] / sh, bboxes[:, 3] / sh, bboxes[:, 2] / sh, scores)).astype(f32)
    top = min(det.shape[0] - 1, int(max_per_image * 1.5))
    keep = np.argsort(det[:, 4], kind="stable")[::-1].astype(np.int64)[:top]     # note: [:-1] when N == 0
    return det[keep, :], keep


def numpy_pairwise_sum(a):
    """numpy's float32 add-reduce over one axis (umath pairwise_sum): < 8 sequential; <= 128 eight interleaved
    accumulators, tree-combined, tail sequential; beyond that split at n/2 rounded down to a multiple of 8."""
    n = len(a)
    if n < 8:
        res = f32(0.)
        for v in a:
            res = f32(res + v)
        return res
    if n <= 128:
        r = [f32(a[i]) for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] = f32(r[k] + a[i + k])
            i += 8
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        while i < n:
            res = f32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f32(numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:]))


def bbox_vote(det, nms_threshold=0.3, max_per_image=750, return_details=False):
    """eval_sfd.py:170-210, restated without the shrinking array: sort by descending score; a detection that no
    earlier HEAD overlaps with IoU >= thr becomes a head, every other one joins the FIRST head (in score order) that
    overlaps it; clusters of fewer than two members are dropped (:191-197); a cluster becomes one box = score-weighted
    mean of its members (float32, the box sums accumulate member by member in score order, the score sum is numpy's
    pairwise sum) with the maximum score (:198-203); the first max_per_image clusters in head order are returned."""
    det = np.asarray(det, dtype=f32).reshape(-1, 5)
    order = np.argsort(det[:, 4], kind="stable")[::-1]
    d = det[order]
    n = d.shape[0]
    area = (d[:, 2] - d[:, 0] + f32(1)) * (d[:, 3] - d[:, 1] + f32(1))
    assign = np.full(n, -1, dtype=np.int64)          # head position (in sorted order) of each detection, -2 = dropped head
    heads = []
    thr = f32(nms_threshold)
    for i in range(n):
        if assign[i] != -1:
            continue
        rest = np.nonzero(assign == -1)[0]
        rest = rest[rest >= i]
        xx1, yy1 = np.maximum(d[i, 0], d[rest, 0]), np.maximum(d[i, 1], d[rest, 1])
        xx2, yy2 = np.minimum(d[i, 2], d[rest, 2]), np.minimum(d[i, 3], d[rest, 3])
        w, h = np.maximum(f32(0.), xx2 - xx1 + f32(1)), np.maximum(f32(0.), yy2 - yy1 + f32(1))
        inter = w * h
        with np.errstate(all="ignore"):
            o = inter / (area[i] + area[rest] - inter)
        merged = rest[o >= thr]
        assign[merged] = i
        if assign[i] == -1:
            assign[i] = -2                            # the head does not even match itself (NaN IoU): deleted alone (:189-190)
        heads.append(i)
    out = []
    for hpos in heads:
        members = np.nonzero(assign == hpos)[0]
        if members.shape[0] <= 1:
            continue
        acc = d[members]
        weighted = acc[:, 0:4] * acc[:, 4:5]
        box_sum = np.zeros(4, dtype=f32)
        for r in weighted:                            # np.sum(axis=0) adds the rows one after the other
            box_sum = (box_sum + r).astype(f32)
        with np.errstate(all="ignore"):
            box = box_sum / numpy_pairwise_sum(acc[:, 4])
        out.append(np.concatenate([box, [np.max(acc[:, 4])]]).astype(f32))
    res = np.stack(out).astype(f32) if out else np.zeros((0, 5), dtype=f32)
    res = res[:min(max_per_image, res.shape[0])]
    if return_details:
        return res, {"order": order, "assign": assign}
    return res


# --------------------------------------------------------------------------
# input hand-off (TF graph code, restated op by op; the random coin of the flip is an input)
# --------------------------------------------------------------------------
def prepare_gt_batch(gt_list, patch_hw, out_shape, mirror=None, min_height=6., min_width=3.):
    """gt_list: per-image [M,4] (ymin, xmin, ymax, xmax) in patch pixels.  sfd_preprocessing.py:482-493 (mirror),
    :529-533 (rescale), :544-550 (small-face filter); dataset_common.py:178,186 (keep_input: images without boxes
    are dropped).  -> (gt_concat [sum,4], gt_offsets int32 [kept+1], image_index int32 [kept])."""
    boxes, offs, idx = [], [0], []
    for b, g in enumerate(gt_list):
        g = np.asarray(g, dtype=f32).reshape(-1, 4)
        ph, pw = f32(patch_hw[b][0]), f32(patch_hw[b][1])
        ymin, xmin, ymax, xmax = g[:, 0], g[:, 1], g[:, 2], g[:, 3]
        if mirror is not None and mirror[b]:
            xmin, xmax = pw - f32(1.) - xmax, pw - f32(1.) - xmin
        th, tw = f32(out_shape[0]), f32(out_shape[1])
        ymin, ymax = ymin * th / ph, ymax * th / ph
        xmin, xmax = xmin * tw / pw, xmax * tw / pw
        keep = ((ymax - ymin) > f32(min_height)) & ((xmax - xmin) > f32(min_width))
        out = np.stack([ymin, xmin, ymax, xmax], -1).astype(f32)[keep]
        if out.shape[0] > 0:
            boxes.append(out)
            offs.append(offs[-1] + out.shape[0])
            idx.append(b)
    cat = np.concatenate(boxes, 0) if boxes else np.zeros((0, 4), dtype=f32)
    return cat, np.asarray(offs, dtype=i32), np.asarray(idx, dtype=i32)
