"""Mirror of the part of the reference's ``utility/custom_op.py`` that is on the anchor hot path.

The reference loads ``cpp/ExtraLib/build/libextra_lib.so`` with ``tf.load_op_library``
(custom_op.py:33-48) and exposes ``small_mining_match`` (:47).  Here the op is a call
into libdan_b200.so; the Deform ops that the reference's module also drags in at import
time (custom_op.py:59-60) are not part of this path and are deliberately not coupled."""
from __future__ import annotations

from .. import _lib as L
from .. import functional as F


def small_mining_match(overlaps, negative_low_thres, negative_high_thres, positive_thres, min_match,
                       stop_positive_thres):
    """SmallMiningMatch (cpp/ExtraLib/small_mining_match.cc:31-54).

    overlaps: [num_anchors, num_ground_truth] fp32 in [0, 1]
    -> (match_indices int32 [num_anchors], match_scores fp32 [num_anchors]);
    match_indices: GT index, -1 negative, -2 ignore.  Attribute violations raise DanError
    with the op's own InvalidArgument conditions (:292-305)."""
    return F.small_mining_match(L.as_f32(overlaps), negative_low_thres, negative_high_thres, positive_thres,
                                min_match, stop_positive_thres)


def dynamic_anchor_routing(anchors, gt_targets, labels, mask_in, feat_height, feat_width, anchor_depth, feat_strides,
                           img_height, img_width, trainging=False, thres=0.03, ignore_thres=0.0):
    """DynamicAnchorRouting (cpp/ExtraLib/dynamic_anchor_routing.cc:31-59), positional order of the generated TF wrapper:
    inputs, then the attrs ``trainging`` (sic), ``thres``, ``ignore_thres`` (call site eval_dan.py:390).

    anchors / gt_targets [n, 4], labels [n] fp32, mask_in [n] int32 with n = feat_height*feat_width*anchor_depth.
    -> (mask_out int32 [n], decode_out fp32 [n, 4]).  Only the evaluation branch exists (trainging=False): the training
    branch draws from an unseeded std::random_device (:196-198) and has no reproducible result."""
    import torch
    if not (0. <= thres < 1.):                      # :527
        raise L.DanError(-1, "Need Attr 1 > thres >= 0., got %g" % thres)
    if not (0. <= ignore_thres < 1.):               # :529
        raise L.DanError(-1, "Need Attr 1 > ignore_thres >= 0., got %g" % ignore_thres)
    if trainging:
        raise NotImplementedError("DynamicAnchorRouting: the training branch is random (std::random_device) and not provided")
    if anchors.dim() != 2 or gt_targets.dim() != 2:  # :546-547
        raise L.DanError(-1, "anchors / gt_targets must be in 'num_anchors x 4' format.")
    if labels.dim() != 1 or mask_in.dim() != 1:      # :548-549
        raise L.DanError(-1, "labels / mask must be in 'num_anchors' format.")
    layers = F.routing_layers([int(feat_height)], [int(feat_width)], [int(anchor_depth)], [int(feat_strides)])
    return F.dynamic_anchor_routing_eval(layers, L.as_f32(anchors), L.as_f32(gt_targets), L.as_f32(labels),
                                         mask_in.to(torch.int32))


def dynamic_anchor_routing_layers(anchors, gt_targets, labels, mask_in, feat_heights, feat_widths, anchors_depth, feat_strides):
    """The per-layer loop of eval_dan.py:386-393 (split -> op per layer -> concat) as ONE call, batched over images:
    anchors / gt_targets [B, N, 4] (or [N, 4]), labels / mask_in [B, N] (or [N]), N = all layers concatenated."""
    import torch
    layers = F.routing_layers(feat_heights, feat_widths, anchors_depth, feat_strides)
    return F.dynamic_anchor_routing_eval(layers, L.as_f32(anchors), L.as_f32(gt_targets), L.as_f32(labels),
                                         mask_in.to(torch.int32))
