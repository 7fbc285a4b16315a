"""The caller side of encode_anchors: what the reference does to the ground-truth boxes of a sampled
patch before `anchor_encoder(gbboxes)` (dataset_common.py:150) and how images reach the batch:

    mirror        preprocessing/sfd_preprocessing.py:482-493  (the random coin is an input here)
    rescale       preprocessing/sfd_preprocessing.py:529-533  patch pixels -> net-input pixels
    small faces   preprocessing/sfd_preprocessing.py:544-550  keep h > 6 and w > 3
    keep_input    dataset/dataset_common.py:178,186           images left without boxes are not batched

One kernel launch for the whole batch; the result is the CSR pair `AnchorEncoder.encode_anchors_batch` takes."""
from __future__ import annotations

from .. import _lib as L
from .. import functional as F


def prepare_gt_batch(gt_boxes, gt_offsets, patch_hw, out_shape, mirror=None, min_height=6., min_width=3., trim=True):
    """gt_boxes [total, 4] (ymin, xmin, ymax, xmax) in patch pixels, gt_offsets int32 [B+1], patch_hw [B, 2].
    -> (gt_boxes, gt_offsets, image_index) of the images that stay in the batch.  trim=True reads the two counts
    back (one synchronisation) and returns exact-size views; trim=False returns the capacity-size buffers + counts."""
    ob, oo, oi, oc = F.gt_handoff(L.as_f32(gt_boxes).reshape(-1, 4), gt_offsets, L.as_f32(patch_hw), out_shape, mirror,
                                  min_height, min_width)
    if not trim:
        return ob, oo, oi, oc
    images, boxes = (int(v) for v in oc.tolist())
    return ob[:boxes], oo[:images + 1], oi[:images]
