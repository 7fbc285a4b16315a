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
