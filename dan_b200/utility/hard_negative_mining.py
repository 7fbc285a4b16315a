"""Drop-in for the reference's hard-negative mining helpers.

    mining_hard_neg                train_dan.py:286-324 (also train_dan_deform.py:286, train_pb.py:310; the inline copy
                                   train_sfd.py:349-384 is the same rule without the `tf.maximum(., 1)` clamp)
    mining_hard_neg_across_batch   train_dan.py:247-284

Same positional arguments and the same 4-tuple as the reference; `FLAGS.negative_ratio` / `FLAGS.num_classes` become
keyword arguments.  All work runs in dan_b200/csrc/mining.cu; the only host synchronisation is reading the two output
lengths (the reference's boolean_mask outputs have data-dependent shapes too)."""
from __future__ import annotations

from .. import functional as F


def _trim(r):
    selected, positives = (int(v) for v in r.counts.tolist())      # the one device->host read
    return r.cls_pred[:selected], r.loc_pred[:positives], r.cls_targets[:selected], r.loc_targets[:positives]


def mining_hard_neg(batch_size, cls_pred, location_pred, cls_targets, match_scores, loc_targets, name=None,
                    negative_ratio=3., num_classes=2, at_least_one=True, return_details=False):
    """-> (cls_pred[final_mask], location_pred[positive_mask], clip(cls_targets)[final_mask], loc_targets[positive_mask]).

    ``match_scores`` is accepted and unused, as in the reference (:288, the masked variant is commented out at :298).
    ``at_least_one=False`` is the train_sfd.py rule: an image without positives selects position -1 of its sorted row,
    which TensorFlow rejects on the CPU -> ValueError here."""
    r = F.hard_negative_mining(cls_pred, location_pred, cls_targets, loc_targets, int(batch_size), negative_ratio, num_classes,
                               at_least_one=at_least_one, strict_greater=False)
    if not at_least_one and int(r.n_neg_select.min()) < 1:
        bad = int((r.n_neg_select < 1).nonzero()[0])
        raise ValueError("indices[%d] = [%d, -1] does not index into param" % (bad, bad))
    out = _trim(r)
    return (out, r) if return_details else out


def mining_hard_neg_across_batch(batch_size, cls_pred, location_pred, cls_targets, match_scores, loc_targets, name=None,
                                 negative_ratio=3., num_classes=2, return_details=False):
    """One selection over the flattened batch with a strict `>` against the k-th value (train_dan.py:274)."""
    r = F.hard_negative_mining(cls_pred, location_pred, cls_targets, loc_targets, 1, negative_ratio, num_classes,
                               at_least_one=False, strict_greater=True)
    if int(r.n_neg_select.min()) < 1:
        raise ValueError("slice index -1 of dimension 0 out of bounds")
    out = _trim(r)
    return (out, r) if return_details else out
