// CPU ORACLE (test infrastructure -- NOT a product path; see oracle/README.md).
//
// Native pieces of the oracle that would be too slow as Python loops:
//   * oracle_small_mining_match : restatement of the SmallMiningMatch CPU kernel
//       /root/reference/cpp/ExtraLib/small_mining_match.cc:67-284
//   * oracle_tf_nms             : restatement of tf.image.non_max_suppression
//       (TensorFlow r1.8 core/kernels/non_max_suppression_op.cc -- TF is a
//        dependency that is absent from /root/reference; call sites
//        utility/bbox_util.py:77,82)
//
// Built by oracle/Makefile with g++ -O2 -ffp-contract=off (no fma contraction:
// every fp32 op is separately rounded, as in the reference build, -O2 on x86-64
// without -mfma).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <vector>

namespace {

// Heap entry for the compensation stage.  Ordering is on the overlap only, the
// same strict-weak order the reference gives its priority_queue element
// (small_mining_match.cc:56-63), so libstdc++'s heap breaks ties identically.
struct Candidate {
  float overlap;
  int32_t anchor;
};
struct ByOverlap {
  bool operator()(const Candidate& a, const Candidate& b) const { return a.overlap < b.overlap; }
};

}  // namespace

extern "C" {

// overlaps: row-major [num_anchors, num_gt].  Outputs: match [num_anchors] int32
// (>=0 GT index, -1 negative, -2 ignore), scores [num_anchors].  Returns 0, or
// -1 when an attribute violates the op's constructor checks (:292-305).
int oracle_small_mining_match(const float* overlaps, int32_t num_anchors, int32_t num_gt,
                              float neg_low, float neg_high, float pos_thres, int32_t min_match,
                              float stop_thres, int32_t* match, float* scores) {
  if (!(neg_low >= 0.f && neg_low < 1.f)) return -1;
  if (!(neg_high > neg_low && neg_high < 1.f)) return -1;
  if (!(pos_thres >= neg_high && pos_thres < 1.f)) return -1;
  if (!(stop_thres >= 0.f && stop_thres < 1.f)) return -1;
  if (min_match < 1) return -1;

  const float kEps = std::numeric_limits<float>::epsilon();
  std::vector<int32_t> matched_per_gt(num_gt, 0);

  // Stage 1 (:72-94): every anchor takes its first strictly-greatest GT.
  for (int32_t a = 0; a < num_anchors; ++a) {
    const float* row = overlaps + static_cast<int64_t>(a) * num_gt;
    float best = std::numeric_limits<float>::lowest();
    int32_t best_gt = 0;
    for (int32_t g = 0; g < num_gt; ++g) {
      if (row[g] > best) { best = row[g]; best_gt = g; }
    }
    scores[a] = best;
    if (best >= neg_low && best < neg_high) {
      match[a] = -1;
    } else if (best >= pos_thres) {
      match[a] = best_gt;
      matched_per_gt[best_gt] += 1;
    } else {
      match[a] = -2;
    }
  }

  // Stage 2 (:160-197): GTs in ascending order force every anchor whose overlap
  // is within FLT_EPSILON of the column maximum; a later GT overwrites an earlier.
  std::vector<int32_t> near_running_max;
  for (int32_t g = 0; g < num_gt; ++g) {
    float col_max = std::numeric_limits<float>::lowest();
    near_running_max.clear();
    for (int32_t a = 0; a < num_anchors; ++a) {
      const float v = overlaps[static_cast<int64_t>(a) * num_gt + g];
      if (v > col_max) col_max = v;
      if (std::abs(v - col_max) < kEps) near_running_max.push_back(a);
    }
    for (int32_t a : near_running_max) {
      const float v = overlaps[static_cast<int64_t>(a) * num_gt + g];
      if (!(std::abs(v - col_max) < kEps)) continue;
      scores[a] = v;
      if (match[a] > -1) matched_per_gt[match[a]] -= 1;
      match[a] = g;
      matched_per_gt[g] += 1;
    }
  }

  // Stage 3 (:199-222): "hard face compensation".  A GT short of min_match pulls
  // still-unmatched anchors with overlap > stop_thres, best first (heap order).
  for (int32_t g = 0; g < num_gt; ++g) {
    if (matched_per_gt[g] >= min_match) continue;
    std::priority_queue<Candidate, std::vector<Candidate>, ByOverlap> heap;
    for (int32_t a = 0; a < num_anchors; ++a) {
      const float v = overlaps[static_cast<int64_t>(a) * num_gt + g];
      if (match[a] < 0 && v > stop_thres) heap.push(Candidate{v, a});
    }
    while (!heap.empty() && matched_per_gt[g] < min_match) {
      const Candidate c = heap.top();
      heap.pop();
      matched_per_gt[g] += 1;
      scores[c.anchor] = c.overlap;
      match[c.anchor] = g;
    }
  }
  return 0;
}

// tf.image.non_max_suppression, TF r1.8 semantics (restated from the published
// kernel; UNPINNED: no reference test covers it):
//   * candidates visited in decreasing score order (std::sort on indices with
//     comparator score[i] > score[j]; tie_mode 0 = stable order instead, which is
//     the tie policy this repository documents; tie_mode 1 = literal std::sort);
//   * corners are min/max normalised, area = (ymax-ymin)*(xmax-xmin) (no +1);
//   * a pair where either area <= 0 never suppresses;
//   * suppressed iff inter / (area_i + area_j - inter) > iou_threshold (strict);
//   * stops once max_output boxes are selected.
// boxes: [num_boxes,4] (y1,x1,y2,x2).  Writes selected indices, returns the count.
int oracle_tf_nms(const float* boxes, const float* scores, int32_t num_boxes, int32_t max_output,
                  float iou_threshold, int32_t tie_mode, int32_t* selected_out) {
  std::vector<int32_t> order(num_boxes);
  for (int32_t i = 0; i < num_boxes; ++i) order[i] = i;
  auto by_score = [scores](int32_t i, int32_t j) { return scores[i] > scores[j]; };
  if (tie_mode == 1) std::sort(order.begin(), order.end(), by_score);
  else std::stable_sort(order.begin(), order.end(), by_score);

  auto suppresses = [boxes, iou_threshold](int32_t i, int32_t j) -> bool {
    const float* bi = boxes + 4 * static_cast<int64_t>(i);
    const float* bj = boxes + 4 * static_cast<int64_t>(j);
    const float ymin_i = std::min(bi[0], bi[2]), xmin_i = std::min(bi[1], bi[3]);
    const float ymax_i = std::max(bi[0], bi[2]), xmax_i = std::max(bi[1], bi[3]);
    const float ymin_j = std::min(bj[0], bj[2]), xmin_j = std::min(bj[1], bj[3]);
    const float ymax_j = std::max(bj[0], bj[2]), xmax_j = std::max(bj[1], bj[3]);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.f || area_j <= 0.f) return false;
    const float iy0 = std::max(ymin_i, ymin_j), ix0 = std::max(xmin_i, xmin_j);
    const float iy1 = std::min(ymax_i, ymax_j), ix1 = std::min(xmax_i, xmax_j);
    const float inter = std::max(iy1 - iy0, 0.f) * std::max(ix1 - ix0, 0.f);
    const float iou = inter / (area_i + area_j - inter);
    return iou > iou_threshold;
  };

  std::vector<int32_t> kept;   // indices into the ORIGINAL arrays
  for (int32_t r = 0; r < num_boxes; ++r) {
    if (static_cast<int32_t>(kept.size()) >= max_output) break;
    const int32_t cand = order[r];
    bool keep = true;
    for (int32_t k = static_cast<int32_t>(kept.size()) - 1; k >= 0; --k) {
      if (suppresses(cand, kept[k])) { keep = false; break; }
    }
    if (keep) kept.push_back(cand);
  }
  for (size_t k = 0; k < kept.size(); ++k) selected_out[k] = kept[k];
  return static_cast<int32_t>(kept.size());
}


// ---------------------------------------------------------------------------------------------
// DynamicAnchorRouting, EVALUATION branch (cpp/ExtraLib/dynamic_anchor_routing.cc:328-408), one layer.
// Our restatement, written as the two phases the CUDA path uses, so that the phase split itself is what gets pinned
// against the reference's sequential loop (oracle/_ref/libdar_ref.so):
//   phase 1  every source anchor i (mask_in >= 1, a box of at least 1x1 px that is not more than one cell outside the
//            feature map) is re-binned to the cell of its rounded centre, same depth slot -> target t (:352-369).
//            The reference keeps, per target, the running STRICT maximum of the label in index order (:370-381), but a
//            target that is itself easy background (mask_in[t] < 1) turns to -1 when the loop reaches i == t (:331-334)
//            and rejects every later source (:371): such a target only sees sources i < t.
//            => winner(t) = max label, ties -> lowest index, over the admissible sources; labels <= 0 never win.
//   phase 2  mask_out[t] = 1 iff mask_in[t] >= 1 and t has a winner (:384); the winner's box (or zeros) is the prior
//            that the stage-2 offsets gt_targets[t] are decoded against, in the reference's mixed float/double
//            arithmetic (:385-406); std::exp(float) is libm's expf.
// ---------------------------------------------------------------------------------------------
int oracle_dynamic_anchor_routing_eval(const float* anchors, const float* gt_targets, const float* labels,
                                       const int32_t* mask_in, int64_t num_anchors, int32_t feat_height,
                                       int32_t feat_width, int32_t anchor_depth, int32_t feat_strides,
                                       int32_t* mask_out, float* decode_out) {
  std::vector<int64_t> winner(num_anchors > 0 ? num_anchors : 1, -1);
  for (int64_t i = 0; i < num_anchors; ++i) {
    if (mask_in[i] < 1) continue;
    const float ymin = anchors[i * 4], xmin = anchors[i * 4 + 1], ymax = anchors[i * 4 + 2], xmax = anchors[i * 4 + 3];
    if (xmax - xmin < 1 || ymax - ymin < 1) continue;
    if (xmin / feat_strides < -1 || xmax / feat_strides > feat_width + 1 - 1.) continue;
    if (ymin / feat_strides < -1 || ymax / feat_strides > feat_height + 1 - 1.) continue;
    int64_t cx = static_cast<int64_t>(std::round((xmin + xmax) / (2. * feat_strides)));
    int64_t cy = static_cast<int64_t>(std::round((ymin + ymax) / (2. * feat_strides)));
    cx = std::max<int64_t>(std::min<int64_t>(cx, feat_width - 1), 0);
    cy = std::max<int64_t>(std::min<int64_t>(cy, feat_height - 1), 0);
    const int64_t t = (cy * feat_width + cx) * anchor_depth + i % anchor_depth;
    if (mask_in[t] < 1 && i > t) continue;                 // the target has already turned to -1
    if (!(labels[i] > 0.f)) continue;
    if (winner[t] < 0 || labels[i] > labels[winner[t]]) winner[t] = i;    // ascending i: ties keep the lowest index
  }
  for (int64_t t = 0; t < num_anchors; ++t) {
    const int64_t w = winner[t];
    mask_out[t] = (w >= 0 && mask_in[t] >= 1) ? 1 : 0;
    const float ymin = w >= 0 ? anchors[w * 4] : 0.f, xmin = w >= 0 ? anchors[w * 4 + 1] : 0.f;
    const float ymax = w >= 0 ? anchors[w * 4 + 2] : 0.f, xmax = w >= 0 ? anchors[w * 4 + 3] : 0.f;
    const float prior_cy = (ymin + ymax) / 2.;
    const float prior_cx = (xmin + xmax) / 2.;
    const float prior_h = (ymax - ymin + 1.);
    const float prior_w = (xmax - xmin + 1.);
    float pred_cy = gt_targets[t * 4], pred_cx = gt_targets[t * 4 + 1];
    float pred_h = gt_targets[t * 4 + 2], pred_w = gt_targets[t * 4 + 3];
    pred_h = std::exp(pred_h) * prior_h;
    pred_w = std::exp(pred_w) * prior_w;
    pred_cy = pred_cy * prior_h + prior_cy;
    pred_cx = pred_cx * prior_w + prior_cx;
    decode_out[t * 4] = pred_cy - (pred_h - 1.) / 2.;
    decode_out[t * 4 + 1] = pred_cx - (pred_w - 1.) / 2.;
    decode_out[t * 4 + 2] = pred_cy + (pred_h - 1.) / 2.;
    decode_out[t * 4 + 3] = pred_cx + (pred_w - 1.) / 2.;
  }
  return 0;
}

}  // extern "C"
