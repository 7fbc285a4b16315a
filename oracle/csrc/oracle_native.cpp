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

}  // extern "C"
