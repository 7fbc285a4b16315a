// CPU ORACLE (test infrastructure).  Lets the reference's OWN DynamicAnchorRouting CPU functor compile outside
// TensorFlow.  oracle/Makefile extracts /root/reference/cpp/ExtraLib/dynamic_anchor_routing.cc lines 188-518
// (DynamicAnchorRoutingFunctor<CPUDevice, T>, both branches) into oracle/_ref/dar_extract.inc at build time -- the
// reference source is never copied into this repository -- and this file supplies the few TensorFlow names it uses.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <limits>
#include <queue>
#include <random>
#include <vector>

namespace tensorflow {}
using namespace tensorflow;

template <typename T>
struct FlatView {
  T* ptr;
  int64_t len;
  T* data() const { return ptr; }
  FlatView& setZero() {
    for (int64_t i = 0; i < len; ++i) ptr[i] = T(0);
    return *this;
  }
};
template <typename T>
struct ConstFlatView {
  const T* ptr;
  int64_t len;
  const T* data() const { return ptr; }
};
template <typename T>
struct TTypes {
  typedef FlatView<T> Flat;
  typedef ConstFlatView<T> ConstFlat;
};
struct OpKernelContext {};
struct CPUDevice {};
template <typename Device, typename T>
struct DynamicAnchorRoutingFunctor;

#include "dar_extract.inc"

// One layer, exactly the op's signature (dynamic_anchor_routing.cc:31-59); matched_num / prior_prob are the op's temporaries.
extern "C" int ref_dynamic_anchor_routing(const float* anchors, const float* gt_targets, const float* labels,
                                          const int32_t* mask_in, int64_t num_anchors, int32_t feat_height,
                                          int32_t feat_width, int32_t anchor_depth, int32_t feat_strides,
                                          int32_t img_height, int32_t img_width, int32_t training, float thres,
                                          float ignore_thres, int32_t* mask_out, float* decode_out) {
  std::vector<int32_t> matched_num(num_anchors > 0 ? num_anchors : 1, 0);
  std::vector<float> prior_prob(num_anchors > 0 ? num_anchors : 1, 0.f);
  OpKernelContext ctx;
  CPUDevice dev;
  DynamicAnchorRoutingFunctor<CPUDevice, float>()(
      &ctx, dev, ConstFlatView<float>{anchors, num_anchors * 4}, ConstFlatView<float>{gt_targets, num_anchors * 4},
      ConstFlatView<float>{labels, num_anchors}, ConstFlatView<int32_t>{mask_in, num_anchors}, feat_height, feat_width,
      anchor_depth, feat_strides, img_height, img_width, FlatView<int32_t>{matched_num.data(), num_anchors},
      FlatView<float>{prior_prob.data(), num_anchors}, FlatView<int32_t>{mask_out, num_anchors},
      FlatView<float>{decode_out, num_anchors * 4}, training != 0, num_anchors, thres, ignore_thres);
  return 0;
}
