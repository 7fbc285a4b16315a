// CPU ORACLE (test infrastructure).  Shims that let the reference's OWN
// SmallMiningMatch CPU functor compile outside TensorFlow.
//
// oracle/Makefile extracts /root/reference/cpp/ExtraLib/small_mining_match.cc
// lines 56-63 (DistancePair) and 67-284 (SmallMiningMatchFunctor<CPUDevice,T>)
// into oracle/_ref/functor_extract.inc at build time -- the reference source is
// never copied into this repository -- and this file supplies just enough of the
// TensorFlow types it names.  Shard() runs the whole range on the calling thread.
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <limits>
#include <queue>
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
struct ThreadPoolStub {};
struct DeviceBase {
  struct CpuWorkerThreads {
    int num_threads = 1;
    ThreadPoolStub* workers = nullptr;
  };
  CpuWorkerThreads threads;
  const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &threads; }
};
struct OpKernelContext {
  DeviceBase dev;
  DeviceBase* device() { return &dev; }
};
struct CPUDevice {};
inline void Shard(int, ThreadPoolStub*, int64_t total, int64_t,
                  std::function<void(int64_t, int64_t)> work) {
  work(0, total);
}
template <typename Device, typename T>
struct SmallMiningMatchFunctor;

#include "functor_extract.inc"

extern "C" int ref_small_mining_match(const float* overlaps, int32_t num_anchors, int32_t num_gt,
                                      float neg_low, float neg_high, float pos_thres,
                                      int32_t min_match, float stop_thres, int32_t* match,
                                      float* scores) {
  std::vector<int32_t> gt_match_num(num_gt > 0 ? num_gt : 1, 0);
  std::vector<int32_t> gt_small_topk(static_cast<size_t>(num_gt > 0 ? num_gt : 1) * min_match, 0);
  OpKernelContext ctx;
  CPUDevice dev;
  SmallMiningMatchFunctor<CPUDevice, float>()(
      &ctx, dev, ConstFlatView<float>{overlaps, static_cast<int64_t>(num_anchors) * num_gt},
      FlatView<int32_t>{match, num_anchors}, FlatView<float>{scores, num_anchors},
      FlatView<int32_t>{gt_match_num.data(), num_gt},
      FlatView<int32_t>{gt_small_topk.data(), static_cast<int64_t>(num_gt) * min_match},
      num_anchors, num_gt, neg_low, neg_high, pos_thres, stop_thres, min_match);
  return 0;
}
