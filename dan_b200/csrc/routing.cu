// DynamicAnchorRouting, EVALUATION branch, for all pyramid layers and images of a batch at once.
//
// Reference: cpp/ExtraLib/dynamic_anchor_routing.cc:328-408 (CPU-only TF op, one call per layer and image, pinned to
// /cpu:0 between the two GPU stages of DAN's evaluation graph, eval_dan.py:383-391).  The reference loop is sequential
// and order dependent; its result has a closed parallel form (pinned against the reference's own compiled functor in
// tests/test_oracle.py):
//   route_scatter_kernel   every admissible source anchor i is re-binned to the cell of its rounded box centre (same
//                          depth slot) -> target t (:352-369); winner(t) = highest label, ties -> lowest index, via one
//                          64-bit atomicMax of (label bits << 32 | ~i).  A target that is easy background itself
//                          (mask_in[t] < 1) turns to -1 when the reference loop reaches it (:331-334) and rejects later
//                          sources (:371): it only admits sources i < t.  Labels <= 0 never win (:370, prior_prob = 0).
//   route_finalize_kernel  mask_out[t] = winner && mask_in[t] >= 1 (:384); the winner's box (or zeros) is the prior the
//                          stage-2 offsets are decoded against, in the reference's mixed float/double arithmetic
//                          (:385-406).  std::exp(float) there is glibc's expf: restated below bit for bit.
// The training branch (:203-327) draws from an unseeded std::random_device and is not reproducible: not built.
#include "common.cuh"

namespace dan {

namespace {

struct RouteArgs {
  const float4* anchors;      // [B, N] decoded stage-1 boxes (ymin, xmin, ymax, xmax)
  const float4* targets;      // [B, N] stage-2 offsets (cy, cx, h, w), already divided by the prior scaling
  const float* labels;        // [B, N] stage-2 face probability
  const int32_t* mask_in;     // [B, N] stage-1 "not easy background"
  int64_t n;
  int batch;
  dan_routing_layers layers;
  unsigned long long* win;    // [B, N] workspace, zeroed
  int32_t* mask_out;          // [B, N]
  float4* decode_out;         // [B, N]
};

// glibc >= 2.27 expf (sysdeps/ieee754/flt-32/e_expf.c: the ARM optimized-routines algorithm): x is widened to double,
// k = round(x * 32 / ln 2) through the 0x1.8p52 shift trick, 2^(k/32) from a 32-entry table, degree-3 polynomial in
// double, ONE rounding to float at the end.  Every double operation is separately rounded (no FMA), like the generic
// x86-64 build.  Checked bit for bit against libm's expf on 500 000 inputs (tests/test_oracle.py).
__constant__ unsigned long long kExp2fTable[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

DAN_D float libm_expf(float x) {
  if (x != x) return x;
  if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);      // overflow
  if (x < -0x1.9fe368p6f) return 0.f;                            // underflow
  const double z = __dmul_rn(0x1.71547652b82fep+5, (double)x);   // x * N / ln 2
  double kd = __dadd_rn(z, 0x1.8p52);
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dsub_rn(kd, 0x1.8p52);
  const double r = __dsub_rn(z, kd);
  const unsigned long long t = kExp2fTable[ki & 31ull] + (ki << 47);
  const double s = __longlong_as_double((long long)t);
  const double p = __dadd_rn(__dmul_rn(0x1.c6af84b912394p-20, r), 0x1.ebfce50fac4f3p-13);
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(0x1.62e42ff0c52d6p-6, r), 1.0);
  y = __dadd_rn(__dmul_rn(p, r2), y);
  y = __dmul_rn(y, s);
  return __double2float_rn(y);
}

DAN_D int layer_of(const dan_routing_layers& L, int64_t g, int64_t& local) {
  int64_t off = 0;
  for (int l = 0; l < L.num_layers; ++l) {
    const int64_t cnt = (int64_t)L.feat_height[l] * L.feat_width[l] * L.anchor_depth[l];
    if (g < off + cnt) {
      local = g - off;
      return l;
    }
    off += cnt;
  }
  local = 0;
  return -1;
}

__global__ void __launch_bounds__(256) route_zero_kernel(unsigned long long* win, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) win[i] = 0ull;
}

// dynamic_anchor_routing.cc:330-382
__global__ void __launch_bounds__(256) route_scatter_kernel(const __grid_constant__ RouteArgs A) {
  const int64_t total = (int64_t)A.batch * A.n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    if (A.mask_in[idx] < 1) continue;                                            // :331 easy background
    const float label = A.labels[idx];
    if (!(label > 0.f)) continue;                                                // :370 against prior_prob = 0
    const int64_t b = idx / A.n, g = idx - b * A.n;
    int64_t i;
    const int l = layer_of(A.layers, g, i);
    if (l < 0) continue;
    const int H = A.layers.feat_height[l], W = A.layers.feat_width[l], D = A.layers.anchor_depth[l];
    const float stride = (float)A.layers.feat_strides[l];
    const float4 a = A.anchors[idx];                                             // ymin, xmin, ymax, xmax
    if (fsub(a.w, a.y) < 1.f || fsub(a.z, a.x) < 1.f) continue;                   // :347 invalid box
    if (fdiv(a.y, stride) < -1.f || fdiv(a.w, stride) > (float)W) continue;       // :352
    if (fdiv(a.x, stride) < -1.f || fdiv(a.z, stride) > (float)H) continue;       // :361
    const double two_s = __dmul_rn(2.0, (double)A.layers.feat_strides[l]);
    long long cx = (long long)round(__ddiv_rn((double)fadd(a.y, a.w), two_s));   // :351 std::round: half away from zero
    long long cy = (long long)round(__ddiv_rn((double)fadd(a.x, a.z), two_s));   // :360
    cx = max(min(cx, (long long)(W - 1)), 0ll);
    cy = max(min(cy, (long long)(H - 1)), 0ll);
    const int64_t t = (cy * W + cx) * D + i % D;                                  // :369
    const int64_t tidx = idx - i + t;
    if (A.mask_in[tidx] < 1 && i > t) continue;                                  // the target has already turned to -1 (:371)
    const unsigned long long key = ((unsigned long long)__float_as_uint(label) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    atomicMax(A.win + tidx, key);
  }
}

// dynamic_anchor_routing.cc:383-407
__global__ void __launch_bounds__(256) route_finalize_kernel(const __grid_constant__ RouteArgs A) {
  const int64_t total = (int64_t)A.batch * A.n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = A.win[idx];
    float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
    int m = 0;
    if (key != 0ull) {
      const int64_t b = idx / A.n, g = idx - b * A.n;
      int64_t t;
      layer_of(A.layers, g, t);
      const int64_t w = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
      pr = A.anchors[idx - t + w];
      m = A.mask_in[idx] >= 1 ? 1 : 0;
    }
    A.mask_out[idx] = m;
    // float sums, then the double constants of the reference source (`/ 2.`, `+ 1.`), rounded to float on assignment
    const float prior_cy = __double2float_rn(__ddiv_rn((double)fadd(pr.x, pr.z), 2.0));
    const float prior_cx = __double2float_rn(__ddiv_rn((double)fadd(pr.y, pr.w), 2.0));
    const float prior_h = __double2float_rn(__dadd_rn((double)fsub(pr.z, pr.x), 1.0));
    const float prior_w = __double2float_rn(__dadd_rn((double)fsub(pr.w, pr.y), 1.0));
    const float4 p = A.targets[idx];                                              // cy, cx, h, w
    const float ph = fmul(libm_expf(p.z), prior_h);
    const float pw = fmul(libm_expf(p.w), prior_w);
    const float pcy = fadd(fmul(p.x, prior_h), prior_cy);
    const float pcx = fadd(fmul(p.y, prior_w), prior_cx);
    const double hh = __ddiv_rn(__dsub_rn((double)ph, 1.0), 2.0), hw = __ddiv_rn(__dsub_rn((double)pw, 1.0), 2.0);
    A.decode_out[idx] = make_float4(__double2float_rn(__dsub_rn((double)pcy, hh)), __double2float_rn(__dsub_rn((double)pcx, hw)),
                                    __double2float_rn(__dadd_rn((double)pcy, hh)), __double2float_rn(__dadd_rn((double)pcx, hw)));
  }
}

}  // namespace

}  // namespace dan

using namespace dan;

extern "C" {

size_t dan_routing_workspace_bytes(int64_t num_anchors, int32_t batch) {
  if (num_anchors < 0 || batch < 0) return 0;
  return align_up((size_t)num_anchors * (size_t)batch * 8, 256);
}

int dan_dynamic_anchor_routing_eval(const dan_routing_layers* h_layers, const float* anchors, const float* gt_targets,
                                    const float* labels, const int32_t* mask_in, int64_t num_anchors, int32_t batch,
                                    int32_t* mask_out, float* decode_out, void* workspace, size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(h_layers != nullptr, DAN_ERR_INVALID_ARGUMENT, "layers is NULL");
  DAN_REQUIRE(h_layers->num_layers >= 1 && h_layers->num_layers <= DAN_MAX_LAYERS, DAN_ERR_INVALID_ARGUMENT, "num_layers must be in [1, %d]",
              DAN_MAX_LAYERS);
  DAN_REQUIRE(num_anchors >= 0 && batch >= 0, DAN_ERR_INVALID_ARGUMENT, "negative size");
  int64_t sum = 0;
  for (int l = 0; l < h_layers->num_layers; ++l) {
    DAN_REQUIRE(h_layers->feat_height[l] >= 1 && h_layers->feat_width[l] >= 1 && h_layers->anchor_depth[l] >= 1 && h_layers->feat_strides[l] >= 1,
                DAN_ERR_INVALID_ARGUMENT, "layer %d: feat_height, feat_width, anchor_depth and feat_strides must be >= 1", l);
    sum += (int64_t)h_layers->feat_height[l] * h_layers->feat_width[l] * h_layers->anchor_depth[l];
  }
  DAN_REQUIRE(sum == num_anchors, DAN_ERR_INVALID_ARGUMENT, "the layers hold %lld anchors, num_anchors is %lld", (long long)sum,
              (long long)num_anchors);
  DAN_REQUIRE(num_anchors < ((int64_t)1 << 32), DAN_ERR_UNSUPPORTED, "more than 2^32 anchors per image");
  const int64_t total = num_anchors * batch;
  if (total == 0) return DAN_OK;
  DAN_REQUIRE(anchors && gt_targets && labels && mask_in && mask_out && decode_out, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(aligned16(anchors) && aligned16(gt_targets) && aligned16(decode_out), DAN_ERR_INVALID_ARGUMENT, "box tensors must be 16-byte aligned");
  const size_t need = dan_routing_workspace_bytes(num_anchors, batch);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= need, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need,
              workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  RouteArgs A = {};
  A.anchors = reinterpret_cast<const float4*>(anchors);
  A.targets = reinterpret_cast<const float4*>(gt_targets);
  A.labels = labels;
  A.mask_in = mask_in;
  A.n = num_anchors;
  A.batch = batch;
  A.layers = *h_layers;
  A.win = static_cast<unsigned long long*>(workspace);
  A.mask_out = mask_out;
  A.decode_out = reinterpret_cast<float4*>(decode_out);
  route_zero_kernel<<<grid_for(total), 256, 0, st>>>(A.win, total);
  DAN_LAUNCH_CHECK("route_zero_kernel");
  route_scatter_kernel<<<grid_for(total), 256, 0, st>>>(A);
  DAN_LAUNCH_CHECK("route_scatter_kernel");
  route_finalize_kernel<<<grid_for(total), 256, 0, st>>>(A);
  DAN_LAUNCH_CHECK("route_finalize_kernel");
  return DAN_OK;
}

}  // extern "C"
