// K2: anchor x GT matching and box encoding.
//
//   fused path  (dan_encode_batch):  IoU is evaluated on the fly, tile-culled, and
//                                    never written to HBM (the reference
//                                    materialises [N,M] fp32 and makes ~26 passes
//                                    over it, anchor_manipulator.py:24-105,287).
//   dense path  (dan_small_mining_match / dan_dual_max_match): the literal custom
//                                    op boundary, overlaps [N,M] given in HBM.
//
// Three launches per batch, all images of the batch in each launch:
//   pass 1  per-GT column maximum (the only cross-anchor quantity of stage 2 / of the dual matcher's "GT side")
//           -> atomicMax on fp32 bit patterns; the per-anchor row maximum / argmax found on the way are cached
//           (8 B/anchor); mining: anchors above stop_positive_thres are pushed to per-GT candidate buckets
//   pass 2  per anchor: stage 1 from the cached row maximum, tie test against the column maxima (stage 2 / dual
//           claim) only for the GTs this warp can reach, labels, encode, ALL outputs (44 B/anchor)
//   pass 3  (mining only) stage 3 "hard face compensation": one CTA per image walks the GTs in ascending order
//           (the stage is order dependent, small_mining_match.cc:199-222) and patches the few affected anchors
//
// Culling (fused path): a warp owns 32 consecutive anchors; it reduces their
// bounding box with redux.sync and tests 32 GT boxes per ballot against it.  Only
// GTs that can intersect some anchor of the warp are evaluated.  A skipped pair
// has an empty intersection, for which the reference computes exactly 0, so the
// result is unchanged.  GTs whose column maximum is 0 (dual: `==` claim) or below
// FLT_EPSILON (mining: tie band) tie with zero-overlap anchors as well and are
// therefore never culled in pass 2.
//
// Precondition shared with the op's own doc string (small_mining_match.cc:42-43):
// overlaps are in [0, 1], i.e. boxes have non-negative (+1 convention) extents.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "heap_order.cuh"

namespace dan {

constexpr int kEncThreads = 256;
constexpr int kGtChunk = 1024;     // GT boxes staged in shared memory at a time
constexpr int kBucketCap = 64;     // compensation candidates bucketed per GT before spilling

struct EncArgs {
  // anchors (fused)
  const float* ay0;
  const float* ax0;
  const float* ay1;
  const float* ax1;
  const uint8_t* mask;
  int n;
  // ground truth (fused), CSR over the batch
  const float4* gt;
  const int32_t* gt_off;
  // dense
  const float* overlaps;
  int m_dense;
  // parameters
  float low, high;       // dual: low/high thresholds; mining: negative_high / positive
  float neg_low, stop;
  int min_match;
  int ignore_between, gt_max_first;
  float ps0, ps1, ps2, ps3;
  float pa_scale;
  int debug;
  // workspace
  uint32_t* colmax;      // [G] fp32 bit patterns (>= 0)
  int32_t* cnt;          // [G] anchors matched per GT after stage 2
  int32_t* haspos;       // [G] GT has a positive anchor-side match (dual, gt_max_first=0)
  int32_t* fill;         // [G] candidates pushed per GT
  HeapItem* bucket;      // [G, kBucketCap]
  HeapItem* spill;       // [B, 3, n] (candidate list, sort buffer, heap of the overflow path)
  float* rowbest;        // [B, n] per-anchor row maximum, written by fused pass 1, read by fused pass 2
  int32_t* rowgt;        // [B, n] its (first) argmax
  // outputs
  float4* targets;
  int64_t* labels;
  float* scores;
  float4* matched;
  int32_t* match32;
  int64_t* match64;
};

// image b of the fused path owns per-GT slots [off[b] + b, off[b+1] + b + 1): one
// extra slot per image holds the dummy box of an empty image (:286).
struct ImageGt {
  int off0;     // first GT row in A.gt
  int m_real;   // GT rows given
  int m_eff;    // max(m_real, 1)
  int slot0;    // first per-GT workspace slot
};

template <bool DENSE>
DAN_D ImageGt image_gt(const EncArgs& A, int b) {
  ImageGt g;
  if (DENSE) {
    g.off0 = 0;
    g.m_real = g.m_eff = A.m_dense;
    g.slot0 = 0;
  } else {
    g.off0 = A.gt_off[b];
    g.m_real = A.gt_off[b + 1] - g.off0;
    g.m_eff = g.m_real > 0 ? g.m_real : 1;
    g.slot0 = g.off0 + b;
  }
  return g;
}

DAN_D float4 gt_box(const EncArgs& A, const ImageGt& ig, int j) {
  return (j < ig.m_real) ? __ldg(A.gt + ig.off0 + j) : make_float4(0.f, 0.f, 1.f, 1.f);
}

// The box an anchor is MATCHED with: itself (encode_anchors) or shrunk about its
// centre by pa_scale (encode_pa_anchors, anchor_manipulator.py:337-342).
struct AnchorBox {
  float y0, x0, y1, x1;  // original anchor
  float my0, mx0, my1, mx1, marea;  // matching box and its area
};

DAN_D AnchorBox load_anchor(const EncArgs& A, int a) {
  AnchorBox ab;
  ab.y0 = A.ay0[a];
  ab.x0 = A.ax0[a];
  ab.y1 = A.ay1[a];
  ab.x1 = A.ax1[a];
  if (A.pa_scale > 0.f) {
    float cy, cx, h, w;
    point2center(ab.y0, ab.x0, ab.y1, ab.x1, cy, cx, h, w);
    center2point(cy, cx, fdiv(h, A.pa_scale), fdiv(w, A.pa_scale), ab.my0, ab.mx0, ab.my1, ab.mx1);
  } else {
    ab.my0 = ab.y0;
    ab.mx0 = ab.x0;
    ab.my1 = ab.y1;
    ab.mx1 = ab.x1;
  }
  ab.marea = box_area(ab.my0, ab.mx0, ab.my1, ab.mx1);
  return ab;
}

// Encode anchor `a` of image `b` against GT box g (anchor_manipulator.py:306-326 /
// :366-387) and write every output row of a POSITIVE anchor.
DAN_D void write_positive(const EncArgs& A, int64_t row, const AnchorBox& ab, float4 g, float score, int gt_index) {
  float4 t;
  if (A.debug) {
    t = make_float4(ab.y0, ab.x0, ab.y1, ab.x1);
  } else {
    float gcy, gcx, gh, gw, acy, acx, ah, aw;
    point2center(g.x, g.y, g.z, g.w, gcy, gcx, gh, gw);
    point2center(ab.y0, ab.x0, ab.y1, ab.x1, acy, acx, ah, aw);
    t.x = fdiv(fdiv(fsub(gcy, acy), ah), A.ps0);
    t.y = fdiv(fdiv(fsub(gcx, acx), aw), A.ps1);
    if (A.pa_scale > 0.f) {
      t.z = fdiv(cephes_logf(fdiv(fmul(gh, A.pa_scale), ah)), A.ps2);
      t.w = fdiv(cephes_logf(fdiv(fmul(gw, A.pa_scale), aw)), A.ps3);
    } else {
      t.z = fdiv(cephes_logf(fdiv(gh, ah)), A.ps2);
      t.w = fdiv(cephes_logf(fdiv(gw, aw)), A.ps3);
    }
  }
  A.targets[row] = t;
  A.labels[row] = 1;
  A.scores[row] = score;
  if (A.matched != nullptr) A.matched[row] = g;
  if (A.match32 != nullptr) A.match32[row] = gt_index;
}

// ---------------------------------------------------------------------------
// per-warp GT iteration shared by pass 1 and pass 2 of the FUSED path
// ---------------------------------------------------------------------------
struct WarpBox {
  float y0, x0, y1, x1;
};

DAN_D WarpBox warp_bbox(bool active, const AnchorBox& ab) {
  const float inf = __int_as_float(0x7f800000);
  WarpBox wb;
  wb.y0 = warp_min_f(active ? ab.my0 : inf);
  wb.x0 = warp_min_f(active ? ab.mx0 : inf);
  wb.y1 = warp_max_f(active ? ab.my1 : -inf);
  wb.x1 = warp_max_f(active ? ab.mx1 : -inf);
  return wb;
}

// conservative "may intersect some anchor of the warp" test (see file header):
// a pair intersects only if fl(min(ymax) - max(ymin)) > -1, and the warp box
// dominates every anchor of the warp, so >= -1 on the warp box never misses.
DAN_D bool may_hit(const WarpBox& wb, float4 g) {
  const float dy = fsub(fminf(wb.y1, g.z), fmaxf(wb.y0, g.x));
  const float dx = fsub(fminf(wb.x1, g.w), fmaxf(wb.x0, g.y));
  return (dy >= -1.f) && (dx >= -1.f);
}

// ---------------------------------------------------------------------------
// pass 1: per-GT column maxima
// ---------------------------------------------------------------------------
template <bool DENSE, bool NEED_ROW>
__global__ void __launch_bounds__(kEncThreads) enc_pass1_kernel(const EncArgs A) {
  __shared__ float4 s_box[DENSE ? 1 : kGtChunk];
  __shared__ float s_area[DENSE ? 1 : kGtChunk];
  __shared__ uint32_t s_cm[DENSE ? 1 : kGtChunk];
  __shared__ float s_tile[DENSE ? (kEncThreads / 32) * 32 * 33 : 1];

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int a = blockIdx.x * kEncThreads + threadIdx.x;
  const bool valid = a < A.n;
  const ImageGt ig = image_gt<DENSE>(A, b);

  float best = 0.f;
  int best_gt = 0;

  if (DENSE) {
    float* tile = s_tile + warp * 32 * 33;
    const int row0 = blockIdx.x * kEncThreads + warp * 32;
    if (row0 >= A.n) return;
    for (int j0 = 0; j0 < ig.m_eff; j0 += 32) {
      const int jc = j0 + lane;
      for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        tile[r * 33 + lane] = (row < A.n && jc < ig.m_eff) ? A.overlaps[(int64_t)row * ig.m_eff + jc] : 0.f;
      }
      __syncwarp();
      // column maximum of the 32x32 tile: lane == column
      float cmx = 0.f;
      for (int r = 0; r < 32; ++r) cmx = fmaxf(cmx, tile[r * 33 + lane]);
      if (jc < ig.m_eff && cmx > 0.f) atomicMax(A.colmax + ig.slot0 + jc, __float_as_uint(cmx));
      if (NEED_ROW) {
        const int lim = min(32, ig.m_eff - j0);
        for (int jj = 0; jj < lim; ++jj) {
          const float ov = tile[lane * 33 + jj];
          if (ov > best) { best = ov; best_gt = j0 + jj; }
        }
      }
      __syncwarp();
    }
  } else {
    AnchorBox ab = {};
    bool active = false;
    if (valid) {
      ab = load_anchor(A, a);
      active = (A.mask == nullptr) || (A.mask[a] != 0);
    }
    const WarpBox wb = warp_bbox(active, ab);
    for (int c0 = 0; c0 < ig.m_eff; c0 += kGtChunk) {
      const int mc = min(kGtChunk, ig.m_eff - c0);
      __syncthreads();
      for (int k = threadIdx.x; k < mc; k += kEncThreads) {
        const float4 g = gt_box(A, ig, c0 + k);
        s_box[k] = g;
        s_area[k] = box_area(g.x, g.y, g.z, g.w);
        s_cm[k] = 0u;
      }
      __syncthreads();
      for (int k0 = 0; k0 < mc; k0 += 32) {
        const int k = k0 + lane;
        const bool test = (k < mc) && may_hit(wb, s_box[k]);
        unsigned hits = __ballot_sync(0xffffffffu, test);
        while (hits) {
          const int kk = k0 + __ffs(hits) - 1;
          hits &= hits - 1;
          const float4 g = s_box[kk];
          bool hit = false;
          float ov = 0.f;
          if (active) ov = pair_iou(ab.my0, ab.mx0, ab.my1, ab.mx1, ab.marea, g.x, g.y, g.z, g.w, s_area[kk], hit);
          const uint32_t key = (ov > 0.f) ? __float_as_uint(ov) : 0u;
          const uint32_t wmax = __reduce_max_sync(0xffffffffu, key);
          if (lane == 0 && wmax != 0u) atomicMax(&s_cm[kk], wmax);
          if (NEED_ROW && ov > best) { best = ov; best_gt = c0 + kk; }
        }
      }
      __syncthreads();
      for (int k = threadIdx.x; k < mc; k += kEncThreads)
        if (s_cm[k] != 0u) atomicMax(A.colmax + ig.slot0 + c0 + k, s_cm[k]);
    }
  }

  if (NEED_ROW && valid) {
    // anchor-side positive (match_indices >= 0 in anchor_manipulator.py:67-76)
    const bool less = best < A.low;
    const bool between = (best < A.high) && (best >= A.low);
    if (!less && !between) A.haspos[ig.slot0 + best_gt] = 1;
  }
}

// ---------------------------------------------------------------------------
// pass 2: per-anchor resolution + outputs
// ---------------------------------------------------------------------------
struct RowState {
  float best;      // row maximum, first strictly-greatest GT wins
  int best_gt;
  float ov0;       // overlap with GT 0
  // dual
  bool claimed;
  float cbest;
  int cbest_gt;
  // mining
  int owner;
  float owner_ov;
};

template <bool MINING>
DAN_D void row_update(RowState& s, int j, float ov, float cm, bool claimable) {
  if (ov > s.best) { s.best = ov; s.best_gt = j; }
  if (j == 0) s.ov0 = ov;
  if (MINING) {
    // stage 2 tie band, small_mining_match.cc:171,179; ascending j => last GT wins
    if (fabsf(fsub(ov, cm)) < FLT_EPSILON) { s.owner = j; s.owner_ov = ov; }
  } else {
    // anchor_manipulator.py:88 exact equality with the column maximum
    if (claimable && ov == cm) {
      s.claimed = true;
      if (ov > s.cbest) { s.cbest = ov; s.cbest_gt = j; }
    }
  }
}

template <bool DENSE, bool MINING>
__global__ void __launch_bounds__(kEncThreads) enc_pass2_kernel(const EncArgs A) {
  __shared__ float4 s_box[DENSE ? 1 : kGtChunk];
  __shared__ float s_area[DENSE ? 1 : kGtChunk];
  __shared__ float s_cm[DENSE ? 1 : kGtChunk];
  __shared__ uint8_t s_claimable[DENSE ? 1 : kGtChunk];
  __shared__ float s_tile[DENSE ? (kEncThreads / 32) * 32 * 33 : 1];

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int a = blockIdx.x * kEncThreads + threadIdx.x;
  const bool valid = a < A.n;
  const ImageGt ig = image_gt<DENSE>(A, b);
  const bool need_haspos = !MINING && !A.gt_max_first;

  RowState s;
  s.best = 0.f; s.best_gt = 0; s.ov0 = 0.f;
  s.claimed = false; s.cbest = 0.f; s.cbest_gt = 0;
  s.owner = -1; s.owner_ov = 0.f;

  AnchorBox ab = {};
  bool active = false;
  WarpBox wb = {};
  if (!DENSE) {
    if (valid) {
      ab = load_anchor(A, a);
      active = (A.mask == nullptr) || (A.mask[a] != 0);
    }
    wb = warp_bbox(active, ab);
  }

  // ---- sweep the GTs: two rounds only when compensation candidates must be pushed
  bool push_round = false;
  bool warp_push = false;
  int match = -1;
  float score = 0.f;
  for (int round = 0; round < 2; ++round) {
    if (DENSE) {
      float* tile = s_tile + warp * 32 * 33;
      const int row0 = blockIdx.x * kEncThreads + warp * 32;
      if (row0 < A.n && (round == 0 || warp_push)) {
        for (int j0 = 0; j0 < ig.m_eff; j0 += 32) {
          const int jc = j0 + lane;
          for (int r = 0; r < 32; ++r) {
            const int row = row0 + r;
            tile[r * 33 + lane] = (row < A.n && jc < ig.m_eff) ? A.overlaps[(int64_t)row * ig.m_eff + jc] : 0.f;
          }
          __syncwarp();
          const int lim = min(32, ig.m_eff - j0);
          for (int jj = 0; jj < lim; ++jj) {
            const int j = j0 + jj;
            const float ov = tile[lane * 33 + jj];
            if (round == 0) {
              const float cm = __uint_as_float(A.colmax[ig.slot0 + j]);
              const bool claimable = !need_haspos || A.haspos[ig.slot0 + j] == 0;
              row_update<MINING>(s, j, ov, cm, claimable);
            } else if (MINING && push_round && ov > A.stop) {
              const int pos = atomicAdd(A.fill + ig.slot0 + j, 1);
              if (pos < kBucketCap) A.bucket[(int64_t)(ig.slot0 + j) * kBucketCap + pos] = HeapItem{ov, a};
            }
          }
          __syncwarp();
        }
      }
    } else {
      for (int c0 = 0; c0 < ig.m_eff; c0 += kGtChunk) {
        const int mc = min(kGtChunk, ig.m_eff - c0);
        __syncthreads();
        for (int k = threadIdx.x; k < mc; k += kEncThreads) {
          const float4 g = gt_box(A, ig, c0 + k);
          s_box[k] = g;
          s_area[k] = box_area(g.x, g.y, g.z, g.w);
          s_cm[k] = __uint_as_float(A.colmax[ig.slot0 + c0 + k]);
          s_claimable[k] = (!need_haspos || A.haspos[ig.slot0 + c0 + k] == 0) ? 1 : 0;
        }
        __syncthreads();
        if (round == 1 && !warp_push) continue;
        for (int k0 = 0; k0 < mc; k0 += 32) {
          const int k = k0 + lane;
          bool test = false;
          if (k < mc) {
            const float cm = s_cm[k];
            const bool wide = MINING ? (cm < FLT_EPSILON) : (cm == 0.f);
            test = may_hit(wb, s_box[k]) || (round == 0 && wide);
          }
          unsigned hits = __ballot_sync(0xffffffffu, test);
          while (hits) {
            const int kk = k0 + __ffs(hits) - 1;
            hits &= hits - 1;
            const float4 g = s_box[kk];
            bool hit = false;
            float ov = 0.f;
            if (active) ov = pair_iou(ab.my0, ab.mx0, ab.my1, ab.mx1, ab.marea, g.x, g.y, g.z, g.w, s_area[kk], hit);
            if (round == 0) {
              row_update<MINING>(s, c0 + kk, ov, s_cm[kk], s_claimable[kk] != 0);
            } else if (MINING && push_round && ov > A.stop) {
              const int pos = atomicAdd(A.fill + ig.slot0 + c0 + kk, 1);
              if (pos < kBucketCap) A.bucket[(int64_t)(ig.slot0 + c0 + kk) * kBucketCap + pos] = HeapItem{ov, a};
            }
          }
        }
      }
    }

    if (round == 1) break;

    // ---- resolve this anchor (end of round 0)
    if (MINING) {
      // stage 1, small_mining_match.cc:85-93
      if (s.best >= A.neg_low && s.best < A.low) match = -1;
      else if (s.best >= A.high) match = s.best_gt;
      else match = -2;
      score = s.best;
      // stage 2, :178-186
      if (s.owner >= 0) { match = s.owner; score = s.owner_ov; }
      if (valid && match >= 0) atomicAdd(A.cnt + ig.slot0 + match, 1);
      push_round = valid && match < 0 && s.best > A.stop;
    } else {
      // anchor_manipulator.py:67-76
      const bool less = s.best < A.low;
      const bool between = (s.best < A.high) && (s.best >= A.low);
      const bool neg = A.ignore_between ? less : between;
      const bool ign = A.ignore_between ? between : less;
      match = s.best_gt;
      if (neg) match = -1;
      if (ign) match = -2;
      score = s.best;
      // :95-104 GT-side claim has priority; argmax over (overlap * claim mask)
      if (s.claimed) {
        if (s.cbest > 0.f) { match = s.cbest_gt; score = s.cbest; }
        else { match = 0; score = s.ov0; }
      }
    }
    // the push round re-stages GT chunks with __syncthreads, so the decision to run
    // it must be CTA-uniform; warps without a pushing lane skip the inner loops
    if (!MINING || !__syncthreads_or(push_round ? 1 : 0)) break;
    warp_push = __any_sync(0xffffffffu, push_round);
  }

  if (!valid) return;
  const int64_t row = (int64_t)b * A.n + a;
  if (DENSE) {
    if (A.match32 != nullptr) A.match32[row] = match;
    if (A.match64 != nullptr) A.match64[row] = match;
    A.scores[row] = score;
    return;
  }
  if (match >= 0) {
    write_positive(A, row, ab, gt_box(A, ig, match), score, match);
  } else {
    A.targets[row] = make_float4(0.f, 0.f, 0.f, 0.f);
    A.labels[row] = (match < -1) ? -1 : 0;   // anchor_manipulator.py:300-302
    A.scores[row] = score;
    if (A.matched != nullptr) A.matched[row] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (A.match32 != nullptr) A.match32[row] = match;
  }
}

// ---------------------------------------------------------------------------
// FUSED path, passes 1 and 2 as warp-autonomous kernels.
//   * a warp owns 32 consecutive anchors and walks kEncImgPerWarp images with them: the anchor loads, the PA
//     transform, the areas and the warp bounding box are paid once per group of images;
//   * no shared memory and no CTA barrier: the ground truth of an image is a few hundred bytes, read through the
//     read-only path (32 boxes per coalesced load for the cull test, broadcast loads for the hits);
//   * per-GT column maxima of pass 1 are reduced in the warp (redux.sync) and the lane that loaded GT k issues one
//     atomicMax for it, so the atomics of a warp go to 32 different addresses.
// The dense-matrix kernels above keep the tiled shared-memory scheme (they have to transpose the matrix).
// ---------------------------------------------------------------------------
// images per warp pass: runtime (DAN_ENC_IMAGES_PER_WARP, default 1: more, smaller CTAs balance better on 148 SMs)

struct WarpAnchors {
  AnchorBox ab;
  WarpBox wb;
  bool valid, active;
  int a;
};

DAN_D WarpAnchors load_warp_anchors(const EncArgs& A) {
  WarpAnchors w;
  // grid = (image groups, anchor chunks): CTAs are dispatched x-fastest, so all images of one anchor chunk start
  // together, and the chunks are walked from the END of the anchor array: the coarse pyramid levels live there, their
  // warps see every GT and run the longest, so they must not be the tail of the launch
  w.a = (gridDim.y - 1 - blockIdx.y) * blockDim.x + threadIdx.x;
  w.valid = w.a < A.n;
  w.ab = AnchorBox{};
  w.active = false;
  if (w.valid) {
    w.ab = load_anchor(A, w.a);
    w.active = (A.mask == nullptr) || (A.mask[w.a] != 0);
  }
  w.wb = warp_bbox(w.active, w.ab);
  return w;
}

template <bool NEED_ROW, bool MINING>
__global__ void __launch_bounds__(kEncThreads, 8) enc_pass1_fused_kernel(const EncArgs A, int batch, int ipw) {
  const int lane = threadIdx.x & 31;
  const WarpAnchors W = load_warp_anchors(A);
  const int b_end = min(batch, (int)(blockIdx.x + 1) * ipw);
  for (int b = blockIdx.x * ipw; b < b_end; ++b) {
    const ImageGt ig = image_gt<false>(A, b);
    float best = 0.f;
    int best_gt = 0;
    for (int k0 = 0; k0 < ig.m_eff; k0 += 32) {
      const int k = k0 + lane;
      const float4 gk = (k < ig.m_eff) ? gt_box(A, ig, k) : make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned hits = __ballot_sync(0xffffffffu, (k < ig.m_eff) && may_hit(W.wb, gk));
      const float gk_area = box_area(gk.x, gk.y, gk.z, gk.w);      // once per GT, broadcast with the box below
      uint32_t my_colmax = 0u;            // column maximum of GT k over this warp's anchors
      while (hits) {
        const int kl = __ffs(hits) - 1;
        hits &= hits - 1;
        // lane kl loaded this GT for the cull test above: broadcast it instead of fetching it again
        const float4 g = make_float4(__shfl_sync(0xffffffffu, gk.x, kl), __shfl_sync(0xffffffffu, gk.y, kl),
                                     __shfl_sync(0xffffffffu, gk.z, kl), __shfl_sync(0xffffffffu, gk.w, kl));
        const float g_area = __shfl_sync(0xffffffffu, gk_area, kl);
        bool hit = false;
        float ov = 0.f;
        if (W.active) ov = pair_iou(W.ab.my0, W.ab.mx0, W.ab.my1, W.ab.mx1, W.ab.marea, g.x, g.y, g.z, g.w, g_area, hit);
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, (ov > 0.f) ? __float_as_uint(ov) : 0u);
        if (lane == kl) my_colmax = wmax;
        if (ov > best) { best = ov; best_gt = k0 + kl; }      // first strictly-greatest GT wins (tf.argmax / :76-83)
        if (MINING && ov > A.stop) {
          // possible stage-3 compensation candidate (small_mining_match.cc:206); whether the anchor is still
          // unmatched is only known after stage 2, pass 3 filters on that
          const int pos = atomicAdd(A.fill + ig.slot0 + k0 + kl, 1);
          if (pos < kBucketCap) A.bucket[(int64_t)(ig.slot0 + k0 + kl) * kBucketCap + pos] = HeapItem{ov, W.a};
        }
      }
      if (my_colmax != 0u) atomicMax(A.colmax + ig.slot0 + k, my_colmax);
    }
    if (W.valid) {
      // the row maximum is final here; pass 2 only needs the overlaps that can tie with a column maximum
      const int64_t row = (int64_t)b * A.n + W.a;
      A.rowbest[row] = best;
      A.rowgt[row] = best_gt;
      if (NEED_ROW) {
        const bool less = best < A.low;
        const bool between = (best < A.high) && (best >= A.low);
        if (!less && !between) A.haspos[ig.slot0 + best_gt] = 1;
      }
    }
  }
}

template <bool MINING>
__global__ void __launch_bounds__(kEncThreads, 8) enc_pass2_fused_kernel(const EncArgs A, int batch, int ipw) {
  const int lane = threadIdx.x & 31;
  // the kernel is a chain of dependent loads (ncu: long_scoreboard): everything that only depends on the thread's
  // coordinates is requested first, so the cached row maximum, the GT offsets and the anchor arrive in ONE round trip
  const int b_first = blockIdx.x * ipw;
  const int a_first = (gridDim.y - 1 - blockIdx.y) * blockDim.x + threadIdx.x;     // = W.a below
  float pre_best = 0.f;
  int pre_gt = 0, pre_off0 = 0, pre_off1 = 0;
  if (b_first < batch) {
    if (a_first < A.n) {
      pre_best = __ldg(A.rowbest + (int64_t)b_first * A.n + a_first);
      pre_gt = __ldg(A.rowgt + (int64_t)b_first * A.n + a_first);
    }
    pre_off0 = __ldg(A.gt_off + b_first);
    pre_off1 = __ldg(A.gt_off + b_first + 1);
  }
  const WarpAnchors W = load_warp_anchors(A);
  const bool need_haspos = !MINING && !A.gt_max_first;
  const int b_end = min(batch, (int)(blockIdx.x + 1) * ipw);
  for (int b = b_first; b < b_end; ++b) {
    ImageGt ig;
    if (b == b_first) {
      ig.off0 = pre_off0;
      ig.m_real = pre_off1 - pre_off0;
      ig.m_eff = ig.m_real > 0 ? ig.m_real : 1;
      ig.slot0 = pre_off0 + b;
    } else {
      ig = image_gt<false>(A, b);
    }
    RowState s;
    s.best = 0.f; s.best_gt = 0; s.ov0 = 0.f;
    s.claimed = false; s.cbest = 0.f; s.cbest_gt = 0;
    s.owner = -1; s.owner_ov = 0.f;
    if (W.valid) {
      if (b == b_first) {
        s.best = pre_best;
        s.best_gt = pre_gt;
      } else {
        const int64_t row = (int64_t)b * A.n + W.a;
        s.best = A.rowbest[row];
        s.best_gt = A.rowgt[row];
      }
    }
    // An anchor can tie with / be claimed by GT k only if its overlap reaches the column maximum cm[k] (to within
    // FLT_EPSILON for the mining matcher).  No overlap of this warp exceeds the largest row maximum of its lanes, so
    // GTs with cm[k] above that bound are skipped without evaluating a single IoU.
    const float wbest = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(s.best)));
    const float reach = MINING ? fadd(wbest, 2.f * FLT_EPSILON) : wbest;
    int match = -1;
    float score = 0.f;
    for (int k0 = 0; k0 < ig.m_eff; k0 += 32) {
      const int k = k0 + lane;
      bool test = false;
      if (k < ig.m_eff) {
        const float cm = __uint_as_float(__ldg(A.colmax + ig.slot0 + k));
        const bool wide = MINING ? (cm < FLT_EPSILON) : (cm == 0.f);
        test = (cm <= reach && may_hit(W.wb, gt_box(A, ig, k))) || wide;
      }
      unsigned hits = __ballot_sync(0xffffffffu, test);
      while (hits) {
        const int kk = k0 + __ffs(hits) - 1;
        hits &= hits - 1;
        const float4 g = gt_box(A, ig, kk);
        bool hit = false;
        float ov = 0.f;
        if (W.active) ov = pair_iou(W.ab.my0, W.ab.mx0, W.ab.my1, W.ab.mx1, W.ab.marea, g.x, g.y, g.z, g.w,
                                    box_area(g.x, g.y, g.z, g.w), hit);
        const float cm = __uint_as_float(__ldg(A.colmax + ig.slot0 + kk));
        if (MINING) {
          // stage 2 tie band, small_mining_match.cc:171,179; ascending GT index => the last GT wins
          if (fabsf(fsub(ov, cm)) < FLT_EPSILON) { s.owner = kk; s.owner_ov = ov; }
        } else {
          // anchor_manipulator.py:88 exact equality with the column maximum
          const bool claimable = !need_haspos || __ldg(A.haspos + ig.slot0 + kk) == 0;
          if (claimable && ov == cm) {
            s.claimed = true;
            if (ov > s.cbest) { s.cbest = ov; s.cbest_gt = kk; }
          }
        }
      }
    }
    if (MINING) {
      // stage 1, small_mining_match.cc:85-93
      if (s.best >= A.neg_low && s.best < A.low) match = -1;
      else if (s.best >= A.high) match = s.best_gt;
      else match = -2;
      score = s.best;
      // stage 2, :178-186
      if (s.owner >= 0) { match = s.owner; score = s.owner_ov; }
      if (W.valid && match >= 0) atomicAdd(A.cnt + ig.slot0 + match, 1);
    } else {
      // anchor_manipulator.py:67-76
      const bool less = s.best < A.low;
      const bool between = (s.best < A.high) && (s.best >= A.low);
      const bool neg = A.ignore_between ? less : between;
      const bool ign = A.ignore_between ? between : less;
      match = s.best_gt;
      if (neg) match = -1;
      if (ign) match = -2;
      score = s.best;
      // :95-104 GT-side claim has priority; argmax over (overlap * claim mask)
      if (s.claimed) {
        if (s.cbest > 0.f) {
          match = s.cbest_gt;
          score = s.cbest;
        } else {
          // claimed only through zero-overlap ties: argmax of an all-zero row is GT 0, the score is its overlap
          const float4 g0 = gt_box(A, ig, 0);
          bool hit0 = false;
          match = 0;
          score = W.active ? pair_iou(W.ab.my0, W.ab.mx0, W.ab.my1, W.ab.mx1, W.ab.marea, g0.x, g0.y, g0.z, g0.w,
                                      box_area(g0.x, g0.y, g0.z, g0.w), hit0) : 0.f;
        }
      }
    }
    if (W.valid) {
      const int64_t row = (int64_t)b * A.n + W.a;
      if (match >= 0) {
        write_positive(A, row, W.ab, gt_box(A, ig, match), score, match);
      } else {
        // streaming stores: the 44 B/anchor of a step are written once and not read again by this kernel
        __stcs(A.targets + row, make_float4(0.f, 0.f, 0.f, 0.f));
        __stcs(reinterpret_cast<long long*>(A.labels) + row, (long long)((match < -1) ? -1 : 0));   // anchor_manipulator.py:300-302
        __stcs(A.scores + row, score);
        if (A.matched != nullptr) __stcs(A.matched + row, make_float4(0.f, 0.f, 0.f, 0.f));
        if (A.match32 != nullptr) __stcs(A.match32 + row, match);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// pass 3: stage 3 "hard face compensation", small_mining_match.cc:199-222.
// One warp per image, GTs in ascending order.
// ---------------------------------------------------------------------------
template <bool DENSE>
DAN_D bool still_unmatched(const EncArgs& A, int64_t row) {
  if (DENSE) return reinterpret_cast<const volatile int32_t*>(A.match32)[row] < 0;
  return reinterpret_cast<const volatile int64_t*>(A.labels)[row] < 1;
}

template <bool DENSE>
DAN_D void apply_compensation(const EncArgs& A, const ImageGt& ig, int b, int a, int j, float ov) {
  const int64_t row = (int64_t)b * A.n + a;
  if (DENSE) {
    A.match32[row] = j;
    A.scores[row] = ov;
  } else {
    write_positive(A, row, load_anchor(A, a), gt_box(A, ig, j), ov, j);
  }
}

constexpr int kP3Threads = 128;
constexpr int kP3Group = 16;        // buckets staged in shared memory at a time
constexpr int kHashSlots = 4096;    // anchors taken by stage 3 of this image (open addressing)
constexpr int kApplyCap = 1024;     // deferred output patches

struct TakenSet {
  int* slots;      // [kHashSlots], -1 = empty
  int* count;
};

// Output patches of stage 3 are deferred: the serial walk over the GTs only needs the taken SET (shared memory),
// so the encode + global stores of each patched anchor run afterwards, in parallel over the whole CTA.
struct ApplyList {
  int* a;
  int* j;
  float* ov;
  int* n;
};

DAN_D bool taken_has(const TakenSet& t, int a) {
  unsigned h = ((unsigned)a * 2654435761u) & (kHashSlots - 1);
  while (true) {
    const int v = t.slots[h];
    if (v == a) return true;
    if (v < 0) return false;
    h = (h + 1) & (kHashSlots - 1);
  }
}

DAN_D void taken_add(const TakenSet& t, int a) {
  unsigned h = ((unsigned)a * 2654435761u) & (kHashSlots - 1);
  while (true) {
    const int old = atomicCAS(&t.slots[h], -1, a);
    if (old < 0 || old == a) break;
    h = (h + 1) & (kHashSlots - 1);
  }
  atomicAdd(t.count, 1);
}

// Select the `need` best of the c entries of `list` (key <= 0 entries are dead) for GT j and apply them.
// Pop order of a max-heap == descending key; only a tie that straddles the cut depends on libstdc++'s heap
// layout, which is then reproduced exactly (heap_order.cuh).  Executed by one full warp.
template <bool DENSE>
DAN_D void flush_applies(const EncArgs& A, const ImageGt& ig, int b, const ApplyList& al, int first, int step) {
  const int n = min(*al.n, kApplyCap);
  for (int e = first; e < n; e += step) apply_compensation<DENSE>(A, ig, b, al.a[e], al.j[e], al.ov[e]);
}

template <bool DENSE>
DAN_D void compensate_from_list(const EncArgs& A, const ImageGt& ig, int b, int j, int need, HeapItem* list, int c,
                                bool ordered, HeapItem* sort_buf, HeapItem* heap, const TakenSet& taken, bool use_hash,
                                const ApplyList& al, bool defer) {
  const int lane = threadIdx.x & 31;
  int live = 0;
  for (int e = lane; e < c; e += 32) live += (list[e].key > 0.f) ? 1 : 0;
  live = __reduce_add_sync(0xffffffffu, live);
  auto apply = [&](int a, float ov) {
    if (use_hash) taken_add(taken, a);
    if (defer) {
      const int slot = atomicAdd(al.n, 1);
      if (slot < kApplyCap) {
        al.a[slot] = a;
        al.j[slot] = j;
        al.ov[slot] = ov;
        return;
      }
    }
    apply_compensation<DENSE>(A, ig, b, a, j, ov);
  };
  if (live <= need) {
    for (int e = lane; e < c; e += 32)
      if (list[e].key > 0.f) apply(list[e].id, list[e].key);
    return;
  }
  int got = 0;
  bool straddle = false;
  if (c <= 64) {
    // one ranking pass: entry e is popped before the cut iff its whole tie group fits (ge <= need), after the
    // cut iff g >= need; a tie group with g < need < ge straddles the cut
    const float k0 = (lane < c) ? list[lane].key : 0.f;
    const float k1 = (lane + 32 < c) ? list[lane + 32].key : 0.f;
    int g0 = 0, ge0 = 0, g1 = 0, ge1 = 0;
    for (int e = 0; e < c; ++e) {
      const float v = list[e].key;
      if (v > 0.f) {
        g0 += (v > k0) ? 1 : 0;
        ge0 += (v >= k0) ? 1 : 0;
        g1 += (v > k1) ? 1 : 0;
        ge1 += (v >= k1) ? 1 : 0;
      }
    }
    const bool st = (k0 > 0.f && g0 < need && ge0 > need) || (k1 > 0.f && g1 < need && ge1 > need);
    straddle = __any_sync(0xffffffffu, st);
    if (!straddle) {
      if (k0 > 0.f && ge0 <= need) apply(list[lane].id, k0);
      if (k1 > 0.f && ge1 <= need) apply(list[lane + 32].id, k1);
    }
    got = need;
  }
  while (got < need) {
    float lmax = 0.f;
    for (int e = lane; e < c; e += 32) lmax = fmaxf(lmax, list[e].key);
    const uint32_t wbits = __reduce_max_sync(0xffffffffu, __float_as_uint(lmax));
    if (wbits == 0u) break;
    const float wmax = __uint_as_float(wbits);
    int eq = 0;
    for (int e = lane; e < c; e += 32) eq += (list[e].key == wmax) ? 1 : 0;
    const int total = __reduce_add_sync(0xffffffffu, eq);
    if (got + total > need) { straddle = true; break; }
    for (int e = lane; e < c; e += 32) {
      if (list[e].key == wmax) {
        apply(list[e].id, wmax);
        list[e].key = -wmax;     // taken; the value is kept for the exact path
      }
    }
    got += total;
    __syncwarp();
  }
  if (straddle) {
    __syncwarp();
    // The exact path: every live or already taken entry was a member of the reference's heap, pushed in ascending
    // anchor order (small_mining_match.cc:204-209).  Bucket lists (c <= 64) are unordered: each lane ranks its two
    // entries by anchor index in parallel; spill lists are already ordered and are compacted by lane 0.
    int n = 0;
    if (c <= 64 && !ordered) {
      const int e0 = lane, e1 = lane + 32;
      const bool l0 = e0 < c && list[e0].key != 0.f, l1 = e1 < c && list[e1].key != 0.f;
      const int id0 = l0 ? list[e0].id : 0x7fffffff, id1 = l1 ? list[e1].id : 0x7fffffff;
      int r0 = 0, r1 = 0;
      for (int f = 0; f < c; ++f) {
        const bool lf = list[f].key != 0.f;
        const int idf = list[f].id;
        r0 += (lf && idf < id0) ? 1 : 0;
        r1 += (lf && idf < id1) ? 1 : 0;
      }
      if (l0) sort_buf[r0] = HeapItem{fabsf(list[e0].key), id0};
      if (l1) sort_buf[r1] = HeapItem{fabsf(list[e1].key), id1};
      n = __popc(__ballot_sync(0xffffffffu, l0)) + __popc(__ballot_sync(0xffffffffu, l1));
      __syncwarp();
    } else if (lane == 0) {
      for (int e = 0; e < c; ++e) {
        if (list[e].key == 0.f) continue;
        HeapItem v = list[e];
        v.key = fabsf(v.key);
        if (ordered) { sort_buf[n++] = v; continue; }
        int p = n - 1;
        while (p >= 0 && sort_buf[p].id > v.id) { sort_buf[p + 1] = sort_buf[p]; --p; }
        sort_buf[p + 1] = v;
        ++n;
      }
    }
    if (lane == 0) {
      int len = 0;
      for (int e = 0; e < n; ++e) heap_push(heap, len, sort_buf[e]);
      for (int p = 0; p < need && len > 0; ++p) {
        const HeapItem it = heap_pop(heap, len);
        apply(it.id, it.key);     // re-applying an entry taken above is idempotent
      }
    }
  }
  __syncwarp();
}

template <bool DENSE>
__global__ void __launch_bounds__(kP3Threads) enc_pass3_kernel(const EncArgs A) {
  __shared__ HeapItem s_group[kP3Group][kBucketCap];
  __shared__ HeapItem s_sort[kBucketCap];
  __shared__ HeapItem s_heap[kBucketCap];
  __shared__ int s_hash[kHashSlots];
  __shared__ int s_taken_n;
  __shared__ int s_apply_a[kApplyCap], s_apply_j[kApplyCap];
  __shared__ float s_apply_ov[kApplyCap];
  __shared__ int s_apply_n;
  __shared__ int s_needy_j[kP3Threads], s_needy_need[kP3Threads], s_needy_fill[kP3Threads];
  __shared__ int s_warp_cnt[kP3Threads / 32];

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const ImageGt ig = image_gt<DENSE>(A, b);
  TakenSet taken{s_hash, &s_taken_n};
  ApplyList al{s_apply_a, s_apply_j, s_apply_ov, &s_apply_n};

  for (int i = tid; i < kHashSlots; i += kP3Threads) s_hash[i] = -1;
  if (tid == 0) { s_taken_n = 0; s_apply_n = 0; }
  __syncthreads();

  for (int j0 = 0; j0 < ig.m_eff; j0 += kP3Threads) {
    // ---- GTs of this range that are short of min_match, in ascending order
    const int j = j0 + tid;
    int need = 0, fill = 0;
    if (j < ig.m_eff) {
      need = A.min_match - A.cnt[ig.slot0 + j];
      fill = A.fill[ig.slot0 + j];
    }
    const unsigned m = __ballot_sync(0xffffffffu, need > 0);
    if (lane == 0) s_warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, nneedy = 0;
    for (int w = 0; w < kP3Threads / 32; ++w) {
      if (w < warp) before += s_warp_cnt[w];
      nneedy += s_warp_cnt[w];
    }
    if (need > 0) {
      const int pos = before + __popc(m & lt_mask);
      s_needy_j[pos] = j;
      s_needy_need[pos] = need;
      s_needy_fill[pos] = fill;
    }
    __syncthreads();

    for (int g0 = 0; g0 < nneedy; g0 += kP3Group) {
      const int ng = min(kP3Group, nneedy - g0);
      // ---- stage the group's candidate buckets (all threads, coalesced)
      for (int e = tid; e < ng * kBucketCap; e += kP3Threads) {
        const int g = e / kBucketCap, k = e % kBucketCap;
        const int f = s_needy_fill[g0 + g];
        if (f <= kBucketCap && k < f) {
          HeapItem it = A.bucket[(int64_t)(ig.slot0 + s_needy_j[g0 + g]) * kBucketCap + k];
          // buckets hold every anchor with overlap > stop; only those still unmatched after stage 2 (and after the
          // patches of the previous groups, already in HBM) are candidates
          if (!still_unmatched<DENSE>(A, (int64_t)b * A.n + it.id)) it.key = 0.f;
          s_group[g][k] = it;
        }
      }
      __syncthreads();
      // ---- the stage is order dependent: warp 0 walks the GTs in ascending order
      if (warp == 0) {
        for (int g = 0; g < ng; ++g) {
          const int gj = s_needy_j[g0 + g];
          const int gneed = s_needy_need[g0 + g];
          const int gfill = s_needy_fill[g0 + g];
          const bool use_hash = s_taken_n < kHashSlots / 2;
          if (!use_hash || gfill > kBucketCap) {
            // this GT reads the labels in HBM as the truth: write the pending patches first
            flush_applies<DENSE>(A, ig, b, al, lane, 32);
            __syncwarp();
            if (lane == 0) s_apply_n = 0;
            __threadfence_block();
            __syncwarp();
          }
          if (gfill <= kBucketCap) {
            HeapItem* list = s_group[g];
            for (int e = lane; e < gfill; e += 32) {
              const int a = list[e].id;
              const bool dead = use_hash ? taken_has(taken, a) : !still_unmatched<DENSE>(A, (int64_t)b * A.n + a);
              if (dead) list[e].key = 0.f;
            }
            __syncwarp();
            // patches can be deferred as long as the taken set is tracked in shared memory
            compensate_from_list<DENSE>(A, ig, b, gj, gneed, list, gfill, false, s_sort, s_heap, taken, use_hash, al, use_hash);
          } else {
            // bucket overflowed: rescan every anchor of the image in index order (patches are immediate here)
            HeapItem* list = A.spill + (int64_t)b * 3 * A.n;
            int c = 0;
            const float4 gb = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : gt_box(A, ig, gj);
            const float garea = box_area(gb.x, gb.y, gb.z, gb.w);
            for (int a0 = 0; a0 < A.n; a0 += 32) {
              const int a = a0 + lane;
              bool ok = false;
              float ov = 0.f;
              if (a < A.n) {
                if (DENSE) {
                  ov = A.overlaps[(int64_t)a * ig.m_eff + gj];
                } else if (A.mask == nullptr || A.mask[a] != 0) {
                  const AnchorBox ab = load_anchor(A, a);
                  bool hit;
                  ov = pair_iou(ab.my0, ab.mx0, ab.my1, ab.mx1, ab.marea, gb.x, gb.y, gb.z, gb.w, garea, hit);
                }
                ok = (ov > A.stop) && still_unmatched<DENSE>(A, (int64_t)b * A.n + a);
              }
              const unsigned mm = __ballot_sync(0xffffffffu, ok);
              if (ok) list[c + __popc(mm & lt_mask)] = HeapItem{ov, a};
              c += __popc(mm);
            }
            __threadfence_block();
            __syncwarp();
            compensate_from_list<DENSE>(A, ig, b, gj, gneed, list, c, true, list + A.n, list + 2 * (int64_t)A.n, taken, use_hash,
                                        al, false);
          }
          __threadfence_block();
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- deferred output patches of the group, all threads
      flush_applies<DENSE>(A, ig, b, al, tid, kP3Threads);
      __syncthreads();
      if (tid == 0) s_apply_n = 0;
    }
  }
}

// M == 0 on the dense mining op: stage 1 leaves every anchor at (-2, lowest())
__global__ void fill_empty_match_kernel(int32_t* match, float* scores, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    match[i] = -2;
    scores[i] = -FLT_MAX;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct WsLayout {
  size_t colmax, cnt, haspos, fill, zero_bytes, bucket, spill, rowbest, rowgt, total;
};

static WsLayout ws_layout(int64_t n, int64_t batch, int64_t slots) {
  WsLayout w;
  size_t off = 0;
  w.colmax = off; off += align_up(slots * 4, 256);
  w.cnt = off;    off += align_up(slots * 4, 256);
  w.haspos = off; off += align_up(slots * 4, 256);
  w.fill = off;   off += align_up(slots * 4, 256);
  w.zero_bytes = off;
  w.bucket = off; off += align_up(slots * kBucketCap * sizeof(HeapItem), 256);
  w.spill = off;  off += align_up(batch * 3 * n * sizeof(HeapItem), 256);
  w.rowbest = off; off += align_up(batch * n * 4, 256);
  w.rowgt = off;   off += align_up(batch * n * 4, 256);
  w.total = off;
  return w;
}

static int check_mining_attrs(float neg_low, float neg_high, float pos, int min_match, float stop) {
  // small_mining_match.cc:292-305
  DAN_REQUIRE(neg_low >= 0.f && neg_low < 1.f, DAN_ERR_INVALID_ARGUMENT, "Need Attr 1 > negative_low_thres >= 0, got %g", neg_low);
  DAN_REQUIRE(neg_high > neg_low && neg_high < 1.f, DAN_ERR_INVALID_ARGUMENT,
              "Need Attr 1 > negative_high_thres > negative_low_thres, got %g", neg_high);
  DAN_REQUIRE(pos >= neg_high && pos < 1.f, DAN_ERR_INVALID_ARGUMENT, "Need Attr 1 > positive_thres >= negative_high_thres, got %g", pos);
  DAN_REQUIRE(stop >= 0.f && stop < 1.f, DAN_ERR_INVALID_ARGUMENT, "Need Attr 1 > stop_positive_thres >= 0., got %g", stop);
  DAN_REQUIRE(min_match >= 1, DAN_ERR_INVALID_ARGUMENT, "Need Attr min_match >= 1, got %d", min_match);
  return DAN_OK;
}

static void bind_workspace(EncArgs& A, void* workspace, const WsLayout& w) {
  char* base = static_cast<char*>(workspace);
  A.colmax = reinterpret_cast<uint32_t*>(base + w.colmax);
  A.cnt = reinterpret_cast<int32_t*>(base + w.cnt);
  A.haspos = reinterpret_cast<int32_t*>(base + w.haspos);
  A.fill = reinterpret_cast<int32_t*>(base + w.fill);
  A.bucket = reinterpret_cast<HeapItem*>(base + w.bucket);
  A.spill = reinterpret_cast<HeapItem*>(base + w.spill);
  A.rowbest = reinterpret_cast<float*>(base + w.rowbest);
  A.rowgt = reinterpret_cast<int32_t*>(base + w.rowgt);
}

// ev (optional, 4 events): recorded before pass 1 and after each pass, for the profile entry point
template <bool DENSE>
static int run_passes(const EncArgs& A, bool mining, bool need_row, int batch, cudaStream_t st, cudaEvent_t* ev = nullptr) {
  const dim3 grid((A.n + kEncThreads - 1) / kEncThreads, batch);
  // sparse path: 128-thread CTAs, one image per CTA (sweeps on B200: 64/256 threads and 2-8 images per warp are
  // within 2 % of this)
  const int ipw = 1;
  const int fthreads = 128;
  const dim3 fgrid((batch + ipw - 1) / ipw, (A.n + fthreads - 1) / fthreads);
  if (ev) DAN_CUDA(cudaEventRecord(ev[0], st));
  if (DENSE) {
    if (need_row) enc_pass1_kernel<true, true><<<grid, kEncThreads, 0, st>>>(A);
    else enc_pass1_kernel<true, false><<<grid, kEncThreads, 0, st>>>(A);
  } else {
    if (mining) enc_pass1_fused_kernel<false, true><<<fgrid, fthreads, 0, st>>>(A, batch, ipw);
    else if (need_row) enc_pass1_fused_kernel<true, false><<<fgrid, fthreads, 0, st>>>(A, batch, ipw);
    else enc_pass1_fused_kernel<false, false><<<fgrid, fthreads, 0, st>>>(A, batch, ipw);
  }
  DAN_LAUNCH_CHECK("enc_pass1_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[1], st));
  if (DENSE) {
    if (mining) enc_pass2_kernel<true, true><<<grid, kEncThreads, 0, st>>>(A);
    else enc_pass2_kernel<true, false><<<grid, kEncThreads, 0, st>>>(A);
  } else {
    if (mining) enc_pass2_fused_kernel<true><<<fgrid, fthreads, 0, st>>>(A, batch, ipw);
    else enc_pass2_fused_kernel<false><<<fgrid, fthreads, 0, st>>>(A, batch, ipw);
  }
  DAN_LAUNCH_CHECK("enc_pass2_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[2], st));
  if (mining) {
    enc_pass3_kernel<DENSE><<<batch, kP3Threads, 0, st>>>(A);
    DAN_LAUNCH_CHECK("enc_pass3_kernel");
  }
  if (ev) DAN_CUDA(cudaEventRecord(ev[3], st));
  return DAN_OK;
}

// CUDA-event timing of a launch sequence (profile entry points only): creates n events, runs fn(ev), synchronises
// on the last one and writes the n-1 intervals in milliseconds.
template <typename Fn>
static int timed_sequence(int n_events, float* h_ms, Fn fn) {
  cudaEvent_t ev[8];
  for (int i = 0; i < n_events; ++i) DAN_CUDA(cudaEventCreate(&ev[i]));
  int rc = fn(ev);
  if (rc == DAN_OK) {
    cudaError_t e = cudaEventSynchronize(ev[n_events - 1]);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize");
  }
  if (rc == DAN_OK)
    for (int i = 0; i + 1 < n_events; ++i) cudaEventElapsedTime(&h_ms[i], ev[i], ev[i + 1]);
  for (int i = 0; i < n_events; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

}  // namespace dan

using namespace dan;

extern "C" {

size_t dan_match_workspace_bytes(int32_t num_anchors, int32_t num_gt) {
  if (num_anchors < 0 || num_gt < 0) return 0;
  return ws_layout(num_anchors, 1, (int64_t)num_gt + 1).total;
}

size_t dan_encode_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t total_gt) {
  if (num_anchors < 0 || batch < 0 || total_gt < 0) return 0;
  return ws_layout(num_anchors, batch, (int64_t)total_gt + batch).total;
}

int dan_small_mining_match(const float* overlaps, int32_t num_anchors, int32_t num_gt, float negative_low_thres,
                           float negative_high_thres, float positive_thres, int32_t min_match, float stop_positive_thres,
                           int32_t* out_match_indices, float* out_match_scores, void* workspace, size_t workspace_bytes,
                           void* stream) {
  int rc = check_mining_attrs(negative_low_thres, negative_high_thres, positive_thres, min_match, stop_positive_thres);
  if (rc != DAN_OK) return rc;
  DAN_REQUIRE(num_anchors >= 0 && num_gt >= 0, DAN_ERR_INVALID_ARGUMENT, "inputs must be in 'num_anchors x num_ground_truth' format.");
  if (num_anchors == 0) return DAN_OK;
  DAN_REQUIRE(out_match_indices && out_match_scores, DAN_ERR_INVALID_ARGUMENT, "NULL output");
  cudaStream_t st = (cudaStream_t)stream;
  if (num_gt == 0) {
    fill_empty_match_kernel<<<(num_anchors + 255) / 256, 256, 0, st>>>(out_match_indices, out_match_scores, num_anchors);
    DAN_LAUNCH_CHECK("fill_empty_match_kernel");
    return DAN_OK;
  }
  DAN_REQUIRE(overlaps != nullptr, DAN_ERR_INVALID_ARGUMENT, "NULL overlaps");
  const WsLayout w = ws_layout(num_anchors, 1, (int64_t)num_gt + 1);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
              w.total, workspace_bytes);
  EncArgs A = {};
  A.n = num_anchors;
  A.overlaps = overlaps;
  A.m_dense = num_gt;
  A.low = negative_high_thres;
  A.high = positive_thres;
  A.neg_low = negative_low_thres;
  A.stop = stop_positive_thres;
  A.min_match = min_match;
  A.gt_max_first = 1;
  A.ignore_between = 1;
  A.match32 = out_match_indices;
  A.scores = out_match_scores;
  bind_workspace(A, workspace, w);
  DAN_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));
  return run_passes<true>(A, true, false, 1, st);
}

int dan_dual_max_match(const float* overlaps, int32_t num_anchors, int32_t num_gt, float low_thres, float high_thres,
                       int32_t ignore_between, int32_t gt_max_first, int64_t* out_match_indices, float* out_match_scores,
                       void* workspace, size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(num_anchors >= 0 && num_gt >= 1, DAN_ERR_INVALID_ARGUMENT,
              "do_dual_max_match needs overlap_matrix [num_anchors, num_gt>=1] (tf.argmax over an empty axis is an error)");
  if (num_anchors == 0) return DAN_OK;
  DAN_REQUIRE(overlaps && out_match_indices && out_match_scores, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  const WsLayout w = ws_layout(num_anchors, 1, (int64_t)num_gt + 1);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
              w.total, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  EncArgs A = {};
  A.n = num_anchors;
  A.overlaps = overlaps;
  A.m_dense = num_gt;
  A.low = low_thres;
  A.high = high_thres;
  A.ignore_between = ignore_between ? 1 : 0;
  A.gt_max_first = gt_max_first ? 1 : 0;
  A.match64 = out_match_indices;
  A.scores = out_match_scores;
  bind_workspace(A, workspace, w);
  DAN_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));
  return run_passes<true>(A, false, !gt_max_first, 1, st);
}

static int encode_core(const dan_encode_params* p, const float* a_ymin, const float* a_xmin, const float* a_ymax,
                       const float* a_xmax, const uint8_t* inside_mask, int32_t num_anchors, const float* gt_boxes,
                       const int32_t* gt_offsets, int32_t batch, int32_t total_gt, float* out_targets, int64_t* out_labels,
                       float* out_scores, float* out_matched_gt, int32_t* out_match, void* workspace, size_t workspace_bytes,
                       void* stream, cudaEvent_t* ev) {
  DAN_REQUIRE(p != nullptr, DAN_ERR_INVALID_ARGUMENT, "params is NULL");
  DAN_REQUIRE(p->matcher == DAN_MATCH_DUAL || p->matcher == DAN_MATCH_MINING, DAN_ERR_INVALID_ARGUMENT, "unknown matcher %d", p->matcher);
  DAN_REQUIRE(num_anchors >= 0 && batch >= 0 && total_gt >= 0, DAN_ERR_INVALID_ARGUMENT, "negative size");
  DAN_REQUIRE(batch <= 65535, DAN_ERR_UNSUPPORTED, "batch > 65535");
  DAN_REQUIRE(num_anchors <= 4000000, DAN_ERR_UNSUPPORTED, "more than 4,000,000 anchors per image");
  const bool mining = p->matcher == DAN_MATCH_MINING;
  if (mining) {
    int rc = check_mining_attrs(p->negative_low_thres, p->ignore_threshold, p->positive_threshold, p->min_match, p->stop_positive_thres);
    if (rc != DAN_OK) return rc;
  }
  DAN_REQUIRE(p->pa_scale >= 0.f, DAN_ERR_INVALID_ARGUMENT, "pa_scale must be >= 0");
  if (num_anchors == 0 || batch == 0) return DAN_OK;
  DAN_REQUIRE(a_ymin && a_xmin && a_ymax && a_xmax && gt_offsets && out_targets && out_labels && out_scores, DAN_ERR_INVALID_ARGUMENT,
              "NULL pointer");
  DAN_REQUIRE(total_gt == 0 || gt_boxes != nullptr, DAN_ERR_INVALID_ARGUMENT, "gt_boxes is NULL");
  DAN_REQUIRE(aligned16(gt_boxes) && aligned16(out_targets) && aligned16(out_matched_gt), DAN_ERR_INVALID_ARGUMENT,
              "gt_boxes / out_targets / out_matched_gt must be 16-byte aligned");
  const WsLayout w = ws_layout(num_anchors, batch, (int64_t)total_gt + batch);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
              w.total, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  EncArgs A = {};
  A.ay0 = a_ymin; A.ax0 = a_xmin; A.ay1 = a_ymax; A.ax1 = a_xmax;
  A.mask = inside_mask;
  A.n = num_anchors;
  A.gt = reinterpret_cast<const float4*>(gt_boxes);
  A.gt_off = gt_offsets;
  A.low = p->ignore_threshold;
  A.high = p->positive_threshold;
  A.neg_low = p->negative_low_thres;
  A.stop = p->stop_positive_thres;
  A.min_match = p->min_match;
  A.ignore_between = mining ? 1 : (p->ignore_between ? 1 : 0);
  A.gt_max_first = mining ? 1 : (p->gt_max_first ? 1 : 0);
  A.ps0 = p->prior_scaling[0]; A.ps1 = p->prior_scaling[1]; A.ps2 = p->prior_scaling[2]; A.ps3 = p->prior_scaling[3];
  A.pa_scale = p->pa_scale;
  A.debug = p->debug;
  A.targets = reinterpret_cast<float4*>(out_targets);
  A.labels = out_labels;
  A.scores = out_scores;
  A.matched = reinterpret_cast<float4*>(out_matched_gt);
  A.match32 = out_match;
  bind_workspace(A, workspace, w);
  DAN_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));
  return run_passes<false>(A, mining, !mining && !A.gt_max_first, batch, st, ev);
}

int dan_encode_batch(const dan_encode_params* p, const float* a_ymin, const float* a_xmin, const float* a_ymax,
                     const float* a_xmax, const uint8_t* inside_mask, int32_t num_anchors, const float* gt_boxes,
                     const int32_t* gt_offsets, int32_t batch, int32_t total_gt, float* out_targets, int64_t* out_labels,
                     float* out_scores, float* out_matched_gt, int32_t* out_match, void* workspace, size_t workspace_bytes,
                     void* stream) {
  return encode_core(p, a_ymin, a_xmin, a_ymax, a_xmax, inside_mask, num_anchors, gt_boxes, gt_offsets, batch, total_gt, out_targets,
                     out_labels, out_scores, out_matched_gt, out_match, workspace, workspace_bytes, stream, nullptr);
}

int dan_encode_batch_profile(const dan_encode_params* p, const float* a_ymin, const float* a_xmin, const float* a_ymax,
                             const float* a_xmax, const uint8_t* inside_mask, int32_t num_anchors, const float* gt_boxes,
                             const int32_t* gt_offsets, int32_t batch, int32_t total_gt, float* out_targets, int64_t* out_labels,
                             float* out_scores, float* out_matched_gt, int32_t* out_match, void* workspace, size_t workspace_bytes,
                             void* stream, float* h_pass_ms) {
  DAN_REQUIRE(h_pass_ms != nullptr, DAN_ERR_INVALID_ARGUMENT, "h_pass_ms is NULL");
  h_pass_ms[0] = h_pass_ms[1] = h_pass_ms[2] = 0.f;
  return timed_sequence(4, h_pass_ms, [&](cudaEvent_t* ev) {
    return encode_core(p, a_ymin, a_xmin, a_ymax, a_xmax, inside_mask, num_anchors, gt_boxes, gt_offsets, batch, total_gt, out_targets,
                       out_labels, out_scores, out_matched_gt, out_match, workspace, workspace_bytes, stream, ev);
  });
}

}  // extern "C"
