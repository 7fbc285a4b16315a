// K3-K5: the evaluation half of the path -- utility/bbox_util.py:103-119
// parse_by_class and its pieces, batched over images and classes.
//
//   pp_filter_kernel   (more than two classes; for two classes the NMS kernel filters its own image)
//                      softmax -> select(threshold) -> decode -> clip -> min-size;
//                      survivors are COMPACTED as 64-bit keys
//                      (ordered score bits << 32 | ~anchor index).  The reference
//                      instead multiplies by 0/1 masks and keeps all N rows
//                      (bbox_util.py:24-59); rows it zeroes can never precede a
//                      positive score in tf.nn.top_k, so dropping them is exact.
//   topk_sort_kernel   sort_bboxes alone (bbox_util.py:61-72): block radix select of the keep_topk largest keys
//                      when more survive, then the bitonic sort of sort.cuh.  Keys are unique, descending key
//                      order == descending score with ties broken by LOWER index, which is tf.nn.top_k's order.
//   nms_greedy_kernel  per (image, class): the same top-k + sort, decode of the survivors, and
//                      tf.image.non_max_suppression (no +1, corners min/max-normalised, strict >, area <= 0 never
//                      suppresses) as a chunked greedy sweep against the kept list, then the zero padded outputs
//                      (bbox_util.py:80-90).  One launch, one CTA per list.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dan {

// phase timestamps (SM cycles) of the per-list kernels, written by thread 0 of the CTA of the longest list when the
// library is built with -DDAN_PHASE_TIMING (tools/phase_timing.py); otherwise the macro is empty
#ifdef DAN_PHASE_TIMING
__device__ long long g_phase[32];
#define DAN_PHASE(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_phase[slot] = clock64(); } while (0)
#else
#define DAN_PHASE(slot) do { } while (0)
#endif

#ifndef DAN_PP_THREADS
#define DAN_PP_THREADS 1024           // threads of the per-list kernels (512 was measured: +20 % latency, no gain from sharing the SM
#endif                                // with the encode CTAs)
#define DAN_SORT_THREADS DAN_PP_THREADS
#include "sort.cuh"

struct PpArgs {
  // inputs
  const float* cls;       // [B, N, C] logits
  const float4* loc;      // [B, N, 4] offsets (or NULL)
  const float4* boxes;    // [B, N, 4] decoded boxes (or NULL)
  const float* ay0;
  const float* ax0;
  const float* ay1;
  const float* ax1;
  int n, batch, num_classes;
  float img_h, img_w;
  float select_thr, min_size_p1;
  float reject_below;         // 2 classes: logit difference below which softmax <= threshold for sure (-inf: off)
  float ps0, ps1, ps2, ps3;
  int keep_topk, nms_topk;
  int nms_cap;                // kept-list capacity = min(nms_topk, keep_topk)
  float nms_thr;
  // shared-memory plan of the NMS kernel (greedy_plan)
  int sort_bytes, kcap, wcap, cand_global, kept_global;
  int kstride;                // per-list stride of the candidate arrays = min(keep_topk, n)
  int dyn_smem_bytes;         // dynamic shared memory of the launch
  // workspace
  unsigned long long* keys;   // [L, n]  L = batch * (C-1) lists
  int32_t* key_count;         // [L]
  int32_t* maybe;             // [L, n] anchors that pass the quick reject (fused filter; beyond shared memory)
  float4* stash;              // [B, n] decoded + clipped box of every anchor that passed the filter, or NULL.  Used when
                              // loc / boxes live in mapped HOST memory: a row then crosses PCIe once, not twice
  float* s_scores;            // sort_bboxes outputs [keep_topk]
  float4* s_boxes;
  int32_t* s_index;
  unsigned long long* s_key;  // [L, kstride] sorted keys
  unsigned long long* s_key2; // [L, kstride] merge buffer of lists longer than the in-shared-memory sort
  float4* s_box;              // [L, keep_topk] normalised corners (y0, x0, y1, x1), rank order (lists too long for shared memory)
  float* s_area;              // [L, keep_topk] (y1-y0)*(x1-x0)
  float4* g_kept_box;         // [L, keep_topk] kept list of the NMS when it does not fit shared memory
  float* g_kept_area;
  int32_t* g_kept_rank;
  uint2* s_mask;              // [L, keep_topk] stripe masks of the candidates
  uint32_t* g_stripes;        // [L, 64, ceil(keep_topk / 32)] stripe bitsets of the kept list
  // outputs
  float4* out_boxes;
  float* out_scores;
  int32_t* out_counts;
  int32_t* out_index;
  int32_t* out_keep;
  int filler;                 // fused parse_by_class: zero-score rows fill up the NMS selection
  // peer exchange (dan_postprocess_batch_peers): every output row of the slab is also stored into the receive buffers of
  // the other ranks over NVLink (mapped peer memory); the last CTA then raises this rank's flag in every peer
  int npeers;                                   // destinations besides out_* (0: no exchange)
  long long peer_delta[DAN_MAX_PEERS];          // byte offset from an address inside the own slab to its copy in destination q
  int32_t* peer_flag[DAN_MAX_PEERS];            // this rank's arrival flag inside destination q
  int32_t* peer_state;                          // [2] device words of this rank: CTAs done, step sequence number
};

// clip_bboxes, bbox_util.py:38-48
DAN_D float4 pp_clip(float4 b, float height, float width) {
  float ymin = fmaxf(b.x, 0.f);
  float xmin = fmaxf(b.y, 0.f);
  const float ymax = fminf(b.z, fsub(height, 1.f));
  const float xmax = fminf(b.w, fsub(width, 1.f));
  ymin = fminf(ymin, ymax);
  xmin = fminf(xmin, xmax);
  return make_float4(ymin, xmin, ymax, xmax);
}

// decode (when offsets are given) + clip for anchor `a` of image `b`, split into the global loads and the
// arithmetic so that a caller can put independent work between the two
struct RawBox {
  float4 v;        // offsets or decoded box
  float4 anchor;   // ymin, xmin, ymax, xmax (only when decoding)
};

DAN_D RawBox pp_load(const PpArgs& A, int b, int a) {
  const int64_t row = (int64_t)b * A.n + a;
  RawBox r;
  if (A.loc != nullptr) {
    r.v = A.loc[row];
    r.anchor = make_float4(A.ay0[a], A.ax0[a], A.ay1[a], A.ax1[a]);
  } else {
    r.v = A.boxes[row];
    r.anchor = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}

DAN_D float4 pp_finish(const PpArgs& A, const RawBox& r) {
  float4 bx = r.v;
  if (A.loc != nullptr) bx = decode_box(r.v, r.anchor.x, r.anchor.y, r.anchor.z, r.anchor.w, A.ps0, A.ps1, A.ps2, A.ps3);
  return pp_clip(bx, A.img_h, A.img_w);
}

DAN_D float4 pp_box(const PpArgs& A, int b, int a) { return pp_finish(A, pp_load(A, b, a)); }

// ---------------------------------------------------------------------------
// K3: filter + compaction.  grid (ceil(N/256), B)
// ---------------------------------------------------------------------------
constexpr int kFilterPerThread = 4;     // anchors per thread: 4 independent logit loads in flight (memory-level parallelism)

// exact per-anchor path: softmax, threshold, decode + clip, min-size, warp-aggregated append of the survivors.
// STAGED (two classes = one list per image): the keys are collected in the CTA's shared memory and appended to the list
// with ONE global atomic per CTA; per-warp atomics on the same counter serialise in L2 (ncu: 20 % of the stall samples).
template <bool STAGED>
DAN_D void filter_one(const PpArgs& A, int b, int a, bool live, int lane, unsigned long long* s_keys, int* s_cnt) {
  const int C = A.num_classes;
  const float* x = A.cls + ((int64_t)b * A.n + (live ? a : 0)) * C;
  // tf.nn.softmax: exp(x - max) * (1 / sum(exp(x - max))), sum in class order
  float mx = 0.f, inv = 0.f;
  if (live) {
    mx = x[0];
    for (int k = 1; k < C; ++k) mx = fmaxf(mx, x[k]);
    float s = 0.f;
    for (int k = 0; k < C; ++k) {
      const float e = cephes_expf(fsub(x[k], mx));
      s = (k == 0) ? e : fadd(s, e);
    }
    inv = fdiv(1.f, s);
  }
  float4 box;
  bool have_box = false;
  for (int c = 1; c < C; ++c) {
    bool pass = false;
    float p = 0.f;
    if (live) {
      p = fmul(cephes_expf(fsub(x[c], mx)), inv);
      if (p > A.select_thr) {                       // select_bboxes :24-36
        if (!have_box) {
          box = pp_box(A, b, a);
          have_box = true;
          if (A.stash != nullptr) A.stash[(int64_t)b * A.n + a] = box;
        }
        const float w = fadd(fsub(box.w, box.y), 1.f);   // filter_bboxes :50-59
        const float h = fadd(fsub(box.z, box.x), 1.f);
        pass = (w > A.min_size_p1) && (h > A.min_size_p1);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m != 0u) {
      const int list = b * (C - 1) + (c - 1);
      const unsigned long long key = ((unsigned long long)score_to_key(p) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
      int base = 0;
      if (lane == 0) base = STAGED ? atomicAdd(s_cnt, __popc(m)) : atomicAdd(A.key_count + list, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (STAGED) s_keys[pos] = key;
        else A.keys[(int64_t)list * A.n + pos] = key;
      }
    }
  }
}

__global__ void __launch_bounds__(256) pp_filter_kernel(const PpArgs A) {
  __shared__ unsigned long long s_keys[256 * kFilterPerThread];
  __shared__ int s_cnt, s_base;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int a0 = blockIdx.x * (256 * kFilterPerThread) + threadIdx.x;
  const bool staged = A.num_classes == 2;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  // Two classes: softmax_1 = sigmoid(x1 - x0).  An anchor whose logit difference is more than 0.05 below
  // logit(threshold) cannot pass (the fp32 evaluation is accurate to ~1e-6 relative), so the ~98 % background anchors
  // leave after one subtraction.  The exact path runs per 32-anchor group when any of its lanes may pass.
  bool maybe[kFilterPerThread];
  if (A.num_classes == 2) {
    float2 xx[kFilterPerThread];
#pragma unroll
    for (int u = 0; u < kFilterPerThread; ++u) {
      const int a = a0 + u * 256;
      xx[u] = (a < A.n) ? __ldcs(reinterpret_cast<const float2*>(A.cls + ((int64_t)b * A.n + a) * 2)) : make_float2(0.f, 0.f);   // read once: streaming
    }
#pragma unroll
    for (int u = 0; u < kFilterPerThread; ++u)
      maybe[u] = (a0 + u * 256 < A.n) &&
                 (!(fsub(xx[u].y, xx[u].x) < A.reject_below) || fabsf(xx[u].x) > 1e5f || fabsf(xx[u].y) > 1e5f);
  } else {
#pragma unroll
    for (int u = 0; u < kFilterPerThread; ++u) maybe[u] = a0 + u * 256 < A.n;
  }
  // compact the warp's few possibly-passing anchors (out of its 128) into dense lanes, then run the exact path on
  // full warps: without this ~60 % of the 32-anchor groups would run it for one or two live lanes
  __shared__ int s_list[8][32 * kFilterPerThread];
  int* mylist = s_list[threadIdx.x >> 5];
  int total = 0;
#pragma unroll
  for (int u = 0; u < kFilterPerThread; ++u) {
    const unsigned m = __ballot_sync(0xffffffffu, maybe[u]);
    if (maybe[u]) mylist[total + __popc(m & ((1u << lane) - 1u))] = a0 + u * 256;
    total += __popc(m);
  }
  __syncwarp();
  for (int j = 0; j < total; j += 32) {
    const bool live = j + lane < total;
    if (staged) filter_one<true>(A, b, live ? mylist[j + lane] : 0, live, lane, s_keys, &s_cnt);
    else filter_one<false>(A, b, live ? mylist[j + lane] : 0, live, lane, s_keys, &s_cnt);
  }
  if (staged) {
    __syncthreads();
    const int cnt = s_cnt;
    if (threadIdx.x == 0 && cnt > 0) s_base = atomicAdd(A.key_count + b, cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += 256) A.keys[(int64_t)b * A.n + s_base + i] = s_keys[i];
  }
}

// standalone sort_bboxes / nms_bboxes: one key per input row, nothing filtered
__global__ void __launch_bounds__(256) key_build_kernel(const float* __restrict__ scores, int64_t n, unsigned long long* __restrict__ keys,
                                                        int32_t* __restrict__ key_count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ((unsigned long long)score_to_key(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
  if (blockIdx.x == 0 && threadIdx.x == 0) key_count[0] = (int32_t)n;
}

DAN_D uint32_t key_index(unsigned long long key) { return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull); }

// sort_bboxes: tf.nn.top_k + gather + zero pad (bbox_util.py:61-72)
__global__ void __launch_bounds__(kSortThreads, 1) topk_sort_kernel(const PpArgs A, const float* __restrict__ src_scores,
                                                                 const float4* __restrict__ src_boxes) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ SortScratch sc;
  const int tid = threadIdx.x;
  const int cnt = min(A.key_count[0], A.n);
  const int k = min(A.keep_topk, cnt);
  const unsigned long long* sorted = s_keys;
  if (k <= kSortCap) {
    select_and_sort(A.keys, cnt, k, s_keys, sc);
  } else {
    select_and_sort_large(A.keys, cnt, k, A.s_key, A.s_key2, s_keys, sc);
    sorted = A.s_key;
  }
  for (int r = tid; r < A.keep_topk; r += kSortThreads) {
    if (r < k) {
      const uint32_t idx = key_index(sorted[r]);
      A.s_scores[r] = src_scores[idx];
      A.s_boxes[r] = src_boxes[idx];
      if (A.s_index != nullptr) A.s_index[r] = (int32_t)idx;
    } else {
      A.s_scores[r] = 0.f;
      A.s_boxes[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A.s_index != nullptr) A.s_index[r] = -1;
    }
  }
}

// ---------------------------------------------------------------------------
// pair tests of tf.image.non_max_suppression (shared by the NMS kernel below)
// ---------------------------------------------------------------------------
struct NmsBox {
  float y0, x0, y1, x1, area;
};

DAN_D NmsBox nms_norm(float4 b) {
  NmsBox r;
  r.y0 = fminf(b.x, b.z);
  r.x0 = fminf(b.y, b.w);
  r.y1 = fmaxf(b.x, b.z);
  r.x1 = fmaxf(b.y, b.w);
  r.area = fmul(fsub(r.y1, r.y0), fsub(r.x1, r.x0));
  return r;
}

// IOUGreaterThanThreshold of TF's non_max_suppression_op.cc for normalised boxes.  Kept out of line: the callers reach
// it only when a pair is within 1e-6 of the threshold (or the threshold is negative), and the chunk loop of the NMS
// kernel has to stay small enough for the instruction cache.
__device__ __noinline__ bool nms_suppresses(float4 a, float a_area, float4 b, float b_area, float thr) {
  if (a_area <= 0.f || b_area <= 0.f) return false;
  const float h = fmaxf(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float w = fmaxf(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = fmul(h, w);
  if (inter == 0.f) return 0.f > thr;     // 0 / (area_a + area_b) == 0 exactly
  return fdiv(inter, fsub(fadd(a_area, b_area), inter)) > thr;
}

// same predicate; for thr >= 0 disjoint boxes are rejected first and the division is only evaluated when
// inter / union is within 1e-6 (relative) of the threshold
DAN_D bool pair_suppresses(const float4& a, float a_area, const float4& b, float b_area, float thr) {
  if (thr >= 0.f) {
    const float h = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float w = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (!(h > 0.f && w > 0.f)) return false;
    if (!(a_area > 0.f && b_area > 0.f)) return false;
    const float inter = fmul(h, w);
    const float t = fmul(thr, fsub(fadd(a_area, b_area), inter));
    if (t > 1e-30f) {
      if (inter > fmul(t, 1.000001f)) return true;
      if (inter < fmul(t, 0.999999f)) return false;
    }
  }
  return nms_suppresses(a, a_area, b, b_area, thr);
}

// two classes: softmax_1 = sigmoid(x1 - x0); an anchor whose logit difference is more than 0.05 below logit(threshold)
// cannot pass (the fp32 evaluation is accurate to ~1e-6 relative), so ~97 % of the anchors leave after one subtraction
DAN_D bool f2_maybe(const PpArgs& A, float x0, float x1) {
  return !(fsub(x1, x0) < A.reject_below) || fabsf(x0) > 1e5f || fabsf(x1) > 1e5f;
}

// ---------------------------------------------------------------------------
// K4+K5: one CTA per (image, class) list does everything after the filter:
//   1. top-k select + sort of the surviving keys in shared memory (select_and_sort, sort.cuh);
//   2. decode + clip + normalise the K best boxes (shared memory, or the workspace for very long lists);
//   3. greedy NMS exactly as tf.image.non_max_suppression runs it - a candidate is tested against the boxes KEPT so
//      far, never against all earlier candidates - but kChunk candidates at a time:
//        a. every candidate of the chunk against the kept list, which is indexed by 32 x 32 stripes of the image
//           (bitsets over the kept boxes, see the kernel): a candidate only meets the few kept boxes that share a
//           stripe with it on both axes and leaves at the first suppressor.  Final for the candidates it suppresses,
//           because the kept list only grows;
//        b. the survivors are compacted in rank order;
//        c. their mutual suppression bits T[v] = { u < v : IoU(u, v) > thr } (<= kChunk^2 / 2 tests);
//        d. kept(v) = no kept u in T[v], evaluated by relaxation (a survivor is decided once one of its suppressors is
//           known kept, or all are known suppressed): the number of sweeps is the longest dependency chain inside the
//           chunk, a handful; chains longer than kMaxSweeps are finished serially by one warp;
//        e. the kept survivors are appended to the kept list in rank order; the loop ends when nms_topk boxes are kept
//           (max_output_size) or the candidates run out.
//      The detections of a trained detector are clusters of near duplicates around each face: almost every candidate
//      is removed in step a by its cluster's head after one or two exact tests, instead of the K^2 / 2 pair tests of a
//      suppression matrix.
//   4. the first nms_topk kept boxes are written out in rank order, zero padded (bbox_util.py:80-90).
// ---------------------------------------------------------------------------
constexpr int kWin = kSortThreads;                     // candidates in the window: one per thread
#ifndef DAN_NMS_SURV
#define DAN_NMS_SURV 128
#endif
constexpr int kSurv = DAN_NMS_SURV;                    // survivors resolved per round
constexpr int kSurvWords = kSurv / 32;
constexpr int kMaxSweeps = 12;
constexpr int kHintCells = 32 * 32;                    // one hint pair per (y stripe, x stripe) cell of a box centre
constexpr size_t kGreedySmemMax = 186 * 1024;          // dynamic part (the kernel has ~38 KB of static shared memory)

struct GreedyPlan {
  size_t smem;          // dynamic shared memory
  int sort_bytes;       // bytes reserved for the keys while they are sorted
  int kcap;             // candidate slots
  int wcap;             // 32-bit words per stripe bitset = ceil(nms_cap / 32)
  int cand_global;      // candidates live in the workspace instead of shared memory
  int kept_global;      // so do the kept list and its stripe bitsets
};

static int next_pow2(int64_t v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static GreedyPlan greedy_plan(int64_t n, int keep_topk, int nms_cap) {
  GreedyPlan g;
  const int sort_n = (int)(n < kSortCap ? (n > 0 ? n : 1) : kSortCap);
  g.sort_bytes = next_pow2(sort_n) * 8;
  g.kcap = (int)(keep_topk < n ? keep_topk : (n > 0 ? n : 1));      // candidate slots: no upper limit (the workspace holds them then)
  g.wcap = (nms_cap + 31) / 32;
  const size_t kept_bytes = align_up((size_t)nms_cap * 16, 16) + 2 * align_up((size_t)nms_cap * 4, 16) + (size_t)64 * g.wcap * 4;
  const size_t cand_box = align_up((size_t)g.kcap * 16, 16);
  const size_t region0 = cand_box > (size_t)g.sort_bytes ? cand_box : (size_t)g.sort_bytes;
  const size_t full = region0 + align_up((size_t)g.kcap * 4, 16) + align_up((size_t)g.kcap * 8, 16) + kept_bytes;
  g.cand_global = g.kept_global = 0;
  g.smem = full;
  if (full > kGreedySmemMax) {
    g.cand_global = 1;
    g.smem = (size_t)g.sort_bytes + kept_bytes;
    if (g.smem > kGreedySmemMax) {
      g.kept_global = 1;
      g.smem = (size_t)g.sort_bytes;
    }
  }
  return g;
}

// block-wide min / max of a float over all threads (every thread gets the result)
DAN_D void block_minmax(float lo, float hi, float& out_lo, float& out_hi) {
  __shared__ int s_lo[kSortThreads / 32], s_hi[kSortThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wl = __reduce_min_sync(0xffffffffu, float_to_ordered(lo));
  const int wh = __reduce_max_sync(0xffffffffu, float_to_ordered(hi));
  if (lane == 0) { s_lo[warp] = wl; s_hi[warp] = wh; }
  __syncthreads();
  int a = (lane < kSortThreads / 32) ? s_lo[lane] : 0x7fffffff, b = (lane < kSortThreads / 32) ? s_hi[lane] : (int)0x80000000;
  a = __reduce_min_sync(0xffffffffu, a);
  b = __reduce_max_sync(0xffffffffu, b);
  out_lo = ordered_to_float(a);
  out_hi = ordered_to_float(b);
  __syncthreads();
}

// stripe masks of the broad phase (see the kernel); monotone in the coordinate, so overlapping intervals share a bit
struct Stripes {
  float ylo, yinv, xlo, xinv;
  bool all;
  DAN_D static uint32_t axis(float v0, float v1, float lo, float inv) {
    const float f0 = (v0 - lo) * inv, f1 = (v1 - lo) * inv;
    if (!(fabsf(f0) < 1e9f) || !(fabsf(f1) < 1e9f)) return 0xffffffffu;      // non-finite or far outside: every stripe
    const int s0 = min(max((int)f0, 0), 31), s1 = min(max((int)f1, 0), 31);
    return (0xffffffffu >> (31 - s1)) & (0xffffffffu << s0);
  }
  // (x mask, y mask) of a normalised box (y0, x0, y1, x1); a box without area neither suppresses nor is suppressed
  DAN_D uint2 masks(float4 bx, float area) const {
    if (!(area > 0.f)) return make_uint2(0u, 0u);
    if (all) return make_uint2(0xffffffffu, 0xffffffffu);
    return make_uint2(axis(bx.y, bx.w, xlo, xinv), axis(bx.x, bx.z, ylo, yinv));
  }
  // cell (y stripe * 32 + x stripe) of the box centre: the key of the suppressor hints (any cell is valid for any box)
  DAN_D int cell(float4 bx) const {
    const float fy = ((bx.x + bx.z) * 0.5f - ylo) * yinv, fx = ((bx.y + bx.w) * 0.5f - xlo) * xinv;
    const int sy = (fabsf(fy) < 1e9f) ? min(max((int)fy, 0), 31) : 0;
    const int sx = (fabsf(fx) < 1e9f) ? min(max((int)fx, 0), 31) : 0;
    return sy * 32 + sx;
  }
};

// barriers among the first kSurv threads of the CTA (hardware barrier 1); the other warps wait at the next __syncthreads
DAN_D void bar_sync_chunk() { asm volatile("bar.sync 1, %0;" ::"n"(kSurv) : "memory"); }
DAN_D bool bar_or_chunk(bool pred) {
  int r;
  asm volatile(
      "{\n"
      "  .reg .pred p, q;\n"
      "  setp.ne.s32 q, %1, 0;\n"
      "  bar.red.or.pred p, 1, %2, q;\n"
      "  selp.s32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(r)
      : "r"((int)pred), "n"(kSurv)
      : "memory");
  return r != 0;
}

// WHERE: 0 = candidates and kept list in shared memory (the usual case), 1 = candidates in the workspace, 2 = both in the
// workspace (very long lists).  A template parameter rather than a run-time pointer choice so that the shared-memory
// accesses compile to LDS / STS / ATOMS instead of generic-address instructions.
// A dependency chain longer than kMaxSweeps inside a chunk: one warp finishes the chunk serially (when it reaches an
// undecided survivor every earlier one is decided); lane w holds word w of the masks.  Out of line (rare).
__device__ __noinline__ void resolve_chunk_serially(const uint32_t* T, uint32_t* keptm, uint32_t* supm, int ns) {
  const int lane = threadIdx.x & 31;
  uint32_t kw = (lane < kSurvWords) ? keptm[lane] : 0u;
  uint32_t sw = (lane < kSurvWords) ? supm[lane] : 0u;
  for (int v = 0; v < ns; ++v) {
    const int w = v >> 5;
    const uint32_t bit = 1u << (v & 31);
    const uint32_t known = __shfl_sync(0xffffffffu, kw | sw, w);
    if (known & bit) continue;                 // warp-uniform
    const uint32_t t = (lane < kSurvWords && (lane << 5) < v) ? T[v * kSurvWords + lane] : 0u;
    const bool hit = __any_sync(0xffffffffu, (t & kw) != 0u);
    if (lane == w) {
      if (hit) sw |= bit;
      else kw |= bit;
    }
  }
  if (lane < kSurvWords) { keptm[lane] = kw; supm[lane] = sw; }
}

// FILTER (two classes, one list per image): the CTA also runs K3 for its image - it streams the image's logits, keeps
// the indices of the few anchors the quick reject lets through, runs them through the exact path (softmax in
// tf.nn.softmax's op order, threshold, decode, clip, min-size) on dense warps, and the surviving keys land directly in
// the shared memory the sort works in: no filter launch, no key list in HBM, no global atomics.
#ifndef DAN_NMS_REGS
#define DAN_NMS_BOUNDS __launch_bounds__(kSortThreads, 1024 / kSortThreads)
#else                                     // experiment: cap the registers so that CTAs of other kernels fit beside this one
#define DAN_NMS_BOUNDS __maxnreg__(DAN_NMS_REGS)
#endif
constexpr int kMaybeCap = 4096;                        // indices kept in shared memory (the rest goes to the workspace)

template <bool DECODE, int WHERE, bool FILTER>
__global__ void DAN_NMS_BOUNDS nms_greedy_kernel(const PpArgs A, const float* __restrict__ src_scores,
                                                                     const float4* __restrict__ src_boxes) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ SortScratch sc;
  __shared__ unsigned char s_alive[kWin];              // ring over the window: candidate r lives in slot r % kWin
  __shared__ unsigned short s_list[kWin];              // step a: window offsets of the candidates that need the full search
  __shared__ int s_surv[kSurv];                        // step b: rank of survivor u
  __shared__ float4 s_sbox[kSurv];
  __shared__ float s_sarea[kSurv];
  __shared__ uint32_t s_T[kSurv][kSurvWords];
  __shared__ uint32_t s_keptm[kSurvWords], s_supm[kSurvWords];
  __shared__ uint2 s_smask[kSurv];
  __shared__ int s_hint[2][kHintCells];                // suppressor hints per cell: [0] first kept box centred there, [1] last suppressor found
  __shared__ int s_wcnt[kSortThreads / 32];
  __shared__ int s_next[2];                            // work counters of steps a and c
  __shared__ int s_nlist, s_cnext;

  const int list = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int b = list / max(A.num_classes - 1, 1);
  const int64_t o = (int64_t)list * A.kstride;
  const float thr = A.nms_thr;

  // ---- carve
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dyn_smem);
  float4* cbox;
  float* carea;
  uint2* cmask;
  unsigned char* p;
  if (WHERE >= 1) {
    cbox = A.s_box + o;
    carea = A.s_area + o;
    cmask = A.s_mask + o;
    p = dyn_smem + A.sort_bytes;
  } else {
    const size_t cand_box = ((size_t)A.kcap * 16 + 15) / 16 * 16;
    cbox = reinterpret_cast<float4*>(dyn_smem);                                    // aliases the keys (see step 2)
    carea = reinterpret_cast<float*>(dyn_smem + (cand_box > (size_t)A.sort_bytes ? cand_box : (size_t)A.sort_bytes));
    cmask = reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(carea) + ((size_t)A.kcap * 4 + 15) / 16 * 16);
    p = reinterpret_cast<unsigned char*>(cmask) + ((size_t)A.kcap * 8 + 15) / 16 * 16;
  }
  float4* kbox;
  float* karea;
  int* krank;
  uint32_t* stripes;                                   // [2][32][wcap]: kept boxes touching x stripe s / y stripe s
  const int wcap = A.wcap;
  if (WHERE >= 2) {
    const int64_t ok = (int64_t)list * A.kstride;
    kbox = A.g_kept_box + ok;
    karea = A.g_kept_area + ok;
    krank = A.g_kept_rank + ok;
    stripes = A.g_stripes + (int64_t)list * 64 * ((A.kstride + 31) / 32);
  } else {
    kbox = reinterpret_cast<float4*>(p); p += ((size_t)A.nms_cap * 16 + 15) / 16 * 16;
    karea = reinterpret_cast<float*>(p); p += ((size_t)A.nms_cap * 4 + 15) / 16 * 16;
    krank = reinterpret_cast<int*>(p); p += ((size_t)A.nms_cap * 4 + 15) / 16 * 16;
    stripes = reinterpret_cast<uint32_t*>(p);
  }
  uint32_t* xs = stripes;
  uint32_t* ys = stripes + 32 * wcap;

  // ---- 0. filter (FILTER) or the number of keys the filter kernel left in the workspace
  DAN_PHASE(0);
  unsigned long long* gkeys = A.keys + (int64_t)list * A.n;
  int cnt;
  if (FILTER) {
    __shared__ int s_maybe[kMaybeCap];
    __shared__ int s_nmaybe, s_nkeys;
    int* gmaybe = A.maybe + (int64_t)list * A.n;
    const int key_cap = A.sort_bytes / 8;              // keys that fit the sort's shared memory
    if (tid == 0) { s_nmaybe = 0; s_nkeys = 0; }
    __syncthreads();
    auto put = [&](int pos, int a) {
      if (pos < kMaybeCap) s_maybe[pos] = a;
      else gmaybe[pos] = a;
    };
    auto push = [&](int a) { put(atomicAdd(&s_nmaybe, 1), a); };
    // the image's logits as 16-byte vectors (two anchors); the row of an image starts 8-byte aligned only
    const float* x = A.cls + (int64_t)b * A.n * 2;
    const int head = (reinterpret_cast<uintptr_t>(x) & 15u) ? 1 : 0;
    const int nvec = (A.n - head) >> 1;
    const float4* xv = reinterpret_cast<const float4*>(x + 2 * head);
    for (int v0 = 0; v0 < nvec; v0 += 4 * kSortThreads) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * kSortThreads + tid;
        q[u] = (v < nvec) ? __ldcs(xv + v) : make_float4(0.f, 0.f, 0.f, 0.f);       // read once: streaming
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // A warp looks at 64 consecutive anchors here.  Its survivors enter the list in ANCHOR order with one atomic per
        // warp: the exact pass below then reads neighbouring box rows from neighbouring lanes, i.e. whole sectors - which
        // matters when the rows live in host memory and every sector is a PCIe read.
        const int v = v0 + u * kSortThreads + tid;
        const bool p0 = v < nvec && f2_maybe(A, q[u].x, q[u].y);
        const bool p1 = v < nvec && f2_maybe(A, q[u].z, q[u].w);
        const unsigned m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
        if ((m0 | m1) != 0u) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_nmaybe, __popc(m0) + __popc(m1));
          base = __shfl_sync(0xffffffffu, base, 0);
          int pos = base + __popc(m0 & lt_mask) + __popc(m1 & lt_mask);
          if (p0) put(pos++, head + 2 * v);
          if (p1) put(pos, head + 2 * v + 1);
        }
      }
    }
    if (tid == 0 && head == 1 && f2_maybe(A, x[0], x[1])) push(0);
    if (tid == 1 && head + 2 * nvec < A.n && f2_maybe(A, x[2 * (A.n - 1)], x[2 * (A.n - 1) + 1])) push(A.n - 1);
    __syncthreads();
    const int nmaybe = s_nmaybe;
    for (int i0 = 0; i0 < nmaybe; i0 += kSortThreads) {
      const int i = i0 + tid;
      if (i < nmaybe) {
        const int a = (i < kMaybeCap) ? s_maybe[i] : gmaybe[i];
        const float2 xx = *reinterpret_cast<const float2*>(x + 2 * a);
        // tf.nn.softmax: exp(x - max) * (1 / sum(exp(x - max))), sum in class order
        const float mx = fmaxf(xx.x, xx.y);
        const float e0 = cephes_expf(fsub(xx.x, mx));
        const float e1 = cephes_expf(fsub(xx.y, mx));
        const float pr = fmul(e1, fdiv(1.f, fadd(e0, e1)));
        if (pr > A.select_thr) {                            // select_bboxes :24-36
          const float4 box = pp_box(A, b, a);
          if (A.stash != nullptr) A.stash[(int64_t)b * A.n + a] = box;
          const float w = fadd(fsub(box.w, box.y), 1.f);    // filter_bboxes :50-59
          const float h = fadd(fsub(box.z, box.x), 1.f);
          if ((w > A.min_size_p1) && (h > A.min_size_p1)) {
            const unsigned long long key = ((unsigned long long)score_to_key(pr) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
            const int pos = atomicAdd(&s_nkeys, 1);
            if (pos < key_cap) keys[pos] = key;
            else gkeys[pos] = key;
          }
        }
      }
    }
    __syncthreads();
    cnt = s_nkeys;
    if (cnt > kSortCap) {                                   // too many survivors for shared memory: the select runs on HBM
      for (int i = tid; i < min(cnt, key_cap); i += kSortThreads) gkeys[i] = keys[i];
      __syncthreads();
    }
  } else {
    cnt = min(A.key_count[list], A.n);
  }

  // ---- 1. top-k + sort
  DAN_PHASE(8);
  const int k_want = min(A.keep_topk, cnt);
  int K;
  const unsigned long long* sorted = keys;            // shared memory, or A.s_key for lists longer than kSortCap
  if (k_want <= kSortCap) {
    // a second key buffer behind the first one when the CTA's shared memory holds it (nothing else lives there yet)
    unsigned long long* spare = ((size_t)A.sort_bytes + (size_t)min(cnt, kSortCap) * 8 <= (size_t)A.dyn_smem_bytes)
                                    ? reinterpret_cast<unsigned long long*>(dyn_smem + A.sort_bytes) : nullptr;
    K = min(select_and_sort(gkeys, cnt, k_want, keys, sc, FILTER, spare), A.keep_topk);
  } else {
    select_and_sort_large(gkeys, cnt, k_want, A.s_key + o, A.s_key2 + o, keys, sc);
    sorted = A.s_key + o;
    K = k_want;
  }
  DAN_PHASE(1);

  // ---- 2. decode + clip + normalise.  The candidate boxes (16 B) reuse the shared memory of the keys (8 B): the ranks
  // are walked in DEscending batches of one key per thread, so a batch's boxes only overwrite key slots of ranks that
  // are already done (box r covers key slots 2r and 2r+1 >= r); the barrier protects the batch's own keys.
  for (int r0 = K > 0 ? ((K - 1) / kSortThreads) * kSortThreads : -1; r0 >= 0; r0 -= kSortThreads) {
    const int r = r0 + tid;
    const unsigned long long key = (r < K) ? sorted[r] : 0ull;
    if (r < K) A.s_key[o + r] = key;
    __syncthreads();
    if (r < K) {
      const uint32_t idx = key_index(key);
      const NmsBox nb = nms_norm(DECODE ? (A.stash != nullptr ? A.stash[(int64_t)b * A.n + idx] : pp_box(A, b, (int)idx)) : src_boxes[idx]);
      cbox[r] = make_float4(nb.y0, nb.x0, nb.y1, nb.x1);
      carea[r] = nb.area;
    }
  }
  __syncthreads();
  DAN_PHASE(2);

  // ---- 3. greedy NMS, kChunk candidates per round
#ifdef DAN_PHASE_TIMING
  long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_a = 0, t_b = 0, n_sweeps = 0, n_surv = 0, n_chunks = 0;
#define DAN_LAP(i) do { const long long now__ = clock64(); acc[i] += now__ - t_b; t_b = now__; } while (0)
#define DAN_TICK() (t_a = clock64())
#define DAN_TOCK(i) (acc[i] += clock64() - t_a)
#else
#define DAN_TICK() do { } while (0)
#define DAN_TOCK(i) do { } while (0)
#define DAN_LAP(i) do { } while (0)
#endif
  // Broad phase: the extent of the candidates is cut into 32 stripes per axis and every box carries the two 32-bit masks
  // of the stripes it touches.  Boxes that intersect share a stripe on both axes, so a pair whose masks do not meet on
  // either axis needs no exact test.
  Stripes sg;
  {
    float ylo = 3.0e38f, yhi = -3.0e38f, xlo = 3.0e38f, xhi = -3.0e38f;
    for (int r = tid; r < K; r += kSortThreads) {
      const float4 bx = cbox[r];
      if (carea[r] > 0.f && fabsf(bx.x) < 1e30f && fabsf(bx.y) < 1e30f && fabsf(bx.z) < 1e30f && fabsf(bx.w) < 1e30f) {
        ylo = fminf(ylo, bx.x); yhi = fmaxf(yhi, bx.z);
        xlo = fminf(xlo, bx.y); xhi = fmaxf(xhi, bx.w);
      }
    }
    float ey, ex;
    block_minmax(ylo, yhi, sg.ylo, ey);
    block_minmax(xlo, xhi, sg.xlo, ex);
    sg.yinv = (ey > sg.ylo) ? 32.f / (ey - sg.ylo) : 0.f;
    sg.xinv = (ex > sg.xlo) ? 32.f / (ex - sg.xlo) : 0.f;
    sg.all = thr < 0.f;                                // a negative threshold: disjoint boxes suppress too
  }
  for (int r = tid; r < K; r += kSortThreads) cmask[r] = sg.masks(cbox[r], carea[r]);
  for (int i = tid; i < 64 * wcap; i += kSortThreads) stripes[i] = 0u;
  for (int i = tid; i < 2 * kHintCells; i += kSortThreads) (&s_hint[0][0])[i] = 0x7fffffff;
  if (tid < 2) s_next[tid] = 0;
  if (tid == 0) s_nlist = 0;
  __syncthreads();
  int L = 0;                                           // kept boxes so far (CTA-uniform)
  int L_prev = 0;                                      // ... when the previous round tested its window
  int w_end_prev = 0;                                  // end of the previous round's window: candidates below it carry state
  for (int c0 = 0; c0 < K && L < A.nms_topk;) {
    const int w_end = min(K, c0 + kWin);
    DAN_TICK();
    // a. window vs kept list.  Thread t owns candidate c0 + t.  A candidate that was in the previous window has met
    // the first L_prev kept boxes already; a fresh one has met none.
    const int my_r = c0 + tid;
    const bool in_win = my_r < w_end;
    {
      bool alive = false, search = false;
      int upto = 0;
      if (in_win) {
        const bool fresh = my_r >= w_end_prev;
        alive = fresh ? true : (s_alive[my_r & (kWin - 1)] != 0);
        upto = fresh ? 0 : L_prev;
      }
      if (alive && L > upto) {
        const uint2 mm = cmask[my_r];
        if (mm.x != 0u && mm.y != 0u) {                // (no area: no stripes, never suppressed)
          // a1. the two hinted kept boxes of the cell of its centre: a member of a cluster of near duplicates dies here,
          // after one or two exact tests by its own thread
          const float4 me = cbox[my_r];
          const float my_area = carea[my_r];
          const int cl = sg.cell(me);
          const int h0 = s_hint[0][cl], h1 = s_hint[1][cl];
          bool dead = false;
          if (h0 >= upto && h0 < L) dead = pair_suppresses(kbox[h0], karea[h0], me, my_area, thr);
          if (!dead && h1 != h0 && h1 >= upto && h1 < L) dead = pair_suppresses(kbox[h1], karea[h1], me, my_area, thr);
          alive = !dead;
          search = !dead;
        }
      }
      if (in_win) s_alive[my_r & (kWin - 1)] = alive ? 1 : 0;
      const unsigned sm = __ballot_sync(0xffffffffu, search);
      int base = 0;
      if (lane == 0 && sm != 0u) base = atomicAdd(&s_nlist, __popc(sm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (search) s_list[base + __popc(sm & lt_mask)] = (unsigned short)tid;
    }
    __syncthreads();
    // a2. the others against the kept list proper, which is indexed by stripe: bit j of xs[s] / ys[s] says that kept box j
    // touches x / y stripe s.  The kept boxes a candidate can intersect are (OR of xs over its x stripes) AND (OR of ys
    // over its y stripes): a group of G lanes owns one candidate, a lane one 32-bit word of the bitset, and only the set
    // bits (beyond the boxes the candidate has met already) go through the exact test; the candidate leaves at the
    // first suppressor, which becomes the cell's second hint.
    const int nlist = s_nlist;
    if (nlist > 0) {
      const int W = (L + 31) >> 5;
      const int G = W <= 8 ? 8 : (W <= 16 ? 16 : 32);        // lanes per candidate
      const int cpw = 32 / G;                                // candidates per warp pass
      const int gl = lane & (G - 1), sub = lane / G;
      const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (sub * G);
      while (true) {                                         // candidates are handed out dynamically: their cost varies a lot
        int i0 = 0;
        if (lane == 0) i0 = atomicAdd(&s_next[0], cpw);
        i0 = __shfl_sync(0xffffffffu, i0, 0);
        if (i0 >= nlist) break;
        const bool have = i0 + sub < nlist;
        const int r = c0 + (have ? (int)s_list[i0 + sub] : 0);
        const float4 me = have ? cbox[r] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float my_area = have ? carea[r] : 0.f;
        const uint2 mm = have ? cmask[r] : make_uint2(0u, 0u);
        const int upto = (r >= w_end_prev) ? 0 : L_prev;
        bool sup = false;
        for (int wb = 0; wb < W; wb += G) {                  // one block of words unless the kept list is very long
          const int w = wb + gl;
          uint32_t poss = 0u;
          if (w < W && w >= (upto >> 5) && mm.x != 0u && mm.y != 0u) {
            // (the stripes of a box are a contiguous range: independent loads, no address chain)
            uint32_t ax = 0u, ay = 0u;
            const int xe = 31 - __clz(mm.x), ye = 31 - __clz(mm.y);
#pragma unroll 2
            for (int st = __ffs(mm.x) - 1; st <= xe; ++st) ax |= xs[st * wcap + w];
#pragma unroll 2
            for (int st = __ffs(mm.y) - 1; st <= ye; ++st) ay |= ys[st * wcap + w];
            poss = ax & ay;
            if (w == (upto >> 5)) poss &= ~((1u << (upto & 31)) - 1u);
          }
          while (__any_sync(0xffffffffu, poss != 0u && !sup)) {
            bool t = false;
            if (poss != 0u && !sup) {
              const int j = (w << 5) + __ffs(poss) - 1;
              poss &= poss - 1u;
              t = pair_suppresses(kbox[j], karea[j], me, my_area, thr);
              if (t) atomicExch(&s_hint[1][sg.cell(me)], j);
            }
            if ((__ballot_sync(0xffffffffu, t) & gmask) != 0u) sup = true;
          }
        }
        if (have && gl == 0 && sup) s_alive[r & (kWin - 1)] = 0;
      }
    }
    __syncthreads();
    DAN_TOCK(0);
    DAN_TICK();
    // b. the first kSurv candidates of the window that are still alive, in rank order; the window of the next round
    // starts behind the last of them
    const bool f = in_win && s_alive[my_r & (kWin - 1)] != 0;
    const unsigned fm = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_wcnt[warp] = __popc(fm);
    if (tid < kSurvWords) { s_keptm[tid] = 0u; s_supm[tid] = 0u; }
    __syncthreads();
    int total, ubase;
    {
      const int cw = (lane < kSortThreads / 32) ? s_wcnt[lane] : 0;
      int incl = cw;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      total = __shfl_sync(0xffffffffu, incl, 31);
      ubase = __shfl_sync(0xffffffffu, incl - cw, warp & 31);
    }
    if (f) {
      const int u = ubase + __popc(fm & lt_mask);
      if (u < kSurv) {
        s_surv[u] = my_r;
        s_sbox[u] = cbox[my_r];
        s_sarea[u] = carea[my_r];
        s_smask[u] = cmask[my_r];
        if (u == kSurv - 1) s_cnext = my_r + 1;
      }
    }
    __syncthreads();
    const int ns = min(total, kSurv);
    const int c_next = (total >= kSurv) ? s_cnext : w_end;
    DAN_TOCK(1);
    DAN_TICK();
    // c. suppression bits among the survivors: warp -> row v, lanes -> 32 earlier survivors per step
    while (true) {                                           // rows handed out dynamically, the longest ones first
      int vi = 0;
      if (lane == 0) vi = atomicAdd(&s_next[1], 1);
      vi = __shfl_sync(0xffffffffu, vi, 0);
      if (vi >= ns) break;
      const int v = ns - 1 - vi;
      const float4 me = s_sbox[v];
      const float my_area = s_sarea[v];
      const uint2 mm = s_smask[v];
      const int nw = (v + 31) >> 5;                    // words that hold some u < v
      for (int w = 0; w < nw; ++w) {
        const int u = (w << 5) + lane;
        const uint2 um = s_smask[u];
        const bool poss = (u < v) && (um.x & mm.x) != 0u && (um.y & mm.y) != 0u;
        unsigned bits = 0u;
        if (__any_sync(0xffffffffu, poss)) {
          const bool t = poss && pair_suppresses(s_sbox[u], s_sarea[u], me, my_area, thr);
          bits = __ballot_sync(0xffffffffu, t);
        }
        if (lane == 0) s_T[v][w] = bits;
      }
    }
    __syncthreads();
    DAN_TOCK(2);
    DAN_TICK();
    // d. relaxation: thread v owns survivor v (the first kSurv threads)
    int sweeps = 0;
    if (tid < kSurv) {
      uint32_t T[kSurvWords];
      const int vw = tid >> 5;
      const uint32_t vbit = 1u << (tid & 31);
      const bool mine = tid < ns;
#pragma unroll
      for (int w = 0; w < kSurvWords; ++w) T[w] = (mine && (w << 5) < tid) ? s_T[tid][w] : 0u;
      bool decided = !mine;
      bool open = true;
      while (open && sweeps < kMaxSweeps) {
        uint32_t hitm = 0u, pendm = 0u;
        if (!decided) {
#pragma unroll
          for (int w = 0; w < kSurvWords; ++w) {
            const uint32_t kw = s_keptm[w], sw = s_supm[w];
            hitm |= T[w] & kw;
            pendm |= T[w] & ~(kw | sw);
          }
        }
        bar_sync_chunk();                              // every thread has read the masks of the previous sweep
        if (!decided) {
          if (hitm != 0u) { atomicOr(&s_supm[vw], vbit); decided = true; }
          else if (pendm == 0u) { atomicOr(&s_keptm[vw], vbit); decided = true; }
        }
        open = bar_or_chunk(!decided);
        ++sweeps;
      }
      if (open) {
        if (warp == 0) resolve_chunk_serially(&s_T[0][0], s_keptm, s_supm, ns);
        bar_sync_chunk();
      }
      DAN_TOCK(3);
    }
    __syncthreads();
    DAN_TICK();
    // e. append the kept survivors in rank order; a kept box becomes the first hint of its cell unless an earlier
    // (higher scoring) one is centred there.  ALL threads take part: kEntryThreads threads share one survivor, each
    // enters the kept box into every kEntryThreads-th stripe of both axes (a big box touches 2 x 20 stripes: in the hands
    // of the survivor's own thread alone this was the longest serial piece of a round).
    {
      constexpr int kEntryThreads = kSortThreads / kSurv;
      const int v = tid / kEntryThreads, part = tid % kEntryThreads;
      const int vw = v >> 5;
      const uint32_t vbit = 1u << (v & 31);
      if (tid < 2) s_next[tid] = 0;
      if (tid == 2) s_nlist = 0;
      if (v < ns && (s_keptm[vw] & vbit)) {
        int before = 0;
#pragma unroll
        for (int w = 0; w < kSurvWords; ++w)
          if (w < vw) before += __popc(s_keptm[w]);
        const int pos = L + before + __popc(s_keptm[vw] & (vbit - 1u));
        if (pos < A.nms_topk) {
          const uint2 sm = s_smask[v];
          const uint32_t bit = 1u << (pos & 31);
          if (part == 0) {
            const float4 bx = s_sbox[v];
            kbox[pos] = bx;
            karea[pos] = s_sarea[v];
            krank[pos] = s_surv[v];
            if (sm.x != 0u) atomicMin(&s_hint[0][sg.cell(bx)], pos);
          }
          for (int st = part; st < 32; st += kEntryThreads) {
            if ((sm.x >> st) & 1u) atomicOr(&xs[st * wcap + (pos >> 5)], bit);
            if ((sm.y >> st) & 1u) atomicOr(&ys[st * wcap + (pos >> 5)], bit);
          }
        }
      }
    }
    __syncthreads();
    {
      int nk = 0;
#pragma unroll
      for (int w = 0; w < kSurvWords; ++w) nk += __popc(s_keptm[w]);
      L_prev = L;
      L = min(L + nk, A.nms_topk);
    }
    w_end_prev = w_end;
    c0 = c_next;
#ifdef DAN_PHASE_TIMING
    DAN_TOCK(4);
    n_sweeps += sweeps; n_surv += ns; ++n_chunks;
#endif
  }
  const int kept_n = L;
  DAN_PHASE(3);
#ifdef DAN_PHASE_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int i = 0; i < 5; ++i) g_phase[10 + i] = acc[i];
    for (int i = 5; i < 12; ++i) g_phase[15 + i] = acc[i];
    g_phase[15] = n_sweeps; g_phase[16] = n_surv; g_phase[17] = n_chunks; g_phase[18] = K; g_phase[19] = kept_n;
  }
#endif

  // ---- 4. outputs, zero padded to nms_topk; with a peer exchange every slab row also goes to the other ranks
  auto put_score = [&](int64_t oo, float v) {
    A.out_scores[oo] = v;
    for (int q = 0; q < A.npeers; ++q) *reinterpret_cast<float*>(reinterpret_cast<char*>(A.out_scores + oo) + A.peer_delta[q]) = v;
  };
  auto put_box = [&](int64_t oo, float4 v) {
    A.out_boxes[oo] = v;
    for (int q = 0; q < A.npeers; ++q) *reinterpret_cast<float4*>(reinterpret_cast<char*>(A.out_boxes + oo) + A.peer_delta[q]) = v;
  };
  for (int t = tid; t < A.nms_topk; t += kSortThreads) {
    const int64_t oo = (int64_t)list * A.nms_topk + t;
    if (t < kept_n) {
      const int pos = krank[t];
      const unsigned long long key = A.s_key[o + pos];
      const uint32_t idx = key_index(key);
      put_score(oo, DECODE ? key_to_score((uint32_t)(key >> 32)) : src_scores[idx]);
      put_box(oo, DECODE ? kbox[t] : src_boxes[idx]);     // clipped boxes are already min/max ordered
      if (A.out_index != nullptr) A.out_index[oo] = (int32_t)idx;
      if (A.out_keep != nullptr) A.out_keep[oo] = A.filler ? pos : (int32_t)idx;
    } else {
      put_score(oo, 0.f);
      put_box(oo, make_float4(0.f, 0.f, 0.f, 0.f));
      if (A.out_index != nullptr) A.out_index[oo] = -1;
      if (A.out_keep != nullptr) {
        // parse_by_class runs NMS on the zero padded top-k list: zero-area filler rows are never suppressed
        // and get selected until nms_topk is reached
        const int fpos = K + (t - kept_n);
        A.out_keep[oo] = (A.filler && fpos < A.keep_topk) ? fpos : -1;
      }
    }
  }
  if (tid == 0 && A.out_counts != nullptr) {
    A.out_counts[list] = kept_n;
    for (int q = 0; q < A.npeers; ++q) *reinterpret_cast<int32_t*>(reinterpret_cast<char*>(A.out_counts + list) + A.peer_delta[q]) = kept_n;
  }
  if (A.npeers > 0) {
    // release: this CTA's rows are visible system-wide before it counts itself done; the last CTA of the launch bumps the
    // rank's step number and writes it into its flag slot in every destination (dan_wait_detections polls those)
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const int done = atomicAdd(A.peer_state, 1);
      if (done == (int)gridDim.x - 1) {
        A.peer_state[0] = 0;
        const int seq = A.peer_state[1] + 1;
        A.peer_state[1] = seq;
        __threadfence_system();
        for (int q = 0; q < A.npeers; ++q) *reinterpret_cast<volatile int32_t*>(A.peer_flag[q]) = seq;
      }
    }
  }
  DAN_PHASE(4);
}

// One warp polls the arrival flags of the `world` ranks in this rank's receive buffer until all have reached the step
// number this rank's own NMS kernel just produced (state[1]); it holds one warp of one SM, nothing else.
__global__ void wait_detections_kernel(const int32_t* flags, const int32_t* state, int world) {
  const int want = *reinterpret_cast<const volatile int32_t*>(state + 1);
  const int q = threadIdx.x;
  if (q < world) {
    while (*reinterpret_cast<const volatile int32_t*>(flags + q) < want) __nanosleep(100);
  }
  __syncwarp();
  __threadfence_system();
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

struct PpLayout {
  size_t key_count, keys, maybe, stash, s_key, s_key2, s_box, s_area, kept_box, kept_area, kept_rank, s_mask, stripes, total;
};

static PpLayout pp_layout(int64_t n, int64_t lists, int64_t keep_topk, bool nms, int64_t stash_rows = 0) {
  PpLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
  const int64_t ks = keep_topk < n ? keep_topk : (n > 0 ? n : 1);      // kstride
  w.key_count = take(lists * 4);
  w.keys = take(lists * n * 8);
  w.maybe = take(nms ? lists * n * 4 : 0);
  w.stash = take(stash_rows * 16);
  w.s_key = take(lists * ks * 8);
  w.s_key2 = take(lists * ks * 8);
  w.s_box = take(nms ? lists * ks * 16 : 0);
  w.s_area = take(nms ? lists * ks * 4 : 0);
  w.kept_box = take(nms ? lists * ks * 16 : 0);
  w.kept_area = take(nms ? lists * ks * 4 : 0);
  w.kept_rank = take(nms ? lists * ks * 4 : 0);
  w.s_mask = take(nms ? lists * ks * 8 : 0);
  w.stripes = take(nms ? lists * 64 * ((ks + 31) / 32) * 4 : 0);
  w.total = off;
  return w;
}

static void pp_bind(PpArgs& A, void* ws, const PpLayout& w) {
  char* base = static_cast<char*>(ws);
  A.key_count = reinterpret_cast<int32_t*>(base + w.key_count);
  A.keys = reinterpret_cast<unsigned long long*>(base + w.keys);
  A.maybe = reinterpret_cast<int32_t*>(base + w.maybe);
  A.s_key = reinterpret_cast<unsigned long long*>(base + w.s_key);
  A.s_key2 = reinterpret_cast<unsigned long long*>(base + w.s_key2);
  A.s_box = reinterpret_cast<float4*>(base + w.s_box);
  A.s_area = reinterpret_cast<float*>(base + w.s_area);
  A.g_kept_box = reinterpret_cast<float4*>(base + w.kept_box);
  A.g_kept_area = reinterpret_cast<float*>(base + w.kept_area);
  A.g_kept_rank = reinterpret_cast<int32_t*>(base + w.kept_rank);
  A.s_mask = reinterpret_cast<uint2*>(base + w.s_mask);
  A.g_stripes = reinterpret_cast<uint32_t*>(base + w.stripes);
}

// opt-in to > 48 KB of dynamic shared memory; the attribute belongs to the (function, device) pair, so it is set on
// every call (a few hundred ns of host time) rather than cached in a process-wide flag
template <bool DECODE, int WHERE, bool FILTER>
static int launch_greedy(const PpArgs& A, int lists, size_t smem, const float* src_scores, const float4* src_boxes, cudaStream_t st) {
  DAN_CUDA(cudaFuncSetAttribute(nms_greedy_kernel<DECODE, WHERE, FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_greedy_kernel<DECODE, WHERE, FILTER><<<lists, kSortThreads, smem, st>>>(A, src_scores, src_boxes);
  DAN_LAUNCH_CHECK("nms_greedy_kernel");
  return DAN_OK;
}

// [filter +] sort + greedy NMS for `lists` lists
template <bool DECODE, bool FILTER>
static int run_sort_nms(const PpArgs& A_in, int lists, const float* src_scores, const float4* src_boxes, cudaStream_t st) {
  PpArgs A = A_in;
  const GreedyPlan g = greedy_plan(A.n, A.keep_topk, A.nms_cap);
  A.sort_bytes = g.sort_bytes;
  A.kcap = g.kcap;
  A.wcap = g.wcap;
  A.kstride = g.kcap;
  A.dyn_smem_bytes = (int)g.smem;
  A.cand_global = g.cand_global;
  A.kept_global = g.kept_global;
  if (g.kept_global) return launch_greedy<DECODE, 2, FILTER>(A, lists, g.smem, src_scores, src_boxes, st);
  if (g.cand_global) return launch_greedy<DECODE, 1, FILTER>(A, lists, g.smem, src_scores, src_boxes, st);
  return launch_greedy<DECODE, 0, FILTER>(A, lists, g.smem, src_scores, src_boxes, st);
}

}  // namespace dan

using namespace dan;

extern "C" {

#ifdef DAN_PHASE_TIMING
int dan_debug_phases(long long* h_out32) { return cudaMemcpyFromSymbol(h_out32, g_phase, sizeof(long long) * 32) == cudaSuccess ? 0 : -3; }
#endif

size_t dan_postprocess_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t num_classes, int32_t keep_topk) {
  if (num_anchors < 0 || batch < 0 || num_classes < 2 || keep_topk < 1) return 0;
  return pp_layout(num_anchors, (int64_t)batch * (num_classes - 1), keep_topk, true, (int64_t)batch * num_anchors).total;
}

size_t dan_sort_workspace_bytes(int64_t n, int32_t keep_topk) {
  if (n < 0 || keep_topk < 1) return 0;
  return pp_layout(n, 1, keep_topk, false).total;
}

size_t dan_nms_workspace_bytes(int64_t n, int32_t nms_topk) {
  if (n < 0 || nms_topk < 0) return 0;
  return pp_layout(n, 1, n > 0 ? n : 1, true).total;
}

static int postprocess_core(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred, const float* boxes_pred,
                            const float* a_ymin, const float* a_xmin, const float* a_ymax, const float* a_xmax, int32_t num_anchors,
                            int32_t batch, float* out_boxes, float* out_scores, int32_t* out_counts, int32_t* out_anchor_index,
                            int32_t* out_keep_pos, void* workspace, size_t workspace_bytes, void* stream, cudaEvent_t* ev,
                            const dan_peer_exchange* peers = nullptr) {
  DAN_REQUIRE(p != nullptr, DAN_ERR_INVALID_ARGUMENT, "params is NULL");
  DAN_REQUIRE(p->num_classes >= 2, DAN_ERR_INVALID_ARGUMENT, "num_classes must be >= 2 (class 0 is background), got %d", p->num_classes);
  DAN_REQUIRE(num_anchors >= 0 && batch >= 0, DAN_ERR_INVALID_ARGUMENT, "negative size");
  DAN_REQUIRE(batch <= 65535, DAN_ERR_UNSUPPORTED, "batch > 65535");
  DAN_REQUIRE(p->select_threshold >= 0.f, DAN_ERR_INVALID_ARGUMENT,
              "select_threshold must be >= 0 (a negative threshold would let zero-score rows carry boxes), got %g", p->select_threshold);
  DAN_REQUIRE(p->keep_topk >= 1 && p->nms_topk >= 1, DAN_ERR_INVALID_ARGUMENT, "keep_topk and nms_topk must be >= 1");
  DAN_REQUIRE((loc_pred != nullptr) != (boxes_pred != nullptr), DAN_ERR_INVALID_ARGUMENT, "exactly one of loc_pred / boxes_pred must be given");
  if (batch == 0) return DAN_OK;
  DAN_REQUIRE(cls_pred && out_boxes && out_scores, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(loc_pred == nullptr || (a_ymin && a_xmin && a_ymax && a_xmax), DAN_ERR_INVALID_ARGUMENT, "anchors needed to decode loc_pred");
  DAN_REQUIRE(aligned16(loc_pred) && aligned16(boxes_pred) && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "box tensors must be 16-byte aligned");
  const int lists = batch * (p->num_classes - 1);
  const PpLayout w = pp_layout(num_anchors, lists, p->keep_topk, true, (int64_t)batch * num_anchors);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  PpArgs A = {};
  A.cls = cls_pred;
  A.loc = reinterpret_cast<const float4*>(loc_pred);
  A.boxes = reinterpret_cast<const float4*>(boxes_pred);
  A.ay0 = a_ymin; A.ax0 = a_xmin; A.ay1 = a_ymax; A.ax1 = a_xmax;
  A.n = num_anchors;
  A.batch = batch;
  A.num_classes = p->num_classes;
  A.img_h = (float)p->image_h;
  A.img_w = (float)p->image_w;
  A.select_thr = p->select_threshold;
  A.min_size_p1 = (float)((double)p->min_size + 1.0);   // python: min_size + 1. then fp32
  {
    const double t = (double)p->select_threshold;
    A.reject_below = (p->num_classes == 2 && t > 0.0 && t <= 0.99) ? (float)(log(t / (1.0 - t)) - 0.05) : -INFINITY;
  }
  A.ps0 = p->prior_scaling[0]; A.ps1 = p->prior_scaling[1]; A.ps2 = p->prior_scaling[2]; A.ps3 = p->prior_scaling[3];
  A.keep_topk = p->keep_topk;
  A.nms_topk = p->nms_topk;
  A.nms_cap = p->nms_topk < p->keep_topk ? p->nms_topk : p->keep_topk;
  A.nms_thr = p->nms_threshold;
  A.out_boxes = reinterpret_cast<float4*>(out_boxes);
  A.out_scores = out_scores;
  A.out_counts = out_counts;
  A.out_index = out_anchor_index;
  A.out_keep = out_keep_pos;
  A.filler = 1;
  if (peers != nullptr && peers->num_destinations > 0) {
    DAN_REQUIRE(peers->num_destinations <= DAN_MAX_PEERS, DAN_ERR_UNSUPPORTED, "more than %d destinations", DAN_MAX_PEERS);
    DAN_REQUIRE(peers->state != nullptr && out_counts != nullptr, DAN_ERR_INVALID_ARGUMENT, "peer exchange needs state words and out_counts");
    A.npeers = peers->num_destinations;
    for (int q = 0; q < A.npeers; ++q) {
      DAN_REQUIRE(peers->flag[q] != nullptr && (peers->delta_bytes[q] & 15) == 0, DAN_ERR_INVALID_ARGUMENT,
                  "destination %d: NULL flag or a slab offset that is not a multiple of 16 bytes", q);
      A.peer_delta[q] = peers->delta_bytes[q];
      A.peer_flag[q] = peers->flag[q];
    }
    A.peer_state = peers->state;
  }
  pp_bind(A, workspace, w);
  // loc_pred / boxes_pred in page-locked HOST memory (cudaHostAlloc / cudaHostRegister; mapped under unified addressing):
  // the kernels read only the rows of the anchors that pass the score threshold (~3 %), straight over PCIe, and keep
  // the decoded box of such a row in the workspace so that the row is fetched once.  Pageable memory is refused.
  {
    const void* geo = loc_pred != nullptr ? static_cast<const void*>(loc_pred) : static_cast<const void*>(boxes_pred);
    cudaPointerAttributes pa;
    const cudaError_t e = cudaPointerGetAttributes(&pa, geo);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
    } else if (pa.type == cudaMemoryTypeHost) {
      DAN_REQUIRE(pa.devicePointer != nullptr, DAN_ERR_INVALID_ARGUMENT, "loc_pred / boxes_pred: host memory that is not mapped for the device");
      if (loc_pred != nullptr) A.loc = reinterpret_cast<const float4*>(pa.devicePointer);
      else A.boxes = reinterpret_cast<const float4*>(pa.devicePointer);
      A.stash = reinterpret_cast<float4*>(static_cast<char*>(workspace) + w.stash);
    } else {
      DAN_REQUIRE(pa.type != cudaMemoryTypeUnregistered, DAN_ERR_INVALID_ARGUMENT,
                  "loc_pred / boxes_pred is pageable host memory: pass device memory or page-locked (pinned) host memory");
    }
  }
  // two classes: the NMS kernel filters its own image (one launch for the whole evaluation side)
  const bool fused = p->num_classes == 2 && (reinterpret_cast<uintptr_t>(cls_pred) & 7u) == 0;
  if (ev) DAN_CUDA(cudaEventRecord(ev[0], st));
  if (!fused) {
    DAN_CUDA(cudaMemsetAsync(A.key_count, 0, (size_t)lists * 4, st));
    if (num_anchors > 0) {
      pp_filter_kernel<<<dim3((num_anchors + 256 * kFilterPerThread - 1) / (256 * kFilterPerThread), batch), 256, 0, st>>>(A);
      DAN_LAUNCH_CHECK("pp_filter_kernel");
    }
  }
  if (ev) DAN_CUDA(cudaEventRecord(ev[1], st));
  int rc = fused ? run_sort_nms<true, true>(A, lists, nullptr, nullptr, st) : run_sort_nms<true, false>(A, lists, nullptr, nullptr, st);
  if (ev && rc == DAN_OK) DAN_CUDA(cudaEventRecord(ev[2], st));
  return rc;
}

int dan_postprocess_batch(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred, const float* boxes_pred,
                          const float* a_ymin, const float* a_xmin, const float* a_ymax, const float* a_xmax, int32_t num_anchors,
                          int32_t batch, float* out_boxes, float* out_scores, int32_t* out_counts, int32_t* out_anchor_index,
                          int32_t* out_keep_pos, void* workspace, size_t workspace_bytes, void* stream) {
  return postprocess_core(p, cls_pred, loc_pred, boxes_pred, a_ymin, a_xmin, a_ymax, a_xmax, num_anchors, batch, out_boxes, out_scores,
                          out_counts, out_anchor_index, out_keep_pos, workspace, workspace_bytes, stream, nullptr);
}

int dan_postprocess_batch_peers(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred, const float* boxes_pred,
                                const float* a_ymin, const float* a_xmin, const float* a_ymax, const float* a_xmax, int32_t num_anchors,
                                int32_t batch, float* out_boxes, float* out_scores, int32_t* out_counts, int32_t* out_anchor_index,
                                int32_t* out_keep_pos, void* workspace, size_t workspace_bytes, const dan_peer_exchange* peers,
                                void* stream) {
  DAN_REQUIRE(batch > 0 || peers == nullptr || peers->num_destinations == 0, DAN_ERR_INVALID_ARGUMENT,
              "a rank without images cannot take part in the peer exchange (its flag would never rise)");
  return postprocess_core(p, cls_pred, loc_pred, boxes_pred, a_ymin, a_xmin, a_ymax, a_xmax, num_anchors, batch, out_boxes, out_scores,
                          out_counts, out_anchor_index, out_keep_pos, workspace, workspace_bytes, stream, nullptr, peers);
}

int dan_wait_detections(const int32_t* flags, const int32_t* state, int32_t world_size, void* stream) {
  DAN_REQUIRE(flags != nullptr && state != nullptr && world_size >= 1 && world_size <= 32, DAN_ERR_INVALID_ARGUMENT, "bad arguments");
  wait_detections_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, state, world_size);
  DAN_LAUNCH_CHECK("wait_detections_kernel");
  return DAN_OK;
}

int dan_peer_alloc(size_t bytes, void** out_ptr, void* out_handle64) {
  DAN_REQUIRE(bytes > 0 && out_ptr != nullptr && out_handle64 != nullptr, DAN_ERR_INVALID_ARGUMENT, "bad arguments");
  void* ptr = nullptr;
  DAN_CUDA(cudaMalloc(&ptr, bytes));
  cudaError_t e = cudaMemset(ptr, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    cudaFree(ptr);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out_handle64, &h, 64);
  *out_ptr = ptr;
  return DAN_OK;
}

int dan_peer_open(const void* handle64, void** out_ptr) {
  DAN_REQUIRE(handle64 != nullptr && out_ptr != nullptr, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DAN_CUDA(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DAN_OK;
}

int dan_peer_close(void* ptr) {
  if (ptr != nullptr) DAN_CUDA(cudaIpcCloseMemHandle(ptr));
  return DAN_OK;
}

int dan_peer_free(void* ptr) {
  if (ptr != nullptr) DAN_CUDA(cudaFree(ptr));
  return DAN_OK;
}

int dan_postprocess_batch_profile(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred,
                                  const float* boxes_pred, const float* a_ymin, const float* a_xmin, const float* a_ymax,
                                  const float* a_xmax, int32_t num_anchors, int32_t batch, float* out_boxes, float* out_scores,
                                  int32_t* out_counts, int32_t* out_anchor_index, int32_t* out_keep_pos, void* workspace,
                                  size_t workspace_bytes, void* stream, float* h_kernel_ms) {
  DAN_REQUIRE(h_kernel_ms != nullptr, DAN_ERR_INVALID_ARGUMENT, "h_kernel_ms is NULL");
  h_kernel_ms[0] = h_kernel_ms[1] = 0.f;
  cudaEvent_t ev[3];
  for (int i = 0; i < 3; ++i) DAN_CUDA(cudaEventCreate(&ev[i]));
  int rc = postprocess_core(p, cls_pred, loc_pred, boxes_pred, a_ymin, a_xmin, a_ymax, a_xmax, num_anchors, batch, out_boxes,
                            out_scores, out_counts, out_anchor_index, out_keep_pos, workspace, workspace_bytes, stream, ev);
  if (rc == DAN_OK && batch > 0) {
    cudaError_t e = cudaEventSynchronize(ev[2]);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize");
    else for (int i = 0; i < 2; ++i) cudaEventElapsedTime(&h_kernel_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

int dan_sort_bboxes(const float* scores, const float* boxes, int64_t n, int32_t keep_topk, float* out_scores, float* out_boxes,
                    int32_t* out_index, void* workspace, size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(n >= 0 && n < 0x7fffffff && keep_topk >= 1, DAN_ERR_INVALID_ARGUMENT, "bad size");
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1, keep_topk, false);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  DAN_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortCap * 8)));
  cudaStream_t st = (cudaStream_t)stream;
  PpArgs A = {};
  A.n = (int)n;
  A.num_classes = 2;
  A.keep_topk = keep_topk;
  pp_bind(A, workspace, w);
  A.s_scores = out_scores;
  A.s_boxes = reinterpret_cast<float4*>(out_boxes);
  A.s_index = out_index;
  key_build_kernel<<<grid_for(n), 256, 0, st>>>(scores, n, A.keys, A.key_count);
  DAN_LAUNCH_CHECK("key_build_kernel");
  topk_sort_kernel<<<1, kSortThreads, kSortCap * 8, st>>>(A, scores, reinterpret_cast<const float4*>(boxes));
  DAN_LAUNCH_CHECK("topk_sort_kernel");
  return DAN_OK;
}

int dan_nms_bboxes(const float* scores, const float* boxes, int64_t n, int32_t nms_topk, float nms_threshold, float* out_scores,
                   float* out_boxes, int32_t* out_keep, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(n >= 0 && nms_topk >= 1, DAN_ERR_INVALID_ARGUMENT, "bad size");
  DAN_REQUIRE(n < 0x7fffffff, DAN_ERR_INVALID_ARGUMENT, "bad size");
  const int n_eff = n > 0 ? (int)n : 1;
  const int cap = nms_topk < n_eff ? nms_topk : n_eff;
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1, n_eff, true);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  PpArgs A = {};
  A.n = (int)n;
  A.num_classes = 2;
  A.keep_topk = n_eff;
  A.nms_topk = nms_topk;
  A.nms_cap = cap;
  A.nms_thr = nms_threshold;
  A.out_boxes = reinterpret_cast<float4*>(out_boxes);
  A.out_scores = out_scores;
  A.out_counts = out_count;
  A.out_index = nullptr;
  A.out_keep = out_keep;
  A.filler = 0;
  pp_bind(A, workspace, w);
  key_build_kernel<<<grid_for(n), 256, 0, st>>>(scores, n, A.keys, A.key_count);
  DAN_LAUNCH_CHECK("key_build_kernel");
  return run_sort_nms<false, false>(A, 1, scores, reinterpret_cast<const float4*>(boxes), st);
}

}  // extern "C"
