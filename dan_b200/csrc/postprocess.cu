// K3-K5: the evaluation half of the path -- utility/bbox_util.py:103-119
// parse_by_class and its pieces, batched over images and classes.
//
//   pp_filter_kernel   softmax -> select(threshold) -> decode -> clip -> min-size;
//                      survivors are COMPACTED as 64-bit keys
//                      (ordered score bits << 32 | ~anchor index).  The reference
//                      instead multiplies by 0/1 masks and keeps all N rows
//                      (bbox_util.py:24-59); rows it zeroes can never precede a
//                      positive score in tf.nn.top_k, so dropping them is exact.
//   topk_sort_kernel   sort_bboxes alone (bbox_util.py:61-72): block radix select of the keep_topk largest keys
//                      when more survive, then the bitonic sort of sort.cuh.  Keys are unique, descending key
//                      order == descending score with ties broken by LOWER index, which is tf.nn.top_k's order.
//   pp_sort_kernel     per (image, class): the same top-k + sort, then decode of the survivors and the broad phase of
//                      the NMS: boxes binned by size class (power of two of the longer side) and centre cell.
//   nms_pairs_kernel   narrow phase: tf.image.non_max_suppression's IoU test (no +1, corners min/max-normalised,
//                      strict >, area <= 0 never suppresses) only for boxes whose grid windows meet -> a sparse list
//                      of suppression EDGES (lower rank, higher rank).
//   nms_resolve_kernel greedy NMS as a relaxation over the edges (kept(i) = no kept lower-ranked neighbour; the depth
//                      of the dependency chains is a handful of sweeps), ordered compaction, zero padded outputs
//                      (bbox_util.py:80-90).  Lists with too many edges, a negative threshold or more candidates
//                      than the pair kernel can stage fall back to rounds of 64 candidates against the kept list.
#include <math.h>
#include <stdlib.h>

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
  int nms_cap;                // kept-list capacity in shared memory = min(nms_topk, keep_topk)
  float nms_thr;
  int force_rounds;           // 1: the pair kernel's staging does not fit shared memory -> round-based NMS for every list
  // workspace
  unsigned long long* keys;   // [L, n]  L = batch * (C-1) lists
  int32_t* key_count;         // [L]
  float* s_scores;            // sort_bboxes outputs [keep_topk]
  float4* s_boxes;
  int32_t* s_index;
  // per list, written by the sort kernel, read by the pair and resolve kernels (all L2 resident)
  unsigned long long* s_key;  // [L, keep_topk] sorted keys
  float4* s_box;              // [L, keep_topk] normalised corners (y0, x0, y1, x1), rank order
  float* s_area;              // [L, keep_topk] (y1-y0)*(x1-x0)
  int32_t* s_len;             // [L] number of sorted candidates K
  float4* grid_info;          // [L] (origin y, origin x, extent, bit mask of the non-empty size classes)
  float* class_amin;          // [L, 16] smallest box area per size class (IoU <= area ratio: whole classes are skipped)
  uint16_t* box_cell;         // [L, keep_topk] flattened (class, cy, cx) of each box, 0xffff = not binned
  uint16_t* box_pos;          // [L, keep_topk] position of each box inside cell_items
  uint16_t* cell_start;       // [L, kCellStride] (kTotalCells + 1 entries used; the stride keeps every list 16-byte aligned)
  uint16_t* cell_items;       // [L, keep_topk] ranks grouped by cell
  uint32_t* edges;            // [L, kEdgeCap] (hi << 16 | lo): lo suppresses hi when lo is kept
  int32_t* edge_n;            // [L]
  int32_t* ovf;               // [L] edge list overflow / negative threshold -> round-based fallback
  // outputs
  float4* out_boxes;
  float* out_scores;
  int32_t* out_counts;
  int32_t* out_index;
  int32_t* out_keep;
  int filler;                 // fused parse_by_class: zero-score rows fill up the NMS selection
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
        if (!have_box) { box = pp_box(A, b, a); have_box = true; }
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
  select_and_sort(A.keys, cnt, k, s_keys, sc);
  for (int r = tid; r < A.keep_topk; r += kSortThreads) {
    if (r < k) {
      const uint32_t idx = key_index(s_keys[r]);
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
// K4+K5: one CTA per (image, class) list does everything after the filter:
//   1. top-k select + sort of the surviving keys in shared memory (select_and_sort above);
//   2. decode + clip + normalise the K best boxes into shared memory;
//   3. broad phase: boxes are binned by size class (max side in [2^c, 2^(c+1))) and by the cell of their centre in a
//      per-class uniform grid whose cell is as large as the class's boxes, so that a box only has to look at <= 3x3
//      cells per class to find everything it can overlap (counting sort in shared memory);
//   4. narrow phase: tf.image.non_max_suppression's IoU test (no +1, corners min/max normalised, area <= 0 never
//      suppresses, strict >) on those few candidates; every pair (lo, hi) with lo ranked above hi and IoU > thr
//      becomes an edge "lo suppresses hi if lo is kept".  ~K*6 edges instead of K*K/2 tests;
//   5. greedy NMS == evaluation of the DAG  kept(i) = !any(kept(j) : edge j -> i)  in rank order.  It is evaluated by
//      parallel relaxation over the edges: a box is decided as soon as one suppressor is known kept (suppressed) or
//      all of them are known suppressed (kept).  The number of sweeps is the longest dependency chain (6 for the
//      benchmark detections), not K;
//   6. the first nms_topk kept boxes in rank order are written out, zero padded (bbox_util.py:80-90); this is what
//      TF's sequential loop selects because a decision never depends on lower ranked boxes.
// Fallback (more edges than fit, or a negative threshold where even disjoint boxes suppress): rounds of 64
// candidates tested against the kept list in shared memory, resolved serially per round (nms_rounds below).
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

// IOUGreaterThanThreshold of TF's non_max_suppression_op.cc for normalised boxes
DAN_D bool nms_suppresses(float4 a, float a_area, float4 b, float b_area, float thr) {
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
  if (thr < 0.f) return nms_suppresses(a, a_area, b, b_area, thr);
  const float h = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float w = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
  if (!(h > 0.f && w > 0.f)) return false;
  if (!(a_area > 0.f && b_area > 0.f)) return false;
  const float inter = fmul(h, w);
  const float uni = fsub(fadd(a_area, b_area), inter);
  const float t = fmul(thr, uni);
  if (t > 1e-30f && inter > fmul(t, 1.000001f)) return true;
  if (t > 1e-30f && inter < fmul(t, 0.999999f)) return false;
  return fdiv(inter, uni) > thr;
}

constexpr int kNmsClasses = 10;                                    // class 9: max side >= 512, one cell
constexpr int kGridDim = 32;                                       // cells per dimension and class
constexpr int kCellsPerClass = kGridDim * kGridDim;
constexpr int kTotalCells = kNmsClasses * kCellsPerClass;
constexpr int kCellStride = (kTotalCells + 1 + 7) / 8 * 8;      // uint16 entries per list, a multiple of 16 bytes
constexpr int kEdgeCap = 1 << 18;                                  // suppression edges per list (1 MB)

// geometry of the per-class grids of one list
struct GridGeom {
  float oy, ox, extent;
  // cell = half the class's largest side (a window of <= 6x6 cells then covers box + reach tightly), but never
  // more than kGridDim cells per dimension
  DAN_D float cell_size(int c) const { return fmaxf((float)(1 << c), extent * (1.f / (kGridDim - 1))); }
  DAN_D static int cell_of(float v, float org, float inv) {
    const int q = (int)((v - org) * inv);
    return min(max(q, 0), kGridDim - 1);
  }
};

// shared memory of the resolve kernel
struct NmsSmem {
  float4* kept_box;      // [nms_cap]   (fallback)
  float* kept_area;      // [nms_cap]   (fallback)
  float4* cand_box;      // [keep_topk] (fallback)
  float* cand_area;      // [keep_topk] (fallback)
  uint8_t* status;       // [keep_topk] 0 undecided, 1 kept, 2 suppressed
  uint8_t* pending;      // [keep_topk]
  int32_t* kept_pos;     // [nms_cap]
};

static size_t nms_smem_bytes(int nms_cap, int keep_topk) {
  return (size_t)nms_cap * 20 + (size_t)keep_topk * 20 + align_up((size_t)keep_topk, 16) * 2 + align_up((size_t)nms_cap * 4, 16);
}
constexpr size_t kNmsSmemMax = 227 * 1024 - 6 * 1024;   // dynamic part; a few KB of static shared memory on top
constexpr size_t kSortSmem = (size_t)kSortCap * 8 > (size_t)kTotalCells * 4 ? (size_t)kSortCap * 8 : (size_t)kTotalCells * 4;   // sort kernel: keys, then (aliased) the cell counters

DAN_D NmsSmem nms_carve(unsigned char* base, int nms_cap, int keep_topk) {
  NmsSmem m;
  unsigned char* p = base;
  m.kept_box = reinterpret_cast<float4*>(p); p += (size_t)nms_cap * 16;
  m.cand_box = reinterpret_cast<float4*>(p); p += (size_t)keep_topk * 16;
  m.kept_area = reinterpret_cast<float*>(p); p += (size_t)nms_cap * 4;
  m.cand_area = reinterpret_cast<float*>(p); p += (size_t)keep_topk * 4;
  m.status = p; p += ((size_t)keep_topk + 15) / 16 * 16;
  m.pending = p; p += ((size_t)keep_topk + 15) / 16 * 16;
  m.kept_pos = reinterpret_cast<int32_t*>(p);
  return m;
}

// ---- fallback: rounds of 64 candidates against the kept list.  In round c:
//   S1  all warps      suppression bits among the candidates of round c (their flags vs the kept list are final)
//   S2  warp 0         serial resolve of round c -> appends nk boxes to the kept list
//       warps 1..30    candidates of round c+1 vs the kept list as it was BEFORE round c
//   S3  all warps      candidates of round c+1 vs the nk boxes round c just appended
// Returns the number of kept boxes; their positions are in m.kept_pos.
DAN_D int nms_rounds(const NmsSmem& m, int K, int nms_topk, float thr) {
  __shared__ int s_flag[2][64];
  __shared__ unsigned long long s_rows[64];
  __shared__ int s_new_n;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nchunks = (K + 63) >> 6;
  float4* kept_box = m.kept_box;
  float* kept_area = m.kept_area;
  const float4* cand_box = m.cand_box;
  const float* cand_area = m.cand_area;
  if (tid < 128) (&s_flag[0][0])[tid] = 0;
  __syncthreads();
  int kept_n = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int base = c << 6;
    const int nvalid = min(64, K - base);
    const int nvalid_next = max(0, min(64, K - base - 64));
    const float4* cur_box = cand_box + base;
    const float* cur_area = cand_area + base;
    const float4* nxt_box = cand_box + base + 64;
    const float* nxt_area = cand_area + base + 64;
    const int fb = c & 1, fb1 = fb ^ 1;
    // ---- S1: warp w -> rows 2w, 2w+1; lane -> cols lane, lane+32
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = 2 * warp + rr;
      const bool row_ok = (r < nvalid) && (s_flag[fb][r] == 0);
      const float4 rb = row_ok ? cur_box[r] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float ra = row_ok ? cur_area[r] : 0.f;
      const int c0 = lane, c1 = lane + 32;
      const bool t0 = row_ok && c0 > r && c0 < nvalid && nms_suppresses(rb, ra, cur_box[c0], cur_area[c0], thr);
      const bool t1 = row_ok && c1 > r && c1 < nvalid && nms_suppresses(rb, ra, cur_box[c1], cur_area[c1], thr);
      const unsigned lo = __ballot_sync(0xffffffffu, t0);
      const unsigned hi = __ballot_sync(0xffffffffu, t1);
      if (lane == 0) s_rows[r] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    // ---- S2
    if (warp == 0) {
      // greedy resolve of the round, 32-bit halves: bit i of cl/ch set <=> candidate i / 32+i is suppressed
      const unsigned long long d0 = s_rows[lane], d1 = s_rows[lane + 32];
      const unsigned d0lo = (unsigned)d0, d0hi = (unsigned)(d0 >> 32), d1hi = (unsigned)(d1 >> 32);
      const unsigned vlo = (nvalid >= 32) ? 0xffffffffu : ((1u << nvalid) - 1u);
      const unsigned vhi = (nvalid >= 64) ? 0xffffffffu : ((nvalid > 32) ? ((1u << (nvalid - 32)) - 1u) : 0u);
      unsigned cl = __ballot_sync(0xffffffffu, s_flag[fb][lane] != 0) | ~vlo;
      unsigned ch = __ballot_sync(0xffffffffu, s_flag[fb][lane + 32] != 0) | ~vhi;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned rl = __shfl_sync(0xffffffffu, d0lo, i);
        const unsigned rh = __shfl_sync(0xffffffffu, d0hi, i);
        const unsigned alive = ((cl >> i) & 1u) - 1u;       // all ones when candidate i survives
        cl |= rl & alive;
        ch |= rh & alive;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned rh = __shfl_sync(0xffffffffu, d1hi, i);
        const unsigned alive = ((ch >> i) & 1u) - 1u;
        ch |= rh & alive;
      }
      unsigned long long kept = (((unsigned long long)(~ch & vhi)) << 32) | (unsigned long long)(~cl & vlo);
      int nk = __popcll(kept);
      if (kept_n + nk > nms_topk) {             // max_output_size reached inside the round
        int drop = kept_n + nk - nms_topk;
        while (drop-- > 0) kept &= ~(1ull << (63 - __clzll((long long)kept)));
        nk = nms_topk - kept_n;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if ((kept >> i) & 1ull) {
          const int pos = kept_n + __popcll(kept & ((1ull << i) - 1ull));
          kept_box[pos] = cur_box[i];
          kept_area[pos] = cur_area[i];
          m.kept_pos[pos] = base + i;
        }
      }
      if (lane == 0) s_new_n = nk;
      s_flag[fb][lane] = 0;          // this flag buffer is reused by round c+2
      s_flag[fb][lane + 32] = 0;
    } else if (warp < 31) {
      // round c+1 vs kept[0, kept_n): warp w in 1..30 -> candidates 32*((w-1)&1)+lane, slice (w-1)>>1 of 15.
      // The loop is kept WARP-UNIFORM (uniform trip count, structured ifs): a per-lane continue/break would let
      // the lanes drift apart for the rest of the loop and multiply the issued instructions.
      const int i = (((warp - 1) & 1) << 5) | lane;
      const bool have = i < nvalid_next;
      const float4 me = have ? nxt_box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float my_area = have ? nxt_area[i] : 0.f;
      bool done = !(have && my_area > 0.f);
      bool sup = false;
      for (int k = kept_n - 1 - ((warp - 1) >> 1); k >= 0; k -= 60) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int kk = k - 15 * u;
          if (kk >= 0 && !done && pair_suppresses(kept_box[kk], kept_area[kk], me, my_area, thr)) { sup = true; done = true; }
        }
        if (__all_sync(0xffffffffu, done)) break;
      }
      if (sup) s_flag[fb1][i] = 1;
    }
    __syncthreads();
    // ---- S3: round c+1 vs the boxes appended by round c: thread -> candidate tid&63, new boxes (tid>>6)+16j
    const int nk = s_new_n;
    if (nvalid_next > 0 && nk > 0) {
      const int i = tid & 63;
      const bool have = i < nvalid_next;
      const float4 me = have ? nxt_box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float my_area = have ? nxt_area[i] : 0.f;
      if (have && my_area > 0.f) {
        bool sup = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = (tid >> 6) + 16 * u;
          if (j < nk && !sup && pair_suppresses(kept_box[kept_n + j], kept_area[kept_n + j], me, my_area, thr)) sup = true;
        }
        if (sup) s_flag[fb1][i] = 1;
      }
    }
    kept_n += nk;
    __syncthreads();
    if (kept_n >= nms_topk) break;
  }
  return kept_n;
}

// block-wide min / max of a float over all threads (every thread gets the result)
DAN_D void block_minmax(float lo, float hi, float& out_lo, float& out_hi) {
  __shared__ int s_lo[kSortThreads / 32], s_hi[kSortThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wl = __reduce_min_sync(0xffffffffu, float_to_ordered(lo));
  const int wh = __reduce_max_sync(0xffffffffu, float_to_ordered(hi));
  if (lane == 0) { s_lo[warp] = wl; s_hi[warp] = wh; }
  __syncthreads();
  int a = s_lo[lane], b = s_hi[lane];
  a = __reduce_min_sync(0xffffffffu, a);
  b = __reduce_max_sync(0xffffffffu, b);
  out_lo = ordered_to_float(a);
  out_hi = ordered_to_float(b);
  __syncthreads();
}

// ---- kernel S: one CTA per list: top-k + sort, decode, broad-phase grid (steps 1-3) -> HBM (L2 resident)
template <bool DECODE>
__global__ void __launch_bounds__(kSortThreads, 1) pp_sort_kernel(const PpArgs A, const float4* __restrict__ src_boxes) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dyn_smem);
  int* cell_cnt = reinterpret_cast<int*>(dyn_smem);            // aliases the keys once they are in HBM
  __shared__ SortScratch sc;
  __shared__ int s_scan[kSortThreads / 32];
  __shared__ int s_class_mask;
  __shared__ int s_class_amin[kNmsClasses];

  const int list = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int b = list / max(A.num_classes - 1, 1);
  const int cnt = min(A.key_count[list], A.n);
  const int64_t o = (int64_t)list * A.keep_topk;

  DAN_PHASE(0);
  const int K = min(select_and_sort(A.keys + (int64_t)list * A.n, cnt, min(A.keep_topk, cnt), keys, sc), A.keep_topk);
  DAN_PHASE(1);
  float ylo = 3.0e38f, yhi = -3.0e38f, xlo = 3.0e38f, xhi = -3.0e38f;
  for (int r = tid; r < K; r += kSortThreads) {
    const unsigned long long key = keys[r];
    const uint32_t idx = key_index(key);
    A.s_key[o + r] = key;
    const NmsBox nb = nms_norm(DECODE ? pp_box(A, b, (int)idx) : src_boxes[idx]);
    A.s_box[o + r] = make_float4(nb.y0, nb.x0, nb.y1, nb.x1);
    A.s_area[o + r] = nb.area;
    if (nb.area > 0.f) {
      ylo = fminf(ylo, nb.y0); yhi = fmaxf(yhi, nb.y1);
      xlo = fminf(xlo, nb.x0); xhi = fmaxf(xhi, nb.x1);
    }
  }
  if (tid == 0) s_class_mask = 0;
  if (tid < kNmsClasses) s_class_amin[tid] = 0x7f7fffff;     // FLT_MAX as ordered int (areas are positive)
  GridGeom g;
  float ey, ex;
  block_minmax(ylo, yhi, g.oy, ey);      // (contains barriers: the keys are dead from here on)
  block_minmax(xlo, xhi, g.ox, ex);
  g.extent = fmaxf(fmaxf(ey - g.oy, ex - g.ox), 1.f);
  DAN_PHASE(2);

  for (int i = tid; i < kTotalCells; i += kSortThreads) cell_cnt[i] = 0;
  __syncthreads();
  uint16_t* box_cell = A.box_cell + o;
  for (int i = tid; i < K; i += kSortThreads) {
    const float4 bx = A.s_box[o + i];
    int cid = 0xffff;
    if (A.s_area[o + i] > 0.f) {
      const float side = fmaxf(bx.z - bx.x, bx.w - bx.y);
      const int c = (side < 1.f) ? 0 : min(kNmsClasses - 1, (int)((__float_as_uint(side) >> 23) & 255u) - 127);
      if (c == kNmsClasses - 1) {
        cid = c * kCellsPerClass;
      } else {
        const float inv = 1.f / g.cell_size(c);
        cid = c * kCellsPerClass + GridGeom::cell_of(0.5f * (bx.x + bx.z), g.oy, inv) * kGridDim +
              GridGeom::cell_of(0.5f * (bx.y + bx.w), g.ox, inv);
      }
      atomicAdd(&cell_cnt[cid], 1);
      atomicOr(&s_class_mask, 1 << c);
      atomicMin(&s_class_amin[c], __float_as_int(A.s_area[o + i]));
    }
    box_cell[i] = (uint16_t)cid;
  }
  __syncthreads();
  DAN_PHASE(3);
  uint16_t* cell_start = A.cell_start + (int64_t)list * kCellStride;
  {  // exclusive scan of the cell counters: consecutive cells per thread
    constexpr int per = (kTotalCells + kSortThreads - 1) / kSortThreads;
    int local[per];
    int sum = 0;
#pragma unroll
    for (int e = 0; e < per; ++e) {
      const int cidx = tid * per + e;
      local[e] = (cidx < kTotalCells) ? cell_cnt[cidx] : 0;
      sum += local[e];
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_scan[w];
    int run = before + incl - sum;
#pragma unroll
    for (int e = 0; e < per; ++e) {
      const int cidx = tid * per + e;
      if (cidx < kTotalCells) {
        cell_cnt[cidx] = run;          // start of the cell; becomes the scatter cursor below
        run += local[e];
      }
    }
    if (tid == kSortThreads - 1) cell_start[kTotalCells] = (uint16_t)run;
  }
  __syncthreads();
  for (int i = tid; i < kTotalCells; i += kSortThreads) cell_start[i] = (uint16_t)cell_cnt[i];   // coalesced copy to HBM
  __syncthreads();
  uint16_t* cell_items = A.cell_items + o;
  for (int i = tid; i < K; i += kSortThreads) {
    const int cid = box_cell[i];
    if (cid != 0xffff) {
      const int pos = atomicAdd(&cell_cnt[cid], 1);
      cell_items[pos] = (uint16_t)i;
      A.box_pos[o + i] = (uint16_t)pos;
    }
  }
  DAN_PHASE(4);
  // the pair kernel stages these arrays with 16-byte copies: define the few entries between K and the next multiple of 8,
  // and the padding of the cell table
  for (int i = K + tid; i < min(A.keep_topk, (K + 7) & ~7); i += kSortThreads) {
    A.s_area[o + i] = 0.f;
    box_cell[i] = 0xffff;
    A.box_pos[o + i] = 0;
    cell_items[i] = 0;
  }
  for (int i = kTotalCells + 1 + tid; i < kCellStride; i += kSortThreads) cell_start[i] = 0;
  if (tid == 0) {
    A.s_len[list] = K;
    A.grid_info[list] = make_float4(g.oy, g.ox, g.extent, __int_as_float(s_class_mask));
    A.edge_n[list] = 0;
    A.ovf[list] = (A.nms_thr < 0.f || A.force_rounds) ? 1 : 0;     // thr < 0: disjoint boxes suppress too, no spatial pruning
  }
  if (tid < kNmsClasses) A.class_amin[list * 16 + tid] = __int_as_float(s_class_amin[tid]);
}

// ---- kernel P: narrow phase (step 4), up to kPairCtas CTAs per list.  Every CTA stages the list's boxes and grid in
// shared memory (the searches are chains of dependent lookups: ~30 cycles there instead of an L2 round trip) and its
// warps take the boxes round-robin.  A box j of class d that overlaps box i has its centre within 2^d (half its largest
// possible side) of i (+1 px and 1e-6 relative for fp32 rounding), i.e. in a window of at most 4x4 cells of class d's
// grid (more when the grid had to be coarsened).  Box i searches the classes d > class(i) in full and only the half
// of its own class's window that follows it (each same-class pair is met exactly once).  Edges are collected in
// warp-private shared-memory buffers and appended to the list's edge array with one atomic per flush.
constexpr int kPairCtas = 16;         // upper bound of CTAs per list
constexpr int kPairEdgeBuf = 8192;

static size_t pairs_smem_bytes(int keep_topk) {
  return (size_t)keep_topk * 16 + align_up((size_t)keep_topk * 4, 16) + align_up((size_t)keep_topk * 2, 16) * 3 +
         (size_t)kCellStride * 2 + (size_t)kPairEdgeBuf * 4;
}

__global__ void __launch_bounds__(kSortThreads, 1) nms_pairs_kernel(const PpArgs A) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ float s_prune[kNmsClasses], s_inv[kNmsClasses];
  const int list = blockIdx.y;
  if (A.ovf[list] != 0) return;
  const int K = A.s_len[list];
  // long lists get all the CTAs of their grid row, short ones only a few (the others exit at once): the launch time
  // is set by the longest list
  // (one CTA per ~320 boxes: every CTA stages the whole list, so more CTAs mostly add staging work; measured at 32 lists
  // of ~2000 boxes: 160 boxes per CTA 45 us / 403 k img/s, 320 boxes per CTA 43 us / 427 k img/s)
  const int my_ctas = min((int)gridDim.x, max(1, K / 320));
  if ((int)blockIdx.x >= my_ctas) return;
  DAN_PHASE(24);
  const int tid = threadIdx.x;
  const int64_t o = (int64_t)list * A.keep_topk;
  unsigned char* p = dyn_smem;
  float4* box = reinterpret_cast<float4*>(p); p += (size_t)A.keep_topk * 16;
  float* area = reinterpret_cast<float*>(p); p += ((size_t)A.keep_topk * 4 + 15) / 16 * 16;
  uint16_t* box_cell = reinterpret_cast<uint16_t*>(p); p += ((size_t)A.keep_topk * 2 + 15) / 16 * 16;
  uint16_t* box_pos = reinterpret_cast<uint16_t*>(p); p += ((size_t)A.keep_topk * 2 + 15) / 16 * 16;
  uint16_t* cell_items = reinterpret_cast<uint16_t*>(p); p += ((size_t)A.keep_topk * 2 + 15) / 16 * 16;
  uint16_t* cell_start = reinterpret_cast<uint16_t*>(p); p += (size_t)kCellStride * 2;
  uint32_t* ebuf = reinterpret_cast<uint32_t*>(p);

  if ((A.keep_topk & 7) == 0) {
    // every list starts 16-byte aligned: 16-byte copies (reading up to the next multiple of 8 <= keep_topk entries)
    for (int i = tid; i < K; i += kSortThreads) box[i] = A.s_box[o + i];
    const int n4 = (K + 3) / 4, n8 = (K + 7) / 8;
    for (int i = tid; i < n4; i += kSortThreads)
      reinterpret_cast<float4*>(area)[i] = reinterpret_cast<const float4*>(A.s_area + o)[i];
    for (int i = tid; i < n8; i += kSortThreads) {
      reinterpret_cast<uint4*>(box_cell)[i] = reinterpret_cast<const uint4*>(A.box_cell + o)[i];
      reinterpret_cast<uint4*>(box_pos)[i] = reinterpret_cast<const uint4*>(A.box_pos + o)[i];
      reinterpret_cast<uint4*>(cell_items)[i] = reinterpret_cast<const uint4*>(A.cell_items + o)[i];
    }
  } else {
    for (int i = tid; i < K; i += kSortThreads) {
      box[i] = A.s_box[o + i];
      area[i] = A.s_area[o + i];
      box_cell[i] = A.box_cell[o + i];
      box_pos[i] = A.box_pos[o + i];
      cell_items[i] = A.cell_items[o + i];
    }
  }
  {  // the grid: 16-byte copies (20 KB per CTA; two-byte loads made this the longest part of the staging)
    const uint4* g_start = reinterpret_cast<const uint4*>(A.cell_start + (int64_t)list * kCellStride);
    uint4* s_start = reinterpret_cast<uint4*>(cell_start);
    for (int i = tid; i < kCellStride / 8; i += kSortThreads) s_start[i] = g_start[i];
  }
  const float4 gi = A.grid_info[list];
  GridGeom g;
  g.oy = gi.x; g.ox = gi.y; g.extent = gi.z;
  if (tid < kNmsClasses) {
    s_prune[tid] = A.nms_thr * A.class_amin[list * 16 + tid] * 0.999f;   // area below which class `tid` cannot suppress
    s_inv[tid] = 1.f / g.cell_size(tid);
  }
  __syncthreads();
  DAN_PHASE(25);

  const int class_mask = __float_as_int(gi.w);
  uint32_t* edges = A.edges + (int64_t)list * kEdgeCap;
  const float thr = A.nms_thr;
  // warp-cooperative search: a half-warp takes one box i at a time and walks its size classes d >= class(i); the item
  // ranges of the grid rows of the class-d windows form one flat candidate list that its 16 lanes test 16 at a time.
  // (A thread-per-query loop is SIMT-hostile here: the windows hold anything from 0 to hundreds of boxes.)
  const int lane = tid & 31;
  const int warps_total = my_ctas * (kSortThreads / 32);
  // every warp collects its edges in a private slice of shared memory (no atomics in the search loop) and appends
  // the slice to the list's edge array with one atomic when it is full and at the end
  constexpr int kWarpEdgeBuf = kPairEdgeBuf / (kSortThreads / 32);
  uint32_t* wbuf = ebuf + (tid >> 5) * kWarpEdgeBuf;
  int wcount = 0;                                            // warp-uniform
  auto flush_warp = [&]() {
    __syncwarp();                                            // the buffer entries were written by other lanes
    int base = 0;
    if (lane == 0) base = atomicAdd(A.edge_n + list, wcount);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + wcount > kEdgeCap) {
      if (lane == 0) A.ovf[list] = 1;
    } else {
      for (int e = lane; e < wcount; e += 32) edges[base + e] = wbuf[e];
    }
    __syncwarp();
    wcount = 0;
  };
#ifdef DAN_PHASE_TIMING
  long long acc_setup = 0, acc_loop = 0, acc_n = 0, acc_boxes = 0, t_a = 0;
#define DAN_TICK() (t_a = clock64())
#define DAN_TOCK(acc) (acc += clock64() - t_a)
#else
#define DAN_TICK() do { } while (0)
#define DAN_TOCK(acc) do { } while (0)
#endif
  // TWO boxes per warp, one per half-warp (consecutive ranks): most boxes have fewer than 16 candidates and search two
  // size classes, so a whole warp per box left half of its lanes idle and paid the window set-up once per box.
  const int half = lane >> 4, hl = lane & 15;
  for (int i0 = 2 * (blockIdx.x * (kSortThreads / 32) + (tid >> 5)); i0 < K; i0 += 2 * warps_total) {
    const int i = i0 + half;
    const int my_cid = (i < K) ? box_cell[i] : 0xffff;
    const bool live_box = my_cid != 0xffff;
    if (!__any_sync(0xffffffffu, live_box)) continue;      // warp-uniform
    const int c = live_box ? my_cid / kCellsPerClass : 0;
    const int own_row = (my_cid - c * kCellsPerClass) / kGridDim;    // grid row of the box's own cell (0 for the last class)
    const int after_me = live_box ? box_pos[i] + 1 : 0;
    const float4 me = live_box ? box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float my_area = live_box ? area[i] : 0.f;
    auto emit = [&](bool edge, int j) {       // append the edges found by this step to the warp's private buffer
      const unsigned em = __ballot_sync(0xffffffffu, edge);
      if (em != 0u) {
        if (wcount + 32 > kWarpEdgeBuf) flush_warp();
        if (edge) wbuf[wcount + __popc(em & ((1u << lane) - 1u))] = ((uint32_t)max(i, j) << 16) | (uint32_t)min(i, j);
        wcount += __popc(em);
      }
    };
    // classes to search (lane hl of a half-warp speaks for class hl): non-empty, d >= c, and not ruled out by the area
    // ratio (IoU <= area_i / area_j: a class whose smallest box is already too large for the threshold cannot
    // suppress i)
    const bool want = live_box && hl >= c && hl < kNmsClasses && ((class_mask >> hl) & 1) &&
                      (hl == c || !(my_area < s_prune[hl < kNmsClasses ? hl : 0]));
    unsigned todo = (__ballot_sync(0xffffffffu, want) >> (16 * half)) & 0xffffu;     // per half-warp
    while (__any_sync(0xffffffffu, todo != 0u)) {          // up to 2 classes per pass and box
      DAN_TICK();
      // lane = (class slot, window row): each lane finds the item range of ONE row of ONE class's window, so the
      // window geometry of the classes is computed in parallel and all candidates of a box form one flat list
      const int slot = hl >> 3, r = hl & 7;
      unsigned rest = todo;                                // (slot+1)-th set bit of todo (no __fns: it is a software loop)
      int dsel = -1;
#pragma unroll
      for (int sidx = 0; sidx < 2; ++sidx) {
        const int dd = rest ? (__ffs(rest) - 1) : -1;
        if (slot == sidx) dsel = dd;
        rest &= rest - 1u;
      }
      const bool has = dsel >= 0;
      const int d = has ? dsel : 0;
      int cy0 = 0, cy1 = 0, cx0 = 0, cx1 = 0;
      if (has && d < kNmsClasses - 1) {
        const float inv = s_inv[d];
        const float reach = (float)(1 << d) + 1.f;
        cy0 = GridGeom::cell_of(me.x - reach - 1e-6f * fabsf(me.x), g.oy, inv);
        cy1 = GridGeom::cell_of(me.z + reach + 1e-6f * fabsf(me.z), g.oy, inv);
        cx0 = GridGeom::cell_of(me.y - reach - 1e-6f * fabsf(me.y), g.ox, inv);
        cx1 = GridGeom::cell_of(me.w + reach + 1e-6f * fabsf(me.w), g.ox, inv);
      }
      // a window taller than 8 rows (possible only when the grid had to be coarsened) is walked in row blocks of 8
      const int nrows = cy1 - cy0 + 1;
      const int nblocks = __reduce_max_sync(0xffffffffu, has ? (nrows + 7) >> 3 : 0);
      for (int rb = 0; rb < nblocks; ++rb) {
        const int row_in_window = rb * 8 + r;
        int p0 = 0, len = 0;
        if (has && row_in_window < nrows) {
          const int row = d * kCellsPerClass + (cy0 + row_in_window) * kGridDim;
          p0 = cell_start[row + cx0];
          // Own class: both boxes of an overlapping pair have each other in their windows, so only the HALF window
          // after the box is searched: the rest of its own cell, the cells to the right in its own row, and the
          // rows below.  (The items of a grid row are contiguous, and so are those of a cell.)
          if (d == c) {
            if (cy0 + row_in_window < own_row) p0 = 0x7fffffff;
            else if (cy0 + row_in_window == own_row) p0 = after_me;
          }
          len = max((int)cell_start[row + cx1 + 1] - p0, 0);
        }
        int incl = len;                                  // inclusive prefix over the 16 (class, row) ranges of the half-warp
#pragma unroll
        for (int sh = 1; sh < 16; sh <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, sh, 16);
          if (hl >= sh) incl += v;
        }
        const int n = __shfl_sync(0xffffffffu, incl, 15, 16);          // candidates of this half-warp's box
        const int n_max = __reduce_max_sync(0xffffffffu, n);
        const int shift = p0 - (incl - len);             // item position = q + shift inside this lane's range
        DAN_TOCK(acc_setup);
        DAN_TICK();
#ifdef DAN_PHASE_TIMING
        acc_n += n; acc_boxes += 1;
#endif
        for (int q0 = 0; q0 < n_max; q0 += 16) {          // warp-uniform trip count
          const int q = q0 + hl;
          int lo = 0;                                      // first range whose inclusive prefix exceeds q
#pragma unroll
          for (int st = 8; st >= 1; st >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, lo + st - 1, 16);
            if (q >= v) lo += st;
          }
          lo = min(lo, 15);
          const int pos = q + __shfl_sync(0xffffffffu, shift, lo, 16);
          bool edge = false;
          int j = 0;
          if (q < n) {
            j = cell_items[pos];
            edge = pair_suppresses(box[j], area[j], me, my_area, thr);
          }
          emit(edge, j);
        }
        DAN_TOCK(acc_loop);
        DAN_TICK();
      }
      todo = rest;                                         // the (up to) 2 classes of this pass are done
    }
  }
  DAN_PHASE(26);
#ifdef DAN_PHASE_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { g_phase[28] = acc_setup; g_phase[29] = acc_loop; g_phase[30] = acc_n; g_phase[31] = acc_boxes; }
#endif
  if (wcount > 0) flush_warp();
  DAN_PHASE(27);
}

// ---- kernel R: one CTA per list: relaxation over the edges (step 5) and the outputs (step 6)
template <bool DECODE>
__global__ void __launch_bounds__(kSortThreads, 1) nms_resolve_kernel(const PpArgs A, const float* __restrict__ src_scores,
                                                                      const float4* __restrict__ src_boxes) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const NmsSmem m = nms_carve(dyn_smem, A.nms_cap, A.keep_topk);
  __shared__ int s_scan[kSortThreads / 32];

  const int list = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int K = A.s_len[list];
  const int64_t o = (int64_t)list * A.keep_topk;
  int kept_n = 0;

  DAN_PHASE(16);
  if (A.ovf[list] == 0) {
    const int n_edges = min(A.edge_n[list], kEdgeCap);
    const uint32_t* edges = A.edges + (int64_t)list * kEdgeCap;
    // the relaxation sweeps the edge list several times: keep it in shared memory when it fits (the fallback's
    // candidate array is unused on this path)
    const int smem_edges = A.keep_topk * 4;          // 16 B per candidate slot
    if (n_edges <= smem_edges) {
      uint32_t* se = reinterpret_cast<uint32_t*>(m.cand_box);
      // 16-byte copies (the list's edge array is 16-byte aligned and kEdgeCap a multiple of 4)
      for (int e = tid; e < (n_edges + 3) / 4; e += kSortThreads) {
        if (4 * e + 3 < n_edges) {
          reinterpret_cast<uint4*>(se)[e] = reinterpret_cast<const uint4*>(edges)[e];
        } else {                                          // the last, partial vector: nothing beyond the list is read
          for (int q = 4 * e; q < n_edges; ++q) se[q] = edges[q];
        }
      }
      edges = se;
    }
    for (int i = tid; i < K; i += kSortThreads) { m.status[i] = 0; m.pending[i] = 0; }
    __syncthreads();
#ifdef DAN_PHASE_TIMING
    int dbg_rounds = 0;
#endif
    // One sweep of the relaxation reads status[] while other threads set entries to 2 (suppressed) in the same sweep:
    // an INTENDED race (compute-sanitizer racecheck reports it).  status only moves 0 -> 2 inside a sweep and 0 -> 1
    // between sweeps (behind the barriers); a stale 0 merely postpones a decision to the next sweep, and the fixed
    // point - the greedy NMS result - does not depend on the interleaving.  Byte stores do not tear.
    while (true) {
#ifdef DAN_PHASE_TIMING
      ++dbg_rounds;
#endif
      for (int e0 = tid; e0 < n_edges; e0 += 4 * kSortThreads) {
        uint32_t ed[4];
        int sh[4], sl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                    // independent lookups in flight
          const int e = e0 + u * kSortThreads;
          ed[u] = (e < n_edges) ? edges[e] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = ed[u] != 0xffffffffu;
          sh[u] = ok ? m.status[ed[u] >> 16] : 1;
          sl[u] = ok ? m.status[ed[u] & 0xffffu] : 2;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (sh[u] == 0) {
            if (sl[u] == 1) m.status[ed[u] >> 16] = 2;
            else if (sl[u] == 0) m.pending[ed[u] >> 16] = 1;
          }
        }
      }
      __syncthreads();
      bool any = false;
      for (int i = tid; i < K; i += kSortThreads) {
        if (m.status[i] == 0) {
          if (m.pending[i] == 0) m.status[i] = 1;
          else { any = true; m.pending[i] = 0; }         // (cleared here for the next sweep: one barrier less per sweep)
        }
      }
      if (!__syncthreads_or(any ? 1 : 0)) break;
    }
    DAN_PHASE(17);
#ifdef DAN_PHASE_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0) { g_phase[20] = n_edges; g_phase[21] = dbg_rounds; }
#endif
    // ordered compaction of the kept boxes: thread t owns ranks [t*E, (t+1)*E)
    const int E = (K + kSortThreads - 1) / kSortThreads;
    int mine_cnt = 0;
    for (int e = 0; e < E; ++e) {
      const int i = tid * E + e;
      if (i < K && m.status[i] == 1) ++mine_cnt;
    }
    int incl = mine_cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) {
      if (w < warp) before += s_scan[w];
      total += s_scan[w];
    }
    int pos = before + incl - mine_cnt;
    for (int e = 0; e < E; ++e) {
      const int i = tid * E + e;
      if (i < K && m.status[i] == 1) {
        if (pos < A.nms_topk) m.kept_pos[pos] = i;
        ++pos;
      }
    }
    kept_n = min(total, A.nms_topk);
    __syncthreads();
  } else {
    for (int r = tid; r < K; r += kSortThreads) {
      m.cand_box[r] = A.s_box[o + r];
      m.cand_area[r] = A.s_area[o + r];
    }
    __syncthreads();
    kept_n = nms_rounds(m, K, A.nms_topk, A.nms_thr);
  }

  DAN_PHASE(18);
  // ---- outputs, zero padded to nms_topk
  for (int t = tid; t < A.nms_topk; t += kSortThreads) {
    const int64_t oo = (int64_t)list * A.nms_topk + t;
    if (t < kept_n) {
      const int pos = m.kept_pos[t];
      const unsigned long long key = A.s_key[o + pos];
      const uint32_t idx = key_index(key);
      A.out_scores[oo] = DECODE ? key_to_score((uint32_t)(key >> 32)) : src_scores[idx];
      A.out_boxes[oo] = DECODE ? A.s_box[o + pos] : src_boxes[idx];     // clipped boxes are already min/max ordered
      if (A.out_index != nullptr) A.out_index[oo] = (int32_t)idx;
      if (A.out_keep != nullptr) A.out_keep[oo] = A.filler ? pos : (int32_t)idx;
    } else {
      A.out_scores[oo] = 0.f;
      A.out_boxes[oo] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A.out_index != nullptr) A.out_index[oo] = -1;
      if (A.out_keep != nullptr) {
        // parse_by_class runs NMS on the zero padded top-k list: zero-area filler rows are never suppressed
        // and get selected until nms_topk is reached
        const int fpos = K + (t - kept_n);
        A.out_keep[oo] = (A.filler && fpos < A.keep_topk) ? fpos : -1;
      }
    }
  }
  if (tid == 0 && A.out_counts != nullptr) A.out_counts[list] = kept_n;
  DAN_PHASE(19);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

struct PpLayout {
  size_t key_count, s_len, edge_n, ovf, grid_info, class_amin, keys, s_key, s_box, s_area, box_cell, box_pos, cell_start, cell_items, edges, total;
};

static PpLayout pp_layout(int64_t n, int64_t lists, int64_t keep_topk, bool nms) {
  PpLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
  w.key_count = take(lists * 4);
  w.keys = take(lists * n * 8);
  w.s_len = take(nms ? lists * 4 : 0);
  w.edge_n = take(nms ? lists * 4 : 0);
  w.ovf = take(nms ? lists * 4 : 0);
  w.grid_info = take(nms ? lists * 16 : 0);
  w.class_amin = take(nms ? lists * 64 : 0);
  w.s_key = take(nms ? lists * keep_topk * 8 : 0);
  w.s_box = take(nms ? lists * keep_topk * 16 : 0);
  w.s_area = take(nms ? lists * keep_topk * 4 : 0);
  w.box_cell = take(nms ? lists * keep_topk * 2 : 0);
  w.box_pos = take(nms ? lists * keep_topk * 2 : 0);
  w.cell_start = take(nms ? lists * (size_t)kCellStride * 2 : 0);
  w.cell_items = take(nms ? lists * keep_topk * 2 : 0);
  w.edges = take(nms ? lists * (size_t)kEdgeCap * 4 : 0);
  w.total = off;
  return w;
}

static void pp_bind(PpArgs& A, void* ws, const PpLayout& w) {
  char* base = static_cast<char*>(ws);
  A.key_count = reinterpret_cast<int32_t*>(base + w.key_count);
  A.keys = reinterpret_cast<unsigned long long*>(base + w.keys);
  A.s_len = reinterpret_cast<int32_t*>(base + w.s_len);
  A.edge_n = reinterpret_cast<int32_t*>(base + w.edge_n);
  A.ovf = reinterpret_cast<int32_t*>(base + w.ovf);
  A.grid_info = reinterpret_cast<float4*>(base + w.grid_info);
  A.class_amin = reinterpret_cast<float*>(base + w.class_amin);
  A.s_key = reinterpret_cast<unsigned long long*>(base + w.s_key);
  A.s_box = reinterpret_cast<float4*>(base + w.s_box);
  A.s_area = reinterpret_cast<float*>(base + w.s_area);
  A.box_cell = reinterpret_cast<uint16_t*>(base + w.box_cell);
  A.box_pos = reinterpret_cast<uint16_t*>(base + w.box_pos);
  A.cell_start = reinterpret_cast<uint16_t*>(base + w.cell_start);
  A.cell_items = reinterpret_cast<uint16_t*>(base + w.cell_items);
  A.edges = reinterpret_cast<uint32_t*>(base + w.edges);
}

static bool nms_fits(int nms_cap, int keep_topk) {
  return nms_smem_bytes(nms_cap, keep_topk) <= kNmsSmemMax;   // (the pair kernel is skipped when ITS staging does not fit)
}

static int enable_big_smem() {
  static bool done = false;
  if (!done) {
    DAN_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DAN_CUDA(cudaFuncSetAttribute(pp_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DAN_CUDA(cudaFuncSetAttribute(pp_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DAN_CUDA(cudaFuncSetAttribute(nms_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemMax));
    DAN_CUDA(cudaFuncSetAttribute(nms_resolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemMax));
    DAN_CUDA(cudaFuncSetAttribute(nms_resolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemMax));
    done = true;
  }
  return DAN_OK;
}

// sort+grid -> pairs -> resolve for `lists` lists whose keys are in the workspace; ev (optional): 3 events, one after
// each kernel
template <bool DECODE>
static int run_sort_nms(const PpArgs& A_in, int lists, const float* src_scores, const float4* src_boxes, cudaStream_t st,
                        cudaEvent_t* ev = nullptr) {
  PpArgs A = A_in;
  A.force_rounds = pairs_smem_bytes(A.keep_topk) > kNmsSmemMax ? 1 : 0;     // very long lists (> ~6 500 candidates)
  pp_sort_kernel<DECODE><<<lists, kSortThreads, kSortSmem, st>>>(A, src_boxes);
  DAN_LAUNCH_CHECK("pp_sort_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[0], st));
  // up to kPairCtas CTAs per list; a CTA exits at once when its list is short (see the kernel)
  if (!A.force_rounds) {
    nms_pairs_kernel<<<dim3(kPairCtas, lists), kSortThreads, pairs_smem_bytes(A.keep_topk), st>>>(A);
    DAN_LAUNCH_CHECK("nms_pairs_kernel");
  }
  if (ev) DAN_CUDA(cudaEventRecord(ev[1], st));
  nms_resolve_kernel<DECODE><<<lists, kSortThreads, nms_smem_bytes(A.nms_cap, A.keep_topk), st>>>(A, src_scores, src_boxes);
  DAN_LAUNCH_CHECK("nms_resolve_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[2], st));
  return DAN_OK;
}

}  // namespace dan

using namespace dan;

extern "C" {

#ifdef DAN_PHASE_TIMING
int dan_debug_phases(long long* h_out32) { return cudaMemcpyFromSymbol(h_out32, g_phase, sizeof(long long) * 32) == cudaSuccess ? 0 : -3; }
#endif

size_t dan_postprocess_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t num_classes, int32_t keep_topk) {
  if (num_anchors < 0 || batch < 0 || num_classes < 2 || keep_topk < 1) return 0;
  return pp_layout(num_anchors, (int64_t)batch * (num_classes - 1), keep_topk, true).total;
}

size_t dan_sort_workspace_bytes(int64_t n, int32_t keep_topk) {
  if (n < 0 || keep_topk < 1) return 0;
  return pp_layout(n, 1, 1, false).total;
}

size_t dan_nms_workspace_bytes(int64_t n, int32_t nms_topk) {
  if (n < 0 || nms_topk < 0) return 0;
  return pp_layout(n, 1, n > 0 ? n : 1, true).total;
}

static int postprocess_core(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred, const float* boxes_pred,
                            const float* a_ymin, const float* a_xmin, const float* a_ymax, const float* a_xmax, int32_t num_anchors,
                            int32_t batch, float* out_boxes, float* out_scores, int32_t* out_counts, int32_t* out_anchor_index,
                            int32_t* out_keep_pos, void* workspace, size_t workspace_bytes, void* stream, cudaEvent_t* ev) {
  DAN_REQUIRE(p != nullptr, DAN_ERR_INVALID_ARGUMENT, "params is NULL");
  DAN_REQUIRE(p->num_classes >= 2, DAN_ERR_INVALID_ARGUMENT, "num_classes must be >= 2 (class 0 is background), got %d", p->num_classes);
  DAN_REQUIRE(num_anchors >= 0 && batch >= 0, DAN_ERR_INVALID_ARGUMENT, "negative size");
  DAN_REQUIRE(batch <= 65535, DAN_ERR_UNSUPPORTED, "batch > 65535");
  DAN_REQUIRE(p->select_threshold >= 0.f, DAN_ERR_INVALID_ARGUMENT,
              "select_threshold must be >= 0 (a negative threshold would let zero-score rows carry boxes), got %g", p->select_threshold);
  DAN_REQUIRE(p->keep_topk >= 1 && p->nms_topk >= 1, DAN_ERR_INVALID_ARGUMENT, "keep_topk and nms_topk must be >= 1");
  DAN_REQUIRE(p->keep_topk <= kSortCap, DAN_ERR_UNSUPPORTED, "keep_topk %d exceeds the in-shared-memory sort capacity %d", p->keep_topk, kSortCap);
  DAN_REQUIRE(nms_fits(p->nms_topk < p->keep_topk ? p->nms_topk : p->keep_topk, p->keep_topk), DAN_ERR_UNSUPPORTED,
              "keep_topk %d / nms_topk %d do not fit the NMS kernel's shared memory (22 B per candidate + 24 B per kept box in "
              "%zu bytes)", p->keep_topk, p->nms_topk, kNmsSmemMax);
  DAN_REQUIRE((loc_pred != nullptr) != (boxes_pred != nullptr), DAN_ERR_INVALID_ARGUMENT, "exactly one of loc_pred / boxes_pred must be given");
  if (batch == 0) return DAN_OK;
  DAN_REQUIRE(cls_pred && out_boxes && out_scores, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(loc_pred == nullptr || (a_ymin && a_xmin && a_ymax && a_xmax), DAN_ERR_INVALID_ARGUMENT, "anchors needed to decode loc_pred");
  DAN_REQUIRE(aligned16(loc_pred) && aligned16(boxes_pred) && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "box tensors must be 16-byte aligned");
  const int lists = batch * (p->num_classes - 1);
  const PpLayout w = pp_layout(num_anchors, lists, p->keep_topk, true);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  int rc = enable_big_smem();
  if (rc != DAN_OK) return rc;
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
  pp_bind(A, workspace, w);
  DAN_CUDA(cudaMemsetAsync(A.key_count, 0, (size_t)lists * 4, st));
  if (ev) DAN_CUDA(cudaEventRecord(ev[0], st));
  if (num_anchors > 0) {
    pp_filter_kernel<<<dim3((num_anchors + 256 * kFilterPerThread - 1) / (256 * kFilterPerThread), batch), 256, 0, st>>>(A);
    DAN_LAUNCH_CHECK("pp_filter_kernel");
  }
  if (ev) DAN_CUDA(cudaEventRecord(ev[1], st));
  return run_sort_nms<true>(A, lists, nullptr, nullptr, st, ev ? ev + 2 : nullptr);
}

int dan_postprocess_batch(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred, const float* boxes_pred,
                          const float* a_ymin, const float* a_xmin, const float* a_ymax, const float* a_xmax, int32_t num_anchors,
                          int32_t batch, float* out_boxes, float* out_scores, int32_t* out_counts, int32_t* out_anchor_index,
                          int32_t* out_keep_pos, void* workspace, size_t workspace_bytes, void* stream) {
  return postprocess_core(p, cls_pred, loc_pred, boxes_pred, a_ymin, a_xmin, a_ymax, a_xmax, num_anchors, batch, out_boxes, out_scores,
                          out_counts, out_anchor_index, out_keep_pos, workspace, workspace_bytes, stream, nullptr);
}

int dan_postprocess_batch_profile(const dan_postprocess_params* p, const float* cls_pred, const float* loc_pred,
                                  const float* boxes_pred, const float* a_ymin, const float* a_xmin, const float* a_ymax,
                                  const float* a_xmax, int32_t num_anchors, int32_t batch, float* out_boxes, float* out_scores,
                                  int32_t* out_counts, int32_t* out_anchor_index, int32_t* out_keep_pos, void* workspace,
                                  size_t workspace_bytes, void* stream, float* h_kernel_ms) {
  DAN_REQUIRE(h_kernel_ms != nullptr, DAN_ERR_INVALID_ARGUMENT, "h_kernel_ms is NULL");
  for (int i = 0; i < 4; ++i) h_kernel_ms[i] = 0.f;
  cudaEvent_t ev[5];
  for (int i = 0; i < 5; ++i) DAN_CUDA(cudaEventCreate(&ev[i]));
  int rc = postprocess_core(p, cls_pred, loc_pred, boxes_pred, a_ymin, a_xmin, a_ymax, a_xmax, num_anchors, batch, out_boxes,
                            out_scores, out_counts, out_anchor_index, out_keep_pos, workspace, workspace_bytes, stream, ev);
  if (rc == DAN_OK && batch > 0) {
    cudaError_t e = cudaEventSynchronize(ev[4]);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize");
    else for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&h_kernel_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

int dan_sort_bboxes(const float* scores, const float* boxes, int64_t n, int32_t keep_topk, float* out_scores, float* out_boxes,
                    int32_t* out_index, void* workspace, size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(n >= 0 && n < 0x7fffffff && keep_topk >= 1, DAN_ERR_INVALID_ARGUMENT, "bad size");
  DAN_REQUIRE(keep_topk <= kSortCap || n <= kSortCap, DAN_ERR_UNSUPPORTED, "min(keep_topk, n) exceeds the sort capacity %d", kSortCap);
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1, 1, false);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  int rc = enable_big_smem();
  if (rc != DAN_OK) return rc;
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
  const int n_eff = n > 0 ? (int)(n < kSortCap ? n : kSortCap) : 1;
  const int cap = nms_topk < n_eff ? nms_topk : n_eff;
  DAN_REQUIRE(n <= kSortCap && nms_fits(cap, n_eff), DAN_ERR_UNSUPPORTED,
              "n %lld (max %d) / nms_topk %d do not fit the NMS kernel's shared memory (%zu bytes)", (long long)n, kSortCap, nms_topk,
              kNmsSmemMax);
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1, n_eff, true);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  int rc = enable_big_smem();
  if (rc != DAN_OK) return rc;
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
  return run_sort_nms<false>(A, 1, scores, reinterpret_cast<const float4*>(boxes), st);
}

}  // extern "C"
