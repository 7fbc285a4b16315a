// K3-K5: the evaluation half of the path -- utility/bbox_util.py:103-119
// parse_by_class and its pieces, batched over images and classes.
//
//   pp_filter_kernel   softmax -> select(threshold) -> decode -> clip -> min-size;
//                      survivors are COMPACTED as 64-bit keys
//                      (ordered score bits << 32 | ~anchor index).  The reference
//                      instead multiplies by 0/1 masks and keeps all N rows
//                      (bbox_util.py:24-59); rows it zeroes can never precede a
//                      positive score in tf.nn.top_k, so dropping them is exact.
//   topk_sort_kernel   per (image, class): block radix select of the keep_topk
//                      largest keys when more survive, then an in-shared-memory
//                      bitonic sort.  Keys are unique, descending key order ==
//                      descending score with ties broken by LOWER index, which is
//                      tf.nn.top_k's documented order (bbox_util.py:64).
//   nms_mask_kernel    64x64 tiles of the upper-triangular suppression bit matrix
//                      with tf.image.non_max_suppression's IoU (no +1, strict >,
//                      area<=0 never suppresses).
//   nms_sweep_kernel   one warp per (image, class): 64-wide chunks; the serial
//                      part is a register-resident shuffle sweep over the diagonal
//                      word, the kept rows are OR-ed into the removed set by all
//                      lanes; writes the zero padded outputs (bbox_util.py:80-90).
#include "common.cuh"

namespace dan {

constexpr int kSortCap = 8192;      // keys sorted in shared memory (64 KB)
constexpr int kSortThreads = 1024;
constexpr int kAdjCap = 64;        // suppressor list capacity per candidate
constexpr int kPairCtasPerList = 32;


DAN_D uint32_t score_to_key(float s) { return (uint32_t)float_to_ordered(s) ^ 0x80000000u; }
DAN_D float key_to_score(uint32_t k) { return ordered_to_float((int)(k ^ 0x80000000u)); }

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
  float ps0, ps1, ps2, ps3;
  int keep_topk, nms_topk;
  int nms_cap;                // kept-list capacity in shared memory = min(nms_topk, keep_topk)
  float nms_thr;
  // workspace
  unsigned long long* keys;   // [L, n]  L = batch * (C-1) lists
  int32_t* key_count;         // [L]
  float* s_scores;            // sort_bboxes outputs [keep_topk]
  float4* s_boxes;
  int32_t* s_index;
  // sorted candidates (sort kernel -> pairs kernel -> resolve kernel), per list
  unsigned long long* s_key;  // [L, keep_topk] sorted keys
  float4* s_box;              // [L, keep_topk] normalised corners (y0, x0, y1, x1)
  float* s_area;              // [L, keep_topk] (y1-y0)*(x1-x0)
  int32_t* s_len;             // [L] number of sorted candidates K
  int32_t* deg;               // [L, keep_topk] number of higher-ranked boxes that suppress candidate i
  uint16_t* adj;              // [L, kAdjCap, keep_topk] their positions
  int32_t* ovf;               // [L] some candidate has more than kAdjCap suppressors -> round-based fallback
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
__global__ void __launch_bounds__(256) pp_filter_kernel(const PpArgs A) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = a < A.n;
  const int C = A.num_classes;
  const float* x = A.cls + ((int64_t)b * A.n + (valid ? a : 0)) * C;

  // tf.nn.softmax: exp(x - max) * (1 / sum(exp(x - max))), sum in class order
  float mx = 0.f, inv = 0.f;
  if (valid) {
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
    if (valid) {
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
      int base = 0;
      if (lane == 0) base = atomicAdd(A.key_count + list, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        A.keys[(int64_t)list * A.n + pos] = ((unsigned long long)score_to_key(p) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
      }
    }
  }
}

// standalone sort_bboxes / nms_bboxes: one key per input row, nothing filtered
__global__ void __launch_bounds__(256) key_build_kernel(const float* __restrict__ scores, int64_t n, unsigned long long* __restrict__ keys,
                                                        int32_t* __restrict__ key_count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ((unsigned long long)score_to_key(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
  if (blockIdx.x == 0 && threadIdx.x == 0) key_count[0] = (int32_t)n;
}

// ---------------------------------------------------------------------------
// K4: per-list top-k (radix select when needed) + bitonic sort, all in shared memory.
// Every thread of the CTA calls it; returns the number of sorted keys m (descending in s_keys[0, m)).
// ---------------------------------------------------------------------------
struct SortScratch {
  int hist[256];
  unsigned long long prefix;
  int remaining;
  int fill;
};

DAN_D int select_and_sort(const unsigned long long* __restrict__ keys, int cnt, int k, unsigned long long* s_keys, SortScratch& sc) {
  const int tid = threadIdx.x;
  int m = cnt;
  if (cnt <= kSortCap) {
    for (int i = tid; i < cnt; i += kSortThreads) s_keys[i] = keys[i];
  } else {
    // block radix select, MSB first, 8 bits per pass: find the k-th largest key
    if (tid == 0) { sc.prefix = 0ull; sc.remaining = k; }
    unsigned long long prefix_mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kSortThreads) sc.hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = sc.prefix;
      for (int i = tid; i < cnt; i += kSortThreads) {
        const unsigned long long key = keys[i];
        if ((key & prefix_mask) == prefix) atomicAdd(&sc.hist[(int)((key >> shift) & 255ull)], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, bin = 255;
        for (; bin > 0; --bin) {
          if (cum + sc.hist[bin] >= sc.remaining) break;
          cum += sc.hist[bin];
        }
        sc.remaining -= cum;
        sc.prefix = prefix | ((unsigned long long)bin << shift);
      }
      prefix_mask |= 255ull << shift;
      __syncthreads();
    }
    const unsigned long long kth = sc.prefix;
    if (tid == 0) sc.fill = 0;
    __syncthreads();
    for (int i = tid; i < cnt; i += kSortThreads) {
      const unsigned long long key = keys[i];
      if (key >= kth) s_keys[atomicAdd(&sc.fill, 1)] = key;   // exactly k keys (keys are unique)
    }
    m = k;
  }
  // pad to a power of two with 0 (smaller than any real key: the low word of a real key is ~index != 0)
  int lp2 = 0;
  while ((1 << lp2) < m) ++lp2;
  const int p2 = 1 << lp2;
  for (int i = m + tid; i < p2; i += kSortThreads) s_keys[i] = 0ull;
  __syncthreads();
  // bitonic sort, descending; strides are powers of two -> shifts only
  for (int lsize = 1; lsize <= lp2; ++lsize) {
    for (int ls = lsize - 1; ls >= 0; --ls) {
      const int stride = 1 << ls;
      for (int t = tid; t < (p2 >> 1); t += kSortThreads) {
        const int lo = ((t >> ls) << (ls + 1)) | (t & (stride - 1));
        const int hi = lo | stride;
        const bool desc = ((lo >> lsize) & 1) == 0;
        const unsigned long long x = s_keys[lo], y = s_keys[hi];
        if ((x < y) == desc) { s_keys[lo] = y; s_keys[hi] = x; }
      }
      __syncthreads();
    }
  }
  return m;
}

DAN_D uint32_t key_index(unsigned long long key) { return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull); }

// sort_bboxes: tf.nn.top_k + gather + zero pad (bbox_util.py:61-72)
__global__ void __launch_bounds__(kSortThreads) topk_sort_kernel(const PpArgs A, const float* __restrict__ src_scores,
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
// K4+K5 fused: one CTA per (image, class) list.
//   1. top-k select + sort of the surviving keys in shared memory (above)
//   2. greedy NMS with tf.image.non_max_suppression's IoU (no +1, corners min/max normalised, area<=0 never
//      suppresses, strict >), 64 candidates per round:
//        a. all warps test the 64 candidates against the boxes kept so far (kept list lives in shared memory,
//           newest first like TF's inner loop; the result does not depend on the order)
//        b. all warps build the 64x64 suppression bits among the candidates (ballots, no atomics)
//        c. warp 0 resolves the round serially with a register/shuffle sweep (bit i of `cur` = candidate i is
//           suppressed), truncates at nms_topk and appends the survivors to the kept list
//      Only the rows of KEPT boxes are ever evaluated, so ~K*kept/2 IoU tests instead of K*K/2, no K x K bit
//      matrix in HBM, and no dependent global loads on the serial path: the next round's candidate boxes are
//      fetched (and decoded) while the current round is being tested.
//   3. zero padded outputs (bbox_util.py:80-90).
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

// ---------------------------------------------------------------------------
// K4: one CTA per list: top-k select + sort, then decode + clip + normalise the K best boxes (sorted order) to HBM
// ---------------------------------------------------------------------------
template <bool DECODE>
__global__ void __launch_bounds__(kSortThreads) pp_sort_kernel(const PpArgs A, const float4* __restrict__ src_boxes) {
  extern __shared__ unsigned long long s_keys[];             // [kSortCap]
  __shared__ SortScratch sc;
  const int list = blockIdx.x;
  const int tid = threadIdx.x;
  const int b = list / max(A.num_classes - 1, 1);
  const int cnt = min(A.key_count[list], A.n);
  const int K = min(select_and_sort(A.keys + (int64_t)list * A.n, cnt, min(A.keep_topk, cnt), s_keys, sc), A.keep_topk);
  const int64_t o = (int64_t)list * A.keep_topk;
  for (int r = tid; r < K; r += kSortThreads) {
    const unsigned long long key = s_keys[r];
    const uint32_t idx = key_index(key);
    const NmsBox nb = nms_norm(DECODE ? pp_box(A, b, (int)idx) : src_boxes[idx]);
    A.s_key[o + r] = key;
    A.s_box[o + r] = make_float4(nb.y0, nb.x0, nb.y1, nb.x1);
    A.s_area[o + r] = nb.area;
    A.deg[o + r] = 0;
  }
  if (tid == 0) {
    A.s_len[list] = K;
    A.ovf[list] = 0;
  }
}

// ---------------------------------------------------------------------------
// K5a: all SMs: for every candidate i the positions j < i (higher rank) with IoU(j, i) > thr, i.e. the boxes that
// would suppress it if they are kept.  64x64 tiles of the strictly-lower triangle, kPairCtasPerList CTAs per list.
// The relation is sparse (a detection overlaps the few other detections of the same face), so it is stored as
// short per-candidate lists instead of a K x K bit matrix.
// ---------------------------------------------------------------------------
DAN_D bool pair_suppresses(const float4& a, float a_area, const float4& b, float b_area, float thr) {
  if (thr < 0.f) return nms_suppresses(a, a_area, b, b_area, thr);
  // boxes that do not overlap cannot exceed thr >= 0
  const float h = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float w = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
  if (!(h > 0.f && w > 0.f)) return false;
  if (!(a_area > 0.f && b_area > 0.f)) return false;
  const float inter = fmul(h, w);
  const float uni = fsub(fadd(a_area, b_area), inter);
  // inter/uni > thr decided without the division unless the ratio is within 1e-6 (relative) of the threshold
  const float t = fmul(thr, uni);
  if (t > 1e-30f && inter > fmul(t, 1.000001f)) return true;
  if (t > 1e-30f && inter < fmul(t, 0.999999f)) return false;
  return fdiv(inter, uni) > thr;
}

__global__ void __launch_bounds__(256) nms_pairs_kernel(const PpArgs A) {
  __shared__ float4 s_cb[64];
  __shared__ float s_ca[64];
  const int list = blockIdx.y;
  const int K = A.s_len[list];
  const int nb = (K + 63) >> 6;
  const int tiles = nb * (nb + 1) / 2;
  const int64_t o = (int64_t)list * A.keep_topk;
  const float4* box = A.s_box + o;
  const float* area = A.s_area + o;
  const int t = threadIdx.x;
  const int r = t & 63, q = t >> 6;
  const bool prune = A.nms_thr >= 0.f;     // boxes that do not overlap cannot exceed thr >= 0
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // row block rb (candidates i), column block cb <= rb; row block rb owns rb + 1 tiles: rb = floor((sqrt(8t+1)-1)/2)
    int rb = (int)((sqrtf(8.f * (float)tile + 1.f) - 1.f) * 0.5f);
    while (rb * (rb + 1) / 2 > tile) --rb;
    while ((rb + 1) * (rb + 2) / 2 <= tile) ++rb;
    const int cb = tile - rb * (rb + 1) / 2;
    __syncthreads();
    if (t < 64) {
      const int j = cb * 64 + t;
      if (j < K) { s_cb[t] = box[j]; s_ca[t] = area[j]; }
    }
    const int i = rb * 64 + r;
    const bool have = i < K;
    const float4 me = have ? box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float my_area = have ? area[i] : 0.f;
    __syncthreads();
    // tight branch-free loop: one bit per column that may suppress (spatial overlap); the rare exact tests and the
    // list appends run afterwards, so that a hit in one lane does not stall the other 31 inside the loop
    const int jmax = min(i, K) - cb * 64;            // columns c < jmax are higher ranked than i
    unsigned bits = 0u;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = q * 16 + u;
      const float4 kb = s_cb[c];
      const float h = fsub(fminf(kb.z, me.z), fmaxf(kb.x, me.x));
      const float w = fsub(fminf(kb.w, me.w), fmaxf(kb.y, me.y));
      const bool cand = (c < jmax) & (!prune | ((h > 0.f) & (w > 0.f)));
      bits |= (cand ? 1u : 0u) << u;
    }
    if (!have) bits = 0u;
    while (__any_sync(0xffffffffu, bits != 0u)) {      // warp-uniform loop
      if (bits != 0u) {
        const int u = __ffs(bits) - 1;
        bits &= bits - 1u;
        const int c = q * 16 + u;
        if (pair_suppresses(s_cb[c], s_ca[c], me, my_area, A.nms_thr)) {
          const int pos = atomicAdd(A.deg + o + i, 1);
          if (pos < kAdjCap) A.adj[((int64_t)list * kAdjCap + pos) * A.keep_topk + i] = (uint16_t)(cb * 64 + c);
          else A.ovf[list] = 1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K5b: one CTA per list: resolve the greedy NMS and write the zero padded outputs (bbox_util.py:80-90).
//   Greedy NMS keeps candidate i iff none of its suppressors j < i is kept: kept(i) = !any(kept(j), j in adj(i)).
//   The dependency graph is a DAG ordered by rank; it is evaluated by parallel relaxation: a candidate is decided
//   as soon as one suppressor is known kept (-> suppressed) or all are known suppressed (-> kept).  The number of
//   sweeps is the longest dependency chain (a handful for detections), not K.  Truncation at nms_topk keeps the
//   first nms_topk kept candidates in rank order, which is what the sequential loop of TF selects.
//   Fallback (a candidate with more than kAdjCap suppressors): rounds of 64 candidates tested on the fly against
//   the kept list in shared memory, resolved serially per round.
// ---------------------------------------------------------------------------
template <bool DECODE>
__global__ void __launch_bounds__(kSortThreads) nms_resolve_kernel(const PpArgs A, const float* __restrict__ src_scores,
                                                                   const float4* __restrict__ src_boxes) {
  extern __shared__ float4 dyn_smem[];
  float4* kept_box = dyn_smem;                                         // [nms_cap]   (fallback)
  float4* cand_box = kept_box + A.nms_cap;                             // [keep_topk] (fallback)
  float* kept_area = reinterpret_cast<float*>(cand_box + A.keep_topk); // [nms_cap]   (fallback)
  float* cand_area = kept_area + A.nms_cap;                            // [keep_topk] (fallback)
  int32_t* kept_pos = reinterpret_cast<int32_t*>(cand_area + A.keep_topk);   // [nms_cap]
  uint8_t* status = reinterpret_cast<uint8_t*>(kept_pos + A.nms_cap);  // [keep_topk] 0 undecided, 1 kept, 2 suppressed
  __shared__ int s_flag[2][64];
  __shared__ unsigned long long s_rows[64];
  __shared__ int s_new_n;
  __shared__ int s_scan[kSortThreads / 32];

  const int list = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int K = A.s_len[list];
  const int64_t o = (int64_t)list * A.keep_topk;
  const int nchunks = (K + 63) >> 6;
  int kept_n = 0;

  if (A.ovf[list] == 0) {
    // ---- parallel relaxation over the suppressor lists
    for (int i = tid; i < K; i += kSortThreads) status[i] = 0;
    __syncthreads();
    const uint16_t* adj = A.adj + (int64_t)list * kAdjCap * A.keep_topk;
    // candidate i = tid + 1024*m (m < 8 since K <= kSortCap); its suppressor count stays in a register
    int dreg[kSortCap / kSortThreads];
#pragma unroll
    for (int m = 0; m < kSortCap / kSortThreads; ++m) {
      const int i = tid + m * kSortThreads;
      dreg[m] = (i < K) ? A.deg[o + i] : 0;
    }
    while (true) {
      bool pending_any = false;
#pragma unroll
      for (int m = 0; m < kSortCap / kSortThreads; ++m) {
        if (m * kSortThreads < K) {                     // CTA-uniform
          const int i = tid + m * kSortThreads;
          const bool mine = (i < K) && (status[i] == 0);
          const int d = mine ? dreg[m] : 0;
          const int dmax = __reduce_max_sync(0xffffffffu, d);
          int res = 1;                                   // kept unless a suppressor says otherwise
          for (int e0 = 0; e0 < dmax; e0 += 8) {         // warp-uniform trip count, structured body
            int js[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)                  // 8 independent loads in flight, one L2 round trip
              js[u] = (e0 + u < d) ? (int)adj[(int64_t)(e0 + u) * A.keep_topk + i] : -1;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (js[u] >= 0) {
                const int sj = status[js[u]];
                res = (sj == 1) ? 2 : ((sj == 0 && res != 2) ? 0 : res);
              }
            }
          }
          if (mine) {
            if (res != 0) status[i] = (uint8_t)res;
            else pending_any = true;
          }
        }
      }
      if (!__syncthreads_or(pending_any ? 1 : 0)) break;
    }
    // ordered compaction of the kept candidates: thread t owns positions [t*E, (t+1)*E)
    const int E = (K + kSortThreads - 1) / kSortThreads;
    int mine_cnt = 0;
    for (int e = 0; e < E; ++e) {
      const int i = tid * E + e;
      if (i < K && status[i] == 1) ++mine_cnt;
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
      if (i < K && status[i] == 1) {
        if (pos < A.nms_topk) kept_pos[pos] = i;
        ++pos;
      }
    }
    kept_n = min(total, A.nms_topk);
    __syncthreads();
  } else {
  // ---- fallback: rounds of 64 candidates against the kept list (shared memory resident)
  for (int r = tid; r < K; r += kSortThreads) {
    cand_box[r] = A.s_box[o + r];
    cand_area[r] = A.s_area[o + r];
  }
  if (tid < 128) (&s_flag[0][0])[tid] = 0;
  __syncthreads();

  // Software pipeline over rounds of 64 candidates (see the kernel comment).  In round c:
  //   S1  all warps      b: suppression bits among the candidates of round c (their flags vs the kept list are final)
  //   S2  warp 0         c: serial resolve of round c -> appends nk boxes to the kept list
  //       warps 1..30    a1: candidates of round c+1 vs the kept list as it was BEFORE round c
  //   S3  all warps      a2: candidates of round c+1 vs the nk boxes round c just appended
  const float thr = A.nms_thr;
  auto kept_suppresses = [&](int kk, const float4& me, float my_area) -> bool {
    return pair_suppresses(kept_box[kk], kept_area[kk], me, my_area, thr);
  };

  for (int c = 0; c < nchunks; ++c) {
    const int base = c << 6;
    const int nvalid = min(64, K - base);
    const int nvalid_next = max(0, min(64, K - base - 64));
    const float4* cur_box = cand_box + base;           // candidates of round c
    const float* cur_area = cand_area + base;
    const float4* nxt_box = cand_box + base + 64;      // candidates of round c+1
    const float* nxt_area = cand_area + base + 64;
    const int fb = c & 1, fb1 = fb ^ 1;

    // ---- S1 (b): warp w -> rows 2w, 2w+1; lane -> cols lane, lane+32
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
      // (c) greedy resolve of the round, 32-bit halves: bit i of cl/ch set <=> candidate i / 32+i is suppressed
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
      if (kept_n + nk > A.nms_topk) {             // max_output_size reached inside the round
        int drop = kept_n + nk - A.nms_topk;
        while (drop-- > 0) kept &= ~(1ull << (63 - __clzll((long long)kept)));
        nk = A.nms_topk - kept_n;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if ((kept >> i) & 1ull) {
          const int pos = kept_n + __popcll(kept & ((1ull << i) - 1ull));
          kept_box[pos] = cur_box[i];
          kept_area[pos] = cur_area[i];
          kept_pos[pos] = base + i;
        }
      }
      if (lane == 0) s_new_n = nk;
      s_flag[fb][lane] = 0;          // this flag buffer is reused by round c+2
      s_flag[fb][lane + 32] = 0;
    } else if (warp < 31) {
      // (a1) round c+1 vs kept[0, kept_n): warp w in 1..30 -> candidates 32*((w-1)&1)+lane, slice (w-1)>>1 of 15.
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
          if (kk >= 0 && !done && kept_suppresses(kk, me, my_area)) { sup = true; done = true; }
        }
        if (__all_sync(0xffffffffu, done)) break;
      }
      if (sup) s_flag[fb1][i] = 1;
    }
    __syncthreads();

    // ---- S3 (a2): round c+1 vs the boxes appended by round c: thread -> candidate tid&63, new boxes (tid>>6)+16j
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
          if (j < nk && !sup && kept_suppresses(kept_n + j, me, my_area)) sup = true;
        }
        if (sup) s_flag[fb1][i] = 1;
      }
    }
    kept_n += nk;
    __syncthreads();
    if (kept_n >= A.nms_topk) break;
  }
  }   // fallback

  // ---- outputs, zero padded to nms_topk
  for (int t = tid; t < A.nms_topk; t += kSortThreads) {
    const int64_t oo = (int64_t)list * A.nms_topk + t;
    if (t < kept_n) {
      const int pos = kept_pos[t];
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
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

struct PpLayout {
  size_t key_count, s_len, ovf, keys, s_key, s_box, s_area, deg, adj, total;
};

static PpLayout pp_layout(int64_t n, int64_t lists, int64_t keep_topk) {
  PpLayout w;
  size_t off = 0;
  w.key_count = off; off += align_up(lists * 4, 256);
  w.s_len = off;     off += align_up(lists * 4, 256);
  w.ovf = off;       off += align_up(lists * 4, 256);
  w.keys = off;      off += align_up(lists * n * 8, 256);
  w.s_key = off;     off += align_up(lists * keep_topk * 8, 256);
  w.s_box = off;     off += align_up(lists * keep_topk * 16, 256);
  w.s_area = off;    off += align_up(lists * keep_topk * 4, 256);
  w.deg = off;       off += align_up(lists * keep_topk * 4, 256);
  w.adj = off;       off += align_up(lists * keep_topk * (size_t)kAdjCap * 2, 256);
  w.total = off;
  return w;
}

static void pp_bind(PpArgs& A, void* ws, const PpLayout& w) {
  char* base = static_cast<char*>(ws);
  A.key_count = reinterpret_cast<int32_t*>(base + w.key_count);
  A.s_len = reinterpret_cast<int32_t*>(base + w.s_len);
  A.ovf = reinterpret_cast<int32_t*>(base + w.ovf);
  A.keys = reinterpret_cast<unsigned long long*>(base + w.keys);
  A.s_key = reinterpret_cast<unsigned long long*>(base + w.s_key);
  A.s_box = reinterpret_cast<float4*>(base + w.s_box);
  A.s_area = reinterpret_cast<float*>(base + w.s_area);
  A.deg = reinterpret_cast<int32_t*>(base + w.deg);
  A.adj = reinterpret_cast<uint16_t*>(base + w.adj);
}

// resolve kernel: fallback arrays (20 B per candidate, 24 B per kept box) + kept positions + 1 status byte per candidate
static size_t nms_smem_bytes(int nms_cap, int keep_topk) { return (size_t)nms_cap * 24 + (size_t)keep_topk * 20 + align_up((size_t)keep_topk, 16); }
constexpr size_t kNmsSmemMax = 227 * 1024 - 8 * 1024;   // dynamic part; a few KB of static shared memory on top

static int enable_big_smem() {
  static bool done = false;
  if (!done) {
    DAN_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8));
    DAN_CUDA(cudaFuncSetAttribute(pp_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8));
    DAN_CUDA(cudaFuncSetAttribute(pp_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8));
    DAN_CUDA(cudaFuncSetAttribute(nms_resolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemMax));
    DAN_CUDA(cudaFuncSetAttribute(nms_resolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemMax));
    done = true;
  }
  return DAN_OK;
}

// sort -> pairs -> resolve for `lists` lists whose keys are already in the workspace; ev (optional): 3 events recorded
// after each kernel
template <bool DECODE>
static int run_sort_nms(const PpArgs& A, int lists, const float* src_scores, const float4* src_boxes, cudaStream_t st,
                        cudaEvent_t* ev = nullptr) {
  pp_sort_kernel<DECODE><<<lists, kSortThreads, kSortCap * 8, st>>>(A, src_boxes);
  DAN_LAUNCH_CHECK("pp_sort_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[0], st));
  nms_pairs_kernel<<<dim3(kPairCtasPerList, lists), 256, 0, st>>>(A);
  DAN_LAUNCH_CHECK("nms_pairs_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[1], st));
  nms_resolve_kernel<DECODE><<<lists, kSortThreads, nms_smem_bytes(A.nms_cap, A.keep_topk), st>>>(A, src_scores, src_boxes);
  DAN_LAUNCH_CHECK("nms_resolve_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[2], st));
  return DAN_OK;
}

}  // namespace dan

using namespace dan;

extern "C" {

size_t dan_postprocess_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t num_classes, int32_t keep_topk) {
  if (num_anchors < 0 || batch < 0 || num_classes < 2 || keep_topk < 1) return 0;
  return pp_layout(num_anchors, (int64_t)batch * (num_classes - 1), keep_topk).total;
}

size_t dan_sort_workspace_bytes(int64_t n, int32_t keep_topk) {
  if (n < 0 || keep_topk < 1) return 0;
  return pp_layout(n, 1, 1).total;
}

size_t dan_nms_workspace_bytes(int64_t n, int32_t nms_topk) {
  if (n < 0 || nms_topk < 0) return 0;
  return pp_layout(n, 1, n > 0 ? n : 1).total;
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
  DAN_REQUIRE(nms_smem_bytes(p->nms_topk < p->keep_topk ? p->nms_topk : p->keep_topk, p->keep_topk) <= kNmsSmemMax, DAN_ERR_UNSUPPORTED,
              "keep_topk %d / nms_topk %d need more than %zu bytes of shared memory (20 B per candidate + 24 B per kept box + 64 KB)",
              p->keep_topk, p->nms_topk, kNmsSmemMax);
  DAN_REQUIRE((loc_pred != nullptr) != (boxes_pred != nullptr), DAN_ERR_INVALID_ARGUMENT, "exactly one of loc_pred / boxes_pred must be given");
  if (batch == 0) return DAN_OK;
  DAN_REQUIRE(cls_pred && out_boxes && out_scores, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(loc_pred == nullptr || (a_ymin && a_xmin && a_ymax && a_xmax), DAN_ERR_INVALID_ARGUMENT, "anchors needed to decode loc_pred");
  DAN_REQUIRE(aligned16(loc_pred) && aligned16(boxes_pred) && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "box tensors must be 16-byte aligned");
  const int lists = batch * (p->num_classes - 1);
  const PpLayout w = pp_layout(num_anchors, lists, p->keep_topk);
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
    pp_filter_kernel<<<dim3((num_anchors + 255) / 256, batch), 256, 0, st>>>(A);
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
  const PpLayout w = pp_layout(n, 1, 1);
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
  DAN_REQUIRE(n <= kSortCap && nms_smem_bytes(cap, n_eff) <= kNmsSmemMax, DAN_ERR_UNSUPPORTED,
              "n %lld (max %d) / nms_topk %d need more than %zu bytes of shared memory", (long long)n, kSortCap, nms_topk, kNmsSmemMax);
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1, n_eff);
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
