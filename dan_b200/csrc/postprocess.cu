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
  float nms_thr;
  // workspace
  unsigned long long* keys;   // [L, n]  L = batch * (C-1) lists
  int32_t* key_count;         // [L]
  float* s_scores;            // sort_bboxes outputs [keep_topk]
  float4* s_boxes;
  int32_t* s_index;
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

// decode (when offsets are given) + clip for anchor `a` of image `b`
DAN_D float4 pp_box(const PpArgs& A, int b, int a) {
  const int64_t row = (int64_t)b * A.n + a;
  float4 bx;
  if (A.loc != nullptr) bx = decode_box(A.loc[row], A.ay0[a], A.ax0[a], A.ay1[a], A.ax1[a], A.ps0, A.ps1, A.ps2, A.ps3);
  else bx = A.boxes[row];
  return pp_clip(bx, A.img_h, A.img_w);
}

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

template <bool DECODE>
__global__ void __launch_bounds__(kSortThreads) pp_nms_kernel(const PpArgs A, const float* __restrict__ src_scores,
                                                              const float4* __restrict__ src_boxes) {
  extern __shared__ unsigned long long s_keys[];             // [kSortCap]
  float4* kept_box = reinterpret_cast<float4*>(s_keys + kSortCap);     // [nms_topk]
  float* kept_area = reinterpret_cast<float*>(kept_box + A.nms_topk);  // [nms_topk]
  int32_t* kept_pos = reinterpret_cast<int32_t*>(kept_area + A.nms_topk);   // [nms_topk]
  __shared__ SortScratch sc;
  __shared__ float4 cand_box[2][64];
  __shared__ float cand_area[2][64];
  __shared__ int s_flag[64];
  __shared__ unsigned long long s_rows[64];
  __shared__ int s_kept_n;

  const int list = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int b = list / max(A.num_classes - 1, 1);
  const int cnt = min(A.key_count[list], A.n);
  const int K = min(select_and_sort(A.keys + (int64_t)list * A.n, cnt, min(A.keep_topk, cnt), s_keys, sc), A.keep_topk);
  const int nchunks = (K + 63) >> 6;

  auto fetch_box = [&](int r) -> float4 {
    const uint32_t idx = key_index(s_keys[r]);
    return DECODE ? pp_box(A, b, (int)idx) : src_boxes[idx];
  };

  if (tid == 0) s_kept_n = 0;
  if (tid < 64) {
    s_flag[tid] = 0;
    if (tid < K) {
      const NmsBox nb = nms_norm(fetch_box(tid));
      cand_box[0][tid] = make_float4(nb.y0, nb.x0, nb.y1, nb.x1);
      cand_area[0][tid] = nb.area;
    }
  }
  __syncthreads();

  int kept_n = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int base = c << 6;
    const int nvalid = min(64, K - base);
    const int buf = c & 1;
    // prefetch + decode the next round's candidates; the global loads overlap phase a
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool do_fetch = (tid < 64) && (base + 64 + tid < K);
    if (do_fetch) nxt = fetch_box(base + 64 + tid);

    // ---- a. candidates vs kept list: warp w -> candidates 32*(w&1)+lane, kept slice (w>>1) of 16
    {
      const int i = ((warp & 1) << 5) | lane;
      const float4 me = cand_box[buf][i];
      const float my_area = cand_area[buf][i];
      bool sup = (i >= nvalid);
      for (int k = kept_n - 1 - (warp >> 1); k >= 0; k -= 16) {
        if (!sup && nms_suppresses(kept_box[k], kept_area[k], me, my_area, A.nms_thr)) sup = true;
        if (__all_sync(0xffffffffu, sup)) break;
      }
      if (sup && i < nvalid) s_flag[i] = 1;
    }
    __syncthreads();
    // ---- b. suppression bits among the round's candidates: warp w -> rows 2w, 2w+1; lane -> cols lane, lane+32
    {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * warp + rr;
        const float4 rb = cand_box[buf][r];
        const float ra = cand_area[buf][r];
        const bool row_ok = (r < nvalid) && (s_flag[r] == 0);
        const int c0 = lane, c1 = lane + 32;
        const bool t0 = row_ok && c0 > r && c0 < nvalid && nms_suppresses(rb, ra, cand_box[buf][c0], cand_area[buf][c0], A.nms_thr);
        const bool t1 = row_ok && c1 > r && c1 < nvalid && nms_suppresses(rb, ra, cand_box[buf][c1], cand_area[buf][c1], A.nms_thr);
        const unsigned lo = __ballot_sync(0xffffffffu, t0);
        const unsigned hi = __ballot_sync(0xffffffffu, t1);
        if (lane == 0) s_rows[r] = ((unsigned long long)hi << 32) | lo;
      }
    }
    __syncthreads();
    // ---- c. serial resolve + append (warp 0); the other warps stage the prefetched candidates
    if (warp == 0) {
      const unsigned long long d0 = s_rows[lane], d1 = s_rows[lane + 32];
      const unsigned f0 = __ballot_sync(0xffffffffu, s_flag[lane] != 0);
      const unsigned f1 = __ballot_sync(0xffffffffu, s_flag[lane + 32] != 0);
      const unsigned long long valid_bits = (nvalid >= 64) ? ~0ull : ((1ull << nvalid) - 1ull);
      unsigned long long cur = (((unsigned long long)f1 << 32) | f0) | ~valid_bits;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned long long row = __shfl_sync(0xffffffffu, d0, i);
        if (!((cur >> i) & 1ull)) cur |= row;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned long long row = __shfl_sync(0xffffffffu, d1, i);
        if (!((cur >> (32 + i)) & 1ull)) cur |= row;
      }
      unsigned long long kept = ~cur & valid_bits;
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
          kept_box[pos] = cand_box[buf][i];
          kept_area[pos] = cand_area[buf][i];
          kept_pos[pos] = base + i;
        }
      }
      if (lane == 0) s_kept_n = kept_n + nk;
      s_flag[lane] = 0;
      s_flag[lane + 32] = 0;
    }
    if (do_fetch) {
      const NmsBox nb = nms_norm(nxt);
      cand_box[buf ^ 1][tid] = make_float4(nb.y0, nb.x0, nb.y1, nb.x1);
      cand_area[buf ^ 1][tid] = nb.area;
    }
    __syncthreads();
    kept_n = s_kept_n;
    if (kept_n >= A.nms_topk) break;
  }

  // ---- outputs, zero padded to nms_topk
  for (int t = tid; t < A.nms_topk; t += kSortThreads) {
    const int64_t o = (int64_t)list * A.nms_topk + t;
    if (t < kept_n) {
      const int pos = kept_pos[t];
      const unsigned long long key = s_keys[pos];
      const uint32_t idx = key_index(key);
      A.out_scores[o] = DECODE ? key_to_score((uint32_t)(key >> 32)) : src_scores[idx];
      A.out_boxes[o] = DECODE ? kept_box[t] : src_boxes[idx];     // clipped boxes are already min/max ordered
      if (A.out_index != nullptr) A.out_index[o] = (int32_t)idx;
      if (A.out_keep != nullptr) A.out_keep[o] = A.filler ? pos : (int32_t)idx;
    } else {
      A.out_scores[o] = 0.f;
      A.out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A.out_index != nullptr) A.out_index[o] = -1;
      if (A.out_keep != nullptr) {
        // parse_by_class runs NMS on the zero padded top-k list: zero-area filler rows are never suppressed
        // and get selected until nms_topk is reached
        const int fpos = K + (t - kept_n);
        A.out_keep[o] = (A.filler && fpos < A.keep_topk) ? fpos : -1;
      }
    }
  }
  if (tid == 0 && A.out_counts != nullptr) A.out_counts[list] = kept_n;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
constexpr int kNmsTopkCap = 6144;   // kept list (24 B / box) + 64 KB of keys must fit in 227 KB of shared memory

struct PpLayout {
  size_t key_count, keys, total;
};

static PpLayout pp_layout(int64_t n, int64_t lists) {
  PpLayout w;
  size_t off = 0;
  w.key_count = off; off += align_up(lists * 4, 256);
  w.keys = off;      off += align_up(lists * n * 8, 256);
  w.total = off;
  return w;
}

static void pp_bind(PpArgs& A, void* ws, const PpLayout& w) {
  char* base = static_cast<char*>(ws);
  A.key_count = reinterpret_cast<int32_t*>(base + w.key_count);
  A.keys = reinterpret_cast<unsigned long long*>(base + w.keys);
}

static size_t nms_smem_bytes(int nms_topk) { return (size_t)kSortCap * 8 + (size_t)nms_topk * 24; }

static int enable_big_smem() {
  static bool done = false;
  if (!done) {
    DAN_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8));
    DAN_CUDA(cudaFuncSetAttribute(pp_nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem_bytes(kNmsTopkCap)));
    DAN_CUDA(cudaFuncSetAttribute(pp_nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem_bytes(kNmsTopkCap)));
    done = true;
  }
  return DAN_OK;
}

}  // namespace dan

using namespace dan;

extern "C" {

size_t dan_postprocess_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t num_classes, int32_t keep_topk) {
  if (num_anchors < 0 || batch < 0 || num_classes < 2 || keep_topk < 1) return 0;
  return pp_layout(num_anchors, (int64_t)batch * (num_classes - 1)).total;
}

size_t dan_sort_workspace_bytes(int64_t n, int32_t keep_topk) {
  if (n < 0 || keep_topk < 1) return 0;
  return pp_layout(n, 1).total;
}

size_t dan_nms_workspace_bytes(int64_t n, int32_t nms_topk) {
  if (n < 0 || nms_topk < 0) return 0;
  return pp_layout(n, 1).total;
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
  DAN_REQUIRE(p->nms_topk <= kNmsTopkCap, DAN_ERR_UNSUPPORTED, "nms_topk %d exceeds %d", p->nms_topk, kNmsTopkCap);
  DAN_REQUIRE((loc_pred != nullptr) != (boxes_pred != nullptr), DAN_ERR_INVALID_ARGUMENT, "exactly one of loc_pred / boxes_pred must be given");
  if (batch == 0) return DAN_OK;
  DAN_REQUIRE(cls_pred && out_boxes && out_scores, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(loc_pred == nullptr || (a_ymin && a_xmin && a_ymax && a_xmax), DAN_ERR_INVALID_ARGUMENT, "anchors needed to decode loc_pred");
  DAN_REQUIRE(aligned16(loc_pred) && aligned16(boxes_pred) && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "box tensors must be 16-byte aligned");
  const int lists = batch * (p->num_classes - 1);
  const PpLayout w = pp_layout(num_anchors, lists);
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
  pp_nms_kernel<true><<<lists, kSortThreads, nms_smem_bytes(A.nms_topk), st>>>(A, nullptr, nullptr);
  DAN_LAUNCH_CHECK("pp_nms_kernel");
  if (ev) DAN_CUDA(cudaEventRecord(ev[2], st));
  return DAN_OK;
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
  DAN_REQUIRE(keep_topk <= kSortCap || n <= kSortCap, DAN_ERR_UNSUPPORTED, "min(keep_topk, n) exceeds the sort capacity %d", kSortCap);
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1);
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
  DAN_REQUIRE(n <= kSortCap && nms_topk <= kNmsTopkCap, DAN_ERR_UNSUPPORTED, "n > %d or nms_topk > %d", kSortCap, kNmsTopkCap);
  DAN_REQUIRE(out_scores && out_boxes && aligned16(out_boxes), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned output");
  DAN_REQUIRE(n == 0 || (scores && boxes && aligned16(boxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const PpLayout w = pp_layout(n, 1);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
              workspace_bytes);
  int rc = enable_big_smem();
  if (rc != DAN_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PpArgs A = {};
  A.n = (int)n;
  A.num_classes = 2;
  A.keep_topk = n > 0 ? (int)n : 1;
  A.nms_topk = nms_topk;
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
  pp_nms_kernel<false><<<1, kSortThreads, nms_smem_bytes(nms_topk), st>>>(A, scores, reinterpret_cast<const float4*>(boxes));
  DAN_LAUNCH_CHECK("pp_nms_kernel");
  return DAN_OK;
}

}  // extern "C"
