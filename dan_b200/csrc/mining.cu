// per-image hard-negative mining, the step right after encode on the training side.
//
// Reference: mining_hard_neg, train_dan.py:286-324 (inline copy train_sfd.py:349-384, pinned to /cpu:0 there) and
// mining_hard_neg_across_batch, train_dan.py:247-284.  Per row (= image, or the whole flattened batch):
//   n_sel   = min(int32(negative_ratio * float(#positives)), #negatives)  [max(., 1) in train_dan.py:302]
//   prob    = negatives: 0 - softmax(cls_pred)[.., 0];   everything else: -1                       (:305-309)
//   cut     = n_sel-th largest prob of the row (tf.nn.top_k over the FULL row, :310-311)
//   final   = (negative & prob >= cut) | positive          [strict > in the across-batch variant, :274]
//   outputs = boolean_mask(cls_pred, final), boolean_mask(location_pred, positive),
//             boolean_mask(clip(cls_targets, 0, num_classes), final), boolean_mask(loc_targets, positive)   (:319-322)
//
// The reference sorts all N values of every row to read ONE of them; here the cut is found by a block radix select
// over order-preserving 32-bit keys (no sort), and the four boolean_masks are one ordered compaction.
//   hnm_key_kernel      softmax (same op order as dan_softmax) -> key per anchor, positives/negatives counted per row
//   hnm_select_kernel   one 8-CTA cluster per row: n_sel, 4 x 8-bit MSB-first radix select of the n_sel-th largest key
//                       (slices staged in shared memory, histograms merged through distributed shared memory)
//   hnm_count_kernel    final mask + per-tile counts
//   hnm_scan_kernel     exclusive scan of the tile counts (one CTA)
//   hnm_scatter_kernel  ordered compaction of the four outputs
#include <cooperative_groups.h>

#include "common.cuh"

namespace dan {

namespace {

constexpr int kTile = 1024;          // elements per CTA in the count / scatter kernels (256 threads x 4)

DAN_D uint32_t prob_to_key(float s) { return (uint32_t)float_to_ordered(s) ^ 0x80000000u; }
DAN_D float key_to_prob(uint32_t k) { return ordered_to_float((int)(k ^ 0x80000000u)); }

struct HnmArgs {
  const float* cls;           // [rows*n, c] logits
  const int64_t* labels;      // [rows*n]
  const float4* loc_pred;     // [rows*n]
  const float4* loc_targets;  // [rows*n]
  int c;
  int rows;
  int64_t n;                  // row length
  float ratio;
  int at_least_one;
  int strict;
  int num_classes;
  // workspace
  uint32_t* keys;             // [rows*n]
  int* counts;                // [rows, 2] positives, negatives (zeroed by the caller side of the launch)
  uint32_t* cut;              // [rows]
  int* tile_counts;           // [tiles, 2]  selected, positives  -> exclusive offsets after the scan
  // outputs
  uint8_t* final_mask;        // [rows*n]
  int32_t* n_sel;             // [rows]
  float* score_at_k;          // [rows]
  float* out_cls;             // [<= rows*n, c]
  int64_t* out_labels;        // [<= rows*n]
  float4* out_loc_pred;       // [<= rows*n]
  float4* out_loc_targets;    // [<= rows*n]
  int32_t* out_counts;        // [2] selected, positives
};

__global__ void __launch_bounds__(256) hnm_zero_kernel(int* counts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = 0;
}

// train_dan.py:293-309
__global__ void __launch_bounds__(256) hnm_key_kernel(const HnmArgs A) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int row = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool pos = false, neg = false;
  if (i < A.n) {
    const int64_t idx = (int64_t)row * A.n + i;
    const int64_t label = A.labels[idx];
    pos = label > 0;
    neg = label == 0;
    float prob = -1.f;                                    // 0. - ones_like
    if (neg) {
      const float* x = A.cls + idx * A.c;
      float mx = x[0];
      for (int k = 1; k < A.c; ++k) mx = fmaxf(mx, x[k]);
      const float e0 = cephes_expf(fsub(x[0], mx));
      float s = e0;
      for (int k = 1; k < A.c; ++k) s = fadd(s, cephes_expf(fsub(x[k], mx)));
      prob = fsub(0.f, fmul(e0, fdiv(1.f, s)));           // 0. - softmax[.., 0]
    }
    A.keys[idx] = prob_to_key(prob);
  }
  const unsigned pm = __ballot_sync(0xffffffffu, pos), nm = __ballot_sync(0xffffffffu, neg);
  if ((threadIdx.x & 31) == 0) {
    if (pm) atomicAdd(&s_cnt[0], __popc(pm));
    if (nm) atomicAdd(&s_cnt[1], __popc(nm));
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x] != 0) atomicAdd(A.counts + row * 2 + threadIdx.x, s_cnt[threadIdx.x]);
}

// train_dan.py:301-302,310-311: the n_sel-th largest key of the row, without sorting it.
// A thread-block CLUSTER of kSelCluster CTAs owns one row: each CTA stages its slice of the keys in its own shared
// memory and histograms it (the shared-memory atomics are the cost of a radix select: spread over 8 SMs), every CTA then
// sums the 8 histograms through distributed shared memory and picks the bin redundantly, so nothing is broadcast.
// STAGED=false (slices beyond kStageCap keys: the across-batch rule at large batch) re-reads the keys from L2.
constexpr int kSelCluster = 8;
constexpr int kSelThreads = 512;
constexpr int kStageCap = 55 * 1024;      // keys per CTA: 220 KB of dynamic shared memory

template <bool STAGED>
__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kSelThreads, 1) hnm_select_kernel(const HnmArgs A) {
  namespace cg = cooperative_groups;
  extern __shared__ uint32_t s_keys[];
  __shared__ int hist[256];
  __shared__ int total[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int row = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n_pos = A.counts[row * 2], n_neg = A.counts[row * 2 + 1];
  int k = (int)fmul(A.ratio, (float)n_pos);               // tf.to_int32 truncates
  k = min(k, n_neg);
  if (A.at_least_one) k = max(k, 1);
  if (rank == 0 && tid == 0) A.n_sel[row] = k;
  if (k < 1) {
    // the reference indexes position -1 of the sorted row here (tf.gather_nd rejects it on the CPU); the python
    // mirror raises on n_sel < 1, the kernel selects no negative.  (Uniform over the cluster: no barrier is skipped
    // by only some of its CTAs.)
    if (rank == 0 && tid == 0) { A.cut[row] = 0xffffffffu; A.score_at_k[row] = __int_as_float(0x7fc00000); }
    return;
  }
  const int64_t per = (A.n + kSelCluster - 1) / kSelCluster;
  const int64_t lo = min(A.n, (int64_t)rank * per);
  const int64_t cnt = min(A.n, lo + per) - lo;
  const uint32_t* keys = A.keys + (int64_t)row * A.n + lo;
  if (STAGED) {
    for (int64_t i0 = 0; i0 < cnt; i0 += 8 * kSelThreads) {
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + u * kSelThreads + tid;
        v[u] = (i < cnt) ? keys[i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + u * kSelThreads + tid;
        if (i < cnt) s_keys[i] = v[u];
      }
    }
  }
  uint32_t prefix = 0u, prefix_mask = 0u;
  int remaining = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    // warp-uniform trip count.  First pass: the keys of a row share their leading bits (sign + high exponent bits of
    // a probability), nearly every lane hits the same counter -> equal digits are merged inside the warp first.
    const bool merge = shift == 24;
    for (int64_t i0 = 0; i0 < cnt; i0 += 8 * kSelThreads) {
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + u * kSelThreads + tid;
        v[u] = (i < cnt) ? (STAGED ? s_keys[i] : keys[i]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + u * kSelThreads + tid;
        const bool in = (i < cnt) && ((v[u] & prefix_mask) == prefix);
        const unsigned act = __ballot_sync(0xffffffffu, in);
        if (in) {
          const int digit = (int)((v[u] >> shift) & 255u);
          if (merge) {
            const unsigned peers = __match_any_sync(act, digit);
            if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
          } else {
            atomicAdd(&hist[digit], 1);
          }
        }
      }
    }
    cluster.sync();                                        // all 8 histograms are complete
    if (tid < 256) {
      int sum = 0;
#pragma unroll
      for (int r = 0; r < kSelCluster; ++r) sum += cluster.map_shared_rank(hist, r)[tid];
      total[tid] = sum;
    }
    cluster.sync();                                        // everybody has read them (they are zeroed next pass)
    if (tid < 32) {
      // the bin holding the remaining-th largest key: lane l owns bins 255-8l .. 248-8l (descending)
      int local[8], sum = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        local[e] = total[255 - 8 * lane - e];
        sum += local[e];
      }
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (incl - sum < remaining && remaining <= incl) {   // exactly one lane: the matching keys number >= remaining
        int cum = incl - sum, bin = 255 - 8 * lane;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (cum + local[e] >= remaining) break;
          cum += local[e];
          --bin;
        }
        s_remaining = remaining - cum;
        s_prefix = prefix | ((uint32_t)bin << shift);
      }
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    prefix_mask |= 255u << shift;
  }
  if (rank == 0 && tid == 0) {
    A.cut[row] = prefix;
    A.score_at_k[row] = key_to_prob(prefix);
  }
}

DAN_D void hnm_flags(const HnmArgs& A, int64_t idx, int64_t total, bool& fin, bool& pos) {
  fin = false;
  pos = false;
  if (idx < total) {
    const int64_t label = A.labels[idx];
    pos = label > 0;
    const uint32_t cut = A.cut[idx / A.n];
    const uint32_t key = A.keys[idx];
    const bool sel = A.strict ? (key > cut) : (key >= cut);
    fin = pos || (label == 0 && sel);                     // train_dan.py:313-316
  }
}

__global__ void __launch_bounds__(256) hnm_count_kernel(const HnmArgs A, int64_t total) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int nf = 0, np = 0;
#pragma unroll
  for (int u = 0; u < kTile / 256; ++u) {
    const int64_t idx = (int64_t)blockIdx.x * kTile + u * 256 + threadIdx.x;
    bool fin, pos;
    hnm_flags(A, idx, total, fin, pos);
    if (idx < total) A.final_mask[idx] = fin ? 1 : 0;
    nf += __popc(__ballot_sync(0xffffffffu, fin));
    np += __popc(__ballot_sync(0xffffffffu, pos));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt[0], nf);
    atomicAdd(&s_cnt[1], np);
  }
  __syncthreads();
  if (threadIdx.x < 2) A.tile_counts[blockIdx.x * 2 + threadIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of the (selected, positives) tile counts, in place; totals -> out_counts
__global__ void __launch_bounds__(1024, 1) hnm_scan_kernel(const HnmArgs A, int tiles) {
  __shared__ int s_warp[32][2];
  __shared__ int s_carry[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 2) s_carry[tid] = 0;
  __syncthreads();
  for (int t0 = 0; t0 < tiles; t0 += 1024) {
    const int t = t0 + tid;
    int v[2], incl[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      v[q] = (t < tiles) ? A.tile_counts[t * 2 + q] : 0;
      incl[q] = v[q];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl[q], d);
        if (lane >= d) incl[q] += o;
      }
      if (lane == 31) s_warp[warp][q] = incl[q];
    }
    __syncthreads();
    int before[2] = {s_carry[0], s_carry[1]};
    int chunk[2] = {0, 0};
    for (int w = 0; w < 32; ++w) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (w < warp) before[q] += s_warp[w][q];
        chunk[q] += s_warp[w][q];
      }
    }
    if (t < tiles) {
      A.tile_counts[t * 2] = before[0] + incl[0] - v[0];
      A.tile_counts[t * 2 + 1] = before[1] + incl[1] - v[1];
    }
    __syncthreads();
    if (tid < 2) s_carry[tid] += chunk[tid];
    __syncthreads();
  }
  if (tid < 2) A.out_counts[tid] = s_carry[tid];
}

// train_dan.py:319-322: the four boolean_masks, rows kept in their original order.  Thread t of a tile owns the 4
// consecutive elements 4t..4t+3, so one warp scan + one 8-entry block scan orders the whole tile.
__global__ void __launch_bounds__(256) hnm_scatter_kernel(const HnmArgs A, int64_t total) {
  __shared__ int s_w[8][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t first = (int64_t)blockIdx.x * kTile + 4 * threadIdx.x;
  bool fin[4], pos[4];
  int64_t label[4];
  int nf = 0, np = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t idx = first + e;
    fin[e] = pos[e] = false;
    label[e] = 0;
    if (idx < total) {
      label[e] = A.labels[idx];
      fin[e] = A.final_mask[idx] != 0;
      pos[e] = label[e] > 0;
    }
    nf += fin[e] ? 1 : 0;
    np += pos[e] ? 1 : 0;
  }
  int fi = nf, pi = np;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, fi, d), b = __shfl_up_sync(0xffffffffu, pi, d);
    if (lane >= d) { fi += a; pi += b; }
  }
  if (lane == 31) { s_w[warp][0] = fi; s_w[warp][1] = pi; }
  __syncthreads();
  int64_t fo = A.tile_counts[blockIdx.x * 2] + fi - nf, po = A.tile_counts[blockIdx.x * 2 + 1] + pi - np;
  for (int w = 0; w < warp; ++w) { fo += s_w[w][0]; po += s_w[w][1]; }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t idx = first + e;
    if (fin[e]) {
      for (int k = 0; k < A.c; ++k) A.out_cls[fo * A.c + k] = A.cls[idx * A.c + k];
      A.out_labels[fo] = label[e] < 0 ? 0 : (label[e] > A.num_classes ? (int64_t)A.num_classes : label[e]);   // tf.clip_by_value
      ++fo;
    }
    if (pos[e]) {
      A.out_loc_pred[po] = A.loc_pred[idx];
      A.out_loc_targets[po] = A.loc_targets[idx];
      ++po;
    }
  }
}

struct HnmLayout {
  size_t keys, counts, cut, tile_counts, total;
};

HnmLayout hnm_layout(int64_t rows, int64_t n) {
  HnmLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
  const int64_t total = rows * n;
  w.keys = take((size_t)total * 4);
  w.counts = take((size_t)rows * 2 * 4);
  w.cut = take((size_t)rows * 4);
  w.tile_counts = take((size_t)((total + kTile - 1) / kTile) * 2 * 4);
  w.total = off;
  return w;
}

}  // namespace

}  // namespace dan

using namespace dan;

extern "C" {

size_t dan_hard_negative_workspace_bytes(int32_t rows, int64_t row_len) {
  if (rows < 0 || row_len < 0) return 0;
  return hnm_layout(rows, row_len).total;
}

int dan_hard_negative_mining(const float* cls_pred, int32_t num_logits, const int64_t* cls_targets, const float* loc_pred,
                             const float* loc_targets, int32_t rows, int64_t row_len, float negative_ratio,
                             int32_t num_classes, int32_t at_least_one, int32_t strict_greater, uint8_t* out_final_mask,
                             int32_t* out_n_neg_select, float* out_score_at_k, float* out_cls_pred, int64_t* out_cls_targets,
                             float* out_loc_pred, float* out_loc_targets, int32_t* out_counts, void* workspace,
                             size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(rows >= 0 && row_len >= 0 && num_logits >= 1, DAN_ERR_INVALID_ARGUMENT, "bad shape");
  DAN_REQUIRE(rows <= 65535, DAN_ERR_UNSUPPORTED, "rows > 65535");
  const int64_t total = (int64_t)rows * row_len;
  DAN_REQUIRE(total < ((int64_t)1 << 31), DAN_ERR_UNSUPPORTED, "rows * row_len must be below 2^31");
  DAN_REQUIRE(out_counts != nullptr, DAN_ERR_INVALID_ARGUMENT, "NULL out_counts");
  cudaStream_t st = (cudaStream_t)stream;
  if (total == 0) {
    hnm_zero_kernel<<<1, 256, 0, st>>>(out_counts, 2);
    DAN_LAUNCH_CHECK("hnm_zero_kernel");
    return DAN_OK;
  }
  DAN_REQUIRE(cls_pred && cls_targets && loc_pred && loc_targets, DAN_ERR_INVALID_ARGUMENT, "NULL input");
  DAN_REQUIRE(out_final_mask && out_n_neg_select && out_score_at_k && out_cls_pred && out_cls_targets && out_loc_pred && out_loc_targets,
              DAN_ERR_INVALID_ARGUMENT, "NULL output");
  DAN_REQUIRE(aligned16(loc_pred) && aligned16(loc_targets) && aligned16(out_loc_pred) && aligned16(out_loc_targets),
              DAN_ERR_INVALID_ARGUMENT, "box tensors must be 16-byte aligned");
  const HnmLayout w = hnm_layout(rows, row_len);
  DAN_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
              w.total, workspace_bytes);
  unsigned char* base = static_cast<unsigned char*>(workspace);
  HnmArgs A = {};
  A.cls = cls_pred;
  A.labels = cls_targets;
  A.loc_pred = reinterpret_cast<const float4*>(loc_pred);
  A.loc_targets = reinterpret_cast<const float4*>(loc_targets);
  A.c = num_logits;
  A.rows = rows;
  A.n = row_len;
  A.ratio = negative_ratio;
  A.at_least_one = at_least_one;
  A.strict = strict_greater;
  A.num_classes = num_classes;
  A.keys = reinterpret_cast<uint32_t*>(base + w.keys);
  A.counts = reinterpret_cast<int*>(base + w.counts);
  A.cut = reinterpret_cast<uint32_t*>(base + w.cut);
  A.tile_counts = reinterpret_cast<int*>(base + w.tile_counts);
  A.final_mask = out_final_mask;
  A.n_sel = out_n_neg_select;
  A.score_at_k = out_score_at_k;
  A.out_cls = out_cls_pred;
  A.out_labels = out_cls_targets;
  A.out_loc_pred = reinterpret_cast<float4*>(out_loc_pred);
  A.out_loc_targets = reinterpret_cast<float4*>(out_loc_targets);
  A.out_counts = out_counts;
  const int tiles = (int)((total + kTile - 1) / kTile);
  hnm_zero_kernel<<<(rows * 2 + 255) / 256, 256, 0, st>>>(A.counts, rows * 2);
  DAN_LAUNCH_CHECK("hnm_zero_kernel");
  // across-batch rows can be longer than 65535 * 256 elements: x carries the chunks, y the rows
  hnm_key_kernel<<<dim3((unsigned)((row_len + 255) / 256), rows), 256, 0, st>>>(A);
  DAN_LAUNCH_CHECK("hnm_key_kernel");
  const int64_t slice = (row_len + kSelCluster - 1) / kSelCluster;
  if (slice <= kStageCap) {
    DAN_CUDA(cudaFuncSetAttribute(hnm_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageCap * 4));
    hnm_select_kernel<true><<<dim3(kSelCluster, rows), kSelThreads, (size_t)slice * 4, st>>>(A);
  } else {
    hnm_select_kernel<false><<<dim3(kSelCluster, rows), kSelThreads, 0, st>>>(A);
  }
  DAN_LAUNCH_CHECK("hnm_select_kernel");
  hnm_count_kernel<<<tiles, 256, 0, st>>>(A, total);
  DAN_LAUNCH_CHECK("hnm_count_kernel");
  hnm_scan_kernel<<<1, 1024, 0, st>>>(A, tiles);
  DAN_LAUNCH_CHECK("hnm_scan_kernel");
  hnm_scatter_kernel<<<tiles, 256, 0, st>>>(A, total);
  DAN_LAUNCH_CHECK("hnm_scatter_kernel");
  return DAN_OK;
}

}  // extern "C"
