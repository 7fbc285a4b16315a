// the input hand-off, i.e. the deterministic tail of the reference's training preprocessing that sits
// between the sampled image patch and encode_anchors, for a whole batch:
//   random_flip_left_right    sfd_preprocessing.py:482-493  boxes (ymin, W-1-xmax, ymax, W-1-xmin) when the image was mirrored
//                                                            (the coin flip itself is an input here)
//   rescale to the net input  sfd_preprocessing.py:529-533  ymin * target_h / patch_h, ... (two separately rounded fp32 ops)
//   small-face filter         sfd_preprocessing.py:544-550  keep (ymax - ymin) > 6 and (xmax - xmin) > 3, order preserved
//   keep_input                dataset_common.py:178,186     an image whose box list is empty afterwards is not batched
// Output = the CSR batch dan_encode_batch consumes (gt_boxes of the kept images back to back, gt_offsets) plus the
// indices of the kept images.  One 1024-thread CTA: a batch holds a few thousand boxes at most.
#include "common.cuh"

namespace dan {

namespace {

struct HandoffArgs {
  const float4* boxes;        // [total] (ymin, xmin, ymax, xmax), pixels of the sampled patch
  const int32_t* offsets;     // [B + 1]
  const float* patch_hw;      // [B, 2] height, width of the patch before resizing
  const uint8_t* mirror;      // [B] or NULL
  int batch;
  float target_h, target_w, min_h, min_w;
  float4* out_boxes;          // [total]
  int32_t* out_offsets;       // [B + 1]
  int32_t* out_image;         // [B]
  int32_t* out_counts;        // [2] kept images, kept boxes
};

DAN_D float4 handoff_box(const HandoffArgs& A, int b, int i, bool& keep) {
  float4 g = A.boxes[i];
  const float ph = A.patch_hw[2 * b], pw = A.patch_hw[2 * b + 1];
  if (A.mirror != nullptr && A.mirror[b]) {
    const float x0 = fsub(fsub(pw, 1.f), g.w), x1 = fsub(fsub(pw, 1.f), g.y);     // float_width - 1. - xmax / xmin
    g.y = x0;
    g.w = x1;
  }
  g.x = fdiv(fmul(g.x, A.target_h), ph);
  g.z = fdiv(fmul(g.z, A.target_h), ph);
  g.y = fdiv(fmul(g.y, A.target_w), pw);
  g.w = fdiv(fmul(g.w, A.target_w), pw);
  keep = (fsub(g.z, g.x) > A.min_h) && (fsub(g.w, g.y) > A.min_w);
  return g;
}

// One warp per image, 32 images per sweep of the CTA: the lanes take the image's boxes 32 at a time (ballot + prefix
// popcount keep their order), warp 0 scans the 32 per-image counts, then every warp writes its survivors.
__global__ void __launch_bounds__(1024, 1) gt_handoff_kernel(const HandoffArgs A) {
  __shared__ int s_cnt[32], s_slot[32], s_at[32];
  __shared__ int s_carry[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 2) s_carry[tid] = 0;
  __syncthreads();
  for (int b0 = 0; b0 < A.batch; b0 += 32) {
    const int b = b0 + warp;
    int lo = 0, hi = 0, kept = 0;
    if (b < A.batch) {
      lo = A.offsets[b];
      hi = A.offsets[b + 1];
      for (int i0 = lo; i0 < hi; i0 += 32) {
        bool keep = false;
        if (i0 + lane < hi) handoff_box(A, b, i0 + lane, keep);
        kept += __popc(__ballot_sync(0xffffffffu, keep));
      }
    }
    if (lane == 0) s_cnt[warp] = kept;
    __syncthreads();
    if (warp == 0) {                                       // exclusive scan over the 32 images of (image kept, boxes kept)
      const int c = s_cnt[lane];
      int img = c > 0 ? 1 : 0, box = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o0 = __shfl_up_sync(0xffffffffu, img, d), o1 = __shfl_up_sync(0xffffffffu, box, d);
        if (lane >= d) { img += o0; box += o1; }
      }
      s_slot[lane] = s_carry[0] + img - 1;                  // position of the image in the batch (if it is kept)
      s_at[lane] = s_carry[1] + box - c;                    // its first box
      __syncwarp();
      if (lane == 31) { s_carry[0] += img; s_carry[1] += box; }
    }
    __syncthreads();
    if (b < A.batch && kept > 0) {
      int at = s_at[warp];
      if (lane == 0) {
        A.out_image[s_slot[warp]] = b;
        A.out_offsets[s_slot[warp]] = at;
      }
      for (int i0 = lo; i0 < hi; i0 += 32) {
        bool keep = false;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i0 + lane < hi) g = handoff_box(A, b, i0 + lane, keep);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) A.out_boxes[at + __popc(m & ((1u << lane) - 1u))] = g;
        at += __popc(m);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    A.out_offsets[s_carry[0]] = s_carry[1];
    A.out_counts[0] = s_carry[0];
    A.out_counts[1] = s_carry[1];
  }
}

}  // namespace

}  // namespace dan

using namespace dan;

extern "C" {

int dan_gt_handoff(const float* gt_boxes, const int32_t* gt_offsets, const float* patch_hw, const uint8_t* mirror,
                   int32_t batch, int32_t total_gt, float target_height, float target_width, float min_height,
                   float min_width, float* out_gt_boxes, int32_t* out_gt_offsets, int32_t* out_image_index,
                   int32_t* out_counts, void* stream) {
  DAN_REQUIRE(batch >= 0 && total_gt >= 0, DAN_ERR_INVALID_ARGUMENT, "negative size");
  DAN_REQUIRE(out_gt_offsets && out_counts, DAN_ERR_INVALID_ARGUMENT, "NULL output");
  DAN_REQUIRE(batch == 0 || (gt_offsets && patch_hw && out_image_index), DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(total_gt == 0 || (gt_boxes && out_gt_boxes && aligned16(gt_boxes) && aligned16(out_gt_boxes)), DAN_ERR_INVALID_ARGUMENT,
              "gt boxes NULL or not 16-byte aligned");
  HandoffArgs A = {};
  A.boxes = reinterpret_cast<const float4*>(gt_boxes);
  A.offsets = gt_offsets;
  A.patch_hw = patch_hw;
  A.mirror = mirror;
  A.batch = batch;
  A.target_h = target_height;
  A.target_w = target_width;
  A.min_h = min_height;
  A.min_w = min_width;
  A.out_boxes = reinterpret_cast<float4*>(out_gt_boxes);
  A.out_offsets = out_gt_offsets;
  A.out_image = out_image_index;
  A.out_counts = out_counts;
  gt_handoff_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A);
  DAN_LAUNCH_CHECK("gt_handoff_kernel");
  return DAN_OK;
}

}  // extern "C"
