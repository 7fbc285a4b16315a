// the evaluation merge the reference's scripts actually use.
//
//   detect_face (eval_sfd.py:95-114, after net.run): boxes / shrink, columns (xmin, ymin, xmax, ymax, score), keep the
//       top min(N - 1, 1.5 * max_per_image) detections by descending score.
//   bbox_vote   (eval_sfd.py:170-210 = eval_dan.py:201-241): a numpy while-loop that repeatedly takes the best remaining
//       detection, removes everything with IoU >= thr to it (+1 pixel convention) and replaces the group by its
//       score-weighted mean; groups of one are dropped; the first max_per_image groups are returned.
//
// bbox_vote without the shrinking array: in descending score order, a detection that no earlier HEAD overlaps with
// IoU >= thr is a head, every other one joins the FIRST head that overlaps it (pinned against the reference's own
// function, tests/golden/vote_reference.npz).  One 1024-thread CTA per image keeps the sorted detections in shared
// memory:
//   1. 64-bit keys (score bits << 32 | index), bitonic sort (sort.cuh); equal scores: higher index first, i.e.
//      numpy's argsort(kind="stable")[::-1] (the reference's default quicksort leaves ties unspecified)
//   2. heads, 32 candidates per round: the round's candidates are resolved against each other with 32x32 pair tests
//      and bit masks, then every thread tests its own (register resident) detections against the round's heads
//   3. the groups with >= 2 members are numbered in head order (block scan); thread q walks the members of group q in
//      score order: box sums accumulate sequentially like np.sum(axis=0), the score sum reproduces numpy's pairwise
//      summation (8 interleaved accumulators up to 128 elements, recursive halving beyond), all in fp32 without FMA.
#include "common.cuh"

namespace dan {

#include "sort.cuh"

namespace {

constexpr uint16_t kFree = 0xFFFFu;      // not assigned yet
constexpr uint16_t kGone = 0xFFFEu;      // a head that does not even match itself (NaN IoU): deleted alone

struct VoteArgs {
  const float* det;           // [B, cap, 5] (xmin, ymin, xmax, ymax, score)
  const int32_t* counts;      // [B] valid rows of each image (NULL: all cap rows)
  int cap;
  float thr;
  int max_out;
  float* out;                 // [B, max_out, 5] zero padded
  int32_t* out_count;         // [B]
  int32_t* out_order;         // [B, cap] optional: sorted position -> input row
  int32_t* out_assign;        // [B, cap] optional: sorted position -> head position (-2: deleted head)
};

// +1 pixel convention, float32, the operation order of eval_sfd.py:176-185
DAN_D float vote_area(const float4& b) { return fmul(fadd(fsub(b.z, b.x), 1.f), fadd(fsub(b.w, b.y), 1.f)); }

DAN_D bool vote_overlaps(const float4& h, float area_h, const float4& b, float thr) {
  const float xx1 = fmaxf(h.x, b.x), yy1 = fmaxf(h.y, b.y), xx2 = fminf(h.z, b.z), yy2 = fminf(h.w, b.w);
  const float w = fmaxf(0.f, fadd(fsub(xx2, xx1), 1.f)), hh = fmaxf(0.f, fadd(fsub(yy2, yy1), 1.f));
  const float inter = fmul(w, hh);
  const float o = fdiv(inter, fsub(fadd(area_h, vote_area(b)), inter));
  return o >= thr;                      // false for NaN, like numpy
}

// Broad phase of the head search: the extent of the detections (+1 pixel on the far sides, the overlap convention above)
// is cut into 32 stripes per axis.  A detection touches a contiguous range of stripes per axis, packed into 20 bits
// (xlo | xhi << 5 | ylo << 10 | yhi << 15); kEverywhere marks a box that must meet every head in the exact test
// (non-finite coordinates, a threshold <= 0 for which disjoint boxes "overlap" too).  vote_overlaps() can only be true
// for thr > 0 when the intersection has positive width and height, i.e. when [x1, x2 + 1] and [y1, y2 + 1] of the two
// boxes intersect - and intervals that intersect share a stripe, because the stripe of a coordinate is a monotone
// function of it.
constexpr uint32_t kEverywhere = 1u << 20;

struct VoteStripes {
  float xlo, xinv, ylo, yinv;
  bool all;
  DAN_D int stripe(float v, float lo, float inv) const { return min(max((int)fmul(fsub(v, lo), inv), 0), 31); }
  DAN_D uint32_t range(const float4& b) const {
    if (all || !(fabsf(b.x) < 1e30f && fabsf(b.y) < 1e30f && fabsf(b.z) < 1e30f && fabsf(b.w) < 1e30f)) return kEverywhere;
    return (uint32_t)stripe(b.x, xlo, xinv) | ((uint32_t)stripe(fadd(b.z, 1.f), xlo, xinv) << 5) |
           ((uint32_t)stripe(b.y, ylo, yinv) << 10) | ((uint32_t)stripe(fadd(b.w, 1.f), ylo, yinv) << 15);
  }
};

// bit s set for every stripe s in [lo, hi]
DAN_D uint32_t stripe_span(uint32_t lo, uint32_t hi) { return ((2u << hi) - 1u) & ~((1u << lo) - 1u); }

// walks the members of one group in score order.  One WARP walks one group, all lanes in lock step (every lane holds
// the same sums): the search for the next member looks at 32 positions at a time (one ballot) instead of one.
struct MemberIter {
  const uint16_t* assign;
  const float4* box;
  const float* score;
  int pos, n;                           // next window of 32 positions to look at
  int base;                             // first position of the window `pending` describes
  unsigned pending;                     // members of the current window that have not been consumed
  uint16_t head;
  float sx, sy, sz, sw, mx;             // sequential sums of box * score (np.sum(axis=0)), running max score
  DAN_D float next() {                  // score of the next member; accumulates its weighted box (warp-uniform)
    const int lane = threadIdx.x & 31;
    while (pending == 0u) {             // (the caller asks for exactly count[head] members, so one more always exists)
      const int p = pos + lane;
      pending = __ballot_sync(0xffffffffu, p < n && assign[p] == head);
      base = pos;
      pos += 32;
    }
    const int i = base + __ffs(pending) - 1;
    pending &= pending - 1u;
    const float4 b = box[i];
    const float s = score[i];
    sx = fadd(sx, fmul(b.x, s));
    sy = fadd(sy, fmul(b.y, s));
    sz = fadd(sz, fmul(b.z, s));
    sw = fadd(sw, fmul(b.w, s));
    mx = fmaxf(mx, s);
    return s;
  }
};

// numpy's pairwise_sum (float32 add-reduce) over the next m member scores.  The leaf (<= 128 elements) is inlined so
// that the iterator stays in registers; the recursive halving above 128 elements runs on a small explicit stack.
DAN_D float numpy_pairwise_leaf(MemberIter& it, int m) {
  if (m < 8) {
    float res = 0.f;
    for (int i = 0; i < m; ++i) res = fadd(res, it.next());
    return res;
  }
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = it.next();
  int i = 8;
  for (; i < m - (m % 8); i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = fadd(r[k], it.next());
  }
  float res = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
  for (; i < m; ++i) res = fadd(res, it.next());
  return res;
}

DAN_D float numpy_pairwise(MemberIter& it, int m) {
  int fm[8], stage[8];                  // depth <= log2(8192 / 128) + 1
  float left[8];
  int sp = 0;
  fm[0] = m; stage[0] = 0; left[0] = 0.f;
  float ret = 0.f;
  while (sp >= 0) {
    const int cm = fm[sp];
    if (cm <= 128) {
      ret = numpy_pairwise_leaf(it, cm);
      --sp;
      continue;
    }
    int m2 = cm / 2;
    m2 -= m2 % 8;
    if (stage[sp] == 0) {
      stage[sp] = 1;
      ++sp; fm[sp] = m2; stage[sp] = 0;
    } else if (stage[sp] == 1) {
      left[sp] = ret;
      stage[sp] = 2;
      ++sp; fm[sp] = cm - m2; stage[sp] = 0;
    } else {
      ret = fadd(left[sp], ret);
      --sp;
    }
  }
  return ret;
}

static size_t vote_smem_bytes(int cap) {
  const size_t pad = (size_t)1 << (32 - __builtin_clz((unsigned)(cap > 1 ? cap - 1 : 1)));     // next power of two >= cap
  const size_t keys = (pad < 256 ? 256 : pad) * 8;
  const size_t boxes = (size_t)cap * 16;
  // [ boxes (aliases the sort keys) | score | assign | count | qlist ]
  return (keys > boxes ? keys : boxes) + (size_t)cap * 4 + align_up((size_t)cap * 2, 16) * 3;
}

__global__ void __launch_bounds__(kSortThreads, 1) bbox_vote_kernel(const VoteArgs A) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ int s_scan[kSortThreads / 32];
  __shared__ int s_total;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = A.counts ? min(max(A.counts[b], 0), A.cap) : A.cap;
  const float* det = A.det + (int64_t)b * A.cap * 5;

  size_t pad = 256;
  while ((int)pad < A.cap) pad <<= 1;
  const size_t front = (pad * 8 > (size_t)A.cap * 16) ? pad * 8 : (size_t)A.cap * 16;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dyn_smem);
  float4* box = reinterpret_cast<float4*>(dyn_smem);                     // aliases the keys once they are consumed
  float* score = reinterpret_cast<float*>(dyn_smem + front);
  uint16_t* assign = reinterpret_cast<uint16_t*>(dyn_smem + front + (size_t)A.cap * 4);
  const size_t u16_span = ((size_t)A.cap * 2 + 15) / 16 * 8;             // elements per 16-byte aligned uint16 array
  uint16_t* count = assign + u16_span;
  uint16_t* qlist = count + u16_span;

  // ---- 1. sort by descending score (:171-172)
  for (int i = tid; i < n; i += kSortThreads)
    keys[i] = ((unsigned long long)score_to_key(det[i * 5 + 4]) << 32) | (unsigned long long)(uint32_t)i;
  __syncthreads();
  sort_smem_keys(keys, n);
  // gather through registers (the boxes overwrite the key storage); thread t keeps the boxes of the score ranks
  // t, t + 1024, ... in registers for the sweeps below
  constexpr int kPer = kSortCap / kSortThreads;
  float4 rb[kPer];
  float rs[kPer];
  int rr[kPer];                                            // score rank of the box in rb[e], -1 = none
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    const int r = tid + e * kSortThreads;
    rr[e] = -1;
    if (r < n) {
      const int i = (int)(uint32_t)(keys[r] & 0xFFFFFFFFull);
      rr[e] = r;
      rb[e] = make_float4(det[i * 5], det[i * 5 + 1], det[i * 5 + 2], det[i * 5 + 3]);
      rs[e] = det[i * 5 + 4];
      if (A.out_order) A.out_order[(int64_t)b * A.cap + r] = i;
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    if (rr[e] >= 0) {
      box[rr[e]] = rb[e];
      score[rr[e]] = rs[e];
      assign[rr[e]] = kFree;
      count[rr[e]] = 0;
    }
  }
  __syncthreads();

  // ---- 2. heads in score order (:173-190), 32 candidate heads per round.
  // Thread t owns the score ranks t, t + 1024, ...: an `alive` bit and the packed stripe ranges of each in registers.
  //   a. warp 0 lists the next (up to) 32 unassigned positions = the candidates of this round
  //   b. thread (k = warp, m = lane) tests candidate k against the earlier candidate m  -> 32 row masks
  //   c. every thread resolves the 32 candidates from the masks (bit operations): candidate k is a head unless an earlier
  //      HEAD of the round overlaps it (then it joins the first such head); warp s publishes which of the round's heads
  //      touch x / y stripe s (two ballots)
  //   d. every thread looks its own alive boxes behind the last candidate up in that index - (OR over the box's x
  //      stripes) AND (OR over its y stripes) = the round's heads it can overlap - and runs the exact test on those
  //      only, in head order (the first hit is the head it joins)
  // Four barriers per 32 heads.
  __shared__ int s_cand[32];
  __shared__ unsigned s_mask[32];
  __shared__ unsigned s_selfok;
  __shared__ int s_ncand;
  __shared__ unsigned s_xs[32], s_ys[32];
  __shared__ float s_ext[kSortThreads / 32][4];
  VoteStripes sg;
  {
    float x0 = 3.0e38f, x1 = -3.0e38f, y0 = 3.0e38f, y1 = -3.0e38f;
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      const float4 bx = rb[e];
      if (rr[e] >= 0 && fabsf(bx.x) < 1e30f && fabsf(bx.y) < 1e30f && fabsf(bx.z) < 1e30f && fabsf(bx.w) < 1e30f) {
        x0 = fminf(x0, bx.x); x1 = fmaxf(x1, fadd(bx.z, 1.f));
        y0 = fminf(y0, bx.y); y1 = fmaxf(y1, fadd(bx.w, 1.f));
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, d)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, d));
      y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, d)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, d));
    }
    if (lane == 0) { s_ext[warp][0] = x0; s_ext[warp][1] = x1; s_ext[warp][2] = y0; s_ext[warp][3] = y1; }
    __syncthreads();
    for (int w = 0; w < kSortThreads / 32; ++w) {
      x0 = fminf(x0, s_ext[w][0]); x1 = fmaxf(x1, s_ext[w][1]);
      y0 = fminf(y0, s_ext[w][2]); y1 = fmaxf(y1, s_ext[w][3]);
    }
    sg.xlo = x0; sg.ylo = y0;
    sg.xinv = (x1 > x0) ? fdiv(32.f, fsub(x1, x0)) : 0.f;
    sg.yinv = (y1 > y0) ? fdiv(32.f, fsub(y1, y0)) : 0.f;
    sg.all = !(A.thr > 0.f);
  }
  uint32_t rq[kPer];                                       // packed stripe ranges of the thread's boxes
  unsigned alive = 0;
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    alive |= (rr[e] >= 0) ? (1u << e) : 0u;
    rq[e] = (rr[e] >= 0) ? sg.range(rb[e]) : 0u;
  }
  auto add_count = [&](int head, unsigned k) {
    atomicAdd(reinterpret_cast<unsigned int*>(count) + (head >> 1), k << ((head & 1) * 16));
  };
  int cur = 0;
  while (true) {
    if (warp == 0) {                                       // a.
      int have = 0, p = cur;
      while (p < n && have < 32) {
        const bool un = (p + lane < n) && assign[p + lane] == kFree;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        const int rank = have + __popc(m & ((1u << lane) - 1u));
        if (un && rank < 32) s_cand[rank] = p + lane;
        have += __popc(m);
        p += 32;
      }
      if (lane == 0) s_ncand = min(have, 32);
    }
    __syncthreads();
    const int ncand = s_ncand;
    if (ncand == 0) break;
    const int last = s_cand[ncand - 1];
    {                                                      // b.
      const int k = warp, m = lane;
      bool ov = false, self_ok = false;
      if (k < ncand && m <= k && m < ncand) {
        const float4 bm = box[s_cand[m]];
        const bool hit = vote_overlaps(bm, vote_area(bm), box[s_cand[k]], A.thr);
        ov = hit && m < k;
        self_ok = hit && m == k;
      }
      const unsigned row = __ballot_sync(0xffffffffu, ov);
      const unsigned so = __ballot_sync(0xffffffffu, self_ok);
      if (lane == 0) s_mask[k] = row;
      if (lane == 0 && k == 0) s_selfok = 0u;
      __syncthreads();
      if (lane == 0 && so) atomicOr(&s_selfok, 1u << k);
    }
    __syncthreads();
    unsigned heads = 0u;                                   // c. (redundantly in every thread)
    int joined = -1;                                       // lane k of warp 0 applies candidate k
    for (int k = 0; k < ncand; ++k) {
      const unsigned mk = s_mask[k] & heads;
      if (mk == 0u) heads |= 1u << k;
      if (warp == 0 && lane == k) joined = mk ? __ffs(mk) - 1 : -1;
    }
    if (warp == 0 && lane < ncand) {
      const int pos = s_cand[lane];
      if (joined >= 0) {
        assign[pos] = (uint16_t)s_cand[joined];
        add_count(s_cand[joined], 1u);
      } else if ((s_selfok >> lane) & 1u) {
        assign[pos] = (uint16_t)pos;
        add_count(pos, 1u);
      } else {
        assign[pos] = kGone;                               // NaN self IoU: deleted alone (:189-190)
      }
    }
    {                                                      // warp s: the round's heads that touch x / y stripe s
      uint32_t q = 0u;
      const bool is_head = lane < ncand && ((heads >> lane) & 1u);
      if (is_head) q = sg.range(box[s_cand[lane]]);
      const bool every = (q & kEverywhere) != 0u;
      const unsigned hx = __ballot_sync(0xffffffffu, is_head && (every || ((stripe_span(q & 31u, (q >> 5) & 31u) >> warp) & 1u)));
      const unsigned hy = __ballot_sync(0xffffffffu, is_head && (every || ((stripe_span((q >> 10) & 31u, (q >> 15) & 31u) >> warp) & 1u)));
      if (lane == 0) { s_xs[warp] = hx; s_ys[warp] = hy; }
    }
#pragma unroll
    for (int e = 0; e < kPer; ++e)                         // the candidates themselves are settled
      if (tid + e * kSortThreads <= last) alive &= ~(1u << e);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kPer; ++e) {                       // d.
      if ((alive >> e) & 1u) {
        const uint32_t q = rq[e];
        unsigned poss = heads;
        if (!(q & kEverywhere)) {
          unsigned ax = 0u, ay = 0u;
          for (uint32_t st = q & 31u; st <= ((q >> 5) & 31u); ++st) ax |= s_xs[st];
          for (uint32_t st = (q >> 10) & 31u; st <= ((q >> 15) & 31u); ++st) ay |= s_ys[st];
          poss = ax & ay;
        }
        const int r = tid + e * kSortThreads;              // (the box itself stays in shared memory: it is needed rarely)
        const float4 me = poss ? box[r] : make_float4(0.f, 0.f, 0.f, 0.f);
        while (poss) {                                     // ascending = head order: the first hit is the head it joins
          const int m = __ffs(poss) - 1;
          poss &= poss - 1u;
          const int hp = s_cand[m];
          const float4 h = box[hp];
          if (vote_overlaps(h, vote_area(h), me, A.thr)) {
            assign[r] = (uint16_t)hp;
            alive &= ~(1u << e);
            add_count(hp, 1u);
            poss = 0u;                                     // (no `break`: the lanes of the warp leave the loop together,
                                                           // racecheck saw them reach the barrier apart otherwise)
          }
        }
      }
    }
    cur = last + 1;
    __syncthreads();
  }
  __syncthreads();

  // ---- 3. groups with at least two members, numbered in head order (:191-197, :205)
  const int E = (n + kSortThreads - 1) / kSortThreads;
  int mine_cnt = 0;
  for (int e = 0; e < E; ++e) {
    const int r = tid * E + e;
    if (r < n && count[r] >= 2) ++mine_cnt;
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
  int q = before + incl - mine_cnt;
  for (int e = 0; e < E; ++e) {
    const int r = tid * E + e;
    if (r < n && count[r] >= 2) {
      if (q < A.max_out) qlist[q] = (uint16_t)r;
      ++q;
    }
  }
  if (tid == 0) s_total = min(total, A.max_out);
  __syncthreads();
  const int groups = s_total;
  float* out = A.out + (int64_t)b * A.max_out * 5;
  for (int g = warp; g < A.max_out; g += kSortThreads / 32) {          // one warp per group (warp-uniform control flow)
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, o4 = 0.f;
    if (g < groups) {
      MemberIter it;
      it.assign = assign; it.box = box; it.score = score; it.n = n;
      it.head = qlist[g];
      it.pos = it.head;                 // the head is the first member of its own group (a deleted head has count 0)
      it.base = 0;
      it.pending = 0u;
      it.sx = it.sy = it.sz = it.sw = 0.f;
      it.mx = __int_as_float(0xff800000);
      const float ssum = numpy_pairwise(it, (int)count[it.head]);       // :199-201
      o0 = fdiv(it.sx, ssum); o1 = fdiv(it.sy, ssum); o2 = fdiv(it.sz, ssum); o3 = fdiv(it.sw, ssum);
      o4 = it.mx;                                                       // :200,202
    }
    if (lane == 0) { out[g * 5] = o0; out[g * 5 + 1] = o1; out[g * 5 + 2] = o2; out[g * 5 + 3] = o3; out[g * 5 + 4] = o4; }
  }
  if (tid == 0) A.out_count[b] = groups;
  if (A.out_assign) {
    for (int r = tid; r < n; r += kSortThreads)
      A.out_assign[(int64_t)b * A.cap + r] = assign[r] == kGone ? -2 : (int)assign[r];
  }
}

struct FaceArgs {
  const float4* boxes;        // [n] (ymin, xmin, ymax, xmax)
  const float* scores;        // [n]
  int n;
  float shrink;
  int top;                    // int(1.5 * max_per_image)
  unsigned long long* keys;   // [n] workspace
  float* out;                 // [top, 5] zero padded
  int32_t* out_index;         // [top] optional, -1 padded
  int32_t* out_count;
};

__global__ void __launch_bounds__(256) face_key_kernel(const FaceArgs A) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x)
    A.keys[i] = ((unsigned long long)score_to_key(A.scores[i]) << 32) | (unsigned long long)(uint32_t)i;
}

// eval_sfd.py:101-112
__global__ void __launch_bounds__(kSortThreads, 1) face_select_kernel(const FaceArgs A) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ SortScratch sc;
  const int tid = threadIdx.x;
  const int k = max(min(A.n - 1, A.top), 0);          // N == 0: the reference slices [:-1] of an empty array
  if (k > 0) select_and_sort(A.keys, A.n, k, s_keys, sc);
  for (int r = tid; r < A.top; r += kSortThreads) {
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, o4 = 0.f;
    int idx = -1;
    if (r < k) {
      idx = (int)(uint32_t)(s_keys[r] & 0xFFFFFFFFull);
      const float4 bx = A.boxes[idx];
      o0 = fdiv(bx.y, A.shrink); o1 = fdiv(bx.x, A.shrink); o2 = fdiv(bx.w, A.shrink); o3 = fdiv(bx.z, A.shrink);
      o4 = A.scores[idx];
    }
    A.out[r * 5] = o0; A.out[r * 5 + 1] = o1; A.out[r * 5 + 2] = o2; A.out[r * 5 + 3] = o3; A.out[r * 5 + 4] = o4;
    if (A.out_index) A.out_index[r] = idx;
  }
  if (tid == 0) A.out_count[0] = k;
}

}  // namespace

}  // namespace dan

using namespace dan;

extern "C" {

size_t dan_detect_face_workspace_bytes(int64_t n) { return n < 0 ? 0 : align_up((size_t)n * 8, 256); }

int dan_detect_face_select(const float* bboxes, const float* scores, int64_t n, float shrink, int32_t top, float* out_det,
                           int32_t* out_index, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  DAN_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), DAN_ERR_INVALID_ARGUMENT, "bad n");
  DAN_REQUIRE(top >= 1 && top <= kSortCap, DAN_ERR_UNSUPPORTED, "top must be in [1, %d]", kSortCap);
  DAN_REQUIRE(out_det && out_count, DAN_ERR_INVALID_ARGUMENT, "NULL output");
  DAN_REQUIRE(n == 0 || (bboxes && scores && aligned16(bboxes)), DAN_ERR_INVALID_ARGUMENT, "NULL / misaligned input");
  const size_t need = dan_detect_face_workspace_bytes(n);
  DAN_REQUIRE(n == 0 || (workspace != nullptr && workspace_bytes >= need), DAN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
              need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  FaceArgs A = {};
  A.boxes = reinterpret_cast<const float4*>(bboxes);
  A.scores = scores;
  A.n = (int)n;
  A.shrink = shrink;
  A.top = top;
  A.keys = static_cast<unsigned long long*>(workspace);
  A.out = out_det;
  A.out_index = out_index;
  A.out_count = out_count;
  if (n > 0) {
    face_key_kernel<<<grid_for(n), 256, 0, st>>>(A);
    DAN_LAUNCH_CHECK("face_key_kernel");
  }
  DAN_CUDA(cudaFuncSetAttribute(face_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8));
  face_select_kernel<<<1, kSortThreads, kSortCap * 8, st>>>(A);
  DAN_LAUNCH_CHECK("face_select_kernel");
  return DAN_OK;
}

int dan_bbox_vote(const float* det, const int32_t* counts, int32_t batch, int32_t capacity, float nms_threshold,
                  int32_t max_per_image, float* out_det, int32_t* out_count, int32_t* out_order, int32_t* out_assign,
                  void* stream) {
  DAN_REQUIRE(batch >= 0 && capacity >= 0, DAN_ERR_INVALID_ARGUMENT, "negative size");
  DAN_REQUIRE(capacity <= kSortCap, DAN_ERR_UNSUPPORTED, "more than %d detections per image", kSortCap);
  DAN_REQUIRE(max_per_image >= 1 && max_per_image <= 65535, DAN_ERR_INVALID_ARGUMENT, "max_per_image must be in [1, 65535]");
  if (batch == 0) return DAN_OK;
  DAN_REQUIRE(out_det && out_count && (capacity == 0 || det), DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  VoteArgs A = {};
  A.det = det;
  A.counts = counts;
  A.cap = capacity;
  A.thr = nms_threshold;
  A.max_out = max_per_image;
  A.out = out_det;
  A.out_count = out_count;
  A.out_order = out_order;
  A.out_assign = out_assign;
  const size_t smem = vote_smem_bytes(capacity);
  DAN_REQUIRE(smem <= 221 * 1024, DAN_ERR_UNSUPPORTED, "capacity %d needs %zu bytes of shared memory", capacity, smem);
  DAN_CUDA(cudaFuncSetAttribute(bbox_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
  bbox_vote_kernel<<<batch, kSortThreads, smem, (cudaStream_t)stream>>>(A);
  DAN_LAUNCH_CHECK("bbox_vote_kernel");
  return DAN_OK;
}

}  // extern "C"
