// Block-wide top-k + sort of 64-bit keys in shared memory (1024-thread CTAs), shared by the postprocess and the
// evaluation-merge kernels.  Include it INSIDE `namespace dan { ... }` of a .cu file, after common.cuh (the optional
// DAN_PHASE timing macro of the including file is used when it is defined).
#pragma once

#ifndef DAN_PHASE
#define DAN_PHASE(slot) do { } while (0)
#endif

constexpr int kSortCap = 8192;      // keys sorted in shared memory (64 KB)
constexpr int kSortThreads = 1024;

// order-preserving float <-> uint32 (descending key order = descending score)
DAN_D uint32_t score_to_key(float s) { return (uint32_t)float_to_ordered(s) ^ 0x80000000u; }
DAN_D float key_to_score(uint32_t k) { return ordered_to_float((int)(k ^ 0x80000000u)); }

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

// compare-exchange stage at distance ST (1, 2 or 4) inside a thread's 8 keys; all register indices are static
template <int ST>
DAN_D void reg_stage(unsigned long long (&r)[8], int lsize, bool desc_t) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if ((e & ST) == 0) {
      const bool desc = (lsize >= 3) ? desc_t : (((e >> lsize) & 1) == 0);
      const unsigned long long x = r[e], y = r[e | ST];
      if ((x < y) == desc) { r[e] = y; r[e | ST] = x; }
    }
  }
}

// Bitonic sort (descending) of P = 2^lp2 >= 256 64-bit keys with the keys held in REGISTERS: thread t owns the 8
// consecutive keys 8t..8t+7.  Compare-exchange partners at distance 1, 2, 4 are in the same thread, at distance
// 8..128 in another lane of the same warp (shfl.xor), and only distances >= 256 go through shared memory, written
// transposed ([e][thread]) so that both the store and the partner's load are conflict free.  The plain shared-memory
// version is bandwidth bound (4 x 8 B accesses per compare-exchange, ~800 wavefronts per stage for 4096 keys).
DAN_D void bitonic_sort_regs(unsigned long long* s_keys, int lp2) {
  const int tid = threadIdx.x;
  const int T = 1 << (lp2 - 3);                 // threads that own keys
  const bool active = tid < T;
  unsigned long long r[8];
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = s_keys[8 * tid + e];
  }
  for (int lsize = 1; lsize <= lp2; ++lsize) {
    // direction of the merge this key takes part in: descending iff bit `lsize` of its index is 0
    const bool desc_t = ((tid >> (lsize >= 3 ? lsize - 3 : 0)) & 1) == 0;
    for (int ls = lsize - 1; ls >= 0; --ls) {
      if (ls >= 8) {
        __syncthreads();
        if (active) {
#pragma unroll
          for (int e = 0; e < 8; ++e) s_keys[e * T + tid] = r[e];
        }
        __syncthreads();
        if (active) {
          const int partner = tid ^ (1 << (ls - 3));
          const bool keep_max = ((tid & (1 << (ls - 3))) == 0) == desc_t;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const unsigned long long o = s_keys[e * T + partner];
            r[e] = ((r[e] < o) == keep_max) ? o : r[e];      // keys are unique: max takes o iff r < o, min iff r > o
          }
        }
      } else if (!active) {
        // warps that own no keys only take part in the barriers above (T is a multiple of 32: warp-uniform)
      } else if (ls >= 3) {
        const int lmask = 1 << (ls - 3);
        const bool keep_max = ((tid & lmask) == 0) == desc_t;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, r[e], lmask);
          r[e] = ((r[e] < o) == keep_max) ? o : r[e];
        }
      } else if (ls == 2) {
        reg_stage<4>(r, lsize, desc_t);
      } else if (ls == 1) {
        reg_stage<2>(r, lsize, desc_t);
      } else {
        reg_stage<1>(r, lsize, desc_t);
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_keys[8 * tid + e] = r[e];
  }
  __syncthreads();
}

// Sorts the m <= kSortCap keys in s_keys[0, m) in place, descending (all threads of the CTA call it; the slots up to
// the next power of two are overwritten with 0).
DAN_D void sort_smem_keys(unsigned long long* s_keys, int m) {
  const int tid = threadIdx.x;
  // pad to a power of two with 0 (smaller than any real key: the low word of a real key is ~index != 0)
  int lp2 = 0;
  while ((1 << lp2) < m) ++lp2;
  const int p2 = 1 << lp2;
  for (int i = m + tid; i < p2; i += kSortThreads) s_keys[i] = 0ull;
  __syncthreads();
  if (lp2 >= 8) {
    bitonic_sort_regs(s_keys, lp2);
    return;
  }
  // small lists: plain bitonic sort in shared memory, descending; strides are powers of two -> shifts only
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
}

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
  DAN_PHASE(8);
  sort_smem_keys(s_keys, m);
  return m;
}



