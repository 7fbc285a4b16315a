// Binary max-heap with exactly the element movement of libstdc++'s
// std::push_heap / std::pop_heap (bits/stl_heap.h: __push_heap, __adjust_heap,
// __pop_heap) for a comparator that orders on the key only.
//
// Why: stage 3 of SmallMiningMatch (cpp/ExtraLib/small_mining_match.cc:199-222)
// pops a std::priority_queue<DistancePair> whose operator< looks at the overlap
// only (:56-63).  Among EQUAL overlaps the pop order is whatever libstdc++'s heap
// happens to produce, so matching the reference bit-for-bit on a tie that
// straddles the min_match cut requires reproducing those moves, not just "a" heap.
// The fast path of the kernel never needs this (distinct keys => pop order is the
// descending sort); this code only runs on straddling ties.
//
// The functions are __host__ __device__ so tests/test_heap_order.py can check them
// on the CPU against the real std::priority_queue.
#pragma once

#ifndef DAN_HEAP_HD
#ifdef __CUDACC__
#define DAN_HEAP_HD __host__ __device__ inline
#else
#define DAN_HEAP_HD inline
#endif
#endif

namespace dan {

struct HeapItem {
  float key;
  int id;
};

// __push_heap: sift `value` up from `hole` while parent.key < value.key (strict).
DAN_HEAP_HD void heap_sift_up(HeapItem* h, int hole, int top, HeapItem value) {
  int parent = (hole - 1) / 2;
  while (hole > top && h[parent].key < value.key) {
    h[hole] = h[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h[hole] = value;
}

// priority_queue::push == push_back + push_heap
DAN_HEAP_HD void heap_push(HeapItem* h, int& len, HeapItem value) {
  ++len;
  heap_sift_up(h, len - 1, 0, value);
}

// priority_queue::pop == pop_heap + pop_back.  Returns the popped (top) item.
DAN_HEAP_HD HeapItem heap_pop(HeapItem* h, int& len) {
  const HeapItem top = h[0];
  if (len > 1) {
    const int n = len - 1;          // heap length after the pop
    const HeapItem value = h[n];    // *result saved, *result = *first
    // __adjust_heap(first, 0, n, value)
    int hole = 0;
    int child = 0;
    while (child < (n - 1) / 2) {
      child = 2 * (child + 1);
      if (h[child].key < h[child - 1].key) --child;   // right < left -> take left
      h[hole] = h[child];
      hole = child;
    }
    if ((n & 1) == 0 && child == (n - 2) / 2) {
      child = 2 * (child + 1);
      h[hole] = h[child - 1];
      hole = child - 1;
    }
    heap_sift_up(h, hole, 0, value);
  }
  --len;
  return top;
}

}  // namespace dan
