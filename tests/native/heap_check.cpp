// CPU check of dan_b200/csrc/heap_order.cuh against the real std::priority_queue
// with the reference's ordering (overlap only; small_mining_match.cc:56-63).
// Prints "OK <cases>" or the first mismatch.  Built and run by tests/test_heap_order.py.
#include <cstdio>
#include <cstdlib>
#include <queue>
#include <random>
#include <vector>

#include "heap_order.cuh"

struct Pair {
  float dist;
  int id;
  bool operator<(const Pair& o) const { return o.dist > dist; }
};

int main(int argc, char** argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 20000;
  std::mt19937 rng(20180817);
  for (int c = 0; c < cases; ++c) {
    const int n = 1 + rng() % 80;
    const int levels = 1 + rng() % 6;   // few distinct keys => many ties
    std::vector<float> keys(n);
    for (auto& k : keys) k = 0.3f + 0.01f * (rng() % levels);
    std::priority_queue<Pair> pq;
    std::vector<dan::HeapItem> h(n);
    int len = 0;
    for (int i = 0; i < n; ++i) {
      pq.push(Pair{keys[i], i});
      dan::heap_push(h.data(), len, dan::HeapItem{keys[i], i});
    }
    const int pops = 1 + rng() % n;
    for (int p = 0; p < pops; ++p) {
      const Pair t = pq.top();
      pq.pop();
      const dan::HeapItem m = dan::heap_pop(h.data(), len);
      if (t.id != m.id || t.dist != m.key) {
        printf("MISMATCH case %d pop %d: std (%g,%d) ours (%g,%d)\n", c, p, t.dist, t.id, m.key, m.id);
        return 1;
      }
    }
  }
  printf("OK %d\n", cases);
  return 0;
}
