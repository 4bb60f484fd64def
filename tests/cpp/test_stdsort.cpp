// Checks pgmo::stdsort::sort against the toolchain's std::sort on tie-heavy inputs: the
// permutation of equal keys must be identical (payload compared), for every size 0..N.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "stdsort_restated.cuh"

struct Entry {
  int idx;
  float dist;
  bool operator<(const Entry& o) const { return dist < o.dist; }
};

int main() {
  std::mt19937 rng(1234);
  long cases = 0;
  for (int n = 0; n <= 400; ++n) {
    for (int rep = 0; rep < 60; ++rep) {
      const int distinct = 1 + rng() % (n / 3 + 2);
      std::vector<Entry> a(n);
      for (int i = 0; i < n; ++i) a[i] = {i, static_cast<float>(rng() % distinct)};
      if (rep % 4 == 1) {  // PCL-like: sorted prefix + appended tail
        const int pre = std::min(n, 6);
        std::stable_sort(a.begin(), a.begin() + pre);
      }
      if (rep % 4 == 2) std::stable_sort(a.begin(), a.end());
      if (rep % 4 == 3) {
        std::stable_sort(a.begin(), a.end());
        std::reverse(a.begin(), a.end());
      }
      std::vector<Entry> b = a;
      std::sort(a.begin(), a.end());
      pgmo::stdsort::sort(b.data(), b.data() + b.size(),
                          [](const Entry& x, const Entry& y) { return x.dist < y.dist; });
      for (int i = 0; i < n; ++i) {
        if (a[i].idx != b[i].idx) {
          std::printf("MISMATCH n=%d rep=%d at %d\n", n, rep, i);
          return 1;
        }
      }
      ++cases;
    }
  }
  // adversarial: organ-pipe / many-duplicates inputs at larger sizes (can hit the heap fallback)
  for (int n : {1000, 5000, 20000}) {
    for (int pattern = 0; pattern < 4; ++pattern) {
      std::vector<Entry> a(n);
      for (int i = 0; i < n; ++i) {
        float v = 0;
        if (pattern == 0) v = static_cast<float>(i < n / 2 ? i : n - i);
        if (pattern == 1) v = static_cast<float>(i % 3);
        if (pattern == 2) v = static_cast<float>((i * 7919) % 11);
        if (pattern == 3) v = static_cast<float>(n - i);
        a[i] = {i, v};
      }
      std::vector<Entry> b = a;
      std::sort(a.begin(), a.end());
      pgmo::stdsort::sort(b.data(), b.data() + b.size(),
                          [](const Entry& x, const Entry& y) { return x.dist < y.dist; });
      for (int i = 0; i < n; ++i) {
        if (a[i].idx != b[i].idx) {
          std::printf("MISMATCH big n=%d pattern=%d at %d\n", n, pattern, i);
          return 1;
        }
      }
      ++cases;
    }
  }
  std::printf("stdsort restatement identical to std::sort on %ld cases\n", cases);
  return 0;
}
