// Host-only parts of the C++ shim (no device needed): detail::parallelFor covers [0, n) exactly
// once whatever the thread count, runs small ranges in place, and rethrows a part's exception
// (the std::out_of_range of a cloud's .at()) on the calling thread; PinnedBlock's carving
// arithmetic (padded / take) never hands out overlapping or misaligned arrays.
#include <atomic>
#include <cstdio>
#include <numeric>
#include <thread>

#include "kimera_pgmo_b200/mesh_deformation.h"

using kimera_pgmo_b200::detail::parallelFor;
using kimera_pgmo_b200::detail::PinnedBlock;

static int g_failures = 0;
#define CHECK(a)                                                      \
  do {                                                                \
    if (!(a)) {                                                       \
      ++g_failures;                                                   \
      std::fprintf(stderr, "%s:%d CHECK(%s)\n", __FILE__, __LINE__, #a); \
    }                                                                 \
  } while (0)

int main() {
  for (size_t n : {size_t(0), size_t(1), size_t(1000), (size_t(1) << 17) - 1, size_t(1) << 18,
                   (size_t(1) << 18) + 7, size_t(3) << 18, (size_t(9) << 17) + 12345}) {
    std::vector<uint8_t> hits(n, 0);
    std::atomic<int> parts{0};
    std::atomic<bool> bad_range{false};
    parallelFor(n, [&](size_t begin, size_t end) {
      ++parts;
      if (begin > end || end > n) bad_range = true;
      for (size_t i = begin; i < end && i < n; ++i) ++hits[i];
    });
    CHECK(!bad_range);
    CHECK(std::accumulate(hits.begin(), hits.end(), size_t(0)) == n);
    bool once = true;
    for (uint8_t h : hits) once = once && h == 1;
    CHECK(once);
    if (n < (size_t(1) << 18)) CHECK(parts == 1);  // below two full shares: in place
    CHECK(parts <= KIMERA_PGMO_B200_SHIM_THREADS);
  }

  // in-place ranges run on the calling thread
  const std::thread::id self = std::this_thread::get_id();
  bool same_thread = false;
  parallelFor(100, [&](size_t, size_t) { same_thread = std::this_thread::get_id() == self; });
  CHECK(same_thread);

  // the first exception of any part reaches the caller, after every part has ended
  const size_t n = size_t(1) << 20;
  std::atomic<size_t> visited{0};
  bool caught = false;
  try {
    parallelFor(n, [&](size_t begin, size_t end) {
      visited += end - begin;
      if (begin <= n / 2 && n / 2 < end) throw std::out_of_range("vertex index");
    });
  } catch (const std::out_of_range& e) {
    caught = std::string(e.what()) == "vertex index";
  }
  CHECK(caught);
  CHECK(visited == n);

  // carving: 256-byte alignment, padded() is what reserve() has to be given per array
  CHECK(PinnedBlock::padded(0) == 0);
  CHECK(PinnedBlock::padded(1) == 256);
  CHECK(PinnedBlock::padded(256) == 256);
  CHECK(PinnedBlock::padded(257) == 512);
  {
    PinnedBlock empty;  // never reserved: nothing to free, take() must refuse
    bool refused = false;
    try {
      empty.take<float>(1);
    } catch (const std::logic_error&) {
      refused = true;
    }
    CHECK(refused);
  }
  std::printf("shim host tests: %d failures\n", g_failures);
  return g_failures ? 1 : 0;
}
