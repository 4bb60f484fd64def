// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Pins oracle/pcl_octree_restated.h against REAL PCL.  The reference delegates the arithmetic
// that SURVEY.md §8(c) lists as "parity unpinned" to pcl::octree::OctreePointCloudSearch
// (kimera_pgmo/src/mesh_deformation.cpp:24-50: addPointFromCloud, deleteVoxelAtPoint,
// nearestKSearch, getLeafCount; kimera_pgmo/src/compression/octree_compression.cpp:17-81 and
// include/kimera_pgmo/compression/octree_compression.h:20-27: getLeafCount, getBoundingBox,
// getResolution, approxNearestSearch, addPointsFromInputCloud).  PCL does not exist in the build
// image, so this program is written to be compiled twice from the same source:
//
//   with real PCL (any ROS box):   cmake -S oracle/pcl_fixtures -B /tmp/pclfix && cmake --build /tmp/pclfix
//                                  /tmp/pclfix/make_fixtures generate tests/golden/pcl_fixtures.txt
//   with the restatement (here):   make -C oracle   (binary oracle/pcl_fixtures/fixtures_restated)
//                                  oracle/pcl_fixtures/fixtures_restated check tests/golden/pcl_fixtures.txt
//
// `generate` runs every case below on the backend it was compiled with and writes inputs AND
// results, floats as bit patterns; `check` replays the inputs of a dump on its own backend and
// compares result by result.  tests/test_pcl_fixtures.py runs `check` on the committed dump when
// there is one (otherwise it skips with the two commands above) and always round-trips a dump
// through the restatement to prove that the format and the replayer work.
//
// Case families = SURVEY.md §8(c) (i)-(iii):
//   ties      lattice control points and lattice / mid-point queries: K-boundary winners and
//             the order of equal distances out of nearestKSearch
//   voxels    several points per voxel (leaf count vs point count), > 16 candidates in one leaf
//             (std::sort's unstable range), deleteVoxelAtPoint of shared voxels, queries after
//   growth    bounding-box growth in every direction, points on / within FLT_EPSILON of the box
//             faces, getBoundingBox after every insertion, the InOctreeBoundingBox test
//   approx    approxNearestSearch greedy descent incl. equidistant children, at resolution 1.0
//             and 0.05, after re-anchoring (a new octree over the survivors)
//   random    seeded random clouds and queries at both resolutions, K = 1 .. 7
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#ifdef PGMO_FIXTURES_WITH_PCL
#include <pcl/octree/octree_search.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
#include "../pcl_octree_restated.h"
#endif

namespace {

struct P3 {
  float x, y, z;
};

uint32_t bitsOf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
float floatOf(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
uint64_t bitsOf(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return u;
}

#ifdef PGMO_FIXTURES_WITH_PCL
class Backend {
 public:
  explicit Backend(double resolution) : tree_(resolution), cloud_(new pcl::PointCloud<pcl::PointXYZ>()) {
    tree_.setInputCloud(cloud_);
  }
  static const char* name() { return "pcl"; }
  int push(const P3& p) {  // SearchTree::Impl::addPoint without the insertion
    cloud_->push_back(pcl::PointXYZ(p.x, p.y, p.z));
    return static_cast<int>(cloud_->size()) - 1;
  }
  void add(int idx) { tree_.addPointFromCloud(idx, nullptr); }
  void addAll() { tree_.addPointsFromInputCloud(); }
  void del(int idx) { tree_.deleteVoxelAtPoint(idx); }
  size_t leaves() { return tree_.getLeafCount(); }
  void bbox(double b[6]) { tree_.getBoundingBox(b[0], b[1], b[2], b[3], b[4], b[5]); }
  void knn(const P3& q, int K, std::vector<int>& idx, std::vector<float>& d2) {
    tree_.nearestKSearch(pcl::PointXYZ(q.x, q.y, q.z), K, idx, d2);
  }
  void approx(const P3& q, int& idx, float& d2) { tree_.approxNearestSearch(pcl::PointXYZ(q.x, q.y, q.z), idx, d2); }

 private:
  pcl::octree::OctreePointCloudSearch<pcl::PointXYZ> tree_;
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_;
};
#else
class Backend {
 public:
  explicit Backend(double resolution) : tree_(resolution) { tree_.setInputCloud(&cloud_); }
  static const char* name() { return "restated"; }
  int push(const P3& p) {
    pgmo_oracle::PointXYZ q;
    q.x = p.x;
    q.y = p.y;
    q.z = p.z;
    cloud_.push_back(q);
    return static_cast<int>(cloud_.size()) - 1;
  }
  void add(int idx) { tree_.addPointFromCloud(idx); }
  void addAll() { tree_.addPointsFromInputCloud(); }
  void del(int idx) { tree_.deleteVoxelAtPoint(idx); }
  size_t leaves() { return tree_.getLeafCount(); }
  void bbox(double b[6]) { tree_.getBoundingBox(b[0], b[1], b[2], b[3], b[4], b[5]); }
  void knn(const P3& q, int K, std::vector<int>& idx, std::vector<float>& d2) {
    pgmo_oracle::PointXYZ p;
    p.x = q.x;
    p.y = q.y;
    p.z = q.z;
    tree_.nearestKSearch(p, static_cast<uint32_t>(K), idx, d2);
  }
  void approx(const P3& q, int& idx, float& d2) {
    pgmo_oracle::PointXYZ p;
    p.x = q.x;
    p.y = q.y;
    p.z = q.z;
    tree_.approxNearestSearch(p, idx, d2);
  }

 private:
  pgmo_oracle::OctreeSearchRestated tree_;
  std::vector<pgmo_oracle::PointXYZ> cloud_;
};
#endif

// ---- operations.  One line each: "<op> <inputs> | <results>"; a case starts with "case".
//   case <name> <resolution f64 bits>
//   push <x> <y> <z>            point appended to the cloud, not inserted
//   add <idx>                   addPointFromCloud
//   addall                      addPointsFromInputCloud
//   del <idx>                   deleteVoxelAtPoint
//   leaves | <n>
//   bbox | <6 f64 bits>
//   inbox <x> <y> <z> | <0/1>   InOctreeBoundingBox (octree_compression.h:20-27)
//   knn <x> <y> <z> <K> | <n> <idx d2bits>*
//   approx <x> <y> <z> | <idx> <d2bits>     (idx -1: the search did not set it)
struct Runner {
  std::unique_ptr<Backend> t;
  std::ostringstream out;
  uint64_t n_results = 0;

  void begin(const std::string& name, double resolution) {
    t.reset(new Backend(resolution));
    out << "case " << name << " " << bitsOf(resolution) << "\n";
  }
  int push(P3 p) {
    out << "push " << bitsOf(p.x) << " " << bitsOf(p.y) << " " << bitsOf(p.z) << "\n";
    return t->push(p);
  }
  void add(int idx) {
    out << "add " << idx << "\n";
    t->add(idx);
  }
  void pushAdd(P3 p) { add(push(p)); }
  void addAll() {
    out << "addall\n";
    t->addAll();
  }
  void del(int idx) {
    out << "del " << idx << "\n";
    t->del(idx);
  }
  void leaves() {
    out << "leaves | " << t->leaves() << "\n";
    ++n_results;
  }
  void bbox() {
    double b[6];
    t->bbox(b);
    out << "bbox |";
    for (double v : b) out << " " << bitsOf(v);
    out << "\n";
    ++n_results;
  }
  void inbox(P3 p) {
    double b[6];
    t->bbox(b);
    const bool in = (p.x >= b[0] && p.x <= b[3]) && (p.y >= b[1] && p.y <= b[4]) && (p.z >= b[2] && p.z <= b[5]);
    out << "inbox " << bitsOf(p.x) << " " << bitsOf(p.y) << " " << bitsOf(p.z) << " | " << (in ? 1 : 0) << "\n";
    ++n_results;
  }
  void knn(P3 q, int K) {
    std::vector<int> idx;
    std::vector<float> d2;
    if (t->leaves() > 0) t->knn(q, K, idx, d2);  // the reference never queries an empty tree
    out << "knn " << bitsOf(q.x) << " " << bitsOf(q.y) << " " << bitsOf(q.z) << " " << K << " | " << idx.size();
    for (size_t i = 0; i < idx.size(); ++i) out << " " << idx[i] << " " << bitsOf(d2[i]);
    out << "\n";
    ++n_results;
  }
  void approx(P3 q) {
    int idx = -1;
    float d2 = -1.0f;
    if (t->leaves() > 0) t->approx(q, idx, d2);  // inOctree bottoms out on an empty tree first
    out << "approx " << bitsOf(q.x) << " " << bitsOf(q.y) << " " << bitsOf(q.z) << " | " << idx << " " << bitsOf(d2) << "\n";
    ++n_results;
  }
};

struct SplitMix {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  float uniform(float lo, float hi) {
    return lo + (hi - lo) * static_cast<float>(next() >> 40) * (1.0f / 16777216.0f);
  }
  int below(int n) { return static_cast<int>(next() % static_cast<uint64_t>(n)); }
};

void generateAll(Runner& r) {
  // ---- ties: 0.5 m lattice, one point per second node, K-boundary and equal distances
  for (double res : {1.0, 0.05}) {
    const float s = static_cast<float>(res) * 0.5f;
    r.begin(res == 1.0 ? "ties_res1" : "ties_res005", res);
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j)
        for (int k = 0; k < 3; ++k) r.pushAdd({s * 2 * i, s * 2 * j, s * 2 * k});
    r.leaves();
    r.bbox();
    for (int i = 0; i < 11; ++i)
      for (int j = 0; j < 11; j += 2)
        for (int k = 0; k < 5; k += 2)
          for (int K : {1, 3, 6, 7}) r.knn({s * i, s * j, s * k}, K);
    for (int i = 0; i < 11; ++i)
      for (int j = 0; j < 11; ++j) r.approx({s * i, s * j, s});
  }
  // ---- voxels: several points per voxel, deletions of shared voxels, > 16 points in one leaf
  {
    r.begin("voxels_shared", 1.0);
    SplitMix g{11};
    for (int v = 0; v < 40; ++v) {
      const float bx = static_cast<float>(v % 5) * 1.0f, by = static_cast<float>(v / 5) * 1.0f;
      const int n_in = 1 + g.below(4);
      for (int p = 0; p < n_in; ++p) r.pushAdd({bx + g.uniform(0.05f, 0.95f), by + g.uniform(0.05f, 0.95f), g.uniform(0.05f, 0.95f)});
    }
    r.leaves();
    for (int q = 0; q < 60; ++q) r.knn({g.uniform(-1.f, 6.f), g.uniform(-1.f, 9.f), g.uniform(-0.5f, 1.5f)}, 6);
    for (int d = 0; d < 25; ++d) {
      r.del(g.below(60));
      r.leaves();
      r.knn({g.uniform(-1.f, 6.f), g.uniform(-1.f, 9.f), g.uniform(-0.5f, 1.5f)}, 6);
    }
    r.bbox();
    r.begin("voxels_crowded_leaf", 1.0);  // 0.125 m sub-lattice inside two voxels: ties + > 16 candidates
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 2; ++k) r.pushAdd({0.125f * i, 0.125f * j, 0.125f * k});
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.pushAdd({3.0f + 0.125f * i, 0.125f * j, 0.f});
    r.leaves();
    for (int i = 0; i < 32; ++i)
      for (int K : {2, 6, 14, 20}) r.knn({0.125f * i, 0.125f * (i % 8), 0.f}, K);
  }
  // ---- growth: box growth in every direction, points on / near the faces
  for (double res : {1.0, 0.05}) {
    const float s = static_cast<float>(res);
    r.begin(res == 1.0 ? "growth_res1" : "growth_res005", res);
    r.pushAdd({0.3f * s, 0.4f * s, 0.2f * s});
    r.bbox();
    const P3 steps[] = {{1.4f, 0.4f, 0.2f},  {-0.9f, 0.4f, 0.2f}, {0.3f, 2.9f, 0.2f},   {0.3f, -3.3f, 0.2f},
                        {0.3f, 0.4f, 5.1f},  {0.3f, 0.4f, -7.7f}, {9.5f, 9.5f, 9.5f},   {-20.f, 15.f, -3.f},
                        {40.f, -40.f, 40.f}, {0.31f, 0.41f, 0.21f}};
    for (const P3& p : steps) {
      r.pushAdd({p.x * s, p.y * s, p.z * s});
      r.bbox();
      r.leaves();
    }
    // points exactly on and within one float step of the current faces
    double b[6];
    r.t->bbox(b);
    const float faces[2][3] = {{static_cast<float>(b[0]), static_cast<float>(b[1]), static_cast<float>(b[2])},
                               {static_cast<float>(b[3]), static_cast<float>(b[4]), static_cast<float>(b[5])}};
    for (int side = 0; side < 2; ++side)
      for (int ax = 0; ax < 3; ++ax)
        for (int nudge = -1; nudge <= 1; ++nudge) {
          P3 p{0.3f * s, 0.4f * s, 0.2f * s};
          float v = faces[side][ax];
          if (nudge < 0) v = std::nextafter(v, -INFINITY);
          if (nudge > 0) v = std::nextafter(v, INFINITY);
          (ax == 0 ? p.x : ax == 1 ? p.y : p.z) = v;
          r.inbox(p);
        }
    // a point in the last FLT_EPSILON of the box: "outside" for InOctreeBoundingBox, inserted next
    {
      P3 p{static_cast<float>(b[3]), 0.4f * s, 0.2f * s};
      r.inbox(p);
      r.pushAdd(p);
      r.bbox();
      r.leaves();
      r.approx(p);
    }
    SplitMix g{5};
    for (int q = 0; q < 40; ++q) {
      const P3 p{g.uniform(-45.f, 45.f) * s, g.uniform(-45.f, 45.f) * s, g.uniform(-10.f, 45.f) * s};
      r.inbox(p);
      r.approx(p);
      r.knn(p, 3);
    }
  }
  // ---- approx: greedy descent, equidistant children, re-anchoring over the survivors
  for (double res : {1.0, 0.05}) {
    const float s = static_cast<float>(res);
    r.begin(res == 1.0 ? "approx_res1" : "approx_res005", res);
    SplitMix g{23};
    std::vector<P3> pts;
    for (int i = 0; i < 300; ++i) {
      pts.push_back({g.uniform(0.f, 12.f) * s, g.uniform(0.f, 12.f) * s, g.uniform(0.f, 3.f) * s});
      r.pushAdd(pts.back());
    }
    for (int q = 0; q < 150; ++q) r.approx({g.uniform(-1.f, 13.f) * s, g.uniform(-1.f, 13.f) * s, g.uniform(-1.f, 4.f) * s});
    // queries on voxel corners / face centres: equidistant children
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) r.approx({s * i * 1.5f, s * j * 1.5f, s * 1.0f});
    // re-anchoring: a new octree over the points that survive a prune (index order)
    r.begin(res == 1.0 ? "reanchor_res1" : "reanchor_res005", res);
    for (size_t i = 100; i < pts.size(); ++i) r.push(pts[i]);
    r.addAll();
    r.bbox();
    r.leaves();
    for (int q = 0; q < 100; ++q) {
      const P3 p{g.uniform(-1.f, 13.f) * s, g.uniform(-1.f, 13.f) * s, g.uniform(-1.f, 4.f) * s};
      r.inbox(p);
      r.approx(p);
    }
  }
  // ---- random: the deformation window pattern (insert in stamp order, delete from the low end)
  for (double res : {1.0, 0.05}) {
    const float s = static_cast<float>(res);
    r.begin(res == 1.0 ? "random_window_res1" : "random_window_res005", res);
    SplitMix g{42};
    int lo = 0, n = 0;
    for (int step = 0; step < 120; ++step) {
      for (int a = 0; a < 10; ++a) {
        r.pushAdd({(step + g.uniform(-5.f, 5.f)) * s, g.uniform(-5.f, 5.f) * s, g.uniform(0.f, 3.f) * s});
        ++n;
      }
      if (step >= 20) {
        for (int d = 0; d < 10; ++d) r.del(lo++);
      }
      r.leaves();
      for (int q = 0; q < 6; ++q) {
        r.knn({(step - 8 + g.uniform(-6.f, 6.f)) * s, g.uniform(-6.f, 6.f) * s, g.uniform(0.f, 3.f) * s}, 1 + (q + step) % 7);
      }
    }
    (void)n;
  }
}

int check(const std::string& path) {
  std::ifstream in(path);
  if (!in) {
    std::fprintf(stderr, "cannot open %s\n", path.c_str());
    return 2;
  }
  Runner r;
  std::string line, generator;
  uint64_t n_checked = 0, n_bad = 0;
  std::string case_name;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    if (line[0] == '#') {
      if (line.rfind("# backend", 0) == 0) generator = line;
      continue;
    }
    std::istringstream ls(line);
    std::string op;
    ls >> op;
    r.out.str("");
    r.out.clear();
    if (op == "case") {
      uint64_t rb;
      ls >> case_name >> rb;
      double res;
      std::memcpy(&res, &rb, 8);
      r.begin(case_name, res);
    } else if (op == "push") {
      uint32_t a, b, c;
      ls >> a >> b >> c;
      r.push({floatOf(a), floatOf(b), floatOf(c)});
    } else if (op == "add") {
      int i;
      ls >> i;
      r.add(i);
    } else if (op == "addall") {
      r.addAll();
    } else if (op == "del") {
      int i;
      ls >> i;
      r.del(i);
    } else if (op == "leaves") {
      r.leaves();
    } else if (op == "bbox") {
      r.bbox();
    } else if (op == "inbox" || op == "approx") {
      uint32_t a, b, c;
      ls >> a >> b >> c;
      if (op == "inbox") r.inbox({floatOf(a), floatOf(b), floatOf(c)});
      else r.approx({floatOf(a), floatOf(b), floatOf(c)});
    } else if (op == "knn") {
      uint32_t a, b, c;
      int K;
      ls >> a >> b >> c >> K;
      r.knn({floatOf(a), floatOf(b), floatOf(c)}, K);
    } else {
      std::fprintf(stderr, "unknown op in dump: %s\n", line.c_str());
      return 2;
    }
    if (line.find('|') != std::string::npos) {
      ++n_checked;
      std::string mine = r.out.str();
      while (!mine.empty() && mine.back() == '\n') mine.pop_back();
      if (mine != line) {
        if (++n_bad <= 10) {
          std::fprintf(stderr, "MISMATCH in case %s\n  dump : %s\n  %-5s: %s\n", case_name.c_str(), line.c_str(),
                       Backend::name(), mine.c_str());
        }
      }
    }
  }
  std::printf("%s\nchecked %" PRIu64 " results with backend %s: %" PRIu64 " mismatches\n", generator.c_str(), n_checked,
              Backend::name(), n_bad);
  return n_bad == 0 ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc == 3 && std::string(argv[1]) == "generate") {
    Runner r;
    generateAll(r);
    std::ofstream f(argv[2]);
    f << "# pcl octree fixtures (oracle/pcl_fixtures/make_fixtures.cpp); floats as bit patterns\n";
    f << "# backend " << Backend::name();
#ifdef PGMO_FIXTURES_WITH_PCL
    f << " " << PCL_VERSION_PRETTY;
#endif
    f << "\n" << r.out.str();
    std::printf("wrote %s: %" PRIu64 " results, backend %s\n", argv[2], r.n_results, Backend::name());
    return 0;
  }
  if (argc == 3 && std::string(argv[1]) == "check") return check(argv[2]);
  std::fprintf(stderr, "usage: %s generate <out.txt> | check <dump.txt>\n", argv[0]);
  return 2;
}
