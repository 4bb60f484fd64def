// deformMesh through the drop-in C++ shim (include/kimera_pgmo_b200/deformation_graph.h) at
// configs[2] size: a 10 M-vertex pcl::PolygonMesh (32-byte XYZRGBA records), 50 k control
// points, k = 4 — the call the reference's optimizeFullMesh makes
// (kimera_pgmo_interface.cpp:558-573 -> deformation_graph.cpp:532-559).  Host wall clock per
// call, everything the caller pays included: blob -> cloud copy, SoA gather, H2D, kernels, D2H,
// scatter, cache copy, cloud -> blob copy, polygon copy.  Prints one JSON line.
// Build: g++ -std=c++17 -O2 -Iinclude scripts/bench_shim.cpp -Lkimera_pgmo_b200/lib -lpgmo_b200
//        -Wl,-rpath,$PWD/kimera_pgmo_b200/lib -lpthread -o build/bench_shim
// Recorded session instead of the synthetic mesh: PGMO_BENCH_PLY=<mesh with stamps .ply>
// PGMO_BENCH_DGRF=<DeformationGraph::save dump> [PGMO_BENCH_PREFIX=<vertex prefix, default 's'>]
// (kimera_pgmo_b200/io.h reads both the way the reference does).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "kimera_pgmo_b200/io.h"

using namespace kimera_pgmo_b200;

static double nowMs() {
  using clock = std::chrono::steady_clock;
  return std::chrono::duration<double, std::milli>(clock::now().time_since_epoch()).count();
}

// closed 100 m x 50 m rectangle walked at 1 m/s (kimera_pgmo_b200/synth.py:_trajectory)
static void keyframe(size_t i, double* x, double* y) {
  const double s = static_cast<double>(i % 300);
  *x = s < 100 ? s : s < 150 ? 100.0 : s < 250 ? 250.0 - s : 0.0;
  *y = s < 100 ? 0.0 : s < 150 ? s - 100.0 : s < 250 ? 50.0 : 300.0 - s;
}

// deformMesh of a recorded mesh under the optimised values of a recorded graph
static int runRecordedSession(const char* ply, const char* dgrf, char prefix, int calls) {
  pcl::PolygonMesh mesh;
  std::vector<Timestamp> stamps;
  ReadMeshWithStampsFromPly(ply, mesh, &stamps);
  DeformationGraphMeshPath graph;
  gtsam::Values values;
  const GraphFileSummary summary = loadGraphFile(dgrf, &graph, &values);
  const size_t n_v = mesh.cloud.width;
  if (stamps.size() != n_v) stamps.assign(n_v, 0);  // a mesh saved without stamps
  const std::vector<int> graph_indices(n_v, 0);
  double best_ms = 1e30, first_ms = 0;
  for (int c = 0; c < calls; ++c) {
    graph.setRecalculateVertices(true);
    const double t = nowMs();
    const pcl::PolygonMesh out = graph.deformMesh(mesh, stamps, graph_indices, prefix, values, 4);
    const double ms = nowMs() - t;
    if (c == 0) first_ms = ms;
    else best_ms = ms < best_ms ? ms : best_ms;
  }
  std::printf("{\"what\": \"deformMesh of a recorded session through the C++ shim\", \"ply\": \"%s\", "
              "\"dgrf\": \"%s\", \"prefix\": \"%c\", \"vertices\": %zu, \"faces\": %zu, "
              "\"controls\": %zu, \"nodes\": %zu, \"first_call_ms\": %.3f, \"best_ms\": %.3f}\n",
              ply, dgrf, prefix, n_v, mesh.polygons.size(), graph.getNumVertices(prefix), summary.nodes,
              first_ms, calls > 1 ? best_ms : first_ms);
  return 0;
}

int main(int argc, char** argv) {
  if (const char* ply = std::getenv("PGMO_BENCH_PLY")) {
    const char* dgrf = std::getenv("PGMO_BENCH_DGRF");
    if (!dgrf) {
      std::printf("{\"error\": \"PGMO_BENCH_PLY needs PGMO_BENCH_DGRF\"}\n");
      return 2;
    }
    if (pgmo_device_count() <= 0) {
      std::printf("{\"error\": \"no CUDA device: the shim has no CPU fallback\"}\n");
      return 77;
    }
    const char* pre = std::getenv("PGMO_BENCH_PREFIX");
    return runRecordedSession(ply, dgrf, pre ? pre[0] : 's', argc > 3 ? std::atoi(argv[3]) : 4);
  }
  const size_t n_v = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 10000000;
  const size_t n_c = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 50000;
  const int calls = argc > 3 ? std::atoi(argv[3]) : 4;
  const bool setup_only = std::getenv("PGMO_BENCH_SETUP_ONLY") != nullptr;  // host-side dry run
  const double t_setup = nowMs();
  if (!setup_only && pgmo_device_count() <= 0) {
    std::printf("{\"error\": \"no CUDA device: the shim has no CPU fallback\"}\n");
    return 77;
  }
  const size_t per_kf_c = 10, n_kf = n_c / per_kf_c, per_kf_v = (n_v + n_kf - 1) / n_kf;
  const Timestamp t0 = 1700000000000000000ull;
  uint32_t lcg = 42u;
  auto unit = [&lcg]() {
    lcg = lcg * 1664525u + 1013904223u;
    return static_cast<float>(lcg >> 8) / 16777216.0f;
  };

  // PGMO_SHIM_DEVICES: 1 (default) = device 0 only, 0 = every visible device, n = devices 0..n-1
  const int n_devices = std::getenv("PGMO_SHIM_DEVICES") ? std::atoi(std::getenv("PGMO_SHIM_DEVICES")) : 1;
  DeformationGraphMeshPath graph(0, n_devices);
  std::vector<gtsam::Point3> ctrl;
  for (size_t j = 0; j < n_c; ++j) {
    double kx, ky;
    keyframe(j / per_kf_c, &kx, &ky);
    const gtsam::Point3 p(kx + 10.0 * unit() - 5.0, ky + 10.0 * unit() - 5.0, 3.0 * unit());
    ctrl.push_back(p);
    graph.addControlPoint('a', p, t0 + (j / per_kf_c) * 1000000000ull);
  }
  auto solve = [&](double scale) {  // a loop-closure-like correction ramping along the path
    gtsam::Values values;
    for (size_t j = 0; j < n_c; ++j) {
      const double ramp = scale * static_cast<double>(j) / static_cast<double>(n_c);
      values.insert(gtsam::Symbol('a', j),
                    gtsam::Pose3(gtsam::Rot3::Rz(0.087 * ramp),
                                 gtsam::Point3(ctrl[j].x() + 0.6 * ramp, ctrl[j].y() + 0.8 * ramp,
                                               ctrl[j].z())));
    }
    return values;
  };

  pcl::PolygonMesh mesh;
  std::vector<Timestamp> stamps(n_v);
  std::vector<int> graph_indices(n_v, 0);
  {
    pcl::PointCloud<pcl::PointXYZRGBA> cloud;
    cloud.points.resize(n_v);
    for (size_t i = 0; i < n_v; ++i) {
      double kx, ky;
      keyframe(i / per_kf_v, &kx, &ky);
      pcl::PointXYZRGBA& p = cloud.points[i];
      p.x = static_cast<float>(kx) + 12.0f * unit() - 6.0f;
      p.y = static_cast<float>(ky) + 12.0f * unit() - 6.0f;
      p.z = 3.0f * unit();
      p.rgba = lcg;
      stamps[i] = t0 + (i / per_kf_v) * 1000000000ull;
    }
    pcl::toPCLPointCloud2(cloud, mesh.cloud);
    mesh.polygons.resize(n_v / 3);
    for (size_t f = 0; f < mesh.polygons.size(); ++f) {
      const uint32_t a = static_cast<uint32_t>(3 * f);
      mesh.polygons[f].vertices = {a, a + 1, a + 2};
    }
  }

  double polygon_copy_ms = nowMs();
  {
    const std::vector<pcl::Vertices> copy = mesh.polygons;  // what deformMesh does per call
    polygon_copy_ms = nowMs() - polygon_copy_ms + 0.0 * copy.size();
  }
  if (setup_only) {
    std::printf("{\"setup_ms\": %.1f, \"polygon_copy_ms\": %.1f}\n", nowMs() - t_setup,
                polygon_copy_ms);
    return 0;
  }
  double first_ms = 0, best_ms = 1e30, sum_ms = 0, moved = 0;
  for (int c = 0; c < calls; ++c) {
    const gtsam::Values values = solve(1.0 + 0.1 * c);
    graph.setRecalculateVertices(true);  // a new solve: full recompute (start index 0)
    const double t = nowMs();
    const pcl::PolygonMesh out = graph.deformMesh(mesh, stamps, graph_indices, 'a', values, 4);
    const double ms = nowMs() - t;
    if (c == 0) {
      first_ms = ms;
    } else {
      best_ms = ms < best_ms ? ms : best_ms;
      sum_ms += ms;
    }
    const auto& a = out.cloud.data[n_v - 1];
    const auto& b = mesh.cloud.data[n_v - 1];
    moved = std::sqrt(double(a.x - b.x) * (a.x - b.x) + double(a.y - b.y) * (a.y - b.y));
  }
  const double mean_ms = calls > 1 ? sum_ms / (calls - 1) : first_ms;

  // incremental update (no new solve since the last call: recalculate_vertices_ is false): only
  // vertices newer than the last control stamp - tol_t are deformed, the rest comes from the cache
  // (deformation_graph.h:715-760).  On a cloud, so that no mesh copy hides the path.
  double incremental_ms = 0;
  size_t incremental_vertices = 0;
  {
    pcl::PointCloud<pcl::PointXYZRGBA> cloud, out;
    pcl::fromPCLPointCloud2(mesh.cloud, cloud);
    out = cloud;
    const gtsam::Values values = solve(1.0);
    graph.setRecalculateVertices(true);
    graph.deformPoints(out, cloud, stamps, 'a', values, 4, 10.0, &graph_indices);  // fills the cache
    const Timestamp newest = t0 + (n_kf - 1) * 1000000000ull;
    for (size_t i = 0; i < n_v; ++i) incremental_vertices += stamps[i] + 10000000000ull > newest;
    std::vector<int> no_graph_index(n_v, -1);  // nothing is rigidly predeformed: all go to the knn
    double best = 1e30;
    for (int c = 0; c < 3; ++c) {
      const double t = nowMs();
      graph.deformPoints(out, cloud, stamps, 'a', values, 4, 10.0, &no_graph_index);
      const double ms = nowMs() - t;
      best = ms < best ? ms : best;
    }
    incremental_ms = best;
  }
  std::printf(
      "{\"what\": \"DeformationGraphMeshPath::deformMesh through the C++ shim, host wall clock\", "
      "\"vertices\": %zu, \"controls\": %zu, \"k\": 4, \"faces\": %zu, \"first_call_ms\": %.2f, "
      "\"mean_ms\": %.2f, \"best_ms\": %.2f, \"polygon_copy_ms\": %.2f, \"vertices_per_s\": %.4g, "
      "\"shim_threads\": %d, \"devices\": %d, \"incremental_deformPoints_ms\": %.2f, "
      "\"incremental_vertices\": %zu, "
      "\"last_vertex_moved_m\": %.4f}\n",
      n_v, n_c, mesh.polygons.size(), first_ms, mean_ms, best_ms, polygon_copy_ms,
      n_v / (mean_ms * 1e-3),
      KIMERA_PGMO_B200_SHIM_THREADS, n_devices == 1 ? 1 : (n_devices == 0 ? pgmo_device_count() : n_devices),
      incremental_ms, incremental_vertices, moved);
  return 0;
}
