// The reference's own unit tests, re-typed against the drop-in C++ shim
// (include/kimera_pgmo_b200/*.h) with the stand-in PCL / GTSAM types.  Scenario and expected
// values of every test come from the reference:
//   kimera_pgmo/test/test_mesh_deformation.cpp:16-112
//   kimera_pgmo/test/test_deformation_graph.cpp:163-309 (solver outputs set by hand)
//   kimera_pgmo/test/test_voxel_clearing_compression.cpp:106-356
//   kimera_pgmo/test/test_octree_compression.cpp:58-364, test_mesh_frontend_interface.cpp:523-607
// Needs a GPU (runs under `pytest -m gpu`); compiled (not run) by the CPU test-suite.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

// SHIM_TEST_HOST_DOUBLE: the CPU suite's build against the oracle-backed test doubles
// tests/cpp/mock_{deform,compression}_backend.cpp (host logic of the shim; see
// tests/test_shim.py): every test of this file runs there too, without a GPU.
#include "kimera_pgmo_b200/compression.h"
#include "kimera_pgmo_b200/deformation_graph.h"
#include "kimera_pgmo_b200/delta_compression.h"
#include "kimera_pgmo_b200/mesh_frontend_interface.h"

using namespace kimera_pgmo_b200;

static int g_failures = 0, g_checks = 0;
#define EXPECT_TRUE(a)                                                             \
  do {                                                                             \
    ++g_checks;                                                                    \
    if (!(a)) {                                                                    \
      ++g_failures;                                                                \
      std::cerr << __FILE__ << ":" << __LINE__ << " EXPECT_TRUE(" #a ") failed\n"; \
    }                                                                              \
  } while (0)
#define EXPECT_EQ(a, b)                                                                      \
  do {                                                                                       \
    ++g_checks;                                                                              \
    if (!((a) == (b))) {                                                                     \
      ++g_failures;                                                                          \
      std::cerr << __FILE__ << ":" << __LINE__ << " EXPECT_EQ(" #a ", " #b "): " << (a)      \
                << " vs " << (b) << "\n";                                                    \
    }                                                                                        \
  } while (0)
#define EXPECT_NEAR(a, b, tol)                                                               \
  do {                                                                                       \
    ++g_checks;                                                                              \
    if (!(std::fabs(double(a) - double(b)) <= (tol))) {                                      \
      ++g_failures;                                                                          \
      std::cerr << __FILE__ << ":" << __LINE__ << " EXPECT_NEAR(" #a ", " #b "): " << (a)    \
                << " vs " << (b) << "\n";                                                    \
    }                                                                                        \
  } while (0)

// ------------------------------------------------------------------ TestMeshDeformation
static void testDeformPoints() {
  typedef pcl::PointXYZ Point;
  pcl::PointCloud<Point> original_points;
  std::vector<gtsam::Point3> control_points;
  gtsam::Values optimized_values;
  char prefix = 'a';
  for (size_t i = 0; i < 100; i++) {
    original_points.push_back(Point(static_cast<float>(i), 0.0, 0.0));
    if (i % 10 == 0) {
      control_points.push_back(gtsam::Point3(static_cast<double>(i), 0.0, 0.0));
      optimized_values.insert(
          gtsam::Symbol(prefix, i / 10),
          gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(static_cast<double>(i), 1.0, 0.0)));
    }
  }
  std::vector<std::set<size_t>> control_point_map;
  deformation::deformPoints(original_points, control_point_map, original_points, prefix,
                            control_points, {}, optimized_values);
  EXPECT_EQ(size_t(100), original_points.size());
  EXPECT_EQ(size_t(100), control_point_map.size());
  for (size_t i = 0; i < 100; i++) {
    EXPECT_NEAR(static_cast<double>(i), original_points.points[i].x, 1.0e-9);
    EXPECT_NEAR(1.0, original_points.points[i].y, 1.0e-9);
    EXPECT_NEAR(0.0, original_points.points[i].z, 1.0e-9);
  }
}

static void testDeformPointsWithTimeCheck() {
  typedef pcl::PointXYZ Point;
  pcl::PointCloud<Point> original_points;
  std::vector<Timestamp> stamps;
  std::vector<gtsam::Point3> control_points;
  std::vector<Timestamp> control_point_stamps;
  gtsam::Values optimized_values;
  char prefix = 'a';
  for (size_t i = 0; i < 100; i++) {
    original_points.push_back(Point(static_cast<float>(i), 0.0, 0.0));
    if (i % 10 == 0) {
      control_points.push_back(gtsam::Point3(static_cast<double>(i), 0.0, 0.0));
      if (i > 50) {
        optimized_values.insert(
            gtsam::Symbol(prefix, i / 10),
            gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(static_cast<double>(i), 1.0, 0.0)));
        control_point_stamps.push_back(stampFromSec(20.0));
      } else {
        optimized_values.insert(
            gtsam::Symbol(prefix, i / 10),
            gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(static_cast<double>(i), -1.0, 0.0)));
        control_point_stamps.push_back(0);
      }
    }
    stamps.push_back(i < 50 ? 0 : stampFromSec(20.0));
  }
  std::vector<std::set<size_t>> control_point_map;
  const ConstStampedCloud<pcl::PointXYZ> cloud{original_points, stamps};
  deformation::deformPoints(original_points, control_point_map, cloud, prefix, control_points,
                            control_point_stamps, optimized_values, 3, 10.0);
  EXPECT_EQ(size_t(100), original_points.size());
  for (size_t i = 0; i < 50; i++) {
    EXPECT_EQ(static_cast<double>(i), double(original_points.points[i].x));
    EXPECT_EQ(-1.0, double(original_points.points[i].y));
    EXPECT_EQ(0.0, double(original_points.points[i].z));
  }
  for (size_t i = 51; i < 100; i++) {
    EXPECT_EQ(static_cast<double>(i), double(original_points.points[i].x));
    EXPECT_EQ(1.0, double(original_points.points[i].y));
    EXPECT_EQ(0.0, double(original_points.points[i].z));
  }
}

// The reference's silent early returns (mesh_deformation.h:78-86, :96-99, :143-152;
// deformation_graph.h:819-823) and its .at() failures, through the shim.
static pcl::PointXYZRGBA colored(float x, float y, float z, uint8_t r, uint8_t g, uint8_t b);
static void testEarlyReturnsAndErrors() {
  typedef pcl::PointXYZ Point;
  const char prefix = 'a';
  pcl::PointCloud<Point> points;
  std::vector<Timestamp> stamps;
  for (size_t i = 0; i < 20; i++) {
    points.push_back(Point(static_cast<float>(i), 0.5f, 0.0f));
    stamps.push_back(stampFromSec(1.0));
  }
  std::vector<gtsam::Point3> control_points;
  std::vector<Timestamp> control_point_stamps;
  gtsam::Values values;
  for (size_t j = 0; j < 3; j++) {
    control_points.push_back(gtsam::Point3(5.0 * static_cast<double>(j), 0.0, 0.0));
    control_point_stamps.push_back(stampFromSec(1.0));
    values.insert(gtsam::Symbol(prefix, j),
                  gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(5.0 * static_cast<double>(j), 2.0, 0.0)));
  }
  const std::vector<std::set<size_t>> sentinel(1, std::set<size_t>{42});
  auto untouched = [&](const pcl::PointCloud<Point>& out, const std::vector<std::set<size_t>>& map) {
    bool same = out.size() == points.size() && map == sentinel;
    for (size_t i = 0; same && i < out.size(); i++) {
      same = std::memcmp(&out.points[i], &points.points[i], 3 * sizeof(float)) == 0;
    }
    return same;
  };
  const ConstStampedCloud<Point> cloud{points, stamps};

  {  // unstamped overload, fewer occupied voxels than k (:96-99): returns AFTER clear() (:88)
    gtsam::Values two;  // only two of the three control points have a pose
    for (size_t j = 0; j < 2; j++) {
      two.insert(gtsam::Symbol(prefix, j), gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(0.0, 2.0, 0.0)));
    }
    pcl::PointCloud<Point> out = points;
    std::vector<std::set<size_t>> map = sentinel;
    deformation::deformPoints(out, map, points, prefix, control_points, control_point_stamps, two, 3);
    EXPECT_TRUE(map.empty());
    bool same = out.size() == points.size();
    for (size_t i = 0; same && i < out.size(); i++) {
      same = std::memcmp(&out.points[i], &points.points[i], 3 * sizeof(float)) == 0;
    }
    EXPECT_TRUE(same);
  }
  {  // empty cloud (:78-80 / :143-145): the map is not even cleared
    pcl::PointCloud<Point> none, out;
    std::vector<std::set<size_t>> map = sentinel;
    deformation::deformPoints(out, map, none, prefix, control_points, {}, values);
    EXPECT_TRUE(map == sentinel);
    EXPECT_EQ(size_t(0), out.size());
  }
  {  // unstamped, no control points (:83-86)
    pcl::PointCloud<Point> out = points;
    std::vector<std::set<size_t>> map = sentinel;
    deformation::deformPoints(out, map, points, prefix, {}, {}, values);
    EXPECT_TRUE(untouched(out, map));
  }
  {  // stamped, fewer control points than k (:148-152)
    pcl::PointCloud<Point> out = points;
    std::vector<std::set<size_t>> map = sentinel;
    deformation::deformPoints(out, map, cloud, prefix, control_points, control_point_stamps, values,
                              4, 10.0);
    EXPECT_TRUE(untouched(out, map));
    deformation::deformPointsOnly(out, cloud, prefix, control_points, control_point_stamps, values,
                                  4, 10.0);
    EXPECT_TRUE(untouched(out, map));
  }
  {  // empty `indices`: nothing to do
    pcl::PointCloud<Point> out = points;
    std::vector<std::set<size_t>> map = sentinel;
    const std::vector<size_t> no_indices;
    deformation::deformPoints(out, map, cloud, prefix, control_points, control_point_stamps, values,
                              2, 10.0, &no_indices);
    EXPECT_TRUE(untouched(out, map));
  }
  {  // control_point_stamps.at(j) past its end
    pcl::PointCloud<Point> out = points;
    std::vector<std::set<size_t>> map = sentinel;
    const std::vector<Timestamp> short_stamps(2, stampFromSec(1.0));
    bool thrown = false;
    try {
      deformation::deformPoints(out, map, cloud, prefix, control_points, short_stamps, values, 2,
                                10.0);
    } catch (const std::out_of_range&) {
      thrown = true;
    }
    EXPECT_TRUE(thrown);
  }
  {  // indices->at() / cloud.at() past the cloud
    pcl::PointCloud<Point> out = points;
    std::vector<std::set<size_t>> map = sentinel;
    const std::vector<size_t> bad_indices = {1, 2, 20};
    bool thrown = false;
    try {
      deformation::deformPoints(out, map, cloud, prefix, control_points, control_point_stamps,
                                values, 2, 10.0, &bad_indices);
    } catch (const std::out_of_range&) {
      thrown = true;
    }
    EXPECT_TRUE(thrown);
  }
  {  // a prefix the graph has never seen (deformation_graph.h:819-823): vertices as given
    DeformationGraphMeshPath graph;
    graph.addControlPoint('b', gtsam::Point3(0, 0, 0), 0);
    pcl::PointCloud<pcl::PointXYZRGBA> verts, out;
    for (size_t i = 0; i < 5; i++) verts.push_back(colored(static_cast<float>(i), 1, 2, 9, 8, 7));
    out = verts;
    const std::vector<Timestamp> five(5, 0);
    graph.deformPoints(out, verts, five, 'c', values, 2, 10.0);
    for (size_t i = 0; i < 5; i++) {
      EXPECT_EQ(verts.points[i].x, out.points[i].x);
      EXPECT_EQ(verts.points[i].rgba, out.points[i].rgba);
    }
  }
}

// No reference counterpart: a cloud large enough for the shim's threaded gather / scatter
// (detail::parallelFor engages above 2^18 vertices).  deformPointsOnly must give the vertices
// of deformPoints bit for bit, for the whole cloud and for an `indices` subset, and leave the
// vertices outside the subset alone.
static void testLargeCloudStaging() {
  typedef pcl::PointXYZ Point;
  const size_t n = 600000, n_ctrl = 300;
  pcl::PointCloud<Point> points;
  std::vector<Timestamp> stamps;
  std::vector<gtsam::Point3> control_points;
  std::vector<Timestamp> control_point_stamps;
  gtsam::Values values;
  const char prefix = 'a';
  uint32_t lcg = 12345u;
  auto unit = [&lcg]() {
    lcg = lcg * 1664525u + 1013904223u;
    return static_cast<float>(lcg >> 8) / 16777216.0f;
  };
  for (size_t j = 0; j < n_ctrl; j++) {
    control_points.push_back(gtsam::Point3(0.5 * static_cast<double>(j), unit(), unit()));
    control_point_stamps.push_back(stampFromSec(0.5 * static_cast<double>(j)));
    values.insert(gtsam::Symbol(prefix, j),
                  gtsam::Pose3(gtsam::Rot3::Rz(0.001 * static_cast<double>(j)),
                               gtsam::Point3(0.5 * static_cast<double>(j) + 0.25, 0.5, -0.1)));
  }
  for (size_t i = 0; i < n; i++) {
    const double s = 150.0 * static_cast<double>(i) / static_cast<double>(n);
    points.push_back(Point(static_cast<float>(s) + unit(), 4.0f * unit() - 2.0f, 2.0f * unit()));
    stamps.push_back(stampFromSec(std::floor(s)));
  }
  const ConstStampedCloud<Point> cloud{points, stamps};

  pcl::PointCloud<Point> with_map = points, only = points;
  std::vector<std::set<size_t>> control_point_map;
  deformation::deformPoints(with_map, control_point_map, cloud, prefix, control_points,
                            control_point_stamps, values, 4, 10.0);
  deformation::deformPointsOnly(only, cloud, prefix, control_points, control_point_stamps, values,
                                4, 10.0);
  EXPECT_EQ(n, control_point_map.size());
  size_t differing = 0, moved = 0, small_sets = 0;
  for (size_t i = 0; i < n; i++) {
    differing += std::memcmp(&with_map.points[i], &only.points[i], 3 * sizeof(float)) != 0;
    moved += std::memcmp(&with_map.points[i], &points.points[i], 3 * sizeof(float)) != 0;
  }
  for (const auto& s : control_point_map) small_sets += s.size() < 2;
  EXPECT_EQ(size_t(0), differing);
  EXPECT_TRUE(moved > n / 2);
  EXPECT_EQ(size_t(0), small_sets);

  // incremental form: every third vertex of the second half
  std::vector<size_t> indices;
  for (size_t i = n / 2; i < n; i += 3) indices.push_back(i);
  pcl::PointCloud<Point> sub_map = points, sub_only = points;
  control_point_map.clear();
  deformation::deformPoints(sub_map, control_point_map, cloud, prefix, control_points,
                            control_point_stamps, values, 4, 10.0, &indices);
  deformation::deformPointsOnly(sub_only, cloud, prefix, control_points, control_point_stamps,
                                values, 4, 10.0, &indices);
  EXPECT_EQ(indices.size(), control_point_map.size());
  size_t sub_differing = 0, touched_outside = 0;
  for (size_t i = 0; i < n; i++) {
    sub_differing += std::memcmp(&sub_map.points[i], &sub_only.points[i], 3 * sizeof(float)) != 0;
    const bool in_subset = i >= n / 2 && (i - n / 2) % 3 == 0;
    if (!in_subset) {
      touched_outside += std::memcmp(&sub_only.points[i], &points.points[i], 3 * sizeof(float)) != 0;
    }
  }
  EXPECT_EQ(size_t(0), sub_differing);
  EXPECT_EQ(size_t(0), touched_outside);

  // the same cloud as a mesh through DeformationGraph::deformMesh (full recompute): vertices of
  // deformPointsOnly, colours and faces of the input (300k faces: the threaded polygon copy)
  DeformationGraphMeshPath graph;
  for (size_t j = 0; j < n_ctrl; j++) {
    graph.addControlPoint(prefix, control_points[j], control_point_stamps[j]);
  }
  pcl::PointCloud<pcl::PointXYZRGBA> colored_points;
  colored_points.points.resize(n);
  for (size_t i = 0; i < n; i++) {
    pcl::PointXYZRGBA& p = colored_points.points[i];
    p.x = points.points[i].x;
    p.y = points.points[i].y;
    p.z = points.points[i].z;
    p.rgba = static_cast<uint32_t>(i * 2654435761u);
  }
  pcl::PolygonMesh mesh;
  pcl::toPCLPointCloud2(colored_points, mesh.cloud);
  mesh.polygons.resize(n / 2);
  for (size_t f = 0; f < mesh.polygons.size(); f++) {
    const uint32_t a = static_cast<uint32_t>(2 * f);
    mesh.polygons[f].vertices = {a, a + 1, static_cast<uint32_t>((a + 7) % n)};
  }
  const std::vector<int> graph_indices(n, 0);
  const pcl::PolygonMesh deformed = graph.deformMesh(mesh, stamps, graph_indices, prefix, values, 4);
  pcl::PointCloud<pcl::PointXYZRGBA> deformed_points;
  pcl::fromPCLPointCloud2(deformed.cloud, deformed_points);
  EXPECT_EQ(n, deformed_points.points.size());
  EXPECT_EQ(mesh.polygons.size(), deformed.polygons.size());
  size_t mesh_differing = 0, colour_differing = 0, face_differing = 0;
  for (size_t i = 0; i < n && i < deformed_points.points.size(); i++) {
    mesh_differing += std::memcmp(&deformed_points.points[i], &only.points[i], 3 * sizeof(float)) != 0;
    colour_differing += deformed_points.points[i].rgba != colored_points.points[i].rgba;
  }
  for (size_t f = 0; f < mesh.polygons.size() && f < deformed.polygons.size(); f++) {
    face_differing += !(mesh.polygons[f].vertices == deformed.polygons[f].vertices);
  }
  EXPECT_EQ(size_t(0), mesh_differing);
  EXPECT_EQ(size_t(0), colour_differing);
  EXPECT_EQ(size_t(0), face_differing);
}

// No reference counterpart: DeformationGraphMeshPath over several devices (vertex slabs,
// pgmo_deform_group_*) gives the single-device mesh bit for bit, for a first deformMesh and for
// a second one after a new solve.  One visible device: two slabs on it.
static void testDeviceGroup() {
  const size_t n = 400000, n_ctrl = 400;
  const char prefix = 'a';
  deformation::Backend::groupMinVertices() = 100000;
  if (pgmo_device_count() < 2) setenv("KIMERA_PGMO_B200_GROUP_DEVICES", "0,0", 1);
  DeformationGraphMeshPath single(0, 1), multi(0, 0);
  gtsam::Values values, values2;
  uint32_t lcg = 777u;
  auto unit = [&lcg]() {
    lcg = lcg * 1664525u + 1013904223u;
    return static_cast<float>(lcg >> 8) / 16777216.0f;
  };
  for (size_t j = 0; j < n_ctrl; j++) {
    const gtsam::Point3 g(0.5 * static_cast<double>(j), unit(), unit());
    const Timestamp t = stampFromSec(0.5 * static_cast<double>(j));
    single.addControlPoint(prefix, g, t);
    multi.addControlPoint(prefix, g, t);
    values.insert(gtsam::Symbol(prefix, j), gtsam::Pose3(gtsam::Rot3::Rz(0.001 * static_cast<double>(j)),
                                                        gtsam::Point3(g.x() + 0.25, 0.5, -0.1)));
    values2.insert(gtsam::Symbol(prefix, j), gtsam::Pose3(gtsam::Rot3::Rz(-0.002 * static_cast<double>(j)),
                                                         gtsam::Point3(g.x() - 0.1, 0.3, 0.2)));
  }
  pcl::PointCloud<pcl::PointXYZRGBA> cloud;
  std::vector<Timestamp> stamps;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; i++) {
    const double s = 200.0 * static_cast<double>(i) / static_cast<double>(n);
    pcl::PointXYZRGBA& p = cloud.points[i];
    p.x = static_cast<float>(s) + unit();
    p.y = 4.0f * unit() - 2.0f;
    p.z = 2.0f * unit();
    p.rgba = static_cast<uint32_t>(i * 2654435761u);
    stamps.push_back(stampFromSec(std::floor(s)));
  }
  cloud.width = static_cast<uint32_t>(n);
  pcl::PolygonMesh mesh;
  pcl::toPCLPointCloud2(cloud, mesh.cloud);
  const std::vector<int> graph_indices(n, -1);
  for (const gtsam::Values* v : {&values, &values2}) {
    single.setRecalculateVertices();
    multi.setRecalculateVertices();
    const pcl::PolygonMesh a = single.deformMesh(mesh, stamps, graph_indices, prefix, *v, 4);
    const pcl::PolygonMesh b = multi.deformMesh(mesh, stamps, graph_indices, prefix, *v, 4);
    pcl::PointCloud<pcl::PointXYZRGBA> pa, pb;
    pcl::fromPCLPointCloud2(a.cloud, pa);
    pcl::fromPCLPointCloud2(b.cloud, pb);
    EXPECT_EQ(n, pa.points.size());
    EXPECT_EQ(n, pb.points.size());
    size_t differing = 0, moved = 0;
    for (size_t i = 0; i < n && i < pa.points.size() && i < pb.points.size(); i++) {
      differing += std::memcmp(&pa.points[i], &pb.points[i], sizeof(pcl::PointXYZRGBA)) != 0;
      moved += std::memcmp(&pa.points[i], &cloud.points[i], 3 * sizeof(float)) != 0;
    }
    EXPECT_EQ(size_t(0), differing);
    EXPECT_TRUE(moved > n / 2);
  }
  unsetenv("KIMERA_PGMO_B200_GROUP_DEVICES");
  deformation::Backend::groupMinVertices() = size_t(1) << 20;
}

// ------------------------------------------------------------------ TestDeformationGraph
static pcl::PointXYZRGBA colored(float x, float y, float z, uint8_t r, uint8_t g, uint8_t b) {
  pcl::PointXYZRGBA p;
  p.x = x;
  p.y = y;
  p.z = z;
  p.r = r;
  p.g = g;
  p.b = b;
  p.a = 255;
  return p;
}

// pgmo_fixtures.cpp:97-158
static pcl::PolygonMesh createSimpleMesh(double scale = 1.0) {
  const float s = static_cast<float>(scale);
  pcl::PointCloud<pcl::PointXYZRGBA> ptcld;
  ptcld.push_back(colored(0, 0, 0, 23, 24, 122));
  ptcld.push_back(colored(s, 0, 0, 33, 34, 52));
  ptcld.push_back(colored(0, s, 0, 12, 144, 22));
  ptcld.push_back(colored(s, s, 0, 0, 14, 0));
  ptcld.push_back(colored(0, 0, s, 144, 0, 12));
  pcl::PolygonMesh mesh;
  pcl::toPCLPointCloud2(ptcld, mesh.cloud);
  const uint32_t tris[4][3] = {{0, 1, 2}, {1, 3, 2}, {0, 1, 4}, {0, 4, 2}};
  for (auto& t : tris) {
    pcl::Vertices v;
    v.vertices = {t[0], t[1], t[2]};
    mesh.polygons.push_back(v);
  }
  return mesh;
}

// SetUpDeformationGraph (test_deformation_graph.cpp:98-130): nodes of createMeshTriangle
static void setUpGraph(DeformationGraphMeshPath* graph, gtsam::Values* values, double shift_x) {
  const double tri[3][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
  for (size_t j = 0; j < 3; ++j) {
    graph->addControlPoint('v', gtsam::Point3(tri[j][0], tri[j][1], tri[j][2]), 0);
    values->insert(gtsam::Symbol('v', j),
                   gtsam::Pose3(gtsam::Rot3(),
                                gtsam::Point3(tri[j][0] + shift_x, tri[j][1], tri[j][2])));
  }
}

static void testReconstructAndTranslate() {
  for (int scenario = 0; scenario < 2; ++scenario) {
    const double shift = scenario == 0 ? 0.0 : 0.5;
    DeformationGraphMeshPath graph;
    gtsam::Values values;
    setUpGraph(&graph, &values, shift);
    pcl::PolygonMesh original_mesh = createSimpleMesh();
    std::vector<Timestamp> stamps(5, 0);
    std::vector<int> inds = {-1, 0, -1, 0, -1};
    for (size_t k : {size_t(1), size_t(2)}) {
      graph.setRecalculateVertices(true);
      pcl::PolygonMesh new_mesh = graph.deformMesh(original_mesh, stamps, inds, 'v', values, k);
      pcl::PointCloud<pcl::PointXYZRGBA> deformed, orig;
      pcl::fromPCLPointCloud2(new_mesh.cloud, deformed);
      pcl::fromPCLPointCloud2(original_mesh.cloud, orig);
      EXPECT_EQ(size_t(5), deformed.points.size());
      for (size_t i = 0; i < 5; ++i) {
        EXPECT_NEAR(orig.points[i].x + shift, deformed.points[i].x, 1e-6);
        EXPECT_NEAR(orig.points[i].y, deformed.points[i].y, 1e-6);
        EXPECT_NEAR(orig.points[i].z, deformed.points[i].z, 1e-6);
        EXPECT_EQ(orig.points[i].rgba, deformed.points[i].rgba);
      }
      EXPECT_EQ(33, int(deformed.points[1].r));
      EXPECT_EQ(144, int(deformed.points[2].g));
      EXPECT_EQ(255, int(deformed.points[3].a));
      EXPECT_EQ(144, int(deformed.points[4].r));
      EXPECT_TRUE(original_mesh.polygons[0].vertices == new_mesh.polygons[0].vertices);
      EXPECT_TRUE(original_mesh.polygons[3].vertices == new_mesh.polygons[3].vertices);
    }
    // second call without a new solve: served from the cache (findStartIndex -> all cached)
    pcl::PolygonMesh again = graph.deformMesh(original_mesh, stamps, inds, 'v', values, 2);
    pcl::PointCloud<pcl::PointXYZRGBA> deformed;
    pcl::fromPCLPointCloud2(again.cloud, deformed);
    EXPECT_NEAR(1.0 + shift, deformed.points[1].x, 1e-6);
  }
}

static void testDeformMeshCube() {
  const float cube[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
                            {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  pcl::PointCloud<pcl::PointXYZRGBA> verts;
  for (auto& v : cube) verts.push_back(colored(v[0], v[1], v[2], 100, 0, 0));
  pcl::PolygonMesh cube_mesh;
  pcl::toPCLPointCloud2(verts, cube_mesh.cloud);
  pcl::Vertices f;
  f.vertices = {0, 2, 1};
  cube_mesh.polygons.push_back(f);
  std::vector<Timestamp> stamps(8, 0);
  std::vector<int> inds(8, -1);

  DeformationGraphMeshPath graph;
  gtsam::Values values;
  setUpGraph(&graph, &values, -0.5);
  graph.setRecalculateVertices(true);
  auto new_mesh = graph.deformMesh(cube_mesh, stamps, inds, 'v', values, 3);
  pcl::PointCloud<pcl::PointXYZRGBA> actual;
  pcl::fromPCLPointCloud2(new_mesh.cloud, actual);
  for (size_t i = 0; i < 8; ++i) {
    EXPECT_NEAR(cube[i][0] - 0.5, actual.points[i].x, 1e-6);
    EXPECT_NEAR(cube[i][1], actual.points[i].y, 1e-6);
    EXPECT_NEAR(cube[i][2], actual.points[i].z, 1e-6);
  }
  // second phase: node 0 at (-0.5,0,0), node 1 at (1.5,0,0) -> 1.273 (test :305-306)
  gtsam::Values values2;
  values2.insert(gtsam::Symbol('v', 0), gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(-0.5, 0, 0)));
  values2.insert(gtsam::Symbol('v', 1), gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(1.5, 0, 0)));
  values2.insert(gtsam::Symbol('v', 2), gtsam::Pose3(gtsam::Rot3(), gtsam::Point3(0.5, 1, 0)));
  graph.setRecalculateVertices(true);
  new_mesh = graph.deformMesh(cube_mesh, stamps, inds, 'v', values2, 2);
  pcl::fromPCLPointCloud2(new_mesh.cloud, actual);
  EXPECT_NEAR(-0.5, actual.points[0].x, 0.001);
  EXPECT_NEAR(1.273, actual.points[1].x, 0.001);
}

static void testIncrementalWithPredeform() {
  // incremental mode: cache + start index + rigid predeform by graph index
  DeformationGraphMeshPath graph;
  gtsam::Values values;
  for (size_t j = 0; j < 8; ++j) {
    graph.addControlPoint('v', gtsam::Point3(double(j), 0, 0), Timestamp(j) * 1000000000ull);
    values.insert(gtsam::Symbol('v', j),
                  gtsam::Pose3(gtsam::Rot3::Rz(0.01 * j), gtsam::Point3(double(j), 0.1 * j, 0)));
  }
  pcl::PointCloud<pcl::PointXYZRGBA> verts;
  std::vector<Timestamp> stamps;
  std::vector<int> inds;
  for (size_t i = 0; i < 80; ++i) {
    verts.push_back(colored(0.1f * i, 0.2f, 0.1f, 1, 2, 3));
    stamps.push_back(Timestamp(i / 10) * 1000000000ull);
    inds.push_back(i % 3 == 0 ? int(i / 10) : -1);
  }
  pcl::PointCloud<pcl::PointXYZRGBA> full = verts, incr = verts;
  graph.setRecalculateVertices(true);
  graph.deformPoints(full, verts, stamps, 'v', values, 2, 3.0, &inds);
  // no loop closure since: only vertices newer than last_ctrl_stamp - tol are recomputed
  std::vector<std::set<size_t>> vgm;
  graph.deformPoints(incr, verts, stamps, 'v', values, 2, 3.0, &inds, -1, &vgm);
  EXPECT_EQ(size_t(80), vgm.size());
  for (size_t i = 0; i < 40; ++i) {  // stamps <= 7 s - 3 s: served from the cache
    EXPECT_EQ(full.points[i].x, incr.points[i].x);
    EXPECT_EQ(full.points[i].y, incr.points[i].y);
  }
  for (size_t i = 50; i < 80; ++i) {
    if (inds[i] >= 0) {  // rigidly predeformed by its graph node: R * (v - g) + t
      const double c = std::cos(0.01 * inds[i]), s = std::sin(0.01 * inds[i]);
      const double px = double(verts.points[i].x) - inds[i], py = double(verts.points[i].y);
      EXPECT_NEAR(c * px - s * py + inds[i], incr.points[i].x, 1e-6);
      EXPECT_NEAR(s * px + c * py + 0.1 * inds[i], incr.points[i].y, 1e-6);
      EXPECT_TRUE(vgm[i].empty());
    } else {
      EXPECT_TRUE(!vgm[i].empty());
    }
  }
  // every new vertex tied to a graph node: nothing is left for deformation::deformPoints (its
  // early return); the control table of the next solve must still reach the device
  std::vector<int> all_tied;
  for (size_t i = 0; i < 80; ++i) all_tied.push_back(int(i / 10));
  graph.deformPoints(incr, verts, stamps, 'v', values, 2, 3.0, &all_tied);
  gtsam::Values lifted;  // the same solve, every node 1 m higher
  for (size_t j = 0; j < 8; ++j) {
    lifted.insert(gtsam::Symbol('v', j),
                  gtsam::Pose3(gtsam::Rot3::Rz(0.01 * j), gtsam::Point3(double(j), 0.1 * j, 1.0)));
  }
  graph.setRecalculateVertices(true);
  pcl::PointCloud<pcl::PointXYZRGBA> after = verts;
  graph.deformPoints(after, verts, stamps, 'v', lifted, 2, 3.0, &inds);
  for (size_t i = 0; i < 80; ++i) {
    EXPECT_NEAR(full.points[i].x, after.points[i].x, 1e-5);
    EXPECT_NEAR(full.points[i].z + 1.0, after.points[i].z, 1e-5);
  }
  // and an incremental update right after it (table kept from predeformPoints to the knn call)
  pcl::PointCloud<pcl::PointXYZRGBA> after_incr = verts;
  graph.deformPoints(after_incr, verts, stamps, 'v', lifted, 2, 3.0, &inds);
  for (size_t i = 0; i < 80; ++i) {
    if (i >= 50 && inds[i] >= 0) continue;  // rigidly predeformed (checked above)
    EXPECT_NEAR(after.points[i].x, after_incr.points[i].x, 1e-5);
    EXPECT_NEAR(after.points[i].z, after_incr.points[i].z, 1e-5);
  }
}

// ------------------------------------------------------ TestVoxelClearingCompression
struct OrderedBlockMesh : MeshInterface {  // pgmo_fixtures.h:28-48
  BlockIndices order;
  std::vector<std::vector<pcl::PointXYZRGBA>> blocks;
  mutable size_t active = 0;
  const BlockIndices& blockIndices() const override { return order; }
  void markBlockActive(const BlockIndex& block) const override {
    for (size_t b = 0; b < order.size(); ++b) {
      if (order[b] == block) active = b;
    }
  }
  size_t activeBlockSize() const override { return blocks[active].size(); }
  pcl::PointXYZRGBA getActiveVertex(size_t i) const override { return blocks[active][i]; }
  // optional semantics (mesh_interface.h:28-33): labels[b][i], empty = none
  std::vector<std::vector<uint32_t>> labels;
  bool hasSemantics() const override { return !labels.empty(); }
  std::optional<uint32_t> getActiveSemantics(size_t i) const override {
    return labels.empty() ? std::nullopt : std::optional<uint32_t>(labels[active][i]);
  }
  Ptr clone() const override { return std::make_shared<OrderedBlockMesh>(*this); }
};

using FaceCoords = std::array<std::array<float, 3>, 3>;
struct BlockConfig {
  std::array<int, 3> index;
  std::vector<FaceCoords> faces;
};
static const BlockConfig block1_empty{{0, 0, 0}, {}};
static const BlockConfig block1_test1{{0, 0, 0},
                                      {{{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                                       {{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                                       {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};
static const BlockConfig block2_empty{{-1, 0, 0}, {}};
static const BlockConfig block2_test1{{-1, 0, 0},
                                      {{{{-0.5, 0.5, 0.5}, {-0.5, 0.75, 0.75}, {-0.5, 0.75, 0.5}}},
                                       {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};

static OrderedBlockMesh createMesh(const std::vector<BlockConfig>& configs) {
  OrderedBlockMesh mesh;
  uint8_t point_index = 0;
  for (const auto& c : configs) {
    mesh.order.push_back(BlockIndex(c.index[0], c.index[1], c.index[2]));
    mesh.blocks.emplace_back();
    for (const auto& f : c.faces) {
      for (const auto& v : f) {
        mesh.blocks.back().push_back(colored(v[0], v[1], v[2], point_index, point_index, point_index));
        ++point_index;
      }
    }
  }
  return mesh;
}

struct CompressionOutput {
  explicit CompressionOutput(VoxelClearingCompression& c)
      : vertices(new pcl::PointCloud<pcl::PointXYZRGBA>()),
        triangles(new std::vector<pcl::Vertices>()),
        timestamps(new std::vector<Timestamp>()) {
    c.getVertices(vertices);
    c.getStoredPolygons(triangles);
    invalidated = c.getInvalidIndices();
    active_indices = c.getActiveVerticesIndex();
    c.getTimestamps(timestamps);
  }
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr vertices;
  std::shared_ptr<std::vector<pcl::Vertices>> triangles;
  std::vector<size_t> invalidated;
  std::vector<size_t> active_indices;
  std::shared_ptr<std::vector<Timestamp>> timestamps;
};

static bool checkTriangles(const std::vector<std::vector<uint32_t>>& expected,
                           const std::vector<pcl::Vertices>& result) {
  std::vector<std::vector<uint32_t>> filtered;
  for (const auto& r : result) {
    bool seen = false;
    for (const auto& t : filtered) seen |= (t == r.vertices);
    if (!seen) filtered.push_back(r.vertices);
  }
  if (expected.size() != filtered.size()) return false;
  for (size_t i = 0; i < expected.size(); ++i) {
    std::set<uint32_t> e(expected[i].begin(), expected[i].end());
    std::set<uint32_t> r(filtered[i].begin(), filtered[i].end());
    if (e != r) return false;
  }
  return true;
}

static void integrate(VoxelClearingCompression& c, const OrderedBlockMesh& mesh, double t) {
  pcl::PointCloud<pcl::PointXYZRGBA> v;
  std::vector<pcl::Vertices> tri;
  std::vector<size_t> ind;
  HashedIndexMapping remap;
  c.compressAndIntegrate(mesh, v, tri, ind, remap, t);
}

static void testVoxelClearing() {
  using V = std::vector<size_t>;
  {  // clearingCorrectBasic
    VoxelClearingCompression c(1.0e-3);
    integrate(c, createMesh({block1_test1}), 100.0);
    CompressionOutput o(c);
    EXPECT_EQ(6u, o.vertices->points.size());
    EXPECT_EQ(0u, o.invalidated.size());
    EXPECT_EQ(o.vertices->size(), o.timestamps->size());
    EXPECT_TRUE(o.active_indices == (V{0, 1, 2, 3, 4, 5}));
    EXPECT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}}, *o.triangles));
    integrate(c, createMesh({block1_empty}), 100.0);
    CompressionOutput o2(c);
    EXPECT_EQ(6u, o2.vertices->points.size());
    EXPECT_EQ(6u, o2.invalidated.size());
    EXPECT_EQ(0u, o2.triangles->size());
    EXPECT_EQ(0u, o2.active_indices.size());
  }
  {  // clearingCorrectMultiblock
    VoxelClearingCompression c(1.0e-3);
    integrate(c, createMesh({block1_test1, block2_test1}), 100.0);
    {
      CompressionOutput o(c);
      EXPECT_EQ(9u, o.vertices->points.size());
      EXPECT_EQ(0u, o.invalidated.size());
      EXPECT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}, {0, 1, 2}}, *o.triangles));
    }
    integrate(c, createMesh({block1_empty, block2_test1}), 100.0);
    {
      CompressionOutput o(c);
      EXPECT_EQ(9u, o.vertices->points.size());
      EXPECT_EQ(3u, o.invalidated.size());
      EXPECT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}}, *o.triangles));
    }
    integrate(c, createMesh({block1_test1, block2_test1}), 100.0);
    {
      CompressionOutput o(c);
      EXPECT_EQ(12u, o.vertices->points.size());
      EXPECT_EQ(3u, o.invalidated.size());
      EXPECT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}, {9, 10, 11}}, *o.triangles));
    }
    integrate(c, createMesh({block1_test1, block2_empty}), 100.0);
    {
      CompressionOutput o(c);
      EXPECT_EQ(12u, o.vertices->points.size());
      EXPECT_EQ(6u, o.invalidated.size());
      EXPECT_TRUE(checkTriangles({{9, 10, 11}, {3, 4, 5}}, *o.triangles));
    }
  }
  {  // pruningCorrectMultiblock
    VoxelClearingCompression c(1.0e-3);
    integrate(c, createMesh({block1_test1}), 100.0);
    integrate(c, createMesh({block2_test1}), 102.0);
    c.pruneStoredMesh(101.0);
    integrate(c, createMesh({block1_empty, block2_empty}), 105.0);
    {
      CompressionOutput o(c);
      EXPECT_EQ(9u, o.vertices->points.size());
      EXPECT_EQ(3u, o.invalidated.size());
      EXPECT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}}, *o.triangles));
    }
    integrate(c, createMesh({block1_test1}), 105.0);
    {
      CompressionOutput o(c);
      EXPECT_EQ(15u, o.vertices->points.size());
      EXPECT_EQ(3u, o.invalidated.size());
      EXPECT_EQ(o.vertices->size(), o.timestamps->size());
      EXPECT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}, {9, 10, 11}, {12, 13, 14}}, *o.triangles));
    }
  }
}

// ------------------------------------------------------------- TestOctreeCompression
static void testOctreeCompression() {
  using V = std::vector<size_t>;
  OctreeCompression compression(0.1);
  pcl::PointCloud<pcl::PointXYZRGBA> new_vertices;
  std::vector<pcl::Vertices> new_triangles;
  std::vector<size_t> new_indices;
  HashedIndexMapping remaps;
  PclMeshInterface mesh1(createSimpleMesh(1.0));
  compression.compressAndIntegrate(mesh1, new_vertices, new_triangles, new_indices, remaps, 100.0);
  EXPECT_EQ(5u, new_vertices.size());
  EXPECT_EQ(4u, new_triangles.size());
  EXPECT_TRUE(new_indices == (V{0, 1, 2, 3, 4}));
  EXPECT_EQ(1, new_vertices.points[1].x);
  EXPECT_EQ(1, new_vertices.points[2].y);
  EXPECT_EQ(1, new_vertices.points[4].z);
  EXPECT_EQ(0u, new_triangles.at(0).vertices[0]);
  EXPECT_EQ(2u, new_triangles.at(3).vertices[2]);
  PclMeshInterface mesh2(createSimpleMesh(2.0));
  remaps.clear();
  compression.compressAndIntegrate(mesh2, new_vertices, new_triangles, new_indices, remaps, 101.0);
  EXPECT_EQ(4u, new_vertices.size());
  EXPECT_EQ(4u, new_triangles.size());
  EXPECT_TRUE(new_indices == (V{0, 5, 6, 7, 8}));
  EXPECT_EQ(2, new_vertices.points[0].x);
  EXPECT_EQ(6u, new_triangles.at(3).vertices[2]);
  compression.pruneStoredMesh(100.5);
  pcl::PointCloud<pcl::PointXYZ>::Ptr active(new pcl::PointCloud<pcl::PointXYZ>());
  compression.getActiveVertices(active);
  EXPECT_EQ(5u, active->size());
  EXPECT_EQ(2, active->points[1].x);
  PclMeshInterface mesh3(createSimpleMesh(1.0));
  remaps.clear();
  compression.compressAndIntegrate(mesh3, new_vertices, new_triangles, new_indices, remaps, 102.0);
  EXPECT_TRUE(new_indices == (V{0, 9, 10, 11, 12}));
  EXPECT_EQ(13u, compression.getNumVertices());
}

// ------------------------------------------------------ VoxbloxMsgInterface
// The packed path (u16 / u8 arrays uploaded as they are, getActiveVertex evaluated on the GPU)
// must give exactly what the generic MeshInterface path gives for the same message.
static void testVoxbloxInterface() {
  VoxbloxMesh msg;
  msg.block_edge_length = 0.8f;
  uint32_t seed = 12345u;
  auto next = [&seed]() {
    seed = seed * 1664525u + 1013904223u;
    return seed >> 8;
  };
  const int64_t indices[3][3] = {{0, 0, 0}, {-1, 2, 0}, {3, -4, 1}};
  for (const auto& index : indices) {
    VoxbloxMeshBlock blk;
    for (int a = 0; a < 3; ++a) blk.index[a] = index[a];
    for (int t = 0; t < 40; ++t) {
      const uint32_t ax = next() % 30000, ay = next() % 30000, az = next() % 30000;
      for (int v = 0; v < 3; ++v) {
        blk.x.push_back(static_cast<uint16_t>(ax + next() % 5000));
        blk.y.push_back(static_cast<uint16_t>(ay + next() % 5000));
        blk.z.push_back(static_cast<uint16_t>(az + next() % 5000));
        blk.r.push_back(static_cast<uint8_t>(next()));
        blk.g.push_back(static_cast<uint8_t>(next()));
        blk.b.push_back(static_cast<uint8_t>(next()));
      }
    }
    msg.mesh_blocks.push_back(blk);
  }
  msg.mesh_blocks[0].x[0] = 0;
  msg.mesh_blocks[0].y[0] = 65535;
  const VoxbloxMsgInterface packed(msg);
  OrderedBlockMesh generic;  // the same message through per-vertex getActiveVertex on the host
  for (const auto& block : packed.blockIndices()) {
    generic.order.push_back(block);
    generic.blocks.emplace_back();
    packed.markBlockActive(block);
    for (size_t i = 0; i < packed.activeBlockSize(); ++i) {
      generic.blocks.back().push_back(packed.getActiveVertex(i));
    }
  }
  EXPECT_EQ(3u, packed.blockIndices().size());
  EXPECT_EQ(120u, generic.blocks[1].size());
  EXPECT_EQ(255, static_cast<int>(generic.blocks[0][0].a));
  EXPECT_NEAR(1.6f, generic.blocks[0][0].y, 1e-6);  // 65535 * (2 / 65535) * edge
  for (int rep = 0; rep < 2; ++rep) {
    VoxelClearingCompression a(0.05), b(0.05);
    pcl::PointCloud<pcl::PointXYZRGBA> v;
    std::vector<pcl::Vertices> tri;
    std::vector<size_t> ind;
    HashedIndexMapping ra, rb;
    a.compressAndIntegrate(packed, v, tri, ind, ra, 10.0 + rep);
    b.compressAndIntegrate(generic, v, tri, ind, rb, 10.0 + rep);
    EXPECT_EQ(ra.size(), rb.size());
    for (const auto& kv : rb) {
      EXPECT_TRUE(ra.count(kv.first) == 1 && ra.at(kv.first) == kv.second);
    }
    CompressionOutput oa(a), ob(b);
    EXPECT_EQ(oa.vertices->size(), ob.vertices->size());
    bool same = oa.vertices->size() == ob.vertices->size();
    for (size_t i = 0; same && i < oa.vertices->size(); ++i) {
      const auto& p = oa.vertices->at(i);
      const auto& q = ob.vertices->at(i);
      same = p.x == q.x && p.y == q.y && p.z == q.z && p.rgba == q.rgba;
    }
    EXPECT_TRUE(same);
    EXPECT_TRUE(oa.vertices->size() > 100);
  }
}

// ------------------------------------------------------ TestDeltaCompression (CompressionCorrect)
// test_delta_compression.cpp:334-435: the two richest configurations, checked the way
// ExpectedDelta::checkOutput does (:189-270) — through absolute point ids in the colour channel.
struct DeltaStep {
  int prune_s;  // < 0: none
  uint64_t stamp_s;
  std::vector<BlockConfig> blocks;
  size_t vertex_start, face_start, num_vertices;
  std::vector<std::array<size_t, 3>> triangles;
  std::vector<size_t> indices;
};

static void runDeltaConfig(const std::vector<DeltaStep>& steps) {
  DeltaCompression compression(1.0e-3);
  uint8_t point_index = 0;
  std::map<size_t, size_t> prev_remapping;
  for (const auto& step : steps) {
    if (step.prune_s >= 0) compression.archiveBlocksByTime(static_cast<uint64_t>(step.prune_s) * 1000000000ull);
    OrderedBlockMesh mesh;
    for (const auto& c : step.blocks) {
      mesh.order.push_back(BlockIndex(c.index[0], c.index[1], c.index[2]));
      mesh.blocks.emplace_back();
      for (const auto& f : c.faces) {
        for (const auto& v : f) {
          mesh.blocks.back().push_back(colored(v[0], v[1], v[2], point_index, point_index, point_index));
          ++point_index;
        }
      }
    }
    HashedIndexMapping remapping;
    const auto output = compression.update(mesh, step.stamp_s * 1000000000ull, &remapping);
    EXPECT_TRUE(output != nullptr);
    EXPECT_EQ(output->vertex_updates->size(), output->stamp_updates.size());
    EXPECT_EQ(output->vertex_start, step.vertex_start);
    EXPECT_EQ(output->face_start, step.face_start);
    EXPECT_EQ(output->vertex_updates->size(), step.num_vertices);
    std::map<size_t, size_t> result_remapping;
    for (size_t i = 0; i < output->vertex_start; ++i) {
      if (prev_remapping.count(i)) result_remapping[i] = prev_remapping.at(i);
    }
    for (size_t i = 0; i < output->vertex_updates->size(); ++i) {
      result_remapping[i + output->vertex_start] = output->vertex_updates->at(i).r;
    }
    prev_remapping = result_remapping;
    std::vector<Face> all(output->face_archive_updates);
    all.insert(all.end(), output->face_updates.begin(), output->face_updates.end());
    EXPECT_EQ(all.size(), step.triangles.size());
    std::vector<std::set<size_t>> absolute;
    for (const auto& f : all) {
      const bool known = result_remapping.count(f.v1) && result_remapping.count(f.v2) && result_remapping.count(f.v3);
      EXPECT_TRUE(known);
      if (!known) continue;
      absolute.push_back({result_remapping.at(f.v1), result_remapping.at(f.v2), result_remapping.at(f.v3)});
      bool found = false;
      for (const auto& e : step.triangles) found = found || std::set<size_t>(e.begin(), e.end()) == absolute.back();
      EXPECT_TRUE(found);
    }
    for (const auto& e : step.triangles) {
      bool found = false;
      for (const auto& a : absolute) found = found || std::set<size_t>(e.begin(), e.end()) == a;
      EXPECT_TRUE(found);
    }
    std::vector<size_t> flat;  // flattenRemapping (:105-126)
    for (const auto& c : step.blocks) {
      const auto iter = remapping.find(BlockIndex(c.index[0], c.index[1], c.index[2]));
      if (iter == remapping.end()) continue;
      std::vector<size_t> block_remapping(iter->second.size());
      for (const auto& kv : iter->second) block_remapping.at(kv.first) = kv.second;
      flat.insert(flat.end(), block_remapping.begin(), block_remapping.end());
    }
    std::vector<size_t> abs_indices;
    for (const auto idx : flat) abs_indices.push_back(result_remapping.count(idx) ? result_remapping.at(idx) : 9999);
    EXPECT_TRUE(abs_indices == step.indices);
  }
}

static void testDeltaCompression() {
  const BlockConfig b1_v1 = block1_test1, b2_v1 = block2_test1;
  const BlockConfig b1_v2{{0, 0, 0},
                          {{{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                           {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};
  const BlockConfig b2_v2{{-1, 0, 0},
                          {{{{-0.5, 0.5, 0.5}, {-0.5, 0.75, 0.75}, {-0.5, 0.75, 0.5}}},
                           {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, -0.5}}}}};
  runDeltaConfig({  // MultiBlockPartialUpdates
      {-1, 100, {b1_v1}, 0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {0, 1, 2, 0, 1, 2, 6, 7, 8}},
      {-1, 102, {b2_v1}, 0, 0, 9, {{0, 1, 2}, {9, 10, 11}, {12, 13, 14}}, {9, 10, 11, 12, 13, 14}},
      {101, 103, {block1_empty, b2_v1}, 0, 0, 9, {{0, 1, 2}, {15, 16, 17}, {18, 19, 20}, {18, 19, 20}},
       {15, 16, 17, 18, 19, 20}},
      {-1, 104, {b2_v1}, 3, 1, 6, {{21, 22, 23}, {24, 25, 26}, {24, 25, 26}}, {21, 22, 23, 24, 25, 26}},
      {-1, 105, {b1_v1, b2_v1}, 3, 1, 9, {{27, 28, 29}, {33, 34, 35}, {36, 37, 38}, {33, 34, 35}},
       {27, 28, 29, 27, 28, 29, 33, 34, 35, 36, 37, 38, 33, 34, 35}},
      {-1, 106, {block1_empty, block2_empty}, 3, 1, 3, {{33, 34, 35}}, {}}});
  runDeltaConfig({  // MultiBlockPartialArchive
      {-1, 100, {b1_v2}, 0, 0, 6, {{0, 1, 2}, {3, 4, 5}}, {0, 1, 2, 3, 4, 5}},
      {-1, 102, {b2_v2}, 0, 0, 10, {{0, 1, 2}, {9, 10, 5}, {6, 7, 8}, {9, 10, 11}}, {6, 7, 8, 9, 10, 11}},
      {101, 103, {block1_empty, b2_v2}, 0, 0, 10, {{0, 1, 2}, {12, 13, 14}, {15, 16, 17}, {15, 16, 5}},
       {12, 13, 14, 15, 16, 17}},
      {-1, 104, {block2_empty}, 4, 1, 2, {{15, 16, 5}}, {}}});
  // archiveBlocks with an arbitrary filter + MeshDelta::updateMesh
  DeltaCompression c(1.0e-3);
  OrderedBlockMesh mesh = createMesh({block1_test1, block2_test1});
  auto d0 = c.update(mesh, 100ull * 1000000000ull);
  c.archiveBlocks([](const BlockIndex& idx, const BlockInfo& info) {
    return idx.x() == -1 && info.update_time_ns == 100ull * 1000000000ull && info.sequence_number == 1;
  });
  OrderedBlockMesh empty = createMesh({block1_empty});
  auto d1 = c.update(empty, 101ull * 1000000000ull);
  EXPECT_EQ(6u, d1->getNumArchivedVertices());  // block 2's 6 voxels (3 shared with the cleared block 1)
  pcl::PointCloud<pcl::PointXYZRGBA> vertices;
  std::vector<Timestamp> stamps;
  std::vector<pcl::Vertices> faces;
  d0->updateMesh(vertices, stamps, faces);
  EXPECT_EQ(9u, vertices.size());
  d1->updateMesh(vertices, stamps, faces);
  EXPECT_EQ(d1->getTotalVertices(), vertices.size());
  EXPECT_EQ(d1->getTotalFaces(), faces.size());

  // Semantic labels (delta_compression.cpp:76, :120, :128, :285, :347): a voxel keeps the label
  // of the vertex that set its point and hands it to the delta; MeshDelta::hasSemantics() iff
  // every vertex of the delta carries one.
  {
    DeltaCompression labelled(1.0e-3);
    OrderedBlockMesh m = createMesh({block1_test1, block2_test1});
    for (const auto& blk : m.blocks) {
      m.labels.emplace_back();
      for (size_t i = 0; i < blk.size(); ++i) m.labels.back().push_back(static_cast<uint32_t>(100 * m.labels.size() + i));
    }
    auto d = labelled.update(m, 100ull * 1000000000ull);
    EXPECT_TRUE(d->hasSemantics());
    EXPECT_EQ(d->vertex_updates->size(), d->semantic_updates.size());
    // the delta's vertices in first-observation order: block 1's three distinct corners of its
    // first face (labels 100..102; its second face repeats them with labels 103..105, which lose),
    // then its third face (106..108), then block 2's first face (200..202); block 2's second face
    // re-observes block 1's third-face voxels, which keep 106..108
    const std::vector<uint32_t> expected{100, 101, 102, 106, 107, 108, 200, 201, 202};
    EXPECT_EQ(expected.size(), d->semantic_updates.size());
    for (size_t i = 0; i < expected.size() && i < d->semantic_updates.size(); ++i) {
      EXPECT_EQ(expected[i], d->semantic_updates[i]);
    }
    // an unlabelled message afterwards: re-observed voxels lose their label, like info.label = nullopt
    OrderedBlockMesh plain = createMesh({block1_test1});
    auto d2 = labelled.update(plain, 101ull * 1000000000ull);
    EXPECT_TRUE(!d2->hasSemantics());
    EXPECT_EQ(size_t(3), d2->semantic_updates.size());  // block 2's own three voxels keep theirs
  }

  // DeltaCompression::update takes any MeshInterface (:278-288), a voxblox message included:
  // same delta as the same vertices handed over through a generic interface
  {
    VoxbloxMesh msg;
    msg.block_edge_length = 0.8f;
    uint32_t seed = 99u;
    auto next = [&seed]() {
      seed = seed * 1664525u + 1013904223u;
      return seed >> 8;
    };
    for (int b = 0; b < 3; ++b) {
      VoxbloxMeshBlock blk;
      blk.index[0] = b;
      blk.index[1] = -b;
      for (int v = 0; v < 60; ++v) {
        blk.x.push_back(static_cast<uint16_t>(next() % 65536));
        blk.y.push_back(static_cast<uint16_t>(next() % 65536));
        blk.z.push_back(static_cast<uint16_t>(next() % 65536));
        blk.r.push_back(static_cast<uint8_t>(next()));
        blk.g.push_back(static_cast<uint8_t>(next()));
        blk.b.push_back(static_cast<uint8_t>(next()));
      }
      msg.mesh_blocks.push_back(blk);
    }
    const VoxbloxMsgInterface voxblox(msg);
    OrderedBlockMesh generic;
    for (const auto& block : voxblox.blockIndices()) {
      generic.order.push_back(block);
      generic.blocks.emplace_back();
      voxblox.markBlockActive(block);
      for (size_t i = 0; i < voxblox.activeBlockSize(); ++i) generic.blocks.back().push_back(voxblox.getActiveVertex(i));
    }
    DeltaCompression a(0.05), b(0.05);
    HashedIndexMapping remap_a, remap_b;
    auto da = a.update(const_cast<VoxbloxMsgInterface&>(voxblox), 5ull * 1000000000ull, &remap_a);
    auto db = b.update(generic, 5ull * 1000000000ull, &remap_b);
    EXPECT_EQ(db->vertex_updates->size(), da->vertex_updates->size());
    EXPECT_EQ(db->face_updates.size(), da->face_updates.size());
    EXPECT_TRUE(da->vertex_updates->size() > 100);
    bool same = da->vertex_updates->size() == db->vertex_updates->size();
    for (size_t i = 0; same && i < da->vertex_updates->size(); ++i) {
      same = std::memcmp(&da->vertex_updates->at(i), &db->vertex_updates->at(i), sizeof(pcl::PointXYZRGBA)) == 0;
    }
    EXPECT_TRUE(same);
    EXPECT_TRUE(remap_a == remap_b);
  }
}

// ------------------------------------------------------------- MeshFrontendInterface (bookkeeping)
// mesh_frontend_interface.cpp:100-115, :197-211.  The expected mesh_to_graph_idx_ is the
// reference's own loop run on the host maps after every update, entries of a block taken in
// ascending msg index (the order the device uses where the reference's is implementation-defined).
// kimera_pgmo/test/test_traits.cpp:11-21 (stampTraitsCorrect): which clouds carry stamps is decided
// at compile time by the same has_get_stamp dispatch (the start-index search of
// test_traits.cpp:23-41 is findStartIndex's upper bound, exercised by testIncrementalWithPredeform)
static_assert(!traits::has_get_stamp<pcl::PointCloud<pcl::PointXYZ>>::value, "pcl::PointXYZ get failed");
static_assert(!traits::has_get_stamp<pcl::PointCloud<pcl::PointXYZRGBA>>::value, "pcl::PointXYZRGBA get failed");
static_assert(traits::has_get_stamp<ConstStampedCloud<pcl::PointXYZ>>::value, "const stamped pcl::PointXYZ failed");
static_assert(traits::has_get_stamp<ConstStampedCloud<pcl::PointXYZRGBA>>::value,
              "const stamped pcl::PointXYZRGBA failed");

// kimera_pgmo/test/test_mesh_delta.cpp:39-103, transcribed (host only: MeshDelta is what callers of
// DeltaCompression consume)
static void testMeshDelta() {
  auto point = [](float x, float y, float z, int r, int g, int b) {
    pcl::PointXYZRGBA p;
    p.x = x;
    p.y = y;
    p.z = z;
    p.r = static_cast<std::uint8_t>(r);
    p.g = static_cast<std::uint8_t>(g);
    p.b = static_cast<std::uint8_t>(b);
    p.a = 255;
    return p;
  };
  // kimera_pgmo/test/test_common_structs.cpp:12-20
  EXPECT_EQ(Timestamp(105867312345678), stampFromSec(105867.312345678));
  EXPECT_EQ(105867.312345678, stampToSec(Timestamp(105867312345678)));
  {  // constructor
    MeshDelta mesh;
    EXPECT_EQ(mesh.vertex_start, 0u);
    EXPECT_EQ(mesh.face_start, 0u);
  }
  {  // updateSimple
    MeshDelta delta1;
    for (int r = 0; r < 4; ++r) delta1.addVertex(100, point(1.0f, 2.0f, 3.0f, r, 0, 0));
    delta1.addFace({0, 1, 2});
    delta1.addFace({1, 2, 3});
    pcl::PointCloud<pcl::PointXYZRGBA> vertices;
    std::vector<pcl::Vertices> faces;
    std::vector<Timestamp> stamps;
    delta1.updateMesh(vertices, stamps, faces);
    EXPECT_EQ(stamps.size(), vertices.size());
    EXPECT_EQ(vertices.size(), 4u);
    for (size_t i = 0; i < vertices.size(); ++i) EXPECT_EQ(static_cast<size_t>(vertices.at(i).r), i);
    EXPECT_EQ(faces.size(), 2u);
    MeshDelta delta2(2, 1);
    for (int r = 2; r < 6; ++r) delta2.addVertex(100, point(1.0f, 2.0f, 3.0f, r, 0, 0));
    delta2.addFace({0, 1, 2});
    delta2.addFace({1, 2, 3});
    delta2.updateMesh(vertices, stamps, faces);
    EXPECT_EQ(stamps.size(), vertices.size());
    EXPECT_EQ(vertices.size(), 6u);
    for (size_t i = 0; i < vertices.size(); ++i) EXPECT_EQ(static_cast<size_t>(vertices.at(i).r), i);
    EXPECT_EQ(faces.size(), 3u);
    // beyond the reference's test: indices of the unarchived vertices, validation, labels
    EXPECT_EQ(delta2.getActiveIndices()->size(), 4u);
    EXPECT_EQ(delta2.getActiveIndices()->front(), 2);
    delta2.validate(vertices, stamps, faces);
    MeshDelta bad(0, 0);
    bad.addVertex(1, point(0, 0, 0, 0, 0, 0));
    bad.addFace({0, 1, 2});
    bool threw = false;
    try {
      bad.checkFaces("test");
    } catch (const std::runtime_error&) {
      threw = true;
    }
    EXPECT_TRUE(threw);
    MeshDelta labelled;
    labelled.addVertex(5, point(1, 1, 1, 0, 0, 0), 7u);
    labelled.addVertex(6, point(2, 2, 2, 0, 0, 0), 9u);
    EXPECT_TRUE(labelled.hasSemantics());
    std::vector<uint32_t> labels;
    labelled.updateMesh(vertices, stamps, faces, &labels);
    EXPECT_EQ(labels.size(), 2u);
    EXPECT_EQ(labels.at(1), 9u);
    EXPECT_EQ(stamps.at(1), Timestamp(6));
  }
  {  // archiveVerticesCorrect
    MeshDelta delta(1, 3);
    EXPECT_EQ(delta.addVertex(100, point(1.0f, 2.0f, 3.0f, 0, 0, 0), std::nullopt, true), 1u);
    EXPECT_EQ(delta.addVertex(100, point(1.0f, 2.0f, 3.0f, 0, 0, 0), std::nullopt, true), 2u);
    EXPECT_EQ(delta.addVertex(100, point(1.0f, 2.0f, 3.0f, 0, 0, 0), std::nullopt, false), 3u);
    EXPECT_EQ(delta.addVertex(100, point(1.0f, 2.0f, 3.0f, 0, 0, 0), std::nullopt, false), 4u);
    EXPECT_EQ(delta.getTotalArchivedVertices(), 3u);
    EXPECT_EQ(delta.getNumArchivedVertices(), 2u);
  }
  {  // archiveFacesCorrect
    MeshDelta delta(1, 3);
    delta.addFace({0, 1, 2}, true);
    delta.addFace({1, 2, 3}, true);
    delta.addFace({0, 1, 2}, false);
    delta.addFace({1, 2, 3}, false);
    EXPECT_EQ(delta.getTotalArchivedFaces(), 5u);
    EXPECT_EQ(delta.getNumArchivedFaces(), 2u);
  }
}

static void testMeshFrontend() {
  MeshFrontendInterface::Config config;
  config.full_compression_method = MeshFrontendInterface::FullCompressionMethod::VOXEL_CLEARING;
  config.graph_compression_method = MeshFrontendInterface::GraphCompressionMethod::OCTREE;
  config.mesh_resolution = 0.05;
  config.d_graph_resolution = 0.5;
  config.time_horizon = 2.0;
  MeshFrontendInterface frontend(config);
  std::map<size_t, size_t> expected;
  uint32_t seed = 777u;
  auto next = [&seed]() {
    seed = seed * 1664525u + 1013904223u;
    return seed >> 8;
  };
  for (int msg = 0; msg < 6; ++msg) {
    OrderedBlockMesh mesh;
    const int n_blocks = 2 + msg % 3;
    for (int b = 0; b < n_blocks; ++b) {
      mesh.order.push_back(BlockIndex((b + msg) % 4, 0, 0));
      mesh.blocks.emplace_back();
      // fewer faces in later versions of a block: stale entries stay in its maps
      const int faces = 60 - 7 * msg + 5 * b;
      for (int f = 0; f < faces; ++f) {
        const float ax = 0.8f * static_cast<float>((b + msg) % 4) + 0.02f * static_cast<float>(next() % 36);
        const float ay = 0.02f * static_cast<float>(next() % 40), az = 0.1f * static_cast<float>(next() % 3);
        const float corner[3][3] = {{ax, ay, az}, {ax + 0.04f, ay, az}, {ax, ay + 0.04f, az}};
        for (const auto& c : corner) {
          pcl::PointXYZRGBA p;
          p.x = c[0];
          p.y = c[1];
          p.z = c[2];
          p.r = 10;
          p.g = 20;
          p.b = 30;
          p.a = 255;
          mesh.blocks.back().push_back(p);
        }
      }
    }
    frontend.update(mesh, 10.0 + msg);
    EXPECT_EQ(frontend.vertices()->size(), frontend.vertexStamps()->size());
    EXPECT_TRUE(frontend.graphVertices()->size() > 0);
    for (const auto& block : mesh.order) {
      const IndexMapping& to_mesh = frontend.getMsgToMeshMapping().at(block);
      const IndexMapping& to_graph = frontend.getVoxbloxMsgToGraphMapping().at(block);
      const std::map<size_t, size_t> ordered(to_mesh.begin(), to_mesh.end());
      for (const auto& remap : ordered) {
        if (to_graph.count(remap.first)) expected.insert({remap.second, to_graph.at(remap.first)});
      }
    }
    const IndexMapping got = frontend.getMeshToGraphMapping();
    EXPECT_EQ(expected.size(), got.size());
    bool same = expected.size() == got.size();
    for (const auto& kv : expected) same = same && got.count(kv.first) == 1 && got.at(kv.first) == kv.second;
    EXPECT_TRUE(same);
  }
  EXPECT_TRUE(expected.size() > 50);
#ifndef SHIM_TEST_HOST_DOUBLE  // a device pointer: nothing the double can stand in for
  const auto view = frontend.meshToGraphDevice();
  EXPECT_TRUE(view.mesh_to_graph != nullptr && view.span >= frontend.vertices()->size());
#endif
}


int main(int argc, char** argv) {
  const bool with_octree = argc > 1 && std::strcmp(argv[1], "--octree") == 0;
  if (pgmo_device_count() <= 0) {
    std::printf("no CUDA device: the shim has no CPU fallback\n");
    return 77;
  }
  testDeformPoints();
  testDeformPointsWithTimeCheck();
  testEarlyReturnsAndErrors();
  testLargeCloudStaging();
  testDeviceGroup();
  testReconstructAndTranslate();
  testDeformMeshCube();
  testIncrementalWithPredeform();
  testVoxelClearing();
  testVoxbloxInterface();
  testDeltaCompression();
  testMeshDelta();
  testMeshFrontend();
  if (with_octree) testOctreeCompression();
  std::printf("shim tests: %d checks, %d failures\n", g_checks, g_failures);
  return g_failures ? 1 : 0;
}
