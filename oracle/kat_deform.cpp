// ORACLE — TEST INFRASTRUCTURE ONLY.
// Known-answer tests: the reference's own unit-test scenarios for the deformation path,
// replayed against the CPU restatement.  Expected values are the reference's:
//   kimera_pgmo/test/test_mesh_deformation.cpp:16-50   (deformPoints)
//   kimera_pgmo/test/test_mesh_deformation.cpp:52-112  (deformPointsWithTimeCheck)
//   kimera_pgmo/test/test_deformation_graph.cpp:163-201 (reconstructMesh)
//   kimera_pgmo/test/test_deformation_graph.cpp:203-248 (deformMeshTranslation)
//   kimera_pgmo/test/test_deformation_graph.cpp:250-309 (deformMesh, incl. 1.273 at :306)
//   kimera_pgmo/test/test_common_structs.cpp:12-15      (stampFromSec)
// The solver (GTSAM/Kimera-RPGO) is out of scope, so the poses the solver would return in the
// three DeformationGraph scenarios are set by hand to the values those tests imply.
#include <cmath>
#include <cstdio>
#include <vector>

#include "deform_oracle.h"
#include "kat_check.h"

using namespace pgmo_oracle;

namespace {

struct Cloud {
  std::vector<float> x, y, z;
  std::vector<Timestamp> stamps;
  void push(float a, float b, float c) {
    x.push_back(a);
    y.push_back(b);
    z.push_back(c);
  }
  DeformIO io(bool stamped) {
    DeformIO v;
    v.num_vertices = x.size();
    v.in_x = x.data();
    v.in_y = y.data();
    v.in_z = z.data();
    v.out_x = x.data();  // in place, like deformation_graph.cpp:545
    v.out_y = y.data();
    v.out_z = z.data();
    v.stamps = stamped ? stamps.data() : nullptr;
    return v;
  }
};

Pose3 translation(double x, double y, double z) {
  Pose3 p;
  p.t[0] = x;
  p.t[1] = y;
  p.t[2] = z;
  return p;
}

void katDeformPoints() {
  Cloud pts;
  std::vector<Point3d> ctrl;
  std::vector<Pose3> poses;
  for (size_t i = 0; i < 100; i++) {
    pts.push(static_cast<float>(i), 0.f, 0.f);
    if (i % 10 == 0) {
      ctrl.push_back({static_cast<double>(i), 0.0, 0.0});
      poses.push_back(translation(static_cast<double>(i), 1.0, 0.0));
    }
  }
  std::vector<uint8_t> valid(ctrl.size(), 1);
  ValuesView values{valid.data(), poses.data(), poses.size()};
  std::vector<std::set<size_t>> map;
  const int rc = deformPoints(pts.io(false), map, ctrl, {}, values, 4, 10.0, nullptr);
  KAT_EQ(rc, 0);
  KAT_EQ(map.size(), size_t(100));
  for (size_t i = 0; i < 100; i++) {
    KAT_NEAR(static_cast<double>(i), pts.x[i], 1.0e-9);
    KAT_NEAR(1.0, pts.y[i], 1.0e-9);
    KAT_NEAR(0.0, pts.z[i], 1.0e-9);
  }
}

void katDeformPointsWithTimeCheck() {
  Cloud pts;
  std::vector<Point3d> ctrl;
  std::vector<Timestamp> ctrl_stamps;
  std::vector<Pose3> poses;
  for (size_t i = 0; i < 100; i++) {
    pts.push(static_cast<float>(i), 0.f, 0.f);
    if (i % 10 == 0) {
      ctrl.push_back({static_cast<double>(i), 0.0, 0.0});
      if (i > 50) {
        poses.push_back(translation(static_cast<double>(i), 1.0, 0.0));
        ctrl_stamps.push_back(stampFromSec(20.0));
      } else {
        poses.push_back(translation(static_cast<double>(i), -1.0, 0.0));
        ctrl_stamps.push_back(0);
      }
    }
    pts.stamps.push_back(i < 50 ? 0 : stampFromSec(20.0));
  }
  std::vector<uint8_t> valid(ctrl.size(), 1);
  ValuesView values{valid.data(), poses.data(), poses.size()};
  std::vector<std::set<size_t>> map;
  const int rc = deformPoints(pts.io(true), map, ctrl, ctrl_stamps, values, 3, 10.0, nullptr);
  KAT_EQ(rc, 0);
  for (size_t i = 0; i < 50; i++) {
    KAT_EQ(static_cast<double>(i), static_cast<double>(pts.x[i]));
    KAT_EQ(-1.0, static_cast<double>(pts.y[i]));
    KAT_EQ(0.0, static_cast<double>(pts.z[i]));
  }
  for (size_t i = 51; i < 100; i++) {
    KAT_EQ(static_cast<double>(i), static_cast<double>(pts.x[i]));
    KAT_EQ(1.0, static_cast<double>(pts.y[i]));
    KAT_EQ(0.0, static_cast<double>(pts.z[i]));
  }
}

// createMeshTriangle (test_deformation_graph.cpp:38-54): the 3 control points of prefix 'v'
std::vector<Point3d> triangleControls() { return {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}}; }

// createSimpleMesh (pgmo_fixtures.cpp:97-158): 5 vertices
Cloud simpleMesh() {
  Cloud c;
  c.push(0, 0, 0);
  c.push(1, 0, 0);
  c.push(0, 1, 0);
  c.push(1, 1, 0);
  c.push(0, 0, 1);
  c.stamps.assign(5, 0);
  return c;
}

void katReconstructAndTranslate() {
  const auto ctrl = triangleControls();
  const std::vector<Timestamp> ctrl_stamps(3, 0);
  const std::vector<int> graph_inds = {-1, 0, -1, 0, -1};  // SetUpOriginalMesh :134-147
  std::vector<uint8_t> valid(3, 1);
  for (int scenario = 0; scenario < 2; ++scenario) {
    const double shift = scenario == 0 ? 0.0 : 0.5;  // reconstructMesh / deformMeshTranslation
    std::vector<Pose3> poses;
    for (const auto& g : ctrl) poses.push_back(translation(g.x + shift, g.y, g.z));
    ValuesView values{valid.data(), poses.data(), poses.size()};
    DeformationGraphMeshPath graph;
    graph.setControlPoints('v', ctrl, ctrl_stamps);
    for (size_t k : {size_t(1), size_t(2)}) {
      Cloud mesh = simpleMesh();
      const Cloud orig = simpleMesh();
      graph.setRecalculate(true);  // optimize() with force_recalculate_ (deformation_graph.cpp:646-648)
      graph.deformPoints(mesh.io(true), 'v', values, k, 10.0, &graph_inds, -1, nullptr);
      for (size_t i = 0; i < 5; ++i) {
        KAT_NEAR(orig.x[i] + shift, mesh.x[i], 1e-6);
        KAT_NEAR(orig.y[i], mesh.y[i], 1e-6);
        KAT_NEAR(orig.z[i], mesh.z[i], 1e-6);
      }
      if (scenario == 0) {
        KAT_EQ(1.0, static_cast<double>(mesh.x[1]));
        KAT_EQ(1.0, static_cast<double>(mesh.y[2]));
        KAT_EQ(1.0, static_cast<double>(mesh.z[4]));
      }
    }
  }
}

void katCube() {
  // test/data/cube.ply vertices
  const float cube[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
                            {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  const auto ctrl = triangleControls();
  const std::vector<Timestamp> ctrl_stamps(3, 0);
  const std::vector<int> graph_inds(8, -1);
  std::vector<uint8_t> valid(3, 1);
  DeformationGraphMeshPath graph;
  graph.setControlPoints('v', ctrl, ctrl_stamps);

  {  // phase 1: all nodes pulled by (-0.5, 0, 0); k = 3 shrinks to 2 (3 control points)
    std::vector<Pose3> poses;
    for (const auto& g : ctrl) poses.push_back(translation(g.x - 0.5, g.y, g.z));
    ValuesView values{valid.data(), poses.data(), poses.size()};
    Cloud mesh;
    for (auto& v : cube) mesh.push(v[0], v[1], v[2]);
    mesh.stamps.assign(8, 0);
    graph.setRecalculate(true);
    DeformTrace trace;
    graph.deformPoints(mesh.io(true), 'v', values, 3, 10.0, &graph_inds, -1, nullptr, &trace);
    KAT_EQ(trace.order.size(), size_t(8));
    for (const auto& o : trace.order) KAT_EQ(o.size(), size_t(2));
    for (size_t i = 0; i < 8; ++i) {
      KAT_NEAR(cube[i][0] - 0.5, mesh.x[i], 1e-6);
      KAT_NEAR(cube[i][1], mesh.y[i], 1e-6);
      KAT_NEAR(cube[i][2], mesh.z[i], 1e-6);
    }
  }
  {  // phase 2: node 0 at (-0.5,0,0), node 1 at (1.5,0,0); k = 2.  Pins the LINEAR weight:
     // x(v1) = (1.5 + (1 - 1/sqrt2) * 0.5) / (2 - 1/sqrt2) = 1.27346 (test expects 1.273 +- 1e-3)
    std::vector<Pose3> poses = {translation(-0.5, 0, 0), translation(1.5, 0, 0),
                                translation(0.5, 1, 0)};
    ValuesView values{valid.data(), poses.data(), poses.size()};
    Cloud mesh;
    for (auto& v : cube) mesh.push(v[0], v[1], v[2]);
    mesh.stamps.assign(8, 0);
    graph.setRecalculate(true);
    graph.deformPoints(mesh.io(true), 'v', values, 2, 10.0, &graph_inds, -1, nullptr);
    KAT_NEAR(-0.5, mesh.x[0], 0.001);
    KAT_NEAR(1.273, mesh.x[1], 0.001);
    KAT_NEAR((1.5 + (1.0 - 1.0 / std::sqrt(2.0)) * 0.5) / (2.0 - 1.0 / std::sqrt(2.0)),
             mesh.x[1], 1e-6);
  }
}

void katStampFromSec() {
  KAT_EQ(stampFromSec(105867.312345678), Timestamp(105867312345678ull));
}

}  // namespace

int runDeformKats() {
  katDeformPoints();
  katDeformPointsWithTimeCheck();
  katReconstructAndTranslate();
  katCube();
  katStampFromSec();
  return kat::failures();
}
