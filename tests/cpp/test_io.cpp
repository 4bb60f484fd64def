// Host-only test of include/kimera_pgmo_b200/io.h: the C++ PLY / .dgrf readers against files the
// Python side wrote (kimera_pgmo_b200/formats.py, itself pinned by the reference's fixtures) and
// the prefix tables of kimera_pgmo/utils/common_functions.h:19-81.
//   test_io <ascii.ply> <binary.ply> <graph.dgrf> <out.ply>
#include <cmath>
#include <cstdio>

#include "kimera_pgmo_b200/io.h"

using namespace kimera_pgmo_b200;

static int g_failures = 0;
#define CHECK(a)                                                      \
  do {                                                                \
    if (!(a)) {                                                       \
      ++g_failures;                                                   \
      std::fprintf(stderr, "%s:%d CHECK(%s)\n", __FILE__, __LINE__, #a); \
    }                                                                 \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  // ---- prefixes (common_functions.h:19-81)
  CHECK(GetRobotPrefix(0) == 'a' && GetRobotPrefix(7) == 'h' && GetRobotPrefix(8) == '\0');
  CHECK(GetVertexPrefix(0) == 's' && GetVertexPrefix(3) == 'v' && GetVertexPrefix(7) == 'z' &&
        GetVertexPrefix(100) == '\0');
  CHECK(robot_id_to_prefix.size() == 8 && robot_id_to_prefix.at(2) == 'c');
  CHECK(robot_prefix_to_id.count('h') == 1 && robot_prefix_to_id.count('i') == 0);
  CHECK(robot_id_to_vertex_prefix.at(1) == 't' && vertex_prefix_to_id.at('z') == 7);
  CHECK(vertex_prefix_to_id.count('a') == 0);
  // test/data/graph.dgrf: 'a' -> 6989586621679009792, 'v' -> 8502796096475496448
  CHECK(gtsam::Symbol('a', 0).key() == 6989586621679009792ull);
  CHECK(gtsam::Symbol(GetVertexPrefix(3), 2).key() == 8502796096475496450ull);

  // ---- ASCII PLY written by formats.write_ply: sphere.ply's vertices + colours, stamps, labels
  const IOData::Ptr a = IOData::load(argv[1]);
  CHECK(a->x.size() == 422 && a->faces.size() == 840);
  CHECK(a->r.size() == 422 && a->a.size() == 422 && a->stamps.size() == 422 && a->labels.size() == 422);
  CHECK(a->z[0] == -127.f && a->x[1] == 5.f && a->x[420] == -1.f && a->z[421] == 127.f);  // test_mesh_conversion.cpp:36-44
  for (size_t i = 0; i < a->x.size(); ++i) {
    CHECK(a->stamps[i] == 1700000000000000000ull + i * 1234567891ull);
    CHECK(a->labels[i] == 3 * i);
    CHECK(a->r[i] == (4 * i) % 251 && a->a[i] == (4 * i + 3) % 251);
    if (g_failures) break;
  }
  // ---- through the PolygonMesh functions and back out (read by Python afterwards)
  pcl::PolygonMesh mesh;
  std::vector<Timestamp> stamps;
  ReadMeshWithStampsFromPly(argv[1], mesh, &stamps);
  CHECK(mesh.polygons.size() == 840 && stamps.size() == 422 && mesh.cloud.width == 422);
  CHECK(mesh.polygons[100].vertices.size() == 3);
  WriteMeshWithStampsToPly(argv[4], mesh, stamps);
  const IOData::Ptr again = IOData::load(argv[4]);
  CHECK(again->x == a->x && again->y == a->y && again->z == a->z);
  CHECK(again->r == a->r && again->g == a->g && again->b == a->b && again->a == a->a);
  CHECK(again->stamps == a->stamps && again->faces == a->faces);
  CHECK(again->labels.empty());  // WriteMesh of the reference carries no labels either

  // ---- binary_little_endian with stamps beyond 2^32 ns and no colours
  const IOData::Ptr b = IOData::load(argv[2]);
  CHECK(b->x.size() == 2 && b->r.empty() && b->labels.empty());
  CHECK(b->x[0] == 0.5f && b->y[0] == -1.25f && b->z[1] == -7.5f && b->x[1] == 1e-3f);
  CHECK(b->stamps.size() == 2 && b->stamps[0] == 10000000005ull &&
        b->stamps[1] == 4000000000ull * 1000000000ull + 999999999ull);
  CHECK(b->faces.size() == 1 && b->faces[0] == (std::vector<uint32_t>{0, 1, 0}));

  // ---- .dgrf (test_deformation_graph.cpp:824-860): 3 pose nodes 'a', 3 mesh vertices 'v'
  DeformationGraphMeshPath graph;
  gtsam::Values values;
  const GraphFileSummary s = loadGraphFile(argv[3], &graph, &values);
  CHECK(s.nodes == 6 && s.vertices == 3 && s.solver_lines > 0);
  CHECK(graph.getNumVertices('v') == 3 && graph.getNumVertices('a') == 0);
  CHECK(values.size() == 6);
  CHECK(values.exists(gtsam::Symbol('a', 2)) && values.exists(gtsam::Symbol('v', 1)) &&
        !values.exists(gtsam::Symbol('v', 3)));
  const gtsam::Pose3& pa = values.at<gtsam::Pose3>(gtsam::Symbol('a', 2));
  CHECK(pa.translation().x() == 3.0 && pa.translation().y() == 2.1 && pa.translation().z() == 2.1);
  CHECK(pa.rotation().matrix()(0, 0) == 1.0 && pa.rotation().matrix()(0, 1) == 0.0 &&
        pa.rotation().matrix()(2, 2) == 1.0);
  // loading as robot 2: 'a' -> 'c', 'v' -> 'u' (rekey, deformation_graph_io.cpp:178-191)
  DeformationGraphMeshPath moved;
  gtsam::Values moved_values;
  loadGraphFile(argv[3], &moved, &moved_values, true, 2);
  CHECK(moved.getNumVertices('u') == 3 && moved.getNumVertices('v') == 0);
  CHECK(moved_values.exists(gtsam::Symbol('c', 0)) && moved_values.exists(gtsam::Symbol('u', 2)) &&
        !moved_values.exists(gtsam::Symbol('a', 0)));
  // a quarter turn about z through the quaternion constructor
  const double h = std::sqrt(0.5);
  const gtsam::Rot3 rz(h, 0.0, 0.0, h);
  CHECK(std::fabs(rz.matrix()(0, 1) + 1.0) < 1e-15 && std::fabs(rz.matrix()(1, 0) - 1.0) < 1e-15 &&
        std::fabs(rz.matrix()(0, 0)) < 1e-15 && rz.matrix()(2, 2) == 1.0);

  bool threw = false;
  try {
    IOData::load(std::string(argv[4]) + ".missing");
  } catch (const std::runtime_error&) {
    threw = true;
  }
  CHECK(threw);
  std::printf("test_io: %d failures\n", g_failures);
  return g_failures ? 1 : 0;
}
