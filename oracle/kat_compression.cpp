// ORACLE — TEST INFRASTRUCTURE ONLY.
// Known-answer tests: the reference's own unit-test scenarios for the compression path,
// replayed against the CPU restatement.  Expected values are the reference's:
//   kimera_pgmo/test/test_voxel_clearing_compression.cpp:106-356 (5 tests)
//   kimera_pgmo/test/test_octree_compression.cpp:18-364          (7 tests)
//   kimera_pgmo/test/test_mesh_frontend_interface.cpp:192-345, :523-607
//     (MeshFrontendTest with OCTREE/OCTREE: the frontend is prune(t - horizon) followed by
//      compressAndIntegrate with a persistent remapping, mesh_frontend_interface.cpp:118-195)
//   kimera_pgmo/test/test_common_functions.cpp:177-198            (SurfaceExists)
#include <set>

#include "compression_oracle.h"
#include "kat_check.h"

using namespace pgmo_oracle;

namespace {

using FaceCoords = std::array<std::array<float, 3>, 3>;

struct BlockConfig {
  BlockIndex index;
  std::vector<FaceCoords> faces;
};

// pgmo_fixtures.cpp:68-95 createMesh
FlatBlockMesh createMesh(const std::vector<BlockConfig>& configs) {
  FlatBlockMesh mesh;
  mesh.offsets.push_back(0);
  uint8_t point_index = 0;  // BlockConfig::point_index is never reset by the tests we replay
  for (const auto& c : configs) {
    mesh.blocks.push_back(c.index);
    for (const auto& f : c.faces) {
      for (const auto& v : f) {
        PointXYZRGBA p;
        p.x = v[0];
        p.y = v[1];
        p.z = v[2];
        p.rgba = 0xFF000000u | (uint32_t(point_index) << 16) | (uint32_t(point_index) << 8) |
                 point_index;
        ++point_index;
        mesh.vertices.push_back(p);
      }
    }
    mesh.offsets.push_back(mesh.vertices.size());
  }
  return mesh;
}

const BlockConfig block1_empty{{0, 0, 0}, {}};
const BlockConfig block1_test1{{0, 0, 0},
                               {{{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                                {{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                                {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};
const BlockConfig block2_empty{{-1, 0, 0}, {}};
const BlockConfig block2_test1{{-1, 0, 0},
                               {{{{-0.5, 0.5, 0.5}, {-0.5, 0.75, 0.75}, {-0.5, 0.75, 0.5}}},
                                {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};

// test_voxel_clearing_compression.cpp:59-104 checkTriangles (order-sensitive after dedup)
bool checkTriangles(const std::vector<std::vector<uint32_t>>& expected,
                    const std::vector<Face>& result) {
  std::vector<Face> filtered;
  for (const auto& f : result) {
    bool seen = false;
    for (const auto& g : filtered) seen |= (g == f);
    if (!seen) filtered.push_back(f);
  }
  if (expected.size() != filtered.size()) return false;
  for (size_t i = 0; i < expected.size(); ++i) {
    std::set<uint32_t> e(expected[i].begin(), expected[i].end());
    std::set<uint32_t> r(filtered[i].begin(), filtered[i].end());
    if (e != r) return false;
  }
  return true;
}

struct Scratch {
  std::vector<PointXYZRGBA> vertices;
  std::vector<Face> triangles;
  std::vector<size_t> indices;
  HashedIndexMapping remappings;
};

void integrate(MeshCompression& c, const FlatBlockMesh& mesh, double t) {
  Scratch s;
  c.compressAndIntegrate(mesh, s.vertices, s.triangles, s.indices, s.remappings, t);
}

const double compression_factor = 1.0e-3;
using V = std::vector<size_t>;

void katVoxelClearingBasic() {
  VoxelClearingCompression c(compression_factor);
  integrate(c, createMesh({block1_test1}), 100.0);
  KAT_EQ(c.getVertices().size(), 6u);
  KAT_EQ(c.getInvalidIndices().size(), 0u);
  KAT_EQ(c.getTimestamps().size(), 6u);
  KAT_TRUE(c.getActiveVerticesIndex() == (V{0, 1, 2, 3, 4, 5}));
  KAT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}}, c.getStoredPolygons()));

  integrate(c, createMesh({block1_empty}), 100.0);
  KAT_EQ(c.getVertices().size(), 6u);
  KAT_EQ(c.getInvalidIndices().size(), 6u);
  KAT_EQ(c.getStoredPolygons().size(), 0u);
  KAT_EQ(c.getActiveVerticesIndex().size(), 0u);
  KAT_EQ(c.getTimestamps().size(), 6u);
}

void katVoxelClearingPruningBasic() {
  VoxelClearingCompression c(compression_factor);
  integrate(c, createMesh({block1_test1}), 100.0);
  KAT_EQ(c.getVertices().size(), 6u);
  KAT_TRUE(c.getActiveVerticesIndex() == (V{0, 1, 2, 3, 4, 5}));
  c.pruneStoredMesh(101.0);
  integrate(c, createMesh({block1_empty}), 105.0);
  KAT_EQ(c.getVertices().size(), 6u);
  KAT_EQ(c.getInvalidIndices().size(), 0u);
  KAT_EQ(c.getActiveVerticesIndex().size(), 0u);
  KAT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}}, c.getStoredPolygons()));
  integrate(c, createMesh({block1_test1}), 105.0);
  KAT_EQ(c.getVertices().size(), 12u);
  KAT_EQ(c.getInvalidIndices().size(), 0u);
  KAT_TRUE(c.getActiveVerticesIndex() == (V{6, 7, 8, 9, 10, 11}));
  KAT_EQ(c.getTimestamps().size(), 12u);
  KAT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {9, 10, 11}}, c.getStoredPolygons()));
}

void katVoxelClearingMultiblock() {
  VoxelClearingCompression c(compression_factor);
  integrate(c, createMesh({block1_test1, block2_test1}), 100.0);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getInvalidIndices().size(), 0u);
  KAT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}, {0, 1, 2}}, c.getStoredPolygons()));

  integrate(c, createMesh({block1_empty, block2_test1}), 100.0);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getInvalidIndices().size(), 3u);
  KAT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}}, c.getStoredPolygons()));

  integrate(c, createMesh({block1_test1, block2_test1}), 100.0);
  KAT_EQ(c.getVertices().size(), 12u);
  KAT_EQ(c.getInvalidIndices().size(), 3u);
  KAT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}, {9, 10, 11}}, c.getStoredPolygons()));

  integrate(c, createMesh({block1_test1, block2_empty}), 100.0);
  KAT_EQ(c.getVertices().size(), 12u);
  KAT_EQ(c.getInvalidIndices().size(), 6u);
  KAT_TRUE(checkTriangles({{9, 10, 11}, {3, 4, 5}}, c.getStoredPolygons()));
}

void katVoxelClearingPruningMultiblock() {
  VoxelClearingCompression c(compression_factor);
  integrate(c, createMesh({block1_test1}), 100.0);
  KAT_EQ(c.getVertices().size(), 6u);
  KAT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}}, c.getStoredPolygons()));
  integrate(c, createMesh({block2_test1}), 102.0);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getInvalidIndices().size(), 0u);
  KAT_TRUE(checkTriangles({{6, 7, 8}, {3, 4, 5}, {0, 1, 2}}, c.getStoredPolygons()));
  c.pruneStoredMesh(101.0);
  integrate(c, createMesh({block1_empty, block2_empty}), 105.0);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getInvalidIndices().size(), 3u);
  KAT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}}, c.getStoredPolygons()));
  integrate(c, createMesh({block1_test1}), 105.0);
  KAT_EQ(c.getVertices().size(), 15u);
  KAT_EQ(c.getInvalidIndices().size(), 3u);
  KAT_EQ(c.getTimestamps().size(), 15u);
  KAT_TRUE(checkTriangles({{0, 1, 2}, {3, 4, 5}, {9, 10, 11}, {12, 13, 14}}, c.getStoredPolygons()));
}

// PclMeshInterface over createSimpleMesh(scale) (pgmo_fixtures.cpp:97-163)
FlatBlockMesh simpleMeshInterface(double scale) {
  const float s = static_cast<float>(scale);
  const float v[5][3] = {{0, 0, 0}, {s, 0, 0}, {0, s, 0}, {s, s, 0}, {0, 0, s}};
  const uint32_t rgba[5] = {0xFF17187Au, 0xFF212234u, 0xFF0C9016u, 0xFF000E00u, 0xFF90000Cu};
  const int faces[4][3] = {{0, 1, 2}, {1, 3, 2}, {0, 1, 4}, {0, 4, 2}};
  FlatBlockMesh mesh;
  mesh.blocks.push_back({0, 0, 0});
  mesh.offsets.push_back(0);
  for (auto& f : faces) {
    for (int idx : f) {
      PointXYZRGBA p;
      p.x = v[idx][0];
      p.y = v[idx][1];
      p.z = v[idx][2];
      p.rgba = rgba[idx];
      mesh.vertices.push_back(p);
    }
  }
  mesh.offsets.push_back(mesh.vertices.size());
  return mesh;
}

void katOctreeBoundingBox() {
  // test_octree_compression.cpp:18-31
  std::vector<PointXYZ> cloud(2);
  cloud[1].x = cloud[1].y = cloud[1].z = 1.1f;
  OctreeSearchRestated octree(1.0);
  octree.setInputCloud(&cloud);
  octree.addPointsFromInputCloud();
  double a, b, c, d, e, f;
  octree.getBoundingBox(a, b, c, d, e, f);
  auto inside = [&](float x, float y, float z) {
    return (x >= a && x <= d) && (y >= b && y <= e) && (z >= c && z <= f);
  };
  KAT_TRUE(inside(0.5f, 0.5f, 0.5f));
  KAT_TRUE(!inside(10.f, 10.f, 10.f));
}

void katOctreeReturnedAndStored() {
  OctreeCompression c(0.1);
  Scratch s;
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 100.0);
  KAT_EQ(s.vertices.size(), 5u);
  KAT_EQ(s.triangles.size(), 4u);
  KAT_TRUE(s.indices == (V{0, 1, 2, 3, 4}));
  KAT_EQ(s.vertices[1].x, 1.f);
  KAT_EQ(s.vertices[2].y, 1.f);
  KAT_EQ(s.vertices[4].z, 1.f);
  KAT_EQ(s.triangles.at(0)[0], 0u);
  KAT_EQ(s.triangles.at(3)[2], 2u);
  KAT_EQ(c.getVertices().size(), 5u);
  KAT_EQ(c.getActiveVertices().size(), 5u);
  KAT_EQ(c.getStoredPolygons().size(), 4u);
  KAT_EQ(c.getActiveVerticesTimestamps().at(4), 100.0);

  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(2.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 101.0);
  KAT_EQ(s.vertices.size(), 4u);
  KAT_EQ(s.triangles.size(), 4u);
  KAT_TRUE(s.indices == (V{0, 5, 6, 7, 8}));
  KAT_EQ(s.vertices[0].x, 2.f);
  KAT_EQ(s.vertices[2].y, 2.f);
  KAT_EQ(s.vertices[3].z, 2.f);
  KAT_EQ(s.triangles.at(0)[0], 0u);
  KAT_EQ(s.triangles.at(3)[2], 6u);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getActiveVertices().size(), 9u);
  KAT_EQ(c.getStoredPolygons().size(), 8u);
  KAT_EQ(c.getVertices()[5].x, 2.f);
  KAT_EQ(c.getVertices()[7].y, 2.f);
  KAT_EQ(c.getVertices()[8].z, 2.f);
  KAT_EQ(c.getStoredPolygons().at(4)[0], 0u);
  KAT_EQ(c.getStoredPolygons().at(7)[2], 6u);
  KAT_EQ(c.getActiveVerticesTimestamps().at(8), 101.0);
}

void katOctreePrune() {
  OctreeCompression c(0.1);
  Scratch s;
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 100.0);
  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(2.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 101.0);
  c.pruneStoredMesh(100.5);
  KAT_EQ(c.getActiveVertices().size(), 5u);
  KAT_EQ(c.getStoredPolygons().size(), 8u);
  KAT_EQ(c.getActiveVerticesTimestamps().size(), 5u);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getActiveVertices()[1].x, 2.f);
  KAT_EQ(c.getActiveVertices()[2].y, 2.f);
  KAT_EQ(c.getActiveVertices()[4].z, 2.f);
  KAT_EQ(c.getActiveVerticesTimestamps().at(4), 101.0);
  KAT_EQ(c.getActiveVerticesTimestamps().at(0), 101.0);

  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 102.0);
  c.pruneStoredMesh(100.9);
  KAT_EQ(c.getActiveVertices().size(), 9u);
  KAT_EQ(c.getStoredPolygons().size(), 12u);
  KAT_EQ(c.getVertices().size(), 13u);
  KAT_EQ(c.getActiveVertices()[5].x, 1.f);
  KAT_EQ(c.getActiveVertices()[7].y, 1.f);
  KAT_EQ(c.getActiveVertices()[8].z, 1.f);
  KAT_EQ(c.getStoredPolygons().at(8)[0], 0u);
  KAT_EQ(c.getStoredPolygons().at(11)[2], 10u);
  KAT_EQ(c.getActiveVerticesTimestamps().at(8), 102.0);
  KAT_EQ(s.vertices.size(), 4u);
  KAT_EQ(s.triangles.size(), 4u);
  KAT_TRUE(s.indices == (V{0, 9, 10, 11, 12}));
  KAT_EQ(s.vertices[0].x, 1.f);
  KAT_EQ(s.vertices[2].y, 1.f);
  KAT_EQ(s.vertices[3].z, 1.f);
  KAT_EQ(s.triangles.at(0)[0], 0u);
  KAT_EQ(s.triangles.at(1)[1], 11u);
  KAT_EQ(s.triangles.at(3)[2], 10u);
}


// kimera_pgmo/test/test_block_compression.cpp:40-160 (returnedValues + storedValues),
// :161-260 (pruneStoredMesh), :262-340 (…Compressed)
template <typename C>
void katSimpleMeshFlow(double res) {
  C c(res);
  Scratch s;
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 100.0);
  KAT_EQ(s.vertices.size(), 5u);
  KAT_EQ(s.triangles.size(), 4u);
  KAT_TRUE(s.indices == (V{0, 1, 2, 3, 4}));
  KAT_EQ(s.vertices[1].x, 1.f);
  KAT_EQ(s.vertices[2].y, 1.f);
  KAT_EQ(s.vertices[4].z, 1.f);
  KAT_EQ(s.triangles.at(0)[0], 0u);
  KAT_EQ(s.triangles.at(3)[2], 2u);
  KAT_EQ(c.getVertices().size(), 5u);
  KAT_EQ(c.getActiveVertices().size(), 5u);
  KAT_EQ(c.getStoredPolygons().size(), 4u);
  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(2.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 101.0);
  KAT_EQ(s.vertices.size(), 4u);
  KAT_EQ(s.triangles.size(), 4u);
  KAT_TRUE(s.indices == (V{0, 5, 6, 7, 8}));
  KAT_EQ(s.vertices[0].x, 2.f);
  KAT_EQ(s.vertices[2].y, 2.f);
  KAT_EQ(s.vertices[3].z, 2.f);
  KAT_EQ(s.triangles.at(0)[0], 0u);
  KAT_EQ(s.triangles.at(3)[2], 6u);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getStoredPolygons().size(), 8u);
  c.pruneStoredMesh(100.5);  // test_block_compression.cpp:190-212
  KAT_EQ(c.getActiveVertices().size(), 5u);
  KAT_EQ(c.getStoredPolygons().size(), 8u);
  KAT_EQ(c.getActiveVerticesTimestamps().size(), 5u);
  KAT_EQ(c.getVertices().size(), 9u);
  KAT_EQ(c.getActiveVertices()[1].x, 2.f);
  KAT_EQ(c.getActiveVertices()[2].y, 2.f);
  KAT_EQ(c.getActiveVertices()[4].z, 2.f);
  KAT_EQ(c.getStoredPolygons().at(0)[0], 0u);
  KAT_EQ(c.getStoredPolygons().at(3)[2], 2u);
  KAT_EQ(c.getActiveVerticesTimestamps().at(4), 101.0);
  KAT_EQ(c.getActiveVerticesTimestamps().at(0), 101.0);
  KAT_EQ(c.getVertices()[5].x, 2.f);
  KAT_EQ(c.getVertices()[7].y, 2.f);
  KAT_EQ(c.getVertices()[8].z, 2.f);
  KAT_EQ(c.getStoredPolygons().at(4)[0], 0u);
  KAT_EQ(c.getStoredPolygons().at(7)[2], 6u);
  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 102.0);
  c.pruneStoredMesh(100.9);  // :214-258
  KAT_EQ(c.getActiveVertices().size(), 9u);
  KAT_EQ(c.getStoredPolygons().size(), 12u);
  KAT_EQ(c.getActiveVerticesTimestamps().size(), 9u);
  KAT_EQ(c.getVertices().size(), 13u);
  KAT_EQ(c.getActiveVertices()[5].x, 1.f);
  KAT_EQ(c.getActiveVertices()[7].y, 1.f);
  KAT_EQ(c.getActiveVertices()[8].z, 1.f);
  KAT_EQ(c.getStoredPolygons().at(4)[0], 0u);
  KAT_EQ(c.getStoredPolygons().at(7)[2], 6u);
  KAT_EQ(c.getStoredPolygons().at(8)[0], 0u);
  KAT_EQ(c.getStoredPolygons().at(11)[2], 10u);
  KAT_EQ(c.getActiveVerticesTimestamps().at(8), 102.0);
  KAT_EQ(s.vertices.size(), 4u);
  KAT_EQ(s.triangles.size(), 4u);
  KAT_EQ(s.indices.size(), 5u);
  KAT_TRUE(s.indices == (V{0, 9, 10, 11, 12}));
  KAT_EQ(s.vertices[0].x, 1.f);
  KAT_EQ(s.vertices[2].y, 1.f);
  KAT_EQ(s.vertices[3].z, 1.f);
  KAT_EQ(s.triangles.at(0)[0], 0u);
  KAT_EQ(s.triangles.at(1)[1], 11u);
  KAT_EQ(s.triangles.at(3)[2], 10u);
}

void katBlockCompression() {
  katSimpleMeshFlow<BlockCompression>(0.1);
  // test_block_compression.cpp:262-300: a 10 m cell swallows the whole mesh, nothing survives
  BlockCompression c(10.0);
  Scratch s;
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 100.0);
  KAT_EQ(s.vertices.size(), 0u);
  KAT_EQ(s.triangles.size(), 0u);
  KAT_EQ(s.indices.size(), 0u);
  KAT_EQ(c.getVertices().size(), 0u);
  KAT_EQ(c.getActiveVertices().size(), 0u);          // :301-330 storedValuesCompressed
  KAT_EQ(c.getStoredPolygons().size(), 0u);
  KAT_EQ(c.getActiveVerticesTimestamps().size(), 0u);
  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(2.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 101.0);
  KAT_EQ(s.vertices.size(), 0u);
  KAT_EQ(s.triangles.size(), 0u);
  KAT_EQ(s.indices.size(), 0u);
  KAT_EQ(c.getVertices().size(), 0u);                // :340-350
  KAT_EQ(c.getActiveVertices().size(), 0u);
  KAT_EQ(c.getActiveVerticesTimestamps().size(), 0u);
  KAT_EQ(c.getStoredPolygons().size(), 0u);
  // :18-38 constructor
  BlockCompression fresh(1.0);
  KAT_EQ(fresh.getVertices().size(), 0u);
  KAT_EQ(fresh.getActiveVertices().size(), 0u);
  KAT_EQ(fresh.getStoredPolygons().size(), 0u);
  KAT_EQ(fresh.getActiveVerticesTimestamps().size(), 0u);
}

void katOctreeFullCollapse() {
  OctreeCompression c(10.0);
  Scratch s;
  c.compressAndIntegrate(simpleMeshInterface(1.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 100.0);
  KAT_EQ(s.vertices.size(), 0u);
  KAT_EQ(s.triangles.size(), 0u);
  KAT_EQ(s.indices.size(), 0u);
  KAT_EQ(c.getVertices().size(), 0u);
  s = Scratch();
  c.compressAndIntegrate(simpleMeshInterface(2.0), s.vertices, s.triangles, s.indices,
                         s.remappings, 101.0);
  KAT_EQ(s.vertices.size(), 0u);
  KAT_EQ(s.triangles.size(), 0u);
  KAT_EQ(s.indices.size(), 0u);
  KAT_EQ(c.getVertices().size(), 0u);
  KAT_EQ(c.getActiveVertices().size(), 0u);
}

// test_mesh_frontend_interface.cpp:17-124 fixtures
FlatBlockMesh blockMesh(const std::vector<std::pair<BlockIndex, std::array<std::vector<float>, 3>>>& blocks) {
  FlatBlockMesh mesh;
  mesh.offsets.push_back(0);
  for (const auto& b : blocks) {
    mesh.blocks.push_back(b.first);
    for (size_t i = 0; i < b.second[0].size(); ++i) {
      PointXYZRGBA p;
      p.x = b.second[0][i];
      p.y = b.second[1][i];
      p.z = b.second[2][i];
      p.rgba = 0xFF7B16FAu;
      mesh.vertices.push_back(p);
    }
    mesh.offsets.push_back(mesh.vertices.size());
  }
  return mesh;
}
using F3 = std::array<std::vector<float>, 3>;
FlatBlockMesh simpleMesh1() {
  return blockMesh({{{0, 0, 0}, F3{{{0, 1, 0, 1, 1, 0}, {0, 0, 1, 0, 1, 1}, {0, 0, 0, 0, 0, 0}}}}});
}
FlatBlockMesh simpleMesh2() {
  return blockMesh({{{0, 0, 0}, F3{{{0, 1, 0, 1, 1, 0}, {0, 0, 1, 0, 1, 1}, {0, 0, 0, 0, 0, 0}}}},
                    {{1, 0, 0}, F3{{{2, 3, 2, 3, 3, 2}, {0, 0, 1, 0, 1, 1}, {0, 0, 0, 0, 0, 0}}}},
                    {{1, 1, 0}, F3{{{2, 3, 2, 3, 3, 2}, {2, 2, 3, 2, 3, 3}, {0, 0, 0, 0, 0, 0}}}}});
}
FlatBlockMesh simpleMesh3() {
  return blockMesh({{{0, 0, 0}, F3{{{0, 1, 0, 1, 1, 0}, {0, 0, 1, 0, 1, 1}, {0, 0, 0, 0, 0, 0}}}},
                    {{1, 0, 1}, F3{{{2, 3, 2, 3, 3, 2}, {0, 0, 1, 0, 1, 1}, {2, 2, 2, 2, 2, 2}}}}});
}
FlatBlockMesh simpleMesh4() {
  return blockMesh(
      {{{2, 2, 0}, F3{{{3.5, 4.5, 3.5, 4.5, 4.5, 3.5}, {3.5, 3.5, 4.5, 3.5, 4.5, 4.5}, {0, 0, 0, 0, 0, 0}}}},
       {{2, 2, 2}, F3{{{3.5, 4.5, 3.5, 4.5, 4.5, 3.5}, {3.5, 3.5, 4.5, 3.5, 4.5, 4.5}, {3.5, 4.5, 3.5, 4.5, 3.5, 3.5}}}}});
}
FlatBlockMesh simpleMesh5() {
  return blockMesh(
      {{{1, 0, 0}, F3{{{1.7f, 3.1f, 3.1f, 1.7f, 1.7f, 3.1f}, {0.0f, 0.0f, 1.5f, 0.0f, 1.5f, 1.5f}, {0, 0, 0, 0, 0, 0}}}},
       {{1, 1, 0}, F3{{{1.7f, 3.1f, 3.1f, 1.7f, 1.7f, 3.1f}, {1.7f, 1.7f, 3.1f, 1.7f, 3.1f, 3.1f}, {0, 0, 0, 0, 0, 0}}}}});
}

// MeshFrontendInterface::updateGraph / updateFullMesh minus the solver-side graph bookkeeping
struct FrontendLeg {
  OctreeCompression comp;
  HashedIndexMapping mapping;
  double horizon;
  FrontendLeg(double res, double h) : comp(res), horizon(h) {}
  void update(const FlatBlockMesh& mesh, double t) {
    comp.pruneStoredMesh(t - horizon);
    Scratch s;
    comp.compressAndIntegrate(mesh, s.vertices, s.triangles, s.indices, mapping, t);
  }
};

void katFrontendMeshes(double res) {
  // simplifiedMeshHighRes (graph leg, res 0.5) and fullMesh (full leg, res 0.1): same expectations
  FrontendLeg leg(res, 1.0);
  leg.update(simpleMesh1(), 10.0);
  {
    const auto& v = leg.comp.getVertices();
    const auto& t = leg.comp.getStoredPolygons();
    KAT_EQ(v.size(), 4u);
    KAT_EQ(t.size(), 2u);
    KAT_NEAR(1, v[1].x, 1e-3);
    KAT_NEAR(1, v[3].x, 1e-3);
    KAT_NEAR(1, v[3].y, 1e-3);
    KAT_EQ(t[0][2], 2u);
    KAT_EQ(t[1][0], 1u);
  }
  leg.update(simpleMesh2(), 10.5);
  {
    const auto& v = leg.comp.getVertices();
    const auto& t = leg.comp.getStoredPolygons();
    KAT_EQ(v.size(), 12u);
    KAT_EQ(t.size(), 6u);
    KAT_NEAR(3, v[5].x, 1e-3);
    KAT_NEAR(3, v[7].x, 1e-3);
    KAT_NEAR(1, v[7].y, 1e-3);
    KAT_EQ(t[2][2], 6u);
    KAT_EQ(t[3][0], 5u);
    KAT_NEAR(2, v[9].y, 1e-3);
    KAT_NEAR(3, v[11].x, 1e-3);
    KAT_NEAR(3, v[11].y, 1e-3);
    KAT_EQ(t[4][2], 10u);
  }
  leg.update(simpleMesh3(), 11.6);
  {
    const auto& v = leg.comp.getVertices();
    const auto& t = leg.comp.getStoredPolygons();
    KAT_EQ(v.size(), 20u);
    KAT_EQ(t.size(), 10u);
    KAT_NEAR(1, v[13].x, 1e-3);
    KAT_NEAR(1, v[15].x, 1e-3);
    KAT_NEAR(1, v[15].y, 1e-3);
    KAT_EQ(t[6][2], 14u);
    KAT_EQ(t[7][0], 13u);
    KAT_NEAR(3, v[17].x, 1e-3);
    KAT_NEAR(2, v[19].z, 1e-3);
    KAT_NEAR(1, v[19].y, 1e-3);
    KAT_EQ(t[8][2], 18u);
    KAT_EQ(t[9][0], 17u);
  }
}

void katFrontendCollapse() {
  // simplifiedMeshLowRes (res 1.5) / compression1, compression2 (res 4.0)
  for (double res : {1.5, 4.0}) {
    FrontendLeg leg(res, 1.0);
    leg.update(simpleMesh1(), 10.0);
    KAT_EQ(leg.comp.getVertices().size(), 0u);
    leg.update(simpleMesh2(), 10.5);
    KAT_EQ(leg.comp.getVertices().size(), 0u);
    KAT_EQ(leg.comp.getStoredPolygons().size(), 0u);
    leg.update(simpleMesh3(), 11.6);
    KAT_EQ(leg.comp.getVertices().size(), 0u);
    KAT_EQ(leg.comp.getStoredPolygons().size(), 0u);
  }
  FrontendLeg leg(4.0, 1.0);
  leg.update(simpleMesh1(), 10.0);
  leg.update(simpleMesh4(), 10.5);
  KAT_EQ(leg.comp.getVertices().size(), 0u);
}

void katFrontendMeshGraphNodes() {
  // meshGraph (graph leg res 0.6): node counts 4 then 8 new (keys 4..11)
  FrontendLeg leg(0.6, 1.0);
  Scratch s;
  leg.comp.pruneStoredMesh(9.0);
  leg.comp.compressAndIntegrate(simpleMesh1(), s.vertices, s.triangles, s.indices, leg.mapping, 10.0);
  KAT_TRUE(s.indices == (V{0, 1, 2, 3}));
  KAT_NEAR(1, leg.comp.getVertices()[3].x, 1e-4);
  KAT_NEAR(1, leg.comp.getVertices()[3].y, 1e-4);
  leg.comp.pruneStoredMesh(9.5);
  leg.comp.compressAndIntegrate(simpleMesh4(), s.vertices, s.triangles, s.indices, leg.mapping, 10.5);
  KAT_TRUE(s.indices == (V{4, 5, 6, 7, 8, 9, 10, 11}));
  KAT_NEAR(3.5, leg.comp.getVertices()[4].x, 1e-4);
  KAT_NEAR(4.5, leg.comp.getVertices()[11].x, 1e-4);
  KAT_NEAR(4.5, leg.comp.getVertices()[11].y, 1e-4);
  KAT_NEAR(3.5, leg.comp.getVertices()[11].z, 1e-4);
}

void katIndexMapping() {
  const BlockIndex block1{0, 0, 0}, block2{1, 0, 0}, block3{1, 1, 0};
  {  // vxblxIndexMapping1: graph leg at d_graph_resolution 0.5
    FrontendLeg leg(0.5, 1.0);
    leg.update(simpleMesh1(), 10.0);
    leg.update(simpleMesh5(), 10.5);
    const auto& m = leg.mapping;
    KAT_EQ(m.size(), 3u);
    KAT_EQ(m.at(block1).size(), 6u);
    KAT_EQ(m.at(block2).size(), 6u);
    KAT_EQ(m.at(block3).size(), 6u);
    const size_t e1[6] = {0, 1, 2, 1, 3, 2};
    const size_t e2[6] = {4, 5, 6, 4, 7, 6};
    const size_t e3[6] = {7, 6, 8, 7, 9, 8};
    for (size_t i = 0; i < 6; ++i) {
      KAT_EQ(m.at(block1).at(i), e1[i]);
      KAT_EQ(m.at(block2).at(i), e2[i]);
      KAT_EQ(m.at(block3).at(i), e3[i]);
    }
  }
  {  // vxblxIndexMapping2: graph leg at d_graph_resolution 1.3
    FrontendLeg leg(1.3, 1.0);
    leg.update(simpleMesh1(), 10.0);
    leg.update(simpleMesh5(), 10.5);
    const auto& m = leg.mapping;
    KAT_EQ(m.size(), 3u);
    KAT_EQ(m.at(block1).size(), 0u);
    KAT_EQ(m.at(block2).size(), 6u);
    KAT_EQ(m.at(block3).size(), 6u);
    const size_t e2[6] = {0, 1, 2, 0, 3, 2};
    const size_t e3[6] = {3, 2, 4, 3, 5, 4};
    for (size_t i = 0; i < 6; ++i) {
      KAT_EQ(m.at(block2).at(i), e2[i]);
      KAT_EQ(m.at(block3).at(i), e3[i]);
    }
  }
}

void katSurfaceExists() {
  // test_common_functions.cpp:177-198 meshSurfaceExist
  const Face poly_1{0, 1, 2}, poly_2{0, 2, 3}, poly_3{0, 5, 1}, poly_4{1, 4, 2};
  std::vector<Face> surfaces = {poly_1, poly_2, poly_3, poly_4};
  std::map<size_t, std::vector<size_t>> adj;
  adj[0] = {0, 1};
  adj[1] = {0, 3};
  adj[2] = {0, 1, 3};
  adj[3] = {1};
  adj[4] = {3};
  KAT_TRUE(!SurfaceExists(poly_3, adj, surfaces));
  adj[0].push_back(2);
  adj[1].push_back(2);
  adj[5] = {2};
  KAT_TRUE(SurfaceExists(poly_3, adj, surfaces));
  // ordered-triple match only: a rotated triple is a different surface
  KAT_TRUE(!SurfaceExists({0, 1, 5}, adj, surfaces));
}

}  // namespace

int runCompressionKats() {
  katVoxelClearingBasic();
  katVoxelClearingPruningBasic();
  katVoxelClearingMultiblock();
  katVoxelClearingPruningMultiblock();
  katOctreeBoundingBox();
  katOctreeReturnedAndStored();
  katOctreePrune();
  katOctreeFullCollapse();
  katBlockCompression();
  katFrontendMeshes(0.5);
  katFrontendMeshes(0.1);
  katFrontendCollapse();
  katFrontendMeshGraphNodes();
  katIndexMapping();
  katSurfaceExists();
  return kat::failures();
}
