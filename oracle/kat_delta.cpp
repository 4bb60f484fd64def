// ORACLE — TEST INFRASTRUCTURE ONLY.
// Replays kimera_pgmo/test/test_delta_compression.cpp against the restatement: the VertexInfo
// reference-count test (:437-470) and every parametrised CompressionCorrect configuration
// (:283-435, checked the way ExpectedDelta::checkOutput / checkTriangles do, :189-270).
#include <array>
#include <optional>
#include <string>

#include "delta_oracle.h"
#include "kat_check.h"

using namespace pgmo_oracle;

namespace {

using FaceCoords = std::array<std::array<float, 3>, 3>;
struct BlockConfig {
  BlockIndex index;
  std::vector<FaceCoords> faces;
};

// pgmo_fixtures.cpp createMesh with BlockConfig::point_index running across the calls of one
// configuration (resetIndex() at its start): colour = absolute point index
FlatBlockMesh createMesh(const std::vector<BlockConfig>& configs, uint8_t& point_index) {
  FlatBlockMesh mesh;
  mesh.offsets.push_back(0);
  for (const auto& c : configs) {
    mesh.blocks.push_back(c.index);
    for (const auto& f : c.faces) {
      for (const auto& v : f) {
        PointXYZRGBA p;
        p.x = v[0];
        p.y = v[1];
        p.z = v[2];
        p.rgba = 0xFF000000u | (uint32_t(point_index) << 16) | (uint32_t(point_index) << 8) | point_index;
        ++point_index;
        mesh.vertices.push_back(p);
      }
    }
    mesh.offsets.push_back(mesh.vertices.size());
  }
  return mesh;
}

struct Tri {
  size_t a, b, c;
};
bool sameFace(const Tri& l, const Tri& r) {
  return std::set<size_t>{l.a, l.b, l.c} == std::set<size_t>{r.a, r.b, r.c};
}

struct Input {
  std::optional<uint64_t> prune_s;
  uint64_t stamp_s;
  std::vector<BlockConfig> blocks;
};
struct Expected {
  size_t vertex_start, face_start, num_vertices;
  std::vector<Tri> triangles;
  std::vector<size_t> indices;
};
struct Config {
  std::string name;
  std::vector<std::pair<Input, Expected>> inputs;
};

const BlockConfig block1_empty{{0, 0, 0}, {}};
const BlockConfig block1_v1{{0, 0, 0},
                            {{{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                             {{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                             {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};
const BlockConfig block1_v2{{0, 0, 0},
                            {{{{0.5, 0.5, 0.5}, {0.5, 0.75, 0.75}, {0.5, 0.75, 0.5}}},
                             {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};
const BlockConfig block2_empty{{-1, 0, 0}, {}};
const BlockConfig block2_v1{{-1, 0, 0},
                            {{{{-0.5, 0.5, 0.5}, {-0.5, 0.75, 0.75}, {-0.5, 0.75, 0.5}}},
                             {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, 0.5}}}}};
const BlockConfig block2_v2{{-1, 0, 0},
                            {{{{-0.5, 0.5, 0.5}, {-0.5, 0.75, 0.75}, {-0.5, 0.75, 0.5}}},
                             {{{0.0, 0.0, 0.0}, {0.0, 0.5, 0.5}, {0.0, 0.0, -0.5}}}}};

const std::nullopt_t none = std::nullopt;

// test_delta_compression.cpp:334-435, transcribed
std::vector<Config> configurations() {
  return {
      {"SingleBlockClearing",
       {{{none, 100, {block1_v1}}, {0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {0, 1, 2, 0, 1, 2, 6, 7, 8}}},
        {{none, 101, {block1_empty}}, {0, 0, 0, {}, {}}}}},
      {"SingleBlockPrune",
       {{{none, 100, {block1_v1}}, {0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {0, 1, 2, 0, 1, 2, 6, 7, 8}}},
        {{101, 102, {block1_empty}}, {0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {}}},
        {{none, 103, {block1_v1}},
         {6, 2, 6, {{9, 10, 11}, {15, 16, 17}}, {9, 10, 11, 9, 10, 11, 15, 16, 17}}}}},
      {"MultiBlockClearing",
       {{{none, 100, {block1_v1, block2_v1}},
         {0, 0, 9, {{0, 1, 2}, {6, 7, 8}, {9, 10, 11}}, {0, 1, 2, 0, 1, 2, 6, 7, 8, 9, 10, 11, 6, 7, 8}}},
        {{none, 101, {block1_empty, block2_v1}},
         {0, 0, 6, {{15, 16, 17}, {18, 19, 20}}, {15, 16, 17, 18, 19, 20}}},
        {{none, 102, {block1_v1, block2_v1}},
         {0, 0, 9, {{21, 22, 23}, {27, 28, 29}, {30, 31, 32}},
          {21, 22, 23, 21, 22, 23, 27, 28, 29, 30, 31, 32, 27, 28, 29}}},
        {{none, 103, {block1_v1, block2_empty}},
         {0, 0, 6, {{36, 37, 38}, {42, 43, 44}}, {36, 37, 38, 36, 37, 38, 42, 43, 44}}}}},
      {"MultiBlockPrune",
       {{{none, 100, {block1_v1}}, {0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {0, 1, 2, 0, 1, 2, 6, 7, 8}}},
        {{none, 102, {block2_v1}},
         {0, 0, 9, {{0, 1, 2}, {9, 10, 11}, {12, 13, 14}}, {9, 10, 11, 12, 13, 14}}},
        {{101, 103, {block1_empty, block2_empty}}, {0, 0, 6, {{0, 1, 2}, {12, 13, 14}}, {}}},
        {{none, 104, {block1_v1}},
         {6, 2, 6, {{15, 16, 17}, {21, 22, 23}}, {15, 16, 17, 15, 16, 17, 21, 22, 23}}},
        {{none, 105, {block1_v1}},
         {6, 2, 6, {{24, 25, 26}, {30, 31, 32}}, {24, 25, 26, 24, 25, 26, 30, 31, 32}}}}},
      {"MultiBlockPartialUpdates",
       {{{none, 100, {block1_v1}}, {0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {0, 1, 2, 0, 1, 2, 6, 7, 8}}},
        {{none, 102, {block2_v1}},
         {0, 0, 9, {{0, 1, 2}, {9, 10, 11}, {12, 13, 14}}, {9, 10, 11, 12, 13, 14}}},
        {{101, 103, {block1_empty, block2_v1}},
         {0, 0, 9, {{0, 1, 2}, {15, 16, 17}, {18, 19, 20}, {18, 19, 20}}, {15, 16, 17, 18, 19, 20}}},
        {{none, 104, {block2_v1}},
         {3, 1, 6, {{21, 22, 23}, {24, 25, 26}, {24, 25, 26}}, {21, 22, 23, 24, 25, 26}}},
        {{none, 105, {block1_v1, block2_v1}},
         {3, 1, 9, {{27, 28, 29}, {33, 34, 35}, {36, 37, 38}, {33, 34, 35}},
          {27, 28, 29, 27, 28, 29, 33, 34, 35, 36, 37, 38, 33, 34, 35}}},
        {{none, 106, {block1_empty, block2_empty}}, {3, 1, 3, {{33, 34, 35}}, {}}}}},
      {"MultiBlockPartialArchive",
       {{{none, 100, {block1_v2}}, {0, 0, 6, {{0, 1, 2}, {3, 4, 5}}, {0, 1, 2, 3, 4, 5}}},
        {{none, 102, {block2_v2}},
         {0, 0, 10, {{0, 1, 2}, {9, 10, 5}, {6, 7, 8}, {9, 10, 11}}, {6, 7, 8, 9, 10, 11}}},
        {{101, 103, {block1_empty, block2_v2}},
         {0, 0, 10, {{0, 1, 2}, {12, 13, 14}, {15, 16, 17}, {15, 16, 5}}, {12, 13, 14, 15, 16, 17}}},
        {{none, 104, {block2_empty}}, {4, 1, 2, {{15, 16, 5}}, {}}}}},
      {"RepeatedTimestamp",
       {{{none, 100, {block1_v1}}, {0, 0, 6, {{0, 1, 2}, {6, 7, 8}}, {0, 1, 2, 0, 1, 2, 6, 7, 8}}},
        {{none, 100, {block2_v1}},
         {0, 0, 9, {{0, 1, 2}, {9, 10, 11}, {12, 13, 14}}, {9, 10, 11, 12, 13, 14}}}}},
  };
}

constexpr uint64_t kSec = 1000000000ull;

void runConfig(const Config& config) {
  DeltaCompression compression(1.0e-3);
  uint8_t point_index = 0;  // BlockConfig::resetIndex()
  std::map<size_t, size_t> prev_remapping;
  for (const auto& [input, expected] : config.inputs) {
    if (input.prune_s) compression.archiveBlocksByTime(*input.prune_s * kSec);
    const FlatBlockMesh mesh = createMesh(input.blocks, point_index);
    std::vector<uint64_t> output_indices;
    const auto output = compression.update(mesh, input.stamp_s * kSec, &output_indices);
    KAT_TRUE(output != nullptr);
    // ExpectedDelta::checkOutput (:189-218)
    KAT_EQ(output->vertex_updates.size(), output->stamp_updates.size());
    KAT_EQ(output->vertex_start, expected.vertex_start);
    KAT_EQ(output->face_start, expected.face_start);
    KAT_EQ(output->vertex_updates.size(), expected.num_vertices);
    std::map<size_t, size_t> result_remapping;
    for (size_t i = 0; i < output->vertex_start; ++i) {
      if (prev_remapping.count(i)) result_remapping[i] = prev_remapping.at(i);
    }
    for (size_t i = 0; i < output->vertex_updates.size(); ++i) {
      result_remapping[i + output->vertex_start] = (output->vertex_updates[i].rgba >> 16) & 0xFFu;  // .r
    }
    prev_remapping = result_remapping;
    std::vector<DeltaFace> all_faces(output->face_archive_updates);
    all_faces.insert(all_faces.end(), output->face_updates.begin(), output->face_updates.end());
    // checkTriangles (:220-270)
    KAT_EQ(all_faces.size(), expected.triangles.size());
    std::vector<Tri> absolute;
    for (const auto& f : all_faces) {
      const bool known = result_remapping.count(f.v1) && result_remapping.count(f.v2) && result_remapping.count(f.v3);
      KAT_TRUE(known);
      if (!known) continue;
      const Tri abs_face{result_remapping.at(f.v1), result_remapping.at(f.v2), result_remapping.at(f.v3)};
      absolute.push_back(abs_face);
      bool found = false;
      for (const auto& e : expected.triangles) found = found || sameFace(e, abs_face);
      KAT_TRUE(found);
    }
    for (const auto& e : expected.triangles) {
      bool found = false;
      for (const auto& a : absolute) found = found || sameFace(e, a);
      KAT_TRUE(found);
    }
    std::vector<size_t> abs_indices;
    bool mapped = true;
    for (const auto idx : output_indices) {
      if (!result_remapping.count(idx)) {
        mapped = false;
        break;
      }
      abs_indices.push_back(result_remapping.at(idx));
    }
    KAT_TRUE(mapped);
    KAT_TRUE(abs_indices == expected.indices);
    if (kat::failures()) std::cerr << "  (in configuration " << config.name << ")\n";
  }
}

}  // namespace

int runDeltaKats() {
  {  // vertexInfoCorrect (:437-470)
    VertexInfo info;
    KAT_TRUE(info.notObserved());
    KAT_TRUE(!info.shouldArchive());
    ++info.active_refs;  // addObservation
    KAT_TRUE(!info.notObserved());
    KAT_TRUE(!info.shouldArchive());
    --info.active_refs;  // removeObservation
    KAT_TRUE(info.notObserved());
    KAT_TRUE(!info.shouldArchive());
    ++info.active_refs;
    info.needs_archive = true;  // archiveObservation
    --info.active_refs;
    KAT_TRUE(!info.notObserved());
    KAT_TRUE(info.shouldArchive());
    ++info.active_refs;
    ++info.active_refs;
    info.needs_archive = true;
    --info.active_refs;
    KAT_TRUE(!info.notObserved());
    KAT_TRUE(!info.shouldArchive());
    --info.active_refs;
    KAT_TRUE(!info.notObserved());
    KAT_TRUE(info.shouldArchive());
  }
  for (const auto& config : configurations()) runConfig(config);
  {
    // kimera_pgmo/test/test_mesh_delta.cpp:39-43, :80-103 (constructor, archiveVerticesCorrect,
    // archiveFacesCorrect; updateSimple is replayed through the shim's MeshDelta, tests/cpp/test_shim.cpp)
    MeshDelta fresh(0, 0);
    KAT_EQ(fresh.vertex_start, 0u);
    KAT_EQ(fresh.face_start, 0u);
    MeshDelta delta(1, 3);
    PointXYZRGBA p;
    p.x = 1.0f;
    p.y = 2.0f;
    p.z = 3.0f;
    KAT_EQ(delta.addVertex(100, p, MeshInterface::kNoLabel, true), 1u);
    KAT_EQ(delta.addVertex(100, p, MeshInterface::kNoLabel, true), 2u);
    KAT_EQ(delta.addVertex(100, p, MeshInterface::kNoLabel, false), 3u);
    KAT_EQ(delta.addVertex(100, p, MeshInterface::kNoLabel, false), 4u);
    KAT_EQ(delta.getTotalArchivedVertices(), 3u);
    KAT_EQ(delta.num_archived_vertices, 2u);
    KAT_TRUE(!delta.hasSemantics());
    MeshDelta faces(1, 3);
    faces.addFace({0, 1, 2}, true);
    faces.addFace({1, 2, 3}, true);
    faces.addFace({0, 1, 2}, false);
    faces.addFace({1, 2, 3}, false);
    KAT_EQ(faces.getTotalArchivedFaces(), 5u);
    KAT_EQ(faces.face_archive_updates.size(), 2u);
  }
  return kat::failures();
}
