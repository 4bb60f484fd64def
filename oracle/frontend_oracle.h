// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
//
// CPU restatement of the mesh -> graph index bookkeeping of kimera_pgmo::MeshFrontendInterface
// (kimera_pgmo/src/mesh_frontend_interface.cpp:100-115 update, :197-211
// updateMeshToGraphMappings; members msg_to_mesh_idx_ / msg_to_graph_idx_ / mesh_to_graph_idx_,
// include/kimera_pgmo/mesh_frontend_interface.h:118-120).
//
// What the reference does per MeshFrontendInterface::update:
//   * the full-mesh compressor and the graph compressor each write into their persistent
//     HashedIndexMapping (block -> {msg vertex index -> stored index}).  The per-block maps are
//     never cleared (`remapping.insert({block, IndexMapping()})` is a no-op for a known block:
//     mesh_compression.cpp:256, voxel_clearing_compression.cpp:124), so entries written by an
//     earlier, larger version of a block stay behind;
//   * updateMeshToGraphMappings then walks, for every block of the message in blockIndices()
//     order, ALL entries of msg_to_mesh_idx_[block] and does mesh_to_graph_idx_->insert(
//     {mesh index, graph index}) for those whose msg index also has a graph entry;
//     unordered_map::insert keeps the first pair of a key.
//
// PARITY UNPINNED: when several msg vertices of one block map to the same mesh vertex but to
// different graph vertices, the reference's winner is the first in the iteration order of a
// std::unordered_map<size_t, size_t>, which is implementation-defined (and, in libstdc++,
// depends on the rehash history).  No reference test observes mesh_to_graph_idx_.  This
// restatement — and the product — walk a block's entries in ascending msg index.  Every pair
// either implementation produces is a pair the reference produces for SOME iteration order.
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <tuple>
#include <vector>

namespace pgmo_oracle {

class MeshToGraphOracle {
 public:
  using Block = std::tuple<int32_t, int32_t, int32_t>;
  static constexpr int64_t kAbsent = -1;

  // One MeshFrontendInterface::update.  The compressors' effect on the two persistent
  // remappings is passed as a dense record per input vertex: kind 0 = no write, 1 =
  // `map[i] = index`, 2 = `map.insert({i, index})` (kind == nullptr: all 1).
  void update(uint32_t n_blocks, const int32_t* block_idx_xyz, const uint64_t* block_offsets,
              const int64_t* mesh_index, const uint8_t* mesh_kind, const int64_t* graph_index,
              const uint8_t* graph_kind) {
    std::vector<Block> order;
    for (uint32_t b = 0; b < n_blocks; ++b) {
      const Block key{block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
      order.push_back(key);
      std::vector<int64_t>& m = msg_to_mesh_[key];    // remapping.insert({block, {}})
      std::vector<int64_t>& g = msg_to_graph_[key];
      const uint64_t first = block_offsets[b], n = block_offsets[b + 1] - first;
      apply(m, n, mesh_index + first, mesh_kind ? mesh_kind + first : nullptr);
      apply(g, n, graph_index + first, graph_kind ? graph_kind + first : nullptr);
    }
    // updateMeshToGraphMappings(latest_blocks_), :197-211
    for (const Block& key : order) {
      const std::vector<int64_t>& m = msg_to_mesh_.at(key);
      const std::vector<int64_t>& g = msg_to_graph_.at(key);
      for (size_t i = 0; i < m.size(); ++i) {
        if (m[i] == kAbsent) continue;
        if (i < g.size() && g[i] != kAbsent) {
          mesh_to_graph_.insert({static_cast<uint64_t>(m[i]), static_cast<uint64_t>(g[i])});
        }
      }
    }
  }

  const std::map<uint64_t, uint64_t>& meshToGraph() const { return mesh_to_graph_; }
  // dense view of msg_to_mesh_idx_[block] (which == 0) / msg_to_graph_idx_[block] (which == 1);
  // nullptr for an unknown block
  const std::vector<int64_t>* blockMapping(const Block& key, int which) const {
    const auto& maps = which == 0 ? msg_to_mesh_ : msg_to_graph_;
    const auto it = maps.find(key);
    return it == maps.end() ? nullptr : &it->second;
  }

 private:
  static void apply(std::vector<int64_t>& map, uint64_t n, const int64_t* index, const uint8_t* kind) {
    if (map.size() < n) map.resize(n, kAbsent);
    for (uint64_t i = 0; i < n; ++i) {
      const uint8_t k = kind ? kind[i] : 1;
      if (k == 1 || (k == 2 && map[i] == kAbsent)) map[i] = index[i];
    }
  }

  // dense stand-ins for the unordered_map<size_t, size_t>s: index = msg vertex index
  std::map<Block, std::vector<int64_t>> msg_to_mesh_, msg_to_graph_;
  std::map<uint64_t, uint64_t> mesh_to_graph_;
};

}  // namespace pgmo_oracle
