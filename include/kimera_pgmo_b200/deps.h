// Dependency switch of the C++ shim.
//
// Upstream build (PCL + GTSAM + spatial_hash + Kimera-PGMO headers installed): the shim uses
// the real types and Kimera-PGMO's own ADL traits layer (mesh_traits.h:153-215,
// pcl_mesh_traits.h:19-143), so any cloud type the reference accepts is accepted unchanged.
//
// This image has none of them: deps_lite.h provides stand-ins with the same member names and
// the traits layer is re-stated below with the same customization points (pgmoNumVertices,
// pgmoGetVertex, pgmoSetVertex, pgmoGetVertexStamp) and the same has_get_stamp dispatch.
#pragma once

#if defined(__has_include)
#if __has_include(<gtsam/nonlinear/Values.h>) && __has_include(<pcl/point_cloud.h>) && \
    __has_include("kimera_pgmo/mesh_traits.h") && !defined(KIMERA_PGMO_B200_FORCE_STANDIN)
#define KIMERA_PGMO_B200_REAL_DEPS 1
#endif
#endif

#ifdef KIMERA_PGMO_B200_REAL_DEPS

#include <gtsam/geometry/Point3.h>
#include <gtsam/geometry/Pose3.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/nonlinear/Values.h>
#include <pcl/PolygonMesh.h>
#include <pcl/conversions.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "kimera_pgmo/hashing.h"
#include "kimera_pgmo/mesh_traits.h"
#include "kimera_pgmo/pcl_mesh_traits.h"
#include "kimera_pgmo/utils/common_functions.h"

namespace kimera_pgmo_b200 {
namespace traits = ::kimera_pgmo::traits;
using ::kimera_pgmo::BlockIndex;
using ::kimera_pgmo::BlockIndices;
using ::kimera_pgmo::ConstStampedCloud;
using ::kimera_pgmo::HashedIndexMapping;
using ::kimera_pgmo::IndexMapping;
using ::kimera_pgmo::StampedCloud;
using Timestamp = ::kimera_pgmo::traits::Timestamp;
using ::kimera_pgmo::GetRobotPrefix;
using ::kimera_pgmo::GetVertexPrefix;
using ::kimera_pgmo::robot_id_to_prefix;
using ::kimera_pgmo::robot_id_to_vertex_prefix;
using ::kimera_pgmo::robot_prefix_to_id;
using ::kimera_pgmo::vertex_prefix_to_id;
}  // namespace kimera_pgmo_b200

#else  // stand-ins

#include <optional>
#include <type_traits>

#include "kimera_pgmo_b200/standin/deps_lite.h"

namespace kimera_pgmo_b200 {

using Timestamp = std::uint64_t;
using IndexMapping = std::unordered_map<size_t, size_t>;
using spatial_hash::BlockIndex;
using spatial_hash::BlockIndices;
using HashedIndexMapping = spatial_hash::IndexHashMap<IndexMapping>;

// Robot id <-> gtsam::Symbol prefix (kimera_pgmo/utils/common_functions.h:19-81): robots 0..7 own
// the pose prefixes 'a'..'h' and the mesh-vertex prefixes 's'..'z'; anything else has no prefix.
constexpr size_t kMaxRobots = 8;
inline char GetRobotPrefix(size_t robot_id) {
  return robot_id < kMaxRobots ? static_cast<char>('a' + robot_id) : '\0';
}
inline char GetVertexPrefix(size_t robot_id) {
  return robot_id < kMaxRobots ? static_cast<char>('s' + robot_id) : '\0';
}
namespace detail {
inline std::map<size_t, char> idToPrefixMap(char first) {
  std::map<size_t, char> m;
  for (size_t id = 0; id < kMaxRobots; ++id) m[id] = static_cast<char>(first + id);
  return m;
}
inline std::map<char, size_t> prefixToIdMap(char first) {
  std::map<char, size_t> m;
  for (size_t id = 0; id < kMaxRobots; ++id) m[static_cast<char>(first + id)] = id;
  return m;
}
}  // namespace detail
// the reference's four lookup tables, same names (callers use .count() / .at())
inline const std::map<size_t, char> robot_id_to_prefix = detail::idToPrefixMap('a');
inline const std::map<char, size_t> robot_prefix_to_id = detail::prefixToIdMap('a');
inline const std::map<size_t, char> robot_id_to_vertex_prefix = detail::idToPrefixMap('s');
inline const std::map<char, size_t> vertex_prefix_to_id = detail::prefixToIdMap('s');

namespace traits {
using Color = std::array<std::uint8_t, 4>;
using Face = std::array<size_t, 3>;
using Pos = Eigen::Vector3f;
using Timestamp = std::uint64_t;
using Label = std::uint32_t;
struct VertexTraits {
  VertexTraits() = default;
  virtual ~VertexTraits() = default;
  std::optional<Color> color;
  std::optional<Timestamp> stamp;
  std::optional<Label> label;
};
}  // namespace traits

}  // namespace kimera_pgmo_b200

// ADL customization points for pcl clouds (pcl_mesh_traits.h:19-77)
namespace pcl {
using PgmoPosB200 = kimera_pgmo_b200::traits::Pos;
using PgmoTraitsB200 = kimera_pgmo_b200::traits::VertexTraits;
template <typename T>
size_t pgmoNumVertices(const PointCloud<T>& cloud) {
  return cloud.size();
}
template <typename T>
void pgmoResizeVertices(PointCloud<T>& cloud, size_t size) {
  cloud.resize(size);
}
template <typename T>
PgmoPosB200 pgmoGetVertex(const PointCloud<T>& cloud, size_t i, PgmoTraitsB200* t) {
  const auto& p = cloud.at(i);
  if constexpr (traits::has_color_v<T>) {
    if (t) t->color = {{p.r, p.g, p.b, p.a}};
  }
  (void)t;
  return {p.x, p.y, p.z};
}
template <typename T>
void pgmoSetVertex(PointCloud<T>& cloud, size_t i, const PgmoPosB200& pos, const PgmoTraitsB200& t) {
  auto& p = cloud.at(i);
  p.x = pos.x();
  p.y = pos.y();
  p.z = pos.z();
  if constexpr (traits::has_color_v<T>) {
    if (t.color) {  // colour untouched unless traits carry one (pcl_mesh_traits.h:59-77)
      const auto& c = *t.color;
      p.r = c[0];
      p.g = c[1];
      p.b = c[2];
      p.a = c[3];
    }
  }
  (void)t;
}
}  // namespace pcl

namespace kimera_pgmo_b200 {

// pcl_mesh_traits.h:81-143
template <typename T>
struct StampedCloud {
  StampedCloud(pcl::PointCloud<T>& points, std::vector<std::uint64_t>& stamps)
      : points(points), stamps(stamps) {}
  pcl::PointCloud<T>& points;
  std::vector<std::uint64_t>& stamps;
};
template <typename T>
struct ConstStampedCloud {
  ConstStampedCloud(const pcl::PointCloud<T>& points, const std::vector<std::uint64_t>& stamps)
      : points(points), stamps(stamps) {}
  const pcl::PointCloud<T>& points;
  const std::vector<std::uint64_t>& stamps;
};
template <typename T>
size_t pgmoNumVertices(const StampedCloud<T>& cloud) {
  return cloud.points.size();
}
template <typename T>
size_t pgmoNumVertices(const ConstStampedCloud<T>& cloud) {
  return cloud.points.size();
}
template <typename T>
traits::Pos pgmoGetVertex(const ConstStampedCloud<T>& cloud, size_t i, traits::VertexTraits* t) {
  if (t && i < cloud.stamps.size()) t->stamp = cloud.stamps[i];
  return pgmoGetVertex(cloud.points, i, t);
}
template <typename T>
void pgmoSetVertex(StampedCloud<T>& cloud, size_t i, const traits::Pos& pos,
                   const traits::VertexTraits& t) {
  pgmoSetVertex(cloud.points, i, pos, t);
  if (t.stamp) cloud.stamps.at(i) = *t.stamp;
}
template <typename T>
std::uint64_t pgmoGetVertexStamp(const ConstStampedCloud<T>& cloud, size_t i) {
  return cloud.stamps.at(i);
}

namespace traits {
// mesh_traits.h:153-215 (ADL dispatch; the reference routes it through function objects, the
// observable behaviour is the same)
namespace detail {
struct vertex_stamp_fn {
  template <typename T>
  constexpr auto operator()(const T& v, size_t i) const -> decltype(pgmoGetVertexStamp(v, i)) {
    return pgmoGetVertexStamp(v, i);
  }
};
}  // namespace detail
template <typename T>
size_t num_vertices(const T& vertices) {
  return pgmoNumVertices(vertices);
}
template <typename T>
Pos get_vertex(const T& vertices, size_t i, VertexTraits* t = nullptr) {
  return pgmoGetVertex(vertices, i, t);
}
template <typename T>
void set_vertex(T& vertices, size_t i, const Pos& pos, const VertexTraits& t = {}) {
  pgmoSetVertex(vertices, i, pos, t);
}
template <typename T>
using has_get_stamp = std::is_invocable<detail::vertex_stamp_fn, const T&, size_t>;
template <typename T, std::enable_if_t<!has_get_stamp<T>::value, bool> = true>
Timestamp get_timestamp(const T&, size_t) {
  return 0;
}
template <typename T, std::enable_if_t<has_get_stamp<T>::value, bool> = true>
Timestamp get_timestamp(const T& vertices, size_t i) {
  return pgmoGetVertexStamp(vertices, i);
}
}  // namespace traits

}  // namespace kimera_pgmo_b200

#endif  // KIMERA_PGMO_B200_REAL_DEPS


namespace kimera_pgmo_b200 {
// robot id of a pose / vertex prefix, -1 when the character is neither
inline int GetRobotIdFromPrefix(char prefix) {
  const auto it = robot_prefix_to_id.find(prefix);
  return it == robot_prefix_to_id.end() ? -1 : static_cast<int>(it->second);
}
inline int GetRobotIdFromVertexPrefix(char prefix) {
  const auto it = vertex_prefix_to_id.find(prefix);
  return it == vertex_prefix_to_id.end() ? -1 : static_cast<int>(it->second);
}
}  // namespace kimera_pgmo_b200
