// Minimal stand-ins for the third-party types Kimera-PGMO's hot-path API speaks (Eigen, GTSAM,
// PCL, spatial_hash).  ONLY used when the real headers are not installed (this build image has
// none of them): they carry exactly the members the shim and the replayed reference unit tests
// touch, with the same names, so that the very same shim code compiles against the real
// libraries in an upstream build (see kimera_pgmo_b200/deps.h for the switch).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>
#include <vector>

// ------------------------------------------------------------------------------------ Eigen
namespace Eigen {
template <typename S>
struct Vec3Lite {
  S v[3] = {0, 0, 0};
  Vec3Lite() = default;
  Vec3Lite(S a, S b, S c) : v{a, b, c} {}
  S& x() { return v[0]; }
  S& y() { return v[1]; }
  S& z() { return v[2]; }
  const S& x() const { return v[0]; }
  const S& y() const { return v[1]; }
  const S& z() const { return v[2]; }
  S& operator()(int i) { return v[i]; }
  const S& operator()(int i) const { return v[i]; }
  S& operator[](int i) { return v[i]; }
  const S& operator[](int i) const { return v[i]; }
  template <typename T>
  Vec3Lite<T> cast() const {
    return Vec3Lite<T>(static_cast<T>(v[0]), static_cast<T>(v[1]), static_cast<T>(v[2]));
  }
  static Vec3Lite Zero() { return Vec3Lite(); }
  bool operator==(const Vec3Lite& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
using Vector3f = Vec3Lite<float>;
using Vector3d = Vec3Lite<double>;
using Vector3i = Vec3Lite<int>;
struct Matrix3d {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  double& operator()(int r, int c) { return m[3 * r + c]; }
  const double& operator()(int r, int c) const { return m[3 * r + c]; }
};
}  // namespace Eigen

// ------------------------------------------------------------------------------------ GTSAM
namespace gtsam {
using Key = std::uint64_t;
using Point3 = Eigen::Vector3d;

class Symbol {
 public:
  Symbol(unsigned char c, std::uint64_t j) : c_(c), j_(j) {}
  // gtsam::Symbol::key(): chr in the top byte (test/data/graph.dgrf: 'a' -> 6989586621679009792)
  Key key() const { return (static_cast<Key>(c_) << 56) | j_; }
  operator Key() const { return key(); }
  unsigned char chr() const { return c_; }
  std::uint64_t index() const { return j_; }

 private:
  unsigned char c_;
  std::uint64_t j_;
};

class Rot3 {
 public:
  Rot3() = default;
  explicit Rot3(const Eigen::Matrix3d& R) : R_(R) {}
  // gtsam::Rot3(w, x, y, z) = Eigen::Quaterniond(w, x, y, z).toRotationMatrix(), operation for
  // operation (what DeformationGraph::load builds from a NODE line)
  Rot3(double w, double x, double y, double z) {
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R_(0, 0) = 1.0 - (tyy + tzz);
    R_(0, 1) = txy - twz;
    R_(0, 2) = txz + twy;
    R_(1, 0) = txy + twz;
    R_(1, 1) = 1.0 - (txx + tzz);
    R_(1, 2) = tyz - twx;
    R_(2, 0) = txz - twy;
    R_(2, 1) = tyz + twx;
    R_(2, 2) = 1.0 - (txx + tyy);
  }
  static Rot3 Rz(double t) {
    Eigen::Matrix3d R;
    R(0, 0) = std::cos(t);
    R(0, 1) = -std::sin(t);
    R(1, 0) = std::sin(t);
    R(1, 1) = std::cos(t);
    return Rot3(R);
  }
  const Eigen::Matrix3d& matrix() const { return R_; }

 private:
  Eigen::Matrix3d R_;
};

class Pose3 {
 public:
  Pose3() = default;
  Pose3(const Rot3& R, const Point3& t) : R_(R), t_(t) {}
  const Rot3& rotation() const { return R_; }
  const Point3& translation() const { return t_; }

 private:
  Rot3 R_;
  Point3 t_;
};

class Values {
 public:
  void insert(Key k, const Pose3& p) {
    if (!map_.emplace(k, p).second) throw std::invalid_argument("key already exists");
  }
  void update(Key k, const Pose3& p) { map_.at(k) = p; }
  bool exists(Key k) const { return map_.find(k) != map_.end(); }
  template <typename T>
  const T& at(Key k) const {
    return map_.at(k);  // std::out_of_range like gtsam::ValuesKeyDoesNotExist : out_of_range
  }
  size_t size() const { return map_.size(); }

 private:
  std::map<Key, Pose3> map_;
};
}  // namespace gtsam

// -------------------------------------------------------------------------------------- PCL
namespace pcl {
struct PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f;
  PointXYZ() = default;
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct PointXYZRGBA {
  float x = 0.f, y = 0.f, z = 0.f;
  union {
    struct {
      std::uint8_t b, g, r, a;
    };
    std::uint32_t rgba;
  };
  PointXYZRGBA() : rgba(0xFF000000u) {}  // r = g = b = 0, a = 255
};

template <typename T>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<T>>;
  std::vector<T> points;
  std::uint32_t width = 0, height = 1;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(size_t n) {
    points.resize(n);
    width = static_cast<std::uint32_t>(n);
  }
  void clear() {
    points.clear();
    width = 0;
  }
  void push_back(const T& p) {
    points.push_back(p);
    width = static_cast<std::uint32_t>(points.size());
  }
  T& at(size_t i) { return points.at(i); }
  const T& at(size_t i) const { return points.at(i); }
  T& operator[](size_t i) { return points[i]; }
  const T& operator[](size_t i) const { return points[i]; }
  auto begin() { return points.begin(); }
  auto end() { return points.end(); }
  auto begin() const { return points.begin(); }
  auto end() const { return points.end(); }
};

struct Vertices {
  std::vector<std::uint32_t> vertices;
};

// the byte blob of the real pcl::PCLPointCloud2, kept typed here
struct PCLPointCloud2 {
  std::vector<PointXYZRGBA> data;
  std::uint32_t width = 0, height = 1;
};

struct PolygonMesh {
  PCLPointCloud2 cloud;
  std::vector<Vertices> polygons;
};

inline void fromPCLPointCloud2(const PCLPointCloud2& blob, PointCloud<PointXYZRGBA>& cloud) {
  cloud.points = blob.data;
  cloud.width = blob.width;
  cloud.height = blob.height;
}
inline void fromPCLPointCloud2(const PCLPointCloud2& blob, PointCloud<PointXYZ>& cloud) {
  cloud.clear();
  for (const auto& p : blob.data) cloud.push_back(PointXYZ(p.x, p.y, p.z));
}
inline void toPCLPointCloud2(const PointCloud<PointXYZRGBA>& cloud, PCLPointCloud2& blob) {
  blob.data = cloud.points;
  blob.width = static_cast<std::uint32_t>(cloud.size());
  blob.height = 1;
}
inline void toPCLPointCloud2(const PointCloud<PointXYZ>& cloud, PCLPointCloud2& blob) {
  blob.data.clear();
  for (const auto& p : cloud.points) {
    PointXYZRGBA q;
    q.x = p.x;
    q.y = p.y;
    q.z = p.z;
    blob.data.push_back(q);
  }
  blob.width = static_cast<std::uint32_t>(cloud.size());
  blob.height = 1;
}

namespace traits {
template <typename T>
struct has_color : std::false_type {};
template <>
struct has_color<PointXYZRGBA> : std::true_type {};
template <typename T>
constexpr bool has_color_v = has_color<T>::value;
}  // namespace traits
}  // namespace pcl

// ------------------------------------------------------------------------------ spatial_hash
namespace spatial_hash {
struct BlockIndex : Eigen::Vector3i {
  BlockIndex() = default;
  BlockIndex(int a, int b, int c) : Eigen::Vector3i(a, b, c) {}
  static BlockIndex Zero() { return BlockIndex(0, 0, 0); }
};
using BlockIndices = std::vector<BlockIndex>;
struct IndexHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  size_t operator()(const BlockIndex& i) const {
    return static_cast<unsigned int>(i.x() + i.y() * sl + i.z() * sl2);
  }
};
template <typename V>
using IndexHashMap = std::unordered_map<BlockIndex, V, IndexHash>;
}  // namespace spatial_hash
