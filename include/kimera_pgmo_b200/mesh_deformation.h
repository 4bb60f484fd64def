// Drop-in for kimera_pgmo/include/kimera_pgmo/mesh_deformation.h: the same
// deformation::deformPoints overloads (unstamped :63-75, stamped :129-141), same argument
// meaning and error behaviour, executed by libpgmo_b200.so on a B200.
//
// What the shim does per call (host side, thin):
//   1. gathers the vertices through the ADL traits into SoA float buffers (+ u64 stamps),
//   2. flattens `values` into poses12[n_c][12] (row-major R then t) and valid[n_c] by probing
//      gtsam::Symbol(prefix, j) for j < control_points.size(),
//   3. one C-ABI call (pgmo_deform_points),
//   4. scatters the processed vertices back with traits::set_vertex (colour untouched) and
//      rebuilds std::vector<std::set<size_t>> from the dense association output.
// The staging buffers of 1. and 4. are page-locked and live in the Backend (PCIe-speed copies
// that overlap with the kernels, no page faults per call); the gather / scatter loops run on a
// few threads for large clouds.  Callers that do not read control_point_map (deformMesh, and
// DeformationGraph::deformPoints without vertex_graph_map) go through deformPointsOnly: no
// association output crosses PCIe and no std::set is built — for a 10 M-vertex mesh that is
// the difference between ~50 M heap allocations and none.
#pragma once

#include <algorithm>
#ifdef KIMERA_PGMO_B200_TRACE
#include <chrono>
#include <cstdio>
#endif
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

// upper bound on the threads of the shim's gather / scatter loops (1 = single-threaded)
#ifndef KIMERA_PGMO_B200_SHIM_THREADS
#define KIMERA_PGMO_B200_SHIM_THREADS 16
#endif

#include "kimera_pgmo_b200/deps.h"
#include "pgmo_b200.h"

namespace kimera_pgmo_b200 {

// common_structs.cpp:21-24
inline Timestamp stampFromSec(double sec) { return pgmo_stamp_from_sec(sec); }
// utils/common_structs.cpp:26-29 (host only; nothing on the device path needs it)
inline double stampToSec(Timestamp stamp) {
  return std::chrono::duration<double>(std::chrono::nanoseconds(stamp)).count();
}

class PgmoB200Error : public std::runtime_error {
 public:
  PgmoB200Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

namespace detail {
// -DKIMERA_PGMO_B200_TRACE: phase timings of the shim's host side on stderr (development aid)
#ifdef KIMERA_PGMO_B200_TRACE
struct Trace {
  explicit Trace(const char* what) : what_(what), t0_(std::chrono::steady_clock::now()), last_(t0_) {}
  void mark(const char* phase) {
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[shim trace] %s / %s: %.2f ms\n", what_, phase,
                 std::chrono::duration<double, std::milli>(now - last_).count());
    last_ = now;
  }
  const char* what_;
  std::chrono::steady_clock::time_point t0_, last_;
};
#else
struct Trace {
  explicit Trace(const char*) {}
  void mark(const char*) {}
};
#endif
inline void throwOnError(int rc) {
  if (rc >= 0) return;
  const std::string msg = std::string("pgmo_b200: ") + pgmo_last_error();
  if (rc == PGMO_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);  // what .at() would have thrown
  throw PgmoB200Error(rc, msg);
}

// fn(begin, end) over [0, n): on up to KIMERA_PGMO_B200_SHIM_THREADS threads when the range is
// large, in place otherwise.  The first exception of any part (e.g. the std::out_of_range of a
// cloud's .at()) is rethrown on the calling thread.
template <typename Fn>
void parallelFor(size_t n, Fn&& fn, size_t kMinPerThread = size_t(1) << 17) {
  const size_t hw = std::thread::hardware_concurrency();
  const size_t n_threads = std::min<size_t>(
      {hw ? hw : 1, static_cast<size_t>(KIMERA_PGMO_B200_SHIM_THREADS), n / kMinPerThread});
  if (n_threads <= 1) {
    fn(size_t(0), n);
    return;
  }
  std::exception_ptr error;
  std::mutex error_mutex;
  auto part = [&](size_t t) {
    try {
      fn(n * t / n_threads, n * (t + 1) / n_threads);
    } catch (...) {
      std::lock_guard<std::mutex> lock(error_mutex);
      if (!error) error = std::current_exception();
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(n_threads - 1);
  for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(part, t);
  part(0);
  for (std::thread& th : pool) th.join();
  if (error) std::rethrow_exception(error);
}

// Touches every page of freshly allocated (reserved, not yet sized) vector storage on the shim's
// threads, so that the single-threaded value-initialisation of the resize() that follows runs
// over mapped memory: first-touch page faults are most of the cost of creating a 320 MB output
// blob (deformMesh returns a new mesh by value, deformation_graph.cpp:532).
template <typename T>
void resizePrefaulted(std::vector<T>& v, size_t n) {
  if (n <= v.capacity() || n * sizeof(T) < (size_t(1) << 24)) {
    v.resize(n);
    return;
  }
  v.reserve(n);
  // raw bytes of the capacity beyond size(): no element lives there yet
  volatile unsigned char* raw = reinterpret_cast<volatile unsigned char*>(v.data());
  const size_t begin_byte = v.size() * sizeof(T), bytes = n * sizeof(T);
  parallelFor((bytes - begin_byte + 4095) / 4096, [&](size_t b, size_t e) {
    for (size_t page = b; page < e; ++page) raw[begin_byte + page * 4096] = 0;
  }, 2048);
  v.resize(n);
}

// Grow-only host block carved into the per-call arrays: page-locked (pgmo_host_alloc) when the
// driver grants it, plain pageable memory otherwise (accepted by every entry point, only slower).
class PinnedBlock {
 public:
  PinnedBlock() = default;
  PinnedBlock(const PinnedBlock&) = delete;
  PinnedBlock& operator=(const PinnedBlock&) = delete;
  ~PinnedBlock() { release(); }
  void reserve(size_t bytes) {
    if (bytes <= bytes_) return;
    release();
    const size_t want = bytes + bytes / 2;  // page-locking is slow: grow rarely
    const int rc = pgmo_host_alloc(want, &ptr_);
    if (rc == PGMO_ERR_CUDA) {  // e.g. the locked-memory limit: stage through pageable memory
      ptr_ = std::malloc(want);
      if (!ptr_) throw std::bad_alloc();
      pinned_ = false;
    } else {
      throwOnError(rc);
      pinned_ = true;
    }
    bytes_ = want;
    used_ = 0;
  }
  bool pinned() const { return pinned_; }
  void rewind() { used_ = 0; }
  // the next `count` elements of T, 256-byte aligned (reserve() must have covered them)
  template <typename T>
  T* take(size_t count) {
    used_ = (used_ + 255) & ~size_t(255);
    T* out = reinterpret_cast<T*>(static_cast<char*>(ptr_) + used_);
    used_ += count * sizeof(T);
    if (used_ > bytes_) throw std::logic_error("pgmo_b200: staging block overrun");
    return out;
  }
  static size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }

 private:
  void release() {
    if (ptr_ && pinned_) pgmo_host_free(ptr_);
    if (ptr_ && !pinned_) std::free(ptr_);
    ptr_ = nullptr;
    bytes_ = used_ = 0;
  }
  void* ptr_ = nullptr;
  size_t bytes_ = 0, used_ = 0;
  bool pinned_ = true;
};
// The x / y / z of the points inside the byte blob of a pcl::PCLPointCloud2, as a cloud the traits
// layer understands (ADL customization points below): deformMesh works on the blobs directly
// instead of converting blob -> PointCloud -> blob (deformation_graph.cpp:540, :557: two copies
// of 32 B per vertex plus a temporary cloud).
struct BlobCloud {
  unsigned char* base = nullptr;  // first point
  size_t step = 0;                // bytes per point
  size_t off[3] = {0, 4, 8};      // byte offsets of the float x / y / z inside a point
  size_t n = 0;
  const std::vector<Timestamp>* stamps = nullptr;
};
inline size_t pgmoNumVertices(const BlobCloud& c) { return c.n; }
inline traits::Pos pgmoGetVertex(const BlobCloud& c, size_t i, traits::VertexTraits* t) {
  if (i >= c.n) throw std::out_of_range("pgmo_b200: vertex index past the cloud");  // cloud.at()
  float v[3];
  const unsigned char* p = c.base + i * c.step;
  for (int a = 0; a < 3; ++a) std::memcpy(&v[a], p + c.off[a], sizeof(float));
  if (t && c.stamps && i < c.stamps->size()) t->stamp = (*c.stamps)[i];
  return traits::Pos(v[0], v[1], v[2]);
}
inline void pgmoSetVertex(BlobCloud& c, size_t i, const traits::Pos& pos, const traits::VertexTraits&) {
  if (i >= c.n) throw std::out_of_range("pgmo_b200: vertex index past the cloud");
  unsigned char* p = c.base + i * c.step;
  const float v[3] = {pos.x(), pos.y(), pos.z()};
  for (int a = 0; a < 3; ++a) std::memcpy(p + c.off[a], &v[a], sizeof(float));  // colour untouched
}
inline Timestamp pgmoGetVertexStamp(const BlobCloud& c, size_t i) { return c.stamps->at(i); }

// Where the points of `blob` are, if the blob has exactly the layout toPCLPointCloud2 gives a
// PointXYZRGBA cloud (then blob -> cloud -> blob is the identity on everything but x / y / z and
// the blob can be deformed in place); false: take the generic route.
#ifdef KIMERA_PGMO_B200_REAL_DEPS
inline bool blobView(const pcl::PCLPointCloud2& blob, BlobCloud* view) {
  static const pcl::PCLPointCloud2 canonical = [] {
    pcl::PCLPointCloud2 c;
    pcl::toPCLPointCloud2(pcl::PointCloud<pcl::PointXYZRGBA>(), c);
    return c;
  }();
  const size_t n = static_cast<size_t>(blob.width) * blob.height;
  if (blob.is_bigendian || blob.point_step != canonical.point_step ||
      blob.fields.size() != canonical.fields.size() || blob.data.size() != n * blob.point_step ||
      (blob.height > 1 && blob.row_step != blob.width * blob.point_step)) {
    return false;
  }
  int found = 0;
  for (size_t f = 0; f < blob.fields.size(); ++f) {
    const auto& a = blob.fields[f];
    const auto& b = canonical.fields[f];
    if (a.name != b.name || a.offset != b.offset || a.datatype != b.datatype || a.count != b.count) return false;
    for (int ax = 0; ax < 3; ++ax) {
      if (a.name == (ax == 0 ? "x" : ax == 1 ? "y" : "z")) {
        if (a.datatype != pcl::PCLPointField::FLOAT32) return false;
        view->off[ax] = a.offset;
        ++found;
      }
    }
  }
  if (found != 3) return false;
  view->base = const_cast<unsigned char*>(blob.data.data());
  view->step = blob.point_step;
  view->n = n;
  return true;
}
// `out` = `in` without the point bytes (header, fields, sizes; data sized, contents unspecified)
inline void blobCloneShape(const pcl::PCLPointCloud2& in, pcl::PCLPointCloud2& out) {
  out.header = in.header;
  out.height = in.height;
  out.width = in.width;
  out.fields = in.fields;
  out.is_bigendian = in.is_bigendian;
  out.point_step = in.point_step;
  out.row_step = in.row_step;
  out.is_dense = in.is_dense;
  resizePrefaulted(out.data, in.data.size());
}
#else
inline bool blobView(const pcl::PCLPointCloud2& blob, BlobCloud* view) {
  // the stand-in keeps the blob typed: x, y, z lead every record
  view->base = reinterpret_cast<unsigned char*>(const_cast<pcl::PointXYZRGBA*>(blob.data.data()));
  view->step = sizeof(pcl::PointXYZRGBA);
  view->off[0] = offsetof(pcl::PointXYZRGBA, x);
  view->off[1] = offsetof(pcl::PointXYZRGBA, y);
  view->off[2] = offsetof(pcl::PointXYZRGBA, z);
  view->n = blob.data.size();
  return true;
}
inline void blobCloneShape(const pcl::PCLPointCloud2& in, pcl::PCLPointCloud2& out) {
  out.width = in.width;
  out.height = in.height;
  out.data.reserve(in.data.size());
  {  // as resizePrefaulted (the stand-in's point type has a constructor)
    volatile unsigned char* raw = reinterpret_cast<volatile unsigned char*>(out.data.data());
    const size_t bytes = in.data.size() * sizeof(pcl::PointXYZRGBA);
    parallelFor((bytes + 4095) / 4096, [&](size_t b, size_t e) {
      for (size_t page = b; page < e; ++page) raw[page * 4096] = 0;
    }, 2048);
  }
  out.data.resize(in.data.size());
}
#endif
}  // namespace detail

namespace deformation {

// One GPU-side control table + window-schedule cache.  Owned by whoever owns the deformation
// graph (one per prefix); not thread-safe, like the reference's DeformationGraph.
class Backend {
 public:
  // n_devices: 1 = `device` only; 0 = every visible device; n > 1 = devices 0 .. n - 1.  With more
  // than one device, full-cloud calls of at least groupMinVertices() vertices that do not ask for
  // the association map are sharded by vertex range over all of them (pgmo_deform_group_*);
  // everything else (indices, small clouds, association output) runs on `device`.
  explicit Backend(int device = 0, int n_devices = 1) {
    pgmo_deform_t* h = nullptr;
    detail::throwOnError(pgmo_deform_create(device, &h));
    handle_.reset(h, [](pgmo_deform_t* p) { pgmo_deform_destroy(p); });
    const int visible = pgmo_device_count();
    group_list_.clear();
    for (int i = 0; i < (n_devices == 0 ? visible : std::min(n_devices, visible)); ++i) group_list_.push_back(i);
    // test / tuning knob: an explicit device list, e.g. "0,0" = two slabs on one device
    if (const char* e = std::getenv("KIMERA_PGMO_B200_GROUP_DEVICES")) {
      if (n_devices != 1) {
        group_list_.clear();
        for (const char* p = e; *p;) {
          char* end = nullptr;
          const long d = std::strtol(p, &end, 10);
          if (end == p) break;
          group_list_.push_back(static_cast<int>(d));
          p = *end ? end + 1 : end;
        }
      }
    }
    group_devices_ = static_cast<int>(group_list_.size());
  }
  // clouds below this size stay on one device (tuning knob; process-wide)
  static size_t& groupMinVertices() {
    static size_t n = size_t(1) << 20;
    return n;
  }
  pgmo_deform_t* get() const { return handle_.get(); }
  int groupDevices() const { return group_devices_; }
  // the device group with the control table of the last setControls (created on first use)
  pgmo_deform_group_t* group() {
    if (!group_) {
      pgmo_deform_group_t* g = nullptr;
      detail::throwOnError(pgmo_deform_group_create(group_devices_, group_list_.data(), &g));
      group_.reset(g, [](pgmo_deform_group_t* p) { pgmo_deform_group_destroy(p); });
      group_stale_ = true;
    }
    if (group_stale_) {
      detail::throwOnError(pgmo_deform_group_set_controls(
          group_.get(), valid_.size(), pos_.data(), has_stamps_ ? stamps_.data() : nullptr,
          valid_.data(), poses_.data()));
      group_stale_ = false;
    }
    return group_.get();
  }
  // page-locked staging shared by the copies of this Backend (calls are serial per Backend)
  kimera_pgmo_b200::detail::PinnedBlock& staging() {
    if (!staging_) staging_ = std::make_shared<kimera_pgmo_b200::detail::PinnedBlock>();
    return *staging_;
  }

  // The next setControls is a repeat of the last one (DeformationGraph::deformPoints uploads the
  // table for predeformPoints, then deformation::deformPoints would flatten and upload it again).
  // Scoped: the flag is dropped when the guard goes, whether or not a setControls consumed it
  // (deformPoints may return early or throw before it gets there).
  class KeepControls {
   public:
    explicit KeepControls(Backend& b, bool keep) : b_(b) { b_.controls_kept_ = keep; }
    ~KeepControls() { b_.controls_kept_ = false; }
    KeepControls(const KeepControls&) = delete;
    KeepControls& operator=(const KeepControls&) = delete;

   private:
    Backend& b_;
  };

  // Uploads control positions / stamps and the flattened `values`.
  void setControls(char prefix, const std::vector<gtsam::Point3>& control_points,
                   const std::vector<Timestamp>* control_point_stamps,
                   const gtsam::Values& values) {
    if (controls_kept_) {  // uploaded a moment ago by the same caller with the same arguments
      controls_kept_ = false;
      return;
    }
    const size_t n_c = control_points.size();
    pos_.resize(3 * n_c);
    valid_.resize(n_c);
    poses_.assign(12 * n_c, 0.0);
    // read-only look-ups of `values` (two ordered-map searches per control point): on threads
    kimera_pgmo_b200::detail::parallelFor(n_c, [&](size_t begin, size_t end) {
      for (size_t j = begin; j < end; ++j) {
        pos_[3 * j + 0] = control_points[j].x();
        pos_[3 * j + 1] = control_points[j].y();
        pos_[3 * j + 2] = control_points[j].z();
        const gtsam::Key key = gtsam::Symbol(prefix, j);
        valid_[j] = values.exists(key) ? 1 : 0;
        if (!valid_[j]) continue;
        const gtsam::Pose3& T = values.at<gtsam::Pose3>(key);
        const auto R = T.rotation().matrix();
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) poses_[12 * j + 3 * r + c] = R(r, c);
        }
        poses_[12 * j + 9] = T.translation().x();
        poses_[12 * j + 10] = T.translation().y();
        poses_[12 * j + 11] = T.translation().z();
      }
    }, 4096);
    const uint64_t* stamps = nullptr;
    if (control_point_stamps && control_point_stamps->size() >= n_c) {  // .at(j), j < n_c
      stamps = control_point_stamps->data();
    }
    detail::throwOnError(pgmo_deform_set_controls(handle_.get(), n_c, pos_.data(), stamps,
                                                  valid_.data(), poses_.data()));
    if (group_devices_ > 1) {  // the group gets the table when (if) a call is sharded
      has_stamps_ = stamps != nullptr;
      if (stamps) stamps_.assign(stamps, stamps + n_c);
      group_stale_ = true;
    }
  }

 private:
  std::shared_ptr<pgmo_deform_t> handle_;
  std::shared_ptr<pgmo_deform_group_t> group_;
  int group_devices_ = 1;
  std::vector<int> group_list_;
  bool group_stale_ = true, has_stamps_ = false;
  std::vector<uint64_t> stamps_;
  std::shared_ptr<kimera_pgmo_b200::detail::PinnedBlock> staging_;
  bool controls_kept_ = false;
  std::vector<double> pos_, poses_;
  std::vector<uint8_t> valid_;
};

namespace detail {

// process-wide default backend for the free functions (the reference's are stateless)
inline Backend& defaultBackend() {
  static thread_local Backend backend(0);
  return backend;
}

// control_point_map == nullptr: the caller does not read the associations (no association
// output, no std::set construction).
template <typename CloudOut, typename CloudIn>
void run(Backend& backend, CloudOut& new_points,
         std::vector<std::set<size_t>>* control_point_map, const CloudIn& points, bool stamped,
         size_t k, double tol_t, const std::vector<size_t>* indices, int precision) {
  using kimera_pgmo_b200::detail::parallelFor;
  using kimera_pgmo_b200::detail::PinnedBlock;
  const size_t n_v = traits::num_vertices(points);
  const size_t n_pos = indices ? indices->size() : n_v;
  // With `indices` the reference sweeps the listed vertices in the listed order and touches no
  // other vertex (mesh_deformation.h:159-161): only those are gathered, in that order, and
  // deformed as a cloud of their own — an incremental update of a few thousand vertices of a
  // 10 M-vertex mesh moves a few thousand vertices across PCIe, not the mesh.
  kimera_pgmo_b200::detail::Trace trace("run");
  PinnedBlock& stage = backend.staging();
  stage.reserve(6 * PinnedBlock::padded(n_pos * sizeof(float)) +
                PinnedBlock::padded(stamped ? n_pos * sizeof(uint64_t) : 0) +
                PinnedBlock::padded(n_pos) + 256);
  stage.rewind();
  trace.mark("staging reserve");
  float* x = stage.take<float>(n_pos);
  float* y = stage.take<float>(n_pos);
  float* z = stage.take<float>(n_pos);
  float* ox = stage.take<float>(n_pos);  // only read back where `processed` is set
  float* oy = stage.take<float>(n_pos);
  float* oz = stage.take<float>(n_pos);
  uint64_t* stamps = stamped ? stage.take<uint64_t>(n_pos) : nullptr;
  uint8_t* processed = stage.take<uint8_t>(n_pos);
  parallelFor(n_pos, [&](size_t begin, size_t end) {
    for (size_t p = begin; p < end; ++p) {
      const size_t ii = indices ? (*indices)[p] : p;
      if (ii >= n_v) throw std::out_of_range("pgmo_b200: vertex index past the cloud");  // .at()
      const traits::Pos v = traits::get_vertex(points, ii);
      x[p] = v.x();
      y[p] = v.y();
      z[p] = v.z();
      if (stamped) stamps[p] = traits::get_timestamp(points, ii);
    }
  });
  trace.mark("gather");
  const uint32_t stride = static_cast<uint32_t>(k + 2 > 8 ? k + 2 : 8);
  std::vector<uint32_t> assoc(control_point_map ? n_pos * stride : 0);
  std::vector<uint8_t> cnt(control_point_map ? n_pos : 0);
  int rc;
  if (!control_point_map && !indices && backend.groupDevices() > 1 &&
      n_pos >= Backend::groupMinVertices()) {
    // vertex slabs over every device of the box; same results, bit for bit
    rc = pgmo_deform_group_points(backend.group(), n_pos, x, y, z, stamps, static_cast<uint32_t>(k),
                                  tol_t, precision, ox, oy, oz, processed);
  } else {
    rc = pgmo_deform_points(backend.get(), n_pos, x, y, z, stamps, nullptr, 0,
                            static_cast<uint32_t>(k), tol_t, precision, ox, oy, oz, nullptr,
                            control_point_map ? assoc.data() : nullptr, stride,
                            control_point_map ? cnt.data() : nullptr, processed);
  }
  kimera_pgmo_b200::detail::throwOnError(rc);
  trace.mark("C ABI call");
  if (rc == PGMO_NOOP) {
    // The early returns that precede control_point_map.clear() (no points :78-80 / :143-146, no /
    // too few control points :83-86 / :149-152) were taken by deformPointsImpl; what is left is
    // "octree has fewer than k leaves" of the unstamped overload (:96-99), which comes AFTER the
    // clear() of :88.
    if (control_point_map) control_point_map->clear();
    return;
  }
  if (!control_point_map) {
    parallelFor(n_pos, [&](size_t begin, size_t end) {
      for (size_t p = begin; p < end; ++p) {
        if (!processed[p]) continue;  // skipped vertices stay as they are (:184-186)
        const size_t ii = indices ? (*indices)[p] : p;
        traits::set_vertex(new_points, ii, traits::Pos(ox[p], oy[p], oz[p]));
      }
    });
    trace.mark("scatter");
    return;
  }
  control_point_map->clear();
  for (size_t p = 0; p < n_pos; ++p) {
    if (!processed[p]) continue;  // skipped vertices own no entry (mesh_deformation.h:184-186)
    const size_t ii = indices ? (*indices)[p] : p;
    control_point_map->emplace_back();
    // neighbour counts above the stride only occur for degenerate k; the set is what callers use
    const uint32_t n = cnt[p] < stride ? cnt[p] : stride;
    for (uint32_t j = 0; j < n; ++j) control_point_map->back().insert(assoc[p * stride + j]);
    traits::set_vertex(new_points, ii, traits::Pos(ox[p], oy[p], oz[p]));
  }
}

// common body of the public overloads (early returns as in the reference)
template <typename CloudOut, typename CloudIn>
void deformPointsImpl(CloudOut& new_points, std::vector<std::set<size_t>>* control_point_map,
                      const CloudIn& points, char prefix,
                      const std::vector<gtsam::Point3>& control_points,
                      const std::vector<Timestamp>& control_point_stamps,
                      const gtsam::Values& values, size_t k, double tol_t,
                      const std::vector<size_t>* indices, Backend* backend, int precision) {
  constexpr bool kStamped = traits::has_get_stamp<CloudIn>::value;
  const size_t num_points = indices ? indices->size() : traits::num_vertices(points);
  if (!num_points) return;
  if (kStamped) {
    if (control_points.size() < k) return;  // "Not enough valid control points to deform points"
    if (control_point_stamps.size() < control_points.size()) {
      throw std::out_of_range("control_point_stamps shorter than control_points");
    }
  } else if (control_points.size() == 0) {
    return;  // "No control points. No deformation"
  }
  Backend& b = backend ? *backend : defaultBackend();
  kimera_pgmo_b200::detail::Trace trace("deformPointsImpl");
  b.setControls(prefix, control_points, kStamped ? &control_point_stamps : nullptr, values);
  trace.mark("setControls");
  run(b, new_points, control_point_map, points, kStamped, k, kStamped ? tol_t : 0.0, indices,
      precision);
}

}  // namespace detail

// mesh_deformation.h:63-113 — clouds without stamps
template <typename CloudOut, typename CloudIn,
          std::enable_if_t<!traits::has_get_stamp<CloudIn>::value, bool> = true>
void deformPoints(CloudOut& new_points, std::vector<std::set<size_t>>& control_point_map,
                  const CloudIn& points, char prefix,
                  const std::vector<gtsam::Point3>& control_points,
                  const std::vector<Timestamp>& control_point_stamps,
                  const gtsam::Values& values, size_t k = 4, double tol_t = 10.0,
                  const std::vector<size_t>* indices = nullptr, Backend* backend = nullptr,
                  int precision = PGMO_PRECISION_FP64) {
  detail::deformPointsImpl(new_points, &control_point_map, points, prefix, control_points,
                           control_point_stamps, values, k, tol_t, indices, backend, precision);
}

// mesh_deformation.h:129-213 — stamped clouds
template <typename CloudOut, typename CloudIn,
          std::enable_if_t<traits::has_get_stamp<CloudIn>::value, bool> = true>
void deformPoints(CloudOut& new_points, std::vector<std::set<size_t>>& control_point_map,
                  const CloudIn& points, char prefix,
                  const std::vector<gtsam::Point3>& control_points,
                  const std::vector<Timestamp>& control_point_stamps,
                  const gtsam::Values& values, size_t k = 4, double tol_t = 10.0,
                  const std::vector<size_t>* indices = nullptr, Backend* backend = nullptr,
                  int precision = PGMO_PRECISION_FP64) {
  detail::deformPointsImpl(new_points, &control_point_map, points, prefix, control_points,
                           control_point_stamps, values, k, tol_t, indices, backend, precision);
}

// Extension (no reference counterpart): either overload above for callers that do not read
// control_point_map — same vertices, same early returns; nothing is computed, copied or
// allocated for the associations.
template <typename CloudOut, typename CloudIn>
void deformPointsOnly(CloudOut& new_points, const CloudIn& points, char prefix,
                      const std::vector<gtsam::Point3>& control_points,
                      const std::vector<Timestamp>& control_point_stamps,
                      const gtsam::Values& values, size_t k = 4, double tol_t = 10.0,
                      const std::vector<size_t>* indices = nullptr, Backend* backend = nullptr,
                      int precision = PGMO_PRECISION_FP64) {
  detail::deformPointsImpl(new_points, nullptr, points, prefix, control_points,
                           control_point_stamps, values, k, tol_t, indices, backend, precision);
}

}  // namespace deformation
}  // namespace kimera_pgmo_b200
