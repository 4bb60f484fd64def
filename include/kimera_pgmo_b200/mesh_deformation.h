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
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <new>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

// upper bound on the threads of the shim's gather / scatter loops (1 = single-threaded)
#ifndef KIMERA_PGMO_B200_SHIM_THREADS
#define KIMERA_PGMO_B200_SHIM_THREADS 8
#endif

#include "kimera_pgmo_b200/deps.h"
#include "pgmo_b200.h"

namespace kimera_pgmo_b200 {

// common_structs.cpp:21-24
inline Timestamp stampFromSec(double sec) { return pgmo_stamp_from_sec(sec); }

class PgmoB200Error : public std::runtime_error {
 public:
  PgmoB200Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

namespace detail {
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
void parallelFor(size_t n, Fn&& fn) {
  constexpr size_t kMinPerThread = size_t(1) << 17;
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
}  // namespace detail

namespace deformation {

// One GPU-side control table + window-schedule cache.  Owned by whoever owns the deformation
// graph (one per prefix); not thread-safe, like the reference's DeformationGraph.
class Backend {
 public:
  explicit Backend(int device = 0) {
    pgmo_deform_t* h = nullptr;
    detail::throwOnError(pgmo_deform_create(device, &h));
    handle_.reset(h, [](pgmo_deform_t* p) { pgmo_deform_destroy(p); });
  }
  pgmo_deform_t* get() const { return handle_.get(); }
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
    for (size_t j = 0; j < n_c; ++j) {
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
    const uint64_t* stamps = nullptr;
    if (control_point_stamps && control_point_stamps->size() >= n_c) {  // .at(j), j < n_c
      stamps = control_point_stamps->data();
    }
    detail::throwOnError(pgmo_deform_set_controls(handle_.get(), n_c, pos_.data(), stamps,
                                                  valid_.data(), poses_.data()));
  }

 private:
  std::shared_ptr<pgmo_deform_t> handle_;
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
  PinnedBlock& stage = backend.staging();
  stage.reserve(6 * PinnedBlock::padded(n_pos * sizeof(float)) +
                PinnedBlock::padded(stamped ? n_pos * sizeof(uint64_t) : 0) +
                PinnedBlock::padded(n_pos) + 256);
  stage.rewind();
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
  const uint32_t stride = static_cast<uint32_t>(k + 2 > 8 ? k + 2 : 8);
  std::vector<uint32_t> assoc(control_point_map ? n_pos * stride : 0);
  std::vector<uint8_t> cnt(control_point_map ? n_pos : 0);
  int rc = pgmo_deform_points(backend.get(), n_pos, x, y, z, stamps, nullptr, 0,
                              static_cast<uint32_t>(k), tol_t, precision, ox, oy, oz, nullptr,
                              control_point_map ? assoc.data() : nullptr, stride,
                              control_point_map ? cnt.data() : nullptr, processed);
  kimera_pgmo_b200::detail::throwOnError(rc);
  if (rc == PGMO_NOOP) return;  // the reference returned before control_point_map.clear()
  if (!control_point_map) {
    parallelFor(n_pos, [&](size_t begin, size_t end) {
      for (size_t p = begin; p < end; ++p) {
        if (!processed[p]) continue;  // skipped vertices stay as they are (:184-186)
        const size_t ii = indices ? (*indices)[p] : p;
        traits::set_vertex(new_points, ii, traits::Pos(ox[p], oy[p], oz[p]));
      }
    });
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
  b.setControls(prefix, control_points, kStamped ? &control_point_stamps : nullptr, values);
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
