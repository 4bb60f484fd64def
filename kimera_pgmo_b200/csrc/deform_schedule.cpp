// Host-side window schedule builder.  See deform_schedule.h for the idea; the arithmetic that
// decides voxel membership follows PCL's OctreePointCloud (adoptBoundingBoxToPoint,
// getKeyBitSize, genOctreeKeyforPoint, deleteVoxelAtPoint) because the reference counts and
// deletes *1 m voxels*, not points (mesh_deformation.cpp:31,40).
#include "deform_schedule.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <limits>
#include <vector>

namespace pgmo {

uint64_t stampFromSec(double sec) {
  // kimera_pgmo/src/utils/common_structs.cpp:21-24
  return static_cast<uint64_t>(static_cast<int64_t>(sec * 1e9));
}

namespace {

struct VoxelKey {
  int32_t x, y, z;
  bool operator==(const VoxelKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelKeyHash {
  size_t operator()(const VoxelKey& k) const {
    uint64_t h = static_cast<uint32_t>(k.x) * 0x9E3779B97F4A7C15ull;
    h ^= (static_cast<uint64_t>(static_cast<uint32_t>(k.y)) + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= (static_cast<uint64_t>(static_cast<uint32_t>(k.z)) + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};

// Leaf bookkeeping of the search octree: which voxels are occupied and by which control points.
// No search structure is kept here — only what getLeafCount / deleteVoxelAtPoint observe.
// Flat open-addressing table (every control point is added at most once, so a table of
// >= 2 n_c slots never fills up even though deletions leave tombstones) with the points of a
// leaf chained through `next_`: no allocation per add / remove, this runs once per changed
// control table on the critical path of the first deformMesh after a solve.
class LeafBook {
 public:
  LeafBook(double resolution, uint64_t n_points) : res_(resolution) {
    uint64_t cap = 64;
    while (cap < 2 * n_points + 16) cap <<= 1;
    mask_ = cap - 1;
    // the storage of the previous build on this thread (a few MB): a fresh allocation would be
    // paid in page faults on the critical path of every first deformMesh after a solve
    slots_.swap(workspace().slots);
    next_.swap(workspace().next);
    slots_.assign(cap, Slot());
    next_.assign(n_points, -1);
  }
  ~LeafBook() {
    slots_.swap(workspace().slots);
    next_.swap(workspace().next);
  }
  LeafBook(const LeafBook&) = delete;
  LeafBook& operator=(const LeafBook&) = delete;

  size_t leafCount() const { return n_leaves_; }
  uint32_t maxLeafPoints() const { return max_leaf_points_; }

  void frame(ScheduleRun* run) const {
    run->min_xyz[0] = min_[0];
    run->min_xyz[1] = min_[1];
    run->min_xyz[2] = min_[2];
    run->depth = depth_;
    run->shift[0] = shift_[0];
    run->shift[1] = shift_[1];
    run->shift[2] = shift_[2];
    run->max_xyz[0] = max_[0];
    run->max_xyz[1] = max_[1];
    run->max_xyz[2] = max_[2];
    run->box_defined = box_defined_ ? 1 : 0;
  }

  // extendStampedSchedule: the box as it was when `run` was emitted (it only changes inside
  // add(), and every step that adds also emits a run), then the alive points re-inserted in
  // index order by restore().  Stable keys do not depend on the growth history.
  void restoreFrame(const ScheduleRun& run, uint32_t max_leaf_points_so_far) {
    box_defined_ = run.box_defined != 0;
    for (int a = 0; a < 3; ++a) {
      min_[a] = run.min_xyz[a];
      max_[a] = run.max_xyz[a];
      shift_[a] = run.shift[a];
    }
    depth_ = run.depth;
    max_leaf_points_ = max_leaf_points_so_far;
  }
  void restore(uint32_t j, const VoxelKey& stable) {
    uint64_t at = VoxelKeyHash()(stable) & mask_;
    while (true) {
      Slot& s = slots_[at];
      if (s.state == kLive && s.key == stable) break;
      if (s.state == kEmpty) {
        s.state = kLive;
        s.key = stable;
        s.head = -1;
        s.count = 0;
        ++n_leaves_;
        break;
      }
      at = (at + 1) & mask_;
    }
    Slot& leaf = slots_[at];
    next_[j] = leaf.head;
    leaf.head = static_cast<int32_t>(j);
    ++leaf.count;
  }

  // addPointFromCloud for control point j at float position p; returns its stable key
  VoxelKey add(uint32_t j, const float p[3]) {
    growToFit(p);
    const VoxelKey stable = stableKey(p);
    uint64_t at = VoxelKeyHash()(stable) & mask_;
    int64_t first_tomb = -1;
    while (true) {
      Slot& s = slots_[at];
      if (s.state == kLive && s.key == stable) break;
      if (s.state == kTomb && first_tomb < 0) first_tomb = static_cast<int64_t>(at);
      if (s.state == kEmpty) {
        if (first_tomb >= 0) at = static_cast<uint64_t>(first_tomb);
        Slot& fresh = slots_[at];
        fresh.state = kLive;
        fresh.key = stable;
        fresh.head = -1;
        fresh.count = 0;
        ++n_leaves_;
        break;
      }
      at = (at + 1) & mask_;
    }
    Slot& leaf = slots_[at];
    next_[j] = leaf.head;
    leaf.head = static_cast<int32_t>(j);
    ++leaf.count;
    max_leaf_points_ = std::max(max_leaf_points_, leaf.count);
    return stable;
  }

  // deleteVoxelAtPoint: removes the whole leaf; `killed` receives the control points in it
  void removeVoxelAt(const float p[3], std::vector<uint32_t>* killed) {
    killed->clear();
    if (!box_defined_) return;
    for (int a = 0; a < 3; ++a) {
      if (p[a] < min_[a] || p[a] >= max_[a]) return;  // isPointWithinBoundingBox
    }
    const VoxelKey stable = stableKey(p);
    uint64_t at = VoxelKeyHash()(stable) & mask_;
    while (true) {
      Slot& s = slots_[at];
      if (s.state == kEmpty) return;
      if (s.state == kLive && s.key == stable) {
        for (int32_t j = s.head; j >= 0; j = next_[j]) killed->push_back(static_cast<uint32_t>(j));
        s.state = kTomb;
        --n_leaves_;
        return;
      }
      at = (at + 1) & mask_;
    }
  }

 private:
  VoxelKey stableKey(const float p[3]) const {
    // genOctreeKeyforPoint: unsigned((p - min) / resolution), float promoted to double
    // (the search tree's resolution is 1 m, mesh_deformation.cpp:22-30: x / 1.0 is x, exactly)
    const bool unit = res_ == 1.0;
    const uint32_t kx = static_cast<uint32_t>(unit ? p[0] - min_[0] : (p[0] - min_[0]) / res_);
    const uint32_t ky = static_cast<uint32_t>(unit ? p[1] - min_[1] : (p[1] - min_[1]) / res_);
    const uint32_t kz = static_cast<uint32_t>(unit ? p[2] - min_[2] : (p[2] - min_[2]) / res_);
    return {static_cast<int32_t>(kx) - shift_[0], static_cast<int32_t>(ky) - shift_[1],
            static_cast<int32_t>(kz) - shift_[2]};
  }

  void growToFit(const float p[3]) {
    const float eps = std::numeric_limits<float>::epsilon();
    while (true) {
      bool lower = false, upper[3];
      for (int a = 0; a < 3; ++a) {
        lower = lower || (p[a] < min_[a]);
        upper[a] = p[a] >= max_[a];
      }
      if (!(lower || upper[0] || upper[1] || upper[2] || !box_defined_)) break;
      if (box_defined_) {
        // one more level: the old root becomes a child of the new root
        double side = static_cast<double>(1u << depth_) * res_;
        for (int a = 0; a < 3; ++a) {
          if (!upper[a]) {
            min_[a] -= side;
            shift_[a] += static_cast<int32_t>(1u << depth_);
          }
        }
        ++depth_;
        side = static_cast<double>(1u << depth_) * res_ - eps;
        for (int a = 0; a < 3; ++a) max_[a] = min_[a] + side;
      } else {
        for (int a = 0; a < 3; ++a) {
          min_[a] = p[a] - res_ / 2;
          max_[a] = p[a] + res_ / 2;
        }
        // getKeyBitSize on an empty tree
        uint32_t max_voxels = 2;
        for (int a = 0; a < 3; ++a) {
          max_voxels = std::max(
              max_voxels, static_cast<uint32_t>(std::ceil((max_[a] - min_[a] - eps) / res_)));
        }
        depth_ = static_cast<uint32_t>(
            std::max(std::min(32.0, std::ceil(std::log2(max_voxels) - eps)), 0.0));
        const double side = static_cast<double>(1u << depth_) * res_;
        for (int a = 0; a < 3; ++a) {
          const double oversize = (side - (max_[a] - min_[a])) / 2.0;
          if (oversize > eps) {
            min_[a] -= oversize;
            max_[a] += oversize;
          }
        }
        box_defined_ = true;
      }
    }
  }

  double res_;
  bool box_defined_ = false;
  double min_[3] = {0, 0, 0};
  double max_[3] = {0, 0, 0};
  uint32_t depth_ = 1;
  int32_t shift_[3] = {0, 0, 0};
  enum : uint8_t { kEmpty = 0, kLive = 1, kTomb = 2 };
  struct Slot {
    VoxelKey key{0, 0, 0};
    int32_t head = -1;
    uint32_t count = 0;
    uint8_t state = kEmpty;
  };
  struct Workspace {
    std::vector<Slot> slots;
    std::vector<int32_t> next;
  };
  static Workspace& workspace() {
    static thread_local Workspace w;
    return w;
  }
  std::vector<Slot> slots_;
  std::vector<int32_t> next_;
  uint64_t mask_ = 0;
  size_t n_leaves_ = 0;
  uint32_t max_leaf_points_ = 0;
};

struct Builder {
  const ControlTableHost& ctrl;
  Schedule* out;
  LeafBook book;
  Builder(const ControlTableHost& c, double resolution, Schedule* o)
      : ctrl(c), out(o), book(resolution, c.n) {
    out->runs.clear();
    out->del_run.assign(c.n, 0);
    out->stable_key.assign(3 * c.n, 0);
    out->max_leaf_points = 0;
    out->max_k = 0;
    out->noop = false;
    out->degenerate = false;
    for (uint64_t j = 0; j < c.n; ++j) out->del_run[j] = isValid(j) ? INT32_MAX : 0;
  }
  bool isValid(uint64_t j) const { return !ctrl.valid || ctrl.valid[j]; }
  void floatPos(uint64_t j, float p[3]) const {
    // SearchTree::Impl::addPoint -> eigenToPcl: double -> float (mesh_deformation.cpp:16-20,34)
    p[0] = static_cast<float>(ctrl.pos_xyz[3 * j + 0]);
    p[1] = static_cast<float>(ctrl.pos_xyz[3 * j + 1]);
    p[2] = static_cast<float>(ctrl.pos_xyz[3 * j + 2]);
  }
  void addPoint(uint64_t j) {
    float p[3];
    floatPos(j, p);
    const VoxelKey key = book.add(static_cast<uint32_t>(j), p);
    out->stable_key[3 * j + 0] = key.x;
    out->stable_key[3 * j + 1] = key.y;
    out->stable_key[3 * j + 2] = key.z;
  }
  void emitRun(uint64_t position, uint32_t lo, uint32_t hi, uint32_t k, bool skip) {
    ScheduleRun run;
    run.first_position = position;
    run.lo = lo;
    run.hi = hi;
    run.k = k;
    run.skip = skip ? 1u : 0u;
    book.frame(&run);
    out->runs.push_back(run);
    out->max_k = std::max(out->max_k, k);
  }
};

}  // namespace

namespace {

struct ReplayState {
  size_t k = 0;             // mutated mid-sweep (:183) and stays changed
  uint64_t hi = 0, lo = 0;  // ctrl_pt_idx / lower_ctrl_pt_idx
  bool dirty = true;        // search state changed since the last emitted run
  uint64_t position = 0;
};

// mesh_deformation.h:159-212 from stretch `s_begin` on
void replayFrom(Builder& b, ReplayState st, uint64_t s_begin, uint64_t n_stretches,
                const uint64_t* stretch_stamp, const uint64_t* stretch_len, double tol_t) {
  const ControlTableHost& ctrl = b.ctrl;
  Schedule* out = b.out;
  const uint64_t tol = stampFromSec(tol_t);
  const uint64_t n_c = ctrl.n;
  size_t k = st.k;
  uint64_t hi = st.hi, lo = st.lo;
  bool dirty = st.dirty;
  uint64_t position = st.position;
  std::vector<uint32_t> killed;
  out->checkpoints.resize(s_begin);
  out->end_touch_stretch = ~0ull;

  for (uint64_t s = s_begin; s < n_stretches; ++s) {
    StretchCheckpoint cp;
    cp.position = position;
    cp.run_count = static_cast<uint32_t>(out->runs.size());
    cp.lo = static_cast<uint32_t>(lo);
    cp.hi = static_cast<uint32_t>(hi);
    cp.k = static_cast<uint32_t>(k);
    cp.dirty = dirty ? 1 : 0;
    out->checkpoints.push_back(cp);
    const uint64_t stamp = stretch_stamp[s];
    const uint64_t len = stretch_len[s];
    for (uint64_t v = 0; v < len; ++v) {
      bool step_changed = false;
      size_t num_ctrl_pts = b.book.leafCount();  // :164
      while (hi < n_c && (ctrl.stamps[hi] <= stamp + tol || num_ctrl_pts < k + 1)) {  // :167-169
        const bool valid = b.isValid(hi);
        if (valid) b.addPoint(hi);
        ++hi;
        step_changed = true;
        if (!valid) continue;
        ++num_ctrl_pts;
      }
      // the loop condition was just evaluated with hi == n_c: appended controls would change it
      if (hi == n_c && out->end_touch_stretch == ~0ull) out->end_touch_stretch = s;
      bool skip = false;
      if (b.book.leafCount() < k + 1) {  // :180
        if (num_ctrl_pts > 1) {
          if (k != num_ctrl_pts - 1) step_changed = true;
          k = num_ctrl_pts - 1;
        } else {
          skip = true;
        }
      }
      dirty = dirty || step_changed;
      if (out->runs.empty() || dirty || (out->runs.back().skip != 0) != skip) {
        b.emitRun(position + v, static_cast<uint32_t>(lo), static_cast<uint32_t>(hi),
                  static_cast<uint32_t>(k), skip);
        dirty = false;
      }
      if (!skip) {
        if (b.book.leafCount() == 0) out->degenerate = true;
        size_t num_leaves = b.book.leafCount();  // :200
        while (lo < n_c && num_leaves > k + 1 && ctrl.stamps[lo] < stamp - tol) {  // :201-202
          step_changed = true;
          dirty = true;
          if (!b.isValid(lo)) {
            ++lo;
            continue;
          }
          float p[3];
          b.floatPos(lo, p);
          b.book.removeVoxelAt(p, &killed);  // whole voxel; no-op when already gone
          for (uint32_t j : killed) out->del_run[j] = static_cast<int32_t>(out->runs.size());
          --num_leaves;
          ++lo;
        }
      }
      if (!step_changed) break;  // fixed point: the rest of the stretch sees this very state
    }
    position += len;
  }
  StretchCheckpoint last;  // behind the last stretch: where appended stretches resume
  last.position = position;
  last.run_count = static_cast<uint32_t>(out->runs.size());
  last.lo = static_cast<uint32_t>(lo);
  last.hi = static_cast<uint32_t>(hi);
  last.k = static_cast<uint32_t>(k);
  last.dirty = dirty ? 1 : 0;
  out->checkpoints.push_back(last);
  out->max_leaf_points = std::max(out->max_leaf_points, b.book.maxLeafPoints());
  out->built_n_controls = n_c;
}

}  // namespace

void buildStampedSchedule(const ControlTableHost& ctrl, uint64_t n_stretches,
                          const uint64_t* stretch_stamp, const uint64_t* stretch_len,
                          uint32_t k_in, double tol_t, double resolution, Schedule* out) {
  Builder b(ctrl, resolution, out);
  out->checkpoints.clear();
  out->end_touch_stretch = ~0ull;
  out->built_n_controls = ctrl.n;
  out->resumed_from_stretch = 0;
  uint64_t total = 0;
  for (uint64_t s = 0; s < n_stretches; ++s) total += stretch_len[s];
  out->total_positions = total;
  // mesh_deformation.h:143-152
  if (total == 0 || ctrl.n < k_in) {
    out->noop = true;
    return;
  }
  ReplayState st;
  st.k = k_in;
  replayFrom(b, st, 0, n_stretches, stretch_stamp, stretch_len, tol_t);
}

void extendStampedSchedule(const ControlTableHost& ctrl, uint64_t n_stretches,
                           const uint64_t* stretch_stamp, const uint64_t* stretch_len,
                           uint64_t common_stretches, uint32_t k_in, double tol_t, double resolution,
                           Schedule* inout) {
  const uint64_t n_old = inout->built_n_controls;
  uint64_t total = 0;
  for (uint64_t s = 0; s < n_stretches; ++s) total += stretch_len[s];
  const bool has_state = !inout->noop && !inout->degenerate && !inout->runs.empty() &&
                         !inout->checkpoints.empty() && n_old <= ctrl.n && inout->del_run.size() == n_old;
  uint64_t s_resume = 0;
  if (has_state) {
    const uint64_t old_stretches = inout->checkpoints.size() - 1;
    s_resume = std::min<uint64_t>({common_stretches, inout->end_touch_stretch, old_stretches, n_stretches});
  }
  if (!has_state || total == 0 || ctrl.n < k_in || inout->checkpoints[s_resume].run_count == 0) {
    buildStampedSchedule(ctrl, n_stretches, stretch_stamp, stretch_len, k_in, tol_t, resolution, inout);
    return;
  }
  const StretchCheckpoint cp = inout->checkpoints[s_resume];  // state at the start of that stretch
  const uint32_t R = cp.run_count;
  std::vector<ScheduleRun> runs(inout->runs.begin(), inout->runs.begin() + R);
  std::vector<int32_t> del_run(inout->del_run);
  std::vector<int32_t> stable_key(inout->stable_key);
  const uint32_t max_leaf = inout->max_leaf_points;
  const uint32_t max_k_prefix = [&] {
    uint32_t m = 0;
    for (const auto& r : runs) m = std::max(m, r.k);
    return m;
  }();
  std::vector<StretchCheckpoint> cps(inout->checkpoints.begin(), inout->checkpoints.begin() + s_resume);

  Builder b(ctrl, resolution, inout);  // resets the output for the new table
  inout->runs = std::move(runs);
  inout->max_k = max_k_prefix;
  inout->max_leaf_points = max_leaf;
  inout->checkpoints = std::move(cps);
  inout->total_positions = total;
  // per control point: dead before the resume point stays dead; everything else is alive again
  for (uint64_t j = 0; j < n_old; ++j) {
    if (!b.isValid(j)) continue;  // Builder set 0
    inout->del_run[j] = del_run[j] <= static_cast<int32_t>(R) && del_run[j] != 0 ? del_run[j] : INT32_MAX;
    inout->stable_key[3 * j + 0] = stable_key[3 * j + 0];
    inout->stable_key[3 * j + 1] = stable_key[3 * j + 1];
    inout->stable_key[3 * j + 2] = stable_key[3 * j + 2];
  }
  b.book.restoreFrame(inout->runs[R - 1], max_leaf);
  for (uint64_t j = cp.lo; j < cp.hi; ++j) {  // added and not deleted yet
    if (!b.isValid(j) || inout->del_run[j] != INT32_MAX) continue;
    b.book.restore(static_cast<uint32_t>(j), VoxelKey{stable_key[3 * j + 0], stable_key[3 * j + 1],
                                                      stable_key[3 * j + 2]});
  }
  ReplayState st;
  st.k = cp.k;
  st.hi = cp.hi;
  st.lo = cp.lo;
  st.dirty = cp.dirty != 0;
  st.position = cp.position;
  inout->resumed_from_stretch = s_resume;
  replayFrom(b, st, s_resume, n_stretches, stretch_stamp, stretch_len, tol_t);
}

void buildUnstampedSchedule(const ControlTableHost& ctrl, uint64_t total_positions, uint32_t k,
                            double resolution, Schedule* out) {
  Builder b(ctrl, resolution, out);
  out->total_positions = total_positions;
  // mesh_deformation.h:78-86
  if (total_positions == 0 || ctrl.n == 0) {
    out->noop = true;
    return;
  }
  for (uint64_t j = 0; j < ctrl.n; ++j) {
    if (b.isValid(j)) b.addPoint(j);
  }
  if (b.book.leafCount() < k) {  // :96-99
    out->noop = true;
    return;
  }
  if (b.book.leafCount() == 0) out->degenerate = true;
  b.emitRun(0, 0, static_cast<uint32_t>(ctrl.n), k, false);
  out->max_leaf_points = b.book.maxLeafPoints();
}

}  // namespace pgmo
