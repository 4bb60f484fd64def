// Window schedule for deformation::deformPoints (host side, pure C++).
//
// The reference sweeps the vertices sequentially and, per vertex, adds control points to /
// removes control-point voxels from a PCL octree (mesh_deformation.h:159-212).  That add/remove
// process is a deterministic state machine driven only by the vertex *stamps* in processing
// order and by the control table (position, stamp, validity) — never by vertex positions — and
// inside a stretch of equal stamps it reaches a fixed point after a few vertices.  The builder
// below replays it once per stretch head and emits *runs*: maximal ranges of processing
// positions that see the identical search structure, k and skip decision.  The big kernel then
// only needs `run(position)` and, per control point, the run at which its voxel was deleted.
#pragma once

#include <cstdint>
#include <vector>

namespace pgmo {

struct ScheduleRun {
  uint64_t first_position = 0;  // first processing position of the run
  uint32_t lo = 0;              // every control index < lo is dead or invalid
  uint32_t hi = 0;              // control indices >= hi have not been added yet
  uint32_t k = 0;               // the (mutated) k in force: k+1 blended, (k+2)-th = d_max
  uint32_t skip = 0;            // 1: reference `continue`s (mesh_deformation.h:184-186)
  // octree frame at query time (tie-order resolver): PCL bounding-box minimum, tree depth and
  // the cumulative key shift so that key_now = stable_key + shift
  double min_xyz[3] = {0, 0, 0};
  uint32_t depth = 0;
  int32_t shift[3] = {0, 0, 0};
  double max_xyz[3] = {0, 0, 0};  // host only: lets extendStampedSchedule restore the box exactly
  uint8_t box_defined = 0;        // host only: a control point had been added by then
};

// State of the replay at the start of a stretch (what extendStampedSchedule resumes from).
struct StretchCheckpoint {
  uint64_t position = 0;   // processing position of the stretch's first vertex
  uint32_t run_count = 0;  // runs emitted so far
  uint32_t lo = 0, hi = 0, k = 0;
  uint8_t dirty = 0;
};

struct Schedule {
  std::vector<ScheduleRun> runs;
  std::vector<int32_t> del_run;       // per control point: alive in run r iff del_run > r (0 = never alive)
  std::vector<int32_t> stable_key;    // 3 per control point: leaf key minus the frame shift at insertion
  uint64_t total_positions = 0;
  uint32_t max_leaf_points = 0;       // most control points ever sharing one voxel
  uint32_t max_k = 0;
  bool noop = false;                  // reference early return
  bool degenerate = false;            // a processed vertex would query an empty tree (reference UB)
  // for the incremental rebuild: one checkpoint per stretch plus one behind the last, and the first stretch whose replay
  // looked at the end of the control table (UINT64_MAX: none did).  Everything before that
  // stretch is independent of control points appended later.
  std::vector<StretchCheckpoint> checkpoints;
  uint64_t end_touch_stretch = ~0ull;
  uint64_t built_n_controls = 0;
  uint64_t resumed_from_stretch = 0;  // diagnostics: 0 = full replay
};

struct ControlTableHost {
  uint64_t n = 0;
  const double* pos_xyz = nullptr;   // n x 3
  const uint64_t* stamps = nullptr;  // n (may be null for the unstamped overload)
  const uint8_t* valid = nullptr;    // n (may be null = all valid)
};

// Stamped overload (mesh_deformation.h:129-213).  stretches: (stamp, length) in processing order.
void buildStampedSchedule(const ControlTableHost& ctrl, uint64_t n_stretches,
                          const uint64_t* stretch_stamp, const uint64_t* stretch_len,
                          uint32_t k, double tol_t, double resolution, Schedule* out);

// The same schedule after control points were APPENDED to the table `inout` was built from
// (positions / stamps / validity of the first inout->built_n_controls entries unchanged) and / or
// the stretch list changed behind its first `common_stretches` entries: the replay resumes at
// the first stretch that can see a difference — the earlier of `common_stretches` and the first
// stretch that looked at the end of the old table — from that stretch's checkpoint, with the
// octree's leaf bookkeeping rebuilt from the alive set recorded in the schedule itself.  Falls
// back to a full replay when nothing can be kept.  Result == buildStampedSchedule on the new
// inputs (tests/test_schedule_cpu.py).
void extendStampedSchedule(const ControlTableHost& ctrl, uint64_t n_stretches,
                           const uint64_t* stretch_stamp, const uint64_t* stretch_len,
                           uint64_t common_stretches, uint32_t k, double tol_t, double resolution,
                           Schedule* inout);

// Unstamped overload (mesh_deformation.h:63-113): one run over all valid control points.
void buildUnstampedSchedule(const ControlTableHost& ctrl, uint64_t total_positions, uint32_t k,
                            double resolution, Schedule* out);

uint64_t stampFromSec(double sec);

}  // namespace pgmo
