// Test-only C wrapper around the product's host-side window schedule builder
// (kimera_pgmo_b200/csrc/deform_schedule.cpp) so that its logic can be checked against the
// oracle on a machine without a GPU.
#include <cstring>

#include "deform_schedule.h"

extern "C" {

// runs6: first_position, lo, hi, k, skip, depth per run.  Returns the number of runs, -1 on noop.
long schedule_build(unsigned long n_c, const double* pos, const unsigned long* stamps,
                    const unsigned char* valid, unsigned long n_stretches,
                    const unsigned long* s_stamp, const unsigned long* s_len,
                    unsigned long total, unsigned k, double tol_t, int stamped,
                    unsigned long cap, unsigned long* runs6, int* del_run) {
  pgmo::ControlTableHost c;
  c.n = n_c;
  c.pos_xyz = pos;
  c.stamps = reinterpret_cast<const uint64_t*>(stamps);
  c.valid = valid;
  pgmo::Schedule s;
  if (stamped) {
    pgmo::buildStampedSchedule(c, n_stretches, reinterpret_cast<const uint64_t*>(s_stamp),
                               reinterpret_cast<const uint64_t*>(s_len), k, tol_t, 1.0, &s);
  } else {
    pgmo::buildUnstampedSchedule(c, total, k, 1.0, &s);
  }
  if (s.noop) return -1;
  for (size_t r = 0; r < s.runs.size() && r < cap; ++r) {
    runs6[6 * r + 0] = s.runs[r].first_position;
    runs6[6 * r + 1] = s.runs[r].lo;
    runs6[6 * r + 2] = s.runs[r].hi;
    runs6[6 * r + 3] = s.runs[r].k;
    runs6[6 * r + 4] = s.runs[r].skip;
    runs6[6 * r + 5] = s.runs[r].depth;
  }
  std::memcpy(del_run, s.del_run.data(), n_c * sizeof(int));
  return static_cast<long>(s.runs.size());
}
}

extern "C" {

// Builds the schedule of (first n_old controls, first ns_old stretches), extends it to the full
// inputs with extendStampedSchedule and compares with a from-scratch build of the full inputs.
// Returns 0 when identical, a positive code naming the first differing field otherwise;
// *resumed receives the stretch the extension resumed from (0 = it fell back to a full replay).
long schedule_extend_check(unsigned long n_old, unsigned long n_new, const double* pos,
                           const unsigned long* stamps, const unsigned char* valid,
                           unsigned long ns_old, const unsigned long* s_stamp_old,
                           const unsigned long* s_len_old, unsigned long ns_new,
                           const unsigned long* s_stamp_new, const unsigned long* s_len_new,
                           unsigned k, double tol_t, unsigned long* resumed) {
  pgmo::ControlTableHost c_old, c_new;
  c_old.n = n_old;
  c_new.n = n_new;
  c_old.pos_xyz = c_new.pos_xyz = pos;
  c_old.stamps = c_new.stamps = reinterpret_cast<const uint64_t*>(stamps);
  c_old.valid = c_new.valid = valid;
  pgmo::Schedule inc, full;
  pgmo::buildStampedSchedule(c_old, ns_old, reinterpret_cast<const uint64_t*>(s_stamp_old),
                             reinterpret_cast<const uint64_t*>(s_len_old), k, tol_t, 1.0, &inc);
  unsigned long common = 0;
  while (common < ns_old && common < ns_new && s_stamp_old[common] == s_stamp_new[common] &&
         s_len_old[common] == s_len_new[common]) {
    ++common;
  }
  pgmo::extendStampedSchedule(c_new, ns_new, reinterpret_cast<const uint64_t*>(s_stamp_new),
                              reinterpret_cast<const uint64_t*>(s_len_new), common, k, tol_t, 1.0, &inc);
  pgmo::buildStampedSchedule(c_new, ns_new, reinterpret_cast<const uint64_t*>(s_stamp_new),
                             reinterpret_cast<const uint64_t*>(s_len_new), k, tol_t, 1.0, &full);
  if (resumed) *resumed = inc.resumed_from_stretch;
  if (inc.noop != full.noop) return 1;
  if (full.noop) return 0;
  if (inc.degenerate != full.degenerate) return 2;
  if (inc.runs.size() != full.runs.size()) return 3;
  for (size_t r = 0; r < full.runs.size(); ++r) {
    const auto& a = inc.runs[r];
    const auto& b = full.runs[r];
    if (a.first_position != b.first_position || a.lo != b.lo || a.hi != b.hi || a.k != b.k ||
        a.skip != b.skip || a.depth != b.depth || std::memcmp(a.min_xyz, b.min_xyz, sizeof(a.min_xyz)) ||
        std::memcmp(a.shift, b.shift, sizeof(a.shift)) || std::memcmp(a.max_xyz, b.max_xyz, sizeof(a.max_xyz))) {
      return 100 + static_cast<long>(r);
    }
  }
  if (inc.del_run != full.del_run) return 4;
  if (inc.stable_key != full.stable_key) return 5;
  if (inc.total_positions != full.total_positions) return 6;
  if (inc.max_k != full.max_k) return 7;
  if (inc.max_leaf_points < full.max_leaf_points) return 8;  // an upper bound is carried over
  if (inc.end_touch_stretch != full.end_touch_stretch) return 9;
  if (inc.checkpoints.size() != full.checkpoints.size()) return 10;
  for (size_t s = 0; s < full.checkpoints.size(); ++s) {
    const auto& a = inc.checkpoints[s];
    const auto& b = full.checkpoints[s];
    if (a.position != b.position || a.run_count != b.run_count || a.lo != b.lo || a.hi != b.hi ||
        a.k != b.k || a.dirty != b.dirty) {
      return 11;
    }
  }
  return 0;
}
}
