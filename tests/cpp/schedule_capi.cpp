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
