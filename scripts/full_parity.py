"""Full-size parity run of BASELINE configs[2]: 10M vertices, 50k control points, k = 4, FP64 and
FP32 modes, CUDA path (through the C ABI, associations + pre-cast doubles requested) against the
CPU oracle on the same inputs.  Takes ~1 min of oracle time; prints a summary for profiles/."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402
from oracle import ffi  # noqa: E402

n_v = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_c = n_v // 200
case = synth.deformation_case(n_v, n_c)
t0 = time.perf_counter()
ref = ffi.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                        case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0)
t_ref = time.perf_counter() - t0
h = capi.DeformHandle(0)
h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0,
                      want_f64=True, want_assoc=True)
ga = gpu["assoc"].astype(np.int64)
ga[ga == 0xFFFFFFFF] = -1
assoc_bad = int((ga != ref["assoc"].astype(np.int64)).any(axis=1).sum())
f64_bad = int((gpu["f64"] != ref["f64"]).any(axis=1).sum())
f32_bad = sum(int((gpu[a] != ref[a]).sum()) for a in "xyz")
g32 = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0, precision=1)
err32 = max(float(np.abs(g32[a].astype(np.float64) - ref[a].astype(np.float64)).max()) for a in "xyz")
print(f"vertices {n_v} controls {n_c} k=4 tol_t=10 s; oracle {t_ref:.1f} s on 1 core")
print(f"FP64 mode: association rows differing {assoc_bad}, pre-cast double rows differing {f64_bad}, "
      f"output floats differing {f32_bad}")
print(f"FP32 mode: max |position error| vs reference {err32:.3e} m (bar 1e-4 m)")
sys.exit(0 if assoc_bad == 0 and f64_bad == 0 and f32_bad == 0 and err32 <= 1e-4 else 1)
