"""Timeline of one pipelined host-pointer step (PGMO_PIPE_TRACE=1) for a few slab layouts."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PGMO_PIPE_TRACE"] = "1"
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n_v, n_c = 10_000_000, 50_000
case = synth.deformation_case(n_v, n_c)
pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
h_in = {k: pin(case[k]) for k in "xyz"}
h_stamps = pin(case["stamps"].view(np.int64))
h_out = [torch.empty(n_v, dtype=torch.float32).pin_memory() for _ in range(3)]
h_poses = pin(case["poses"])
for pipe, taper in [(8, 0), (12, 2)]:
    os.environ["PGMO_PIPE"], os.environ["PGMO_PIPE_TAPER"] = str(pipe), str(taper)
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    capi._check(h._lib.pgmo_deform_points(
        h._h, n_v, capi._ptr(h_in["x"]), capi._ptr(h_in["y"]), capi._ptr(h_in["z"]),
        capi._ptr(h_stamps), None, 0, 4, 10.0, 0, capi._ptr(h_out[0]), capi._ptr(h_out[1]),
        capi._ptr(h_out[2]), None, None, 0, None, None))
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.update_poses(h_poses.numpy())
        t1 = time.perf_counter()
        h.run_host_slab(0, n_v, h_in["x"], h_in["y"], h_in["z"], h_out, precision=0)
        t2 = time.perf_counter()
        print(f"pipe {pipe} taper {taper} step {i}: update_poses {1e3 * (t1 - t0):.3f} ms, run_host_slab {1e3 * (t2 - t1):.3f} ms",
              file=sys.stderr, flush=True)
    h.close()
