"""Short deterministic kernel sequence for ncu (one GPU): configs[2], FP64 mode.
  2 steps that search (K3a deformSearchSortedKernel + K3b deformBlendCachedKernel<.,0>), then
  3 pose-only steps (K3b<.,1> validating + the near-empty stale-list launches).
ncu: -k regex:'deformBlendCachedKernel|deformSearchSortedKernel'
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vertices", type=int, default=10_000_000)
ap.add_argument("--controls", type=int, default=50_000)
ap.add_argument("--precision", default="fp64")
ap.add_argument("--search-steps", type=int, default=2)
ap.add_argument("--warm-steps", type=int, default=3)
args = ap.parse_args()

dev = torch.device("cuda", 0)
case = synth.deformation_slab(args.vertices, args.controls, 0, args.vertices)
d = {k: torch.from_numpy(case[k]).to(dev) for k in ("x", "y", "z")}
d_stamps = torch.from_numpy(case["stamps"].view(np.int64)).to(dev)
out = [torch.empty(args.vertices, dtype=torch.float32, device=dev) for _ in range(3)]
d_poses = torch.from_numpy(case["poses"]).to(dev)
stream = torch.cuda.current_stream().cuda_stream
h = capi.DeformHandle(0)
h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
h.set_profiling(True)
prec = capi.FP64 if args.precision == "fp64" else capi.FP32
ss, sl = h.stamp_stretches_device(d_stamps, args.vertices, stream)
assert h.build_schedule(ss, sl, args.vertices, 4, 10.0) == 0
h.set_assoc_cache(False)
for i in range(args.search_steps):
    h.update_poses_device(d_poses, stream)
    h.run_device(0, args.vertices, d["x"], d["y"], d["z"], out[0], out[1], out[2], precision=prec, stream=stream)
    print("search step", h.last_step_info())
h.set_assoc_cache(True)
for i in range(args.warm_steps):
    h.update_poses_device(d_poses, stream)
    h.run_device(0, args.vertices, d["x"], d["y"], d["z"], out[0], out[1], out[2], precision=prec, stream=stream)
    print("pose-only step", h.last_step_info())
torch.cuda.synchronize()
