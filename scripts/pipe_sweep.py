"""End-to-end pose-only step (host buffers in / out, configs[2]) against the slab count and the
taper of the pipelined host entry point (PGMO_PIPE, PGMO_PIPE_TAPER) and with the kernels working
straight on the page-locked host buffers (PGMO_HOST_DIRECT).
    python scripts/pipe_sweep.py [n_vertices] [n_controls]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n_v = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_c = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
case = synth.deformation_case(n_v, n_c)
pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
h_in = {k: pin(case[k]) for k in "xyz"}
# the same buffers as planes of one page-locked allocation each (3-row DMA operations)
h_in3 = torch.empty((3, n_v), dtype=torch.float32).pin_memory()
for i, k in enumerate("xyz"):
    h_in3[i].copy_(h_in[k])
h_out3 = torch.empty((3, n_v), dtype=torch.float32).pin_memory()
n_pad = (n_v + 1023) // 1024 * 1024
h_in3a = torch.empty((3, n_pad), dtype=torch.float32).pin_memory()
for i, k in enumerate("xyz"):
    h_in3a[i, :n_v].copy_(h_in[k])
h_out3a = torch.empty((3, n_pad), dtype=torch.float32).pin_memory()
h_stamps = pin(case["stamps"].view(np.int64))
h_out = [torch.empty(n_v, dtype=torch.float32).pin_memory() for _ in range(3)]
h_poses = pin(case["poses"])
combos = [(8, 0, 0, 0), (8, 0, 0, 1), (8, 0, 0, 2), (8, 0, 0, 3), (8, 0, 0, 4)]
for rep in range(2):
    for pipe, taper, direct, planes in combos:
        os.environ["PGMO_PIPE"], os.environ["PGMO_PIPE_TAPER"] = str(pipe), str(taper)
        os.environ["PGMO_HOST_DIRECT"] = str(direct)
        os.environ["PGMO_COPY_2D"] = "0" if planes in (2, 3) else "1"
        h = capi.DeformHandle(0)
        h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
        capi._check(h._lib.pgmo_deform_points(
            h._h, n_v, capi._ptr(h_in["x"]), capi._ptr(h_in["y"]), capi._ptr(h_in["z"]),
            capi._ptr(h_stamps), None, 0, 4, 10.0, 0, capi._ptr(h_out[0]), capi._ptr(h_out[1]),
            capi._ptr(h_out[2]), None, None, 0, None, None))
        ref = [o.clone() for o in h_out]

        if planes in (3, 4):
            xin = [h_in3a[0, :n_v], h_in3a[1, :n_v], h_in3a[2, :n_v]]
            xout = [h_out3a[0, :n_v], h_out3a[1, :n_v], h_out3a[2, :n_v]]
        elif planes:
            xin, xout = [h_in3[0], h_in3[1], h_in3[2]], [h_out3[0], h_out3[1], h_out3[2]]
        else:
            xin, xout = [h_in["x"], h_in["y"], h_in["z"]], h_out

        def step():
            h.update_poses(h_poses.numpy())
            h.run_host_slab(0, n_v, xin[0], xin[1], xin[2], xout, precision=0)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / 10
        ok = all(torch.equal(a, b) for a, b in zip(ref, xout))
        print(f"pipe {pipe:2d} taper {taper} direct {direct} planes-of-one-allocation {planes} (1: planes + 2D copies, 2: planes + plain copies, 3: 4 KB-aligned planes + plain copies, 4: aligned planes + 2D copies): {ms:.3f} ms per step, outputs equal to the first deformMesh: {ok}", flush=True)
        h.close()
