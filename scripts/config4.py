"""BASELINE configs[3]: a multi-robot mesh (robot ids 0..R-1: pose prefixes 'a'.., vertex prefixes
's'.., common_functions.h:19-61) deformed with vertex-range sharding over the GPUs of one box.

Every robot is an independent deformation problem (own control table, own window schedule); the
robots' vertices are laid end to end and cut into one contiguous slab per rank
(sharding.plan_segments), so a rank owns one or two (robot, range) segments.  Control tables and
poses come from rank 0 by NCCL broadcast; the only other exchange is each robot's stretch list.
Run with python (1 GPU) or torchrun (N GPUs):

  python scripts/config4.py [--vertices-per-robot 25000000] [--robots 4] [--check 1000000]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/config4.py

Prints one JSON line: step time (CUDA events, max over ranks), vertices/s, per-robot checksums of
the output bits (equal for every N = shard-count invariance) and the oracle comparison: for every
segment, the first `--check` vertices are compared bit for bit with the CPU oracle swept from the
robot's first vertex (the window state at a shard boundary depends on every earlier stamp)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, sharding, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vertices-per-robot", type=int, default=25_000_000)
ap.add_argument("--controls-per-robot", type=int, default=50_000)
ap.add_argument("--robots", type=int, default=4)
ap.add_argument("--check", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
flag_group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    flag_group = dist.new_group(backend="gloo")
stream = torch.cuda.current_stream().cuda_stream

R, n_v, n_c = args.robots, args.vertices_per_robot, args.controls_per_robot
K, TOL = 4, 10.0
plan = sharding.plan_segments([n_v] * R, world)
mine = plan[rank]

# ---- control tables: generated on rank 0, replicated by broadcast --------------------------
ctrl = []
for r in range(R):
    if rank == 0:
        c = synth.robot_range(r, n_v, n_c, 0, 0)
        t = [torch.from_numpy(c["ctrl_xyz"]), torch.from_numpy(c["ctrl_stamps"].view(np.int64)),
             torch.from_numpy(c["valid"]), torch.from_numpy(c["poses"])]
    else:
        t = [torch.zeros((n_c, 3), dtype=torch.float64), torch.zeros(n_c, dtype=torch.int64),
             torch.zeros(n_c, dtype=torch.uint8), torch.zeros((n_c, 12), dtype=torch.float64)]
    t = [x.to(dev) for x in t]
    sharding.broadcast_controls(t, src=0)
    ctrl.append(t)

handles = {}
for r, _, _ in mine:
    h = capi.DeformHandle(local_rank)
    pos, st, va, po = ctrl[r]
    h.set_controls_device(n_c, pos, st, va, po, stream)
    handles[r] = h

# ---- this rank's segments -------------------------------------------------------------------
segs = []
for r, first, count in mine:
    case = synth.robot_range(r, n_v, n_c, first, count, with_controls=False)
    d = {k: torch.from_numpy(case[k]).to(dev) for k in ("x", "y", "z")}
    d["stamps"] = torch.from_numpy(case["stamps"].view(np.int64)).to(dev)
    d["out"] = [torch.empty(count, dtype=torch.float32, device=dev) for _ in range(3)]
    segs.append((r, first, count, d))
exchange = sharding.MultiRobotExchange(R, world, dev)
state = {"valid": False}
flag = torch.zeros(1, dtype=torch.int32)
empty = np.zeros(0, np.uint64)


def run_segments():
    for r, first, count, d in segs:
        handles[r].run_device(first, count, d["x"], d["y"], d["z"], *d["out"], stream=stream)


def step():
    for r in range(R):  # a new solve: every robot's poses from rank 0
        if world > 1:
            dist.broadcast(ctrl[r][3], src=0)
    for r, h in handles.items():
        h.update_poses_device(ctrl[r][3], stream)
    changed = True
    if state["valid"]:
        for r, first, count, d in segs:
            handles[r].stamp_check_begin(d["stamps"], count, stream)
            handles[r].run_device(first, count, d["x"], d["y"], d["z"], *d["out"], stream=stream)
        changed = False
        for r, _, _, _ in segs:
            changed = handles[r].stamp_check_finish() or changed
        if world > 1:
            flag[0] = int(changed)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=flag_group)
            changed = bool(flag[0])
    if changed:
        local = {r: (empty, empty) for r in range(R)}
        for r, first, count, d in segs:
            local[r] = handles[r].stamp_stretches_device(d["stamps"], count, stream)
        for r in range(R):
            gs, gl = exchange(r, *local[r])
            if r in handles:
                assert handles[r].build_schedule(gs, gl, n_v, K, TOL) == 0
        run_segments()
        state["valid"] = True


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


t_cold0 = time.perf_counter()
step()
barrier()
cold_ms = 1e3 * (time.perf_counter() - t_cold0)
for _ in range(max(args.warmup, 3)):
    step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())

# ---- shard-count invariant checksums of the output bits, per robot ------------------------
sums = torch.zeros((R, 2), dtype=torch.int64, device=dev)
for r, first, count, d in segs:
    idx = torch.arange(first + 1, first + count + 1, dtype=torch.int64, device=dev)
    for a, o in enumerate(d["out"]):
        bits = o.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        sums[r, 0] += bits.sum() * (a + 1)
        sums[r, 1] += (bits * idx).sum() * (a + 1)  # wraps mod 2^64: position-sensitive
if world > 1:
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)

# ---- oracle at every shard boundary ----------------------------------------------------------
checked = mismatched = 0
oracle_s = 0.0
if args.check > 0:
    from oracle import ffi
    for r, first, count, d in segs:
        n_chk = min(args.check, count)
        ref_case = synth.robot_range(r, n_v, n_c, 0, first + n_chk)
        ref = ffi.deform_points(ref_case["x"], ref_case["y"], ref_case["z"], ref_case["stamps"],
                                ref_case["ctrl_xyz"], ref_case["ctrl_stamps"], ref_case["valid"],
                                ref_case["poses"], k=K, tol_t=TOL, want_trace=False)
        oracle_s = max(oracle_s, ref["elapsed_s"])
        for a, ax in enumerate("xyz"):
            got = d["out"][a][:n_chk].cpu().numpy()
            mismatched += int((got.view(np.uint32) != ref[ax][first:first + n_chk].view(np.uint32)).sum())
        checked += n_chk
tot = torch.tensor([checked, mismatched], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)

if rank == 0:
    total = R * n_v
    print(json.dumps({
        "scenario": f"configs[3]: {R} robots x {n_v} vertices, {n_c} control points each, k=4, "
                    f"vertex-range sharding over {world} GPU(s)",
        "n_gpus": world, "ms_per_step": ms, "value": total / (ms * 1e-3), "unit": "vertices/s",
        "first_step_ms_schedules_built": cold_ms,
        "segments_per_rank": [[list(s) for s in p] for p in plan],
        "checksums_per_robot": {synth.ROBOT_ID_TO_VERTEX_PREFIX[r]: [int(sums[r, 0]), int(sums[r, 1])]
                                for r in range(R)},
        "oracle": {"vertices_checked_at_segment_starts": int(tot[0]), "output_floats_differing": int(tot[1]),
                   "slowest_oracle_sweep_s": oracle_s},
    }))
for h in handles.values():
    h.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if int(tot[1]) == 0 else 1)
