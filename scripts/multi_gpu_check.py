"""Correctness of the sharded path on N GPUs of one box (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_check.py
Every rank deforms its vertex slab with poses pulled from rank 0's pose source (CUDA IPC over
NVLink) — through the device entry point and through the host-pointer slab entry point — for two
different solves, and compares its slab bit for bit with the CPU oracle's sweep of the whole mesh.
Exit code 0 iff every rank matches."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, sharding, synth  # noqa: E402
from oracle import ffi  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
torch.cuda.set_stream(torch.cuda.Stream(dev))  # one stream for torch, NCCL waits and the library
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream().cuda_stream

n_total, n_ctrl = 400_000, 4_000
first, n_local = sharding.slab_bounds(n_total, world, rank)
case = synth.deformation_case(n_total, n_ctrl, seed=77, invalid_fraction=0.01)
rng = np.random.default_rng(5)
solves = [case["poses"]]
p2 = case["poses"].copy()
p2[:, 9:] += rng.normal(0, 0.05, p2[:, 9:].shape)
solves.append(p2)

h = capi.DeformHandle(local_rank)
# every rank has position / stamp / validity (replicated once); poses only on rank 0
zeros = np.zeros_like(case["poses"])
h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"] if rank == 0 else zeros)
src = sharding.PoseSource(h, n_ctrl, rank, world, dev)
sync = torch.zeros(1, dtype=torch.int32, device=dev)
d = {k: torch.from_numpy(case[k][first:first + n_local]).to(dev) for k in ("x", "y", "z")}
d_stamps = torch.from_numpy(case["stamps"][first:first + n_local].view(np.int64)).to(dev)
out = [torch.zeros(n_local, dtype=torch.float32, device=dev) for _ in range(3)]
ss, sl = h.stamp_stretches_device(d_stamps, n_local, stream)
ss, sl = sharding.StretchExchange(world, dev)(ss, sl)
assert h.build_schedule(ss, sl, n_total, 4, 10.0) == 0
lo, hi = h.slab_control_range(first, n_local)
bad = 0
for it, poses in enumerate(solves + solves):
    d_poses = torch.from_numpy(poses).to(dev) if rank == 0 else None
    src.write(d_poses, stream)
    if world > 1:
        dist.all_reduce(sync)
    src.pull(lo, hi, stream)
    ref = ffi.deform_points(case["x"][:first + n_local], case["y"][:first + n_local], case["z"][:first + n_local],
                            case["stamps"][:first + n_local], case["ctrl_xyz"], case["ctrl_stamps"],
                            case["valid"], poses, k=4, tol_t=10.0, want_trace=False)
    if it % 2 == 0:
        h.run_device(first, n_local, d["x"], d["y"], d["z"], out[0], out[1], out[2], stream=stream)
        torch.cuda.synchronize()
        got = [o.cpu().numpy() for o in out]
        how = "run_device"
    else:
        hin = [torch.from_numpy(case[k][first:first + n_local].copy()).pin_memory() for k in ("x", "y", "z")]
        hout = [torch.zeros(n_local, dtype=torch.float32).pin_memory() for _ in range(3)]
        h.run_host_slab(first, n_local, hin[0], hin[1], hin[2], hout)
        got = [o.numpy() for o in hout]
        how = "run_host_slab"
    diff = sum(int((g.view(np.uint32) != ref[ax][first:].view(np.uint32)).sum()) for g, ax in zip(got, "xyz"))
    info = h.last_step_info()
    print(f"rank {rank} solve {it} via {how}: rows pulled [{lo}, {hi}) of {n_ctrl}, warm={info['warm']}, "
          f"{diff} of {3 * n_local} floats differ", flush=True)
    bad += diff
t = torch.tensor([bad], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 0 else 1)
