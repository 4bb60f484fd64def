"""CPU, world_size 2 over gloo: the host-side logic of the sharded deformation path
(kimera_pgmo_b200/sharding.py): slab bounds, stretch extraction, all_gather exchange, and that
every rank ends up with the identical window schedule as an unsharded run."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from kimera_pgmo_b200 import sharding, synth

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_total, n_ctrl = 60_000, 600
first, count = sharding.slab_bounds(n_total, world, rank)
case = synth.deformation_slab(n_total, n_ctrl, first, count)

# control table: only rank 0 has it, everybody else receives it
ctrl = [torch.from_numpy(case["ctrl_xyz"].copy()), torch.from_numpy(case["ctrl_stamps"].view(np.int64).copy()),
        torch.from_numpy(case["valid"].copy()), torch.from_numpy(case["poses"].copy())]
if rank != 0:
    for t in ctrl:
        t.zero_()
sharding.broadcast_controls(ctrl, src=0)
assert np.array_equal(ctrl[0].numpy(), case["ctrl_xyz"])
assert np.array_equal(ctrl[3].numpy(), case["poses"])

ss, sl = sharding.stretches_of(case["stamps"])
ex = sharding.StretchExchange(world, torch.device("cpu"))
gs, gl = ex(ss, sl)
assert int(gl.sum()) == n_total

lib = C.CDLL(sys.argv[2])
lib.schedule_build.restype = C.c_long
lib.schedule_build.argtypes = [C.c_ulong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong,
                               C.c_void_p, C.c_void_p, C.c_ulong, C.c_uint, C.c_double,
                               C.c_int, C.c_ulong, C.c_void_p, C.c_void_p]
p = lambda a: a.ctypes.data_as(C.c_void_p)
pos = np.ascontiguousarray(ctrl[0].numpy()); cst = np.ascontiguousarray(ctrl[1].numpy().view(np.uint64))
valid = np.ascontiguousarray(ctrl[2].numpy())
cap = 2 * gs.shape[0] + 2 * n_ctrl + 16
runs = np.zeros((cap, 6), dtype=np.uint64); del_run = np.zeros(n_ctrl, dtype=np.int32)
n = lib.schedule_build(n_ctrl, p(pos), p(cst), p(valid), gs.shape[0], p(gs), p(gl), n_total, 4, 10.0,
                       1, cap, p(runs), p(del_run))
np.savez(os.path.join(sys.argv[3], f"rank{rank}.npz"), gs=gs, gl=gl, runs=runs[:n], del_run=del_run,
         first=first, count=count)
dist.barrier()
dist.destroy_process_group()
'''


def test_slab_bounds_cover():
    from kimera_pgmo_b200 import sharding
    for n, w in ((10, 3), (100_000_000, 8), (7, 8)):
        spans = [sharding.slab_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1


def test_merge_fuses_boundary_stretch():
    from kimera_pgmo_b200 import sharding
    stamps = np.repeat(np.arange(5, dtype=np.uint64), 7)
    whole = sharding.stretches_of(stamps)
    for cuts in ([10], [7, 14], [1, 2, 3, 30]):
        parts = np.split(stamps, cuts)
        merged = sharding.merge_stretch_lists([sharding.stretches_of(p) for p in parts])
        np.testing.assert_array_equal(merged[0], whole[0])
        np.testing.assert_array_equal(merged[1], whole[1])


@pytest.mark.timeout(300)
def test_two_ranks_build_identical_schedule(tmp_path):
    from kimera_pgmo_b200 import sharding, synth
    build = os.path.join(ROOT, "build", "tests")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libschedule_test.so")
    srcs = [os.path.join(ROOT, "tests", "cpp", "schedule_capi.cpp"),
            os.path.join(ROOT, "kimera_pgmo_b200", "csrc", "deform_schedule.cpp")]
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I",
                    os.path.join(ROOT, "kimera_pgmo_b200", "csrc"), "-o", so] + srcs, check=True)
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                        "29653", str(worker), ROOT, so, str(tmp_path)], capture_output=True,
                       text=True, env=env, timeout=280)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    for key in ("gs", "gl", "runs", "del_run"):
        np.testing.assert_array_equal(r0[key], r1[key])
    assert int(r0["first"]) == 0 and int(r1["first"]) == int(r0["count"])
    # and identical to the unsharded stretch list
    whole = synth.deformation_slab(60_000, 600, 0, 60_000)
    ws, wl = sharding.stretches_of(whole["stamps"])
    np.testing.assert_array_equal(r0["gs"], ws)
    np.testing.assert_array_equal(r0["gl"], wl)
    assert r0["runs"].shape[0] > 0


def test_plan_segments_cover_every_robot_once():
    from kimera_pgmo_b200 import sharding
    for counts, world in (([25_000_000] * 4, 8), ([25_000_000] * 4, 3), ([10, 30, 5], 4), ([7], 8),
                          ([0, 5, 0, 9], 2)):
        plan = sharding.plan_segments(counts, world)
        assert len(plan) == world
        owned = [np.zeros(c, dtype=np.int32) for c in counts]
        sizes = []
        for segs in plan:
            sizes.append(sum(c for _, _, c in segs))
            for robot, first, count in segs:
                assert count > 0
                owned[robot][first:first + count] += 1
        assert all((o == 1).all() for o in owned)
        assert max(sizes) - min(sizes) <= 1  # equal slabs over the concatenated robots


def test_robot_readers_are_dense_and_cover_every_segment():
    """readers of a robot's pose source = the ranks owning a segment of the robot, rank order"""
    from kimera_pgmo_b200 import sharding
    for world in (1, 2, 3, 4, 8):
        plan = sharding.plan_segments([25_000_000] * 4, world)
        readers = sharding.robot_readers(plan, 4)
        for r in range(4):
            assert readers[r] == sorted(readers[r]) and len(set(readers[r])) == len(readers[r])
            assert readers[r], "every robot is deformed by somebody"
            for rank, segs in enumerate(plan):
                assert (rank in readers[r]) == any(seg[0] == r for seg in segs)
    assert sharding.robot_readers(sharding.plan_segments([10, 10, 10, 10], 8), 4) == {
        0: [0, 1], 1: [2, 3], 2: [4, 5], 3: [6, 7]}


def test_range_generator_is_partition_invariant():
    from kimera_pgmo_b200 import synth
    n, n_c = 2_300_000, 2300
    whole = synth.robot_range(1, n, n_c, 0, n)
    for first, count in ((0, 17), (1_048_570, 20), (2_000_000, 300_000), (123_456, 1_500_000)):
        part = synth.robot_range(1, n, n_c, first, count, with_controls=False)
        for key in ("x", "y", "z", "stamps"):
            np.testing.assert_array_equal(whole[key][first:first + count], part[key])
    other = synth.robot_range(2, n, n_c, 0, 1000)
    assert not np.array_equal(other["x"], whole["x"][:1000])
    assert other["ctrl_stamps"][0] != whole["ctrl_stamps"][0]
    # gtsam::Symbol keys of the reference's test data (test/data/graph.dgrf:1,4)
    assert synth.symbol_key("a", 0) == 6989586621679009792
    assert synth.symbol_key("v", 0) == 8502796096475496448
    assert synth.ROBOT_ID_TO_VERTEX_PREFIX[3] == "v" and synth.ROBOT_ID_TO_PREFIX[7] == "h"


WORKER_ROBOTS = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from kimera_pgmo_b200 import sharding, synth

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
R, n_v, n_c = 3, 40_000, 400
plan = sharding.plan_segments([n_v] * R, world)
ex = sharding.MultiRobotExchange(R, world, torch.device("cpu"))
empty = np.zeros(0, np.uint64)
for attempt in range(2):  # second pass: nothing changed, cached lists come back
    local = {r: (empty, empty) for r in range(R)}
    for r, first, count in plan[rank]:
        local[r] = sharding.stretches_of(synth.robot_range(r, n_v, n_c, first, count, with_controls=False)["stamps"])
    for r in range(R):
        gs, gl = ex(r, *local[r])
        ws, wl = sharding.stretches_of(synth.robot_range(r, n_v, n_c, 0, n_v, with_controls=False)["stamps"])
        assert np.array_equal(gs, ws) and np.array_equal(gl, wl), (rank, r, attempt)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.timeout(300)
def test_two_ranks_three_robots_exchange(tmp_path):
    worker = tmp_path / "worker_robots.py"
    worker.write_text(WORKER_ROBOTS)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                        "29654", str(worker), ROOT], capture_output=True, text=True, env=env,
                       timeout=280)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


POSE_WORKER = r'''
import json, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from kimera_pgmo_b200 import sharding

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()


class RecordingHandle:
    """stands in for capi.DeformHandle: records what the pose-exchange protocol asks of the C ABI"""
    def __init__(self):
        self.calls = []
    def pose_source_create(self, n_c):
        self.calls.append(("create", n_c))
        return np.arange(64, dtype=np.uint8), 0
    def pose_source_attach(self, n_c, ipc_handle=None, d_base=None, owner_device=0):
        self.calls.append(("attach", n_c, [int(b) for b in ipc_handle[:4]]))
    def pose_source_publish(self, seq, n_readers, poses, stream=0):
        self.calls.append(("publish", seq, n_readers))
    def update_poses_from_source_seq(self, seq, reader, row_begin, row_end, stream=0):
        self.calls.append(("pull", seq, reader, row_begin, row_end))
    def pose_source_write(self, buffer, poses, stream=0):
        self.calls.append(("write", buffer))
    def update_poses_from_source(self, buffer, row_begin, row_end, stream=0):
        self.calls.append(("pull_buffer", buffer, row_begin, row_end))


h = RecordingHandle()
src = sharding.PoseSource(h, 1000, rank, world, torch.device("cpu"))
poses = np.zeros((1000, 12))
lo, hi = 100 * rank, 600 + 100 * rank
for _ in range(3):
    src.publish_and_pull(poses, lo, hi)
src.pull_again(0, 1000)
# the caller-ordered variant keeps working next to it
src.write(poses)
src.pull(lo, hi)
json.dump(h.calls, open(os.path.join(sys.argv[2], f"pose_rank{rank}.json"), "w"))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_pose_exchange_protocol(tmp_path):
    """sharding.PoseSource on two gloo ranks with a recording handle: the owner creates the source
    and publishes sequence numbers 1, 2, 3 for `world` readers, every rank attaches with the
    owner's IPC handle and pulls its rows as reader `rank` with the same sequence numbers
    (pgmo_deform_pose_source_publish / pgmo_deform_update_poses_from_source_seq)."""
    import json
    worker = tmp_path / "pose_worker.py"
    worker.write_text(POSE_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                        "29655", str(worker), ROOT, str(tmp_path)], capture_output=True, text=True,
                       env=env, timeout=280)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    c0 = json.load(open(tmp_path / "pose_rank0.json"))
    c1 = json.load(open(tmp_path / "pose_rank1.json"))
    assert c0[0] == ["create", 1000]
    assert c1[0] == ["attach", 1000, [0, 1, 2, 3]]           # the owner's handle bytes arrived
    assert [c for c in c0 if c[0] == "publish"] == [["publish", s, 2] for s in (1, 2, 3)]
    assert not [c for c in c1 if c[0] == "publish"]            # only the owner writes
    assert [c for c in c0 if c[0] == "pull"] == [["pull", s, 0, 0, 600] for s in (1, 2, 3)] + [["pull", 3, 0, 0, 1000]]
    assert [c for c in c1 if c[0] == "pull"] == [["pull", s, 1, 100, 700] for s in (1, 2, 3)] + [["pull", 3, 1, 0, 1000]]
    assert c0[-2:] == [["write", 0], ["pull_buffer", 0, 0, 600]]
    assert c1[-1] == ["pull_buffer", 0, 100, 700]
