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
