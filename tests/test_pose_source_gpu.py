"""GPU tests of the pose source (include/pgmo_b200.h: pgmo_deform_pose_source_*): the replication of
a solve result by peer pull that replaces DeformationGraph's "every process holds the optimised
values" for the sharded path.  One device is enough for the protocol: an owner handle
and a second handle attached to the owner's memory through the same-process path (d_base).

Bar: a handle whose pose tables were pulled from the source deforms bit-identically to the oracle
with those poses (mesh_deformation.cpp:73-106) — with the caller ordering writer and readers, and
with the sequence numbers kept in the source's own memory — for partial row ranges too; waits
that cannot be satisfied give up and are counted instead of hanging the device.
"""
import time

import numpy as np
import pytest

from kimera_pgmo_b200 import synth
from test_warm_gpu import _new_poses

pytestmark = pytest.mark.gpu

N_V, N_C = 120_000, 1_200


def _pair(capi, case):
    owner, reader = capi.DeformHandle(0), capi.DeformHandle(0)
    for h in (owner, reader):
        h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    _, base = owner.pose_source_create(N_C)
    reader.pose_source_attach(N_C, d_base=base, owner_device=0)
    return owner, reader


def _check(oracle, h, case, poses):
    got = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], poses, k=4, tol_t=10.0)
    assert got["rc"] == ref["rc"] == 0
    for ax in "xyz":
        np.testing.assert_array_equal(got[ax], ref[ax])


def test_caller_ordered_write_and_pull(capi, oracle):
    torch = pytest.importorskip("torch")
    case = synth.deformation_case(N_V, N_C)
    owner, reader = _pair(capi, case)
    for step in range(3):
        poses = _new_poses(case, 10 + step)
        owner.pose_source_write(step % 2, poses)
        torch.cuda.synchronize()  # the caller's ordering
        for h in (owner, reader):
            h.update_poses_from_source(step % 2, 0, N_C)
        _check(oracle, reader, case, poses)
    _check(oracle, owner, case, poses)
    owner.close()
    reader.close()


def test_sequenced_publish_and_pull(capi, oracle):
    case = synth.deformation_case(N_V, N_C)
    owner, reader = _pair(capi, case)
    poses = None
    for seq in range(1, 6):
        poses = _new_poses(case, 20 + seq)
        # no synchronisation of any kind between the three calls: the flags order them
        owner.pose_source_publish(seq, 2, poses)
        reader.update_poses_from_source_seq(seq, 1, 0, N_C)
        owner.update_poses_from_source_seq(seq, 0, 0, N_C)
        if seq in (1, 4, 5):
            _check(oracle, reader, case, poses)
    _check(oracle, owner, case, poses)
    assert owner.pose_source_timeouts() == 0 and reader.pose_source_timeouts() == 0
    owner.close()
    reader.close()


def test_sequenced_partial_rows(capi, oracle):
    """a slab pulls only the rows its windows reference (pgmo_deform_slab_control_range)"""
    case = synth.deformation_case(N_V, N_C)
    owner, reader = _pair(capi, case)
    first = reader.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    assert first["rc"] == 0
    lo, hi = reader.slab_control_range(0, N_V)
    assert 0 <= lo < hi <= N_C
    poses = _new_poses(case, 31)
    owner.pose_source_publish(1, 2, poses)
    reader.update_poses_from_source_seq(1, 1, lo, hi)
    owner.update_poses_from_source_seq(1, 0, 0, N_C)
    _check(oracle, reader, case, poses)
    owner.close()
    reader.close()


def test_unsatisfiable_wait_gives_up(capi):
    case = synth.deformation_case(20_000, 200)
    owner = capi.DeformHandle(0)
    owner.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    owner.pose_source_create(200)
    t0 = time.perf_counter()
    owner.update_poses_from_source_seq(7, 0, 0, 200)  # nobody ever publishes sequence 7
    assert owner.pose_source_timeouts() == 1
    assert time.perf_counter() - t0 < 10.0
    # a writer waiting for a reader that never acknowledges gives up too
    owner.pose_source_publish(3, 2, case["poses"])  # needs ack >= 1 from readers 0 and 1
    assert owner.pose_source_timeouts() >= 2
    owner.close()
