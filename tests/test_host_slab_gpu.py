"""GPU parity of pgmo_deform_run_host_slab (include/pgmo_b200.h): the pipelined host-pointer entry
point the end-to-end numbers go through.

After a first deformMesh (pgmo_deform_points: stamps up, schedule built, association records
written) every further solve is pgmo_deform_update_poses + pgmo_deform_run_host_slab.  Checked
against the oracle with the new poses (mesh_deformation.cpp:73-106), bit-exact floats, for
  * pageable arrays (three plain copies per chunk),
  * x / y / z and the outputs as planes of ONE page-locked allocation each (with PGMO_COPY_2D=1:
    one 3-row DMA operation per chunk),
  * PGMO_PIPE / PGMO_PIPE_TAPER slab layouts,
  * the kernels reading / writing the page-locked buffers themselves (PGMO_HOST_DIRECT = 1, 2),
  * vertices that moved since the records were written (stale items searched again).
"""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth
from test_warm_gpu import _new_poses

pytestmark = pytest.mark.gpu

N_V, N_C = 300_000, 1_500


def _first_deform(capi, case):
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    first = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    assert first["rc"] == 0
    return h, first


def _oracle_xyz(oracle, case, poses):
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], poses, k=4, tol_t=10.0)
    assert ref["rc"] == 0
    return ref


def _buffers(torch, case, planes):
    n = case["x"].shape[0]
    if planes:
        h_in = torch.empty((3, n), dtype=torch.float32).pin_memory()
        for i, k in enumerate("xyz"):
            h_in[i].copy_(torch.from_numpy(case[k]))
        h_out = torch.zeros((3, n), dtype=torch.float32).pin_memory()
        return [h_in[0], h_in[1], h_in[2]], [h_out[0], h_out[1], h_out[2]]
    return [case[k].copy() for k in "xyz"], [np.zeros(n, dtype=np.float32) for _ in range(3)]


def _as_numpy(a):
    return a.numpy() if hasattr(a, "numpy") else a


@pytest.mark.parametrize("planes", [False, True])
@pytest.mark.parametrize("env", [{}, {"PGMO_PIPE": "5", "PGMO_PIPE_TAPER": "1"}, {"PGMO_PIPE": "1"},
                                 {"PGMO_COPY_2D": "1"}, {"PGMO_HOST_DIRECT": "1"}, {"PGMO_HOST_DIRECT": "2"}])
def test_pose_only_steps_through_host_pointers(capi, oracle, monkeypatch, planes, env):
    torch = pytest.importorskip("torch")
    for k, v in env.items():
        monkeypatch.setenv(k, v)  # read at pgmo_deform_create
    case = synth.deformation_case(N_V, N_C)
    h, _ = _first_deform(capi, case)
    xin, xout = _buffers(torch, case, planes)
    for seed in (1, 2):
        poses = _new_poses(case, seed)
        h.update_poses(poses)
        processed = np.zeros(N_V, dtype=np.uint8)
        h.run_host_slab(0, N_V, xin[0], xin[1], xin[2], xout, processed=processed)
        info = h.last_step_info()
        assert info["warm"] and info["stale_items"] == 0, info
        ref = _oracle_xyz(oracle, case, poses)
        np.testing.assert_array_equal(processed, ref["processed"])
        for got, ax in zip(xout, "xyz"):
            np.testing.assert_array_equal(_as_numpy(got), ref[ax])
    h.close()


@pytest.mark.parametrize("direct", ["0", "1", "2"])
def test_moved_vertices_are_searched_again(capi, oracle, monkeypatch, direct):
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("PGMO_HOST_DIRECT", direct)
    case = synth.deformation_case(N_V, N_C)
    h, _ = _first_deform(capi, case)
    rng = np.random.default_rng(7)
    moved = rng.choice(N_V, 500, replace=False)
    for ax in "xy":
        case[ax][moved] += rng.uniform(-1.0, 1.0, moved.shape[0]).astype(np.float32)
    xin, xout = _buffers(torch, case, True)
    poses = _new_poses(case, 3)
    h.update_poses(poses)
    h.run_host_slab(0, N_V, xin[0], xin[1], xin[2], xout)
    info = h.last_step_info()
    assert info["warm"] and 0 < info["stale_items"] <= moved.shape[0], info
    ref = _oracle_xyz(oracle, case, poses)
    for got, ax in zip(xout, "xyz"):
        np.testing.assert_array_equal(_as_numpy(got), ref[ax])
    h.close()
