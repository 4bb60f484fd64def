"""GPU parity of the pose-only re-deformation path (association cache + blend-only kernel).

A second deformMesh on the same handle after pgmo_deform_update_poses runs the blend kernel
alone (the search result is cached per vertex; see include/pgmo_b200.h,
pgmo_deform_set_assoc_cache).  Bar: bit-identical to the oracle evaluated with the NEW poses
(mesh_deformation.cpp:73-106) — doubles before the cast, floats after it, associations — and
the records must be dropped, item by item, as soon as a vertex changes.
"""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4  # metres, north_star


def _new_poses(case, seed):
    """another solve: different smooth correction + per-node noise"""
    rng = np.random.default_rng(seed)
    n = case["poses"].shape[0]
    ramp = np.arange(n, dtype=np.float64) / max(n - 1, 1)
    theta = np.deg2rad(-3.0) * ramp + rng.normal(0, 1e-3, n)
    phi = np.deg2rad(2.0) * ramp + rng.normal(0, 1e-3, n)
    poses = np.empty_like(case["poses"])
    poses[:, :9] = synth._rot_zy(theta, phi)
    poses[:, 9:] = case["ctrl_xyz"] + np.stack([-0.4 * ramp, 0.3 * ramp, 0.1 * ramp], axis=1) \
        + rng.normal(0, 1e-2, (n, 3))
    return poses


def _oracle(oracle, case, poses, k, tol_t, subset=None, stamped=True, stride=8):
    return oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"] if stamped else None,
                                case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], poses, k=k,
                                tol_t=tol_t, subset=subset, assoc_stride=stride)


def _gpu(h, case, k, tol_t, precision=0, subset=None, stamped=True, stride=8):
    return h.deform_points(case["x"], case["y"], case["z"], case["stamps"] if stamped else None, k=k,
                           tol_t=tol_t, precision=precision, subset=subset, want_f64=True,
                           want_assoc=True, assoc_stride=stride)


def _same(gpu, ref, exact=True):
    assert gpu["rc"] == ref["rc"] == 0
    np.testing.assert_array_equal(gpu["processed"], ref["processed"])
    done = ref["processed"].astype(bool)
    ga = gpu["assoc"].astype(np.int64)
    ga[ga == 0xFFFFFFFF] = -1
    np.testing.assert_array_equal(ga[done], ref["assoc"][done].astype(np.int64))
    np.testing.assert_array_equal(gpu["assoc_cnt"][done].astype(np.int32), ref["assoc_cnt"][done])
    if exact:
        np.testing.assert_array_equal(gpu["f64"][done], ref["f64"][done])
        for ax in "xyz":
            np.testing.assert_array_equal(gpu[ax], ref[ax])
    else:
        for ax in "xyz":
            np.testing.assert_allclose(gpu[ax].astype(np.float64), ref[ax].astype(np.float64), rtol=0,
                                       atol=FP32_TOL)


def _two_solves(capi, oracle, case, k, tol_t, precision=0, subset=None, stamped=True, stride=8,
                expect_warm=True):
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"] if stamped else None, case["valid"], case["poses"])
    first = _gpu(h, case, k, tol_t, precision, subset, stamped, stride)
    _same(first, _oracle(oracle, case, case["poses"], k, tol_t, subset, stamped, stride), precision == 0)
    assert not h.last_step_info()["warm"]
    for seed in (1, 2):
        poses = _new_poses(case, seed)
        h.update_poses(poses)
        again = _gpu(h, case, k, tol_t, precision, subset, stamped, stride)
        info = h.last_step_info()
        if expect_warm:
            assert info["warm"] and info["stale_items"] == 0, info
        elif info["warm"] and k + 1 <= 14:
            assert info["stale_items"] == 0, info
        _same(again, _oracle(oracle, case, poses, k, tol_t, subset, stamped, stride), precision == 0)
    h.close()


@pytest.mark.parametrize("invalid", [0.0, 0.01])
def test_config_a_second_solve(capi, oracle, invalid):
    case = synth.deformation_case(200_000, 2_000, invalid_fraction=invalid)
    _two_solves(capi, oracle, case, 4, 10.0)


def test_fp32_second_solve(capi, oracle):
    case = synth.deformation_case(100_000, 1_000)
    _two_solves(capi, oracle, case, 4, 10.0, precision=1)


@pytest.mark.parametrize("k", [1, 2, 3, 6, 9, 12])
def test_k_sweep_second_solve(capi, oracle, k):
    case = synth.deformation_case(20_000, 500, seed=3)
    _two_solves(capi, oracle, case, k, 10.0, stride=14)


def test_large_k_is_not_cached_but_right(capi, oracle):
    # k + 1 > 14 neighbours: no records (the resolver's blend is final), every call searches
    case = synth.deformation_case(5_000, 400, seed=4)
    _two_solves(capi, oracle, case, 15, 10.0, stride=18, expect_warm=False)


def test_unstamped_and_wrap_second_solve(capi, oracle):
    _two_solves(capi, oracle, synth.deformation_case(20_000, 300, seed=5), 4, 10.0, stamped=False)
    _two_solves(capi, oracle, synth.deformation_case(50_000, 1_000, t0_ns=0), 4, 10.0)


def test_subset_second_solve(capi, oracle):
    case = synth.deformation_case(30_000, 600, seed=9)
    rng = np.random.default_rng(1)
    subset = np.sort(rng.choice(30_000, 7_000, replace=False)).astype(np.uint64)
    subset = subset[subset >= 12_000]
    _two_solves(capi, oracle, case, 4, 10.0, subset=subset)


@pytest.mark.parametrize("ppv", [1, 3])
def test_ties_second_solve(capi, oracle, ppv):
    # every vertex goes through the tie resolver: its records feed the blend kernel too
    case = synth.lattice_case(n_side=12, points_per_voxel=ppv)
    _two_solves(capi, oracle, case, 4, 6.0, stride=12, expect_warm=False)  # tiny windows: brute force
    # 144 / 432 control points in one window: the list kernel finds ties for (almost) every vertex
    case = synth.lattice_case(n_side=24, points_per_voxel=ppv)
    _two_solves(capi, oracle, case, 2, 100.0, stamped=False)
    _two_solves(capi, oracle, case, 4, 100.0, stamped=False)
    _two_solves(capi, oracle, case, 6, 100.0, stamped=False, stride=12)


def test_k_shrink_and_skip_second_solve(capi, oracle):
    case = synth.deformation_case(3_000, 30, ctrl_per_kf=1, seed=2)
    case["valid"][1:4] = 0
    _two_solves(capi, oracle, case, 4, 0.4, expect_warm=False)  # tiny windows: brute force


@pytest.mark.parametrize("n_ctrl", [7, 127, 256, 257, 512, 513])
def test_alive_set_sizes_second_solve(capi, oracle, n_ctrl):
    case = synth.deformation_case(6_000, n_ctrl, ctrl_per_kf=n_ctrl, seed=100 + n_ctrl, ctrl_spread=9.0,
                                  vertex_spread=11.0)
    # > 512 alive points: the grid-walk kernel's items are never cached (they blend by themselves)
    # <= 48: brute force over the window, no records either
    _two_solves(capi, oracle, case, 4, 1.0e6, expect_warm=48 < n_ctrl <= 512)


def test_moved_vertices_invalidate_their_items_only(capi, oracle):
    """Same stamps, same control table, a few vertices moved: the blend kernel's fingerprint
    check rejects exactly the work items that hold them; the result is the oracle's."""
    case = synth.deformation_case(120_000, 1_200, seed=21)
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    _gpu(h, case, 4, 10.0)
    moved = dict(case)
    for ax in "xyz":
        moved[ax] = case[ax].copy()
    moved["x"][[5, 60_000, 60_001]] += np.float32(0.5)
    moved["z"][119_999] = np.float32(2.5)
    poses = _new_poses(case, 7)
    h.update_poses(poses)
    again = _gpu(h, moved, 4, 10.0)
    info = h.last_step_info()
    assert info["warm"] and 1 <= info["stale_items"] <= 4, info
    _same(again, _oracle(oracle, moved, poses, 4, 10.0))
    # the records of the redone items are valid again
    h.update_poses(case["poses"])
    third = _gpu(h, moved, 4, 10.0)
    info = h.last_step_info()
    assert info["warm"] and info["stale_items"] == 0, info
    _same(third, _oracle(oracle, moved, case["poses"], 4, 10.0))
    # swapping two vertices of one item keeps the multiset of coordinates: must still be noticed
    swapped = dict(moved)
    for ax in "xyz":
        swapped[ax] = moved[ax].copy()
        swapped[ax][[10, 11]] = moved[ax][[11, 10]]
    fourth = _gpu(h, swapped, 4, 10.0)
    assert h.last_step_info()["stale_items"] >= 1
    _same(fourth, _oracle(oracle, swapped, case["poses"], 4, 10.0))
    h.close()


def test_cache_off_and_control_change(capi, oracle):
    case = synth.deformation_case(60_000, 1_200, seed=31, invalid_fraction=0.02)
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    _gpu(h, case, 4, 10.0)
    h.set_assoc_cache(False)
    poses = _new_poses(case, 3)
    h.update_poses(poses)
    off = _gpu(h, case, 4, 10.0)
    assert not h.last_step_info()["warm"]
    _same(off, _oracle(oracle, case, poses, 4, 10.0))
    h.set_assoc_cache(True)
    # a control point moved: new schedule, the records are dropped
    ctrl = case["ctrl_xyz"].copy()
    ctrl[7, 1] += 0.25
    h.set_controls(ctrl, case["ctrl_stamps"], case["valid"], poses)
    changed = dict(case)
    changed["ctrl_xyz"] = ctrl
    got = _gpu(h, case, 4, 10.0)
    assert not h.last_step_info()["warm"]
    _same(got, _oracle(oracle, changed, poses, 4, 10.0))
    # a validity flag flipped
    valid = case["valid"].copy()
    valid[100] = 0
    h.set_controls(ctrl, case["ctrl_stamps"], valid, poses)
    changed["valid"] = valid
    got = _gpu(h, case, 4, 10.0)
    assert not h.last_step_info()["warm"]
    _same(got, _oracle(oracle, changed, poses, 4, 10.0))
    # precision switch: the weights are stored in the blend's precision
    got = _gpu(h, case, 4, 10.0, precision=1)
    assert not h.last_step_info()["warm"]
    _same(got, _oracle(oracle, changed, poses, 4, 10.0), exact=False)
    h.close()


def test_pipelined_host_entry_second_solve(capi, oracle):
    """>= 2^18 vertices: pgmo_deform_points runs slab by slab (H2D / kernels / D2H overlapped);
    every slab of the second call starts from the records of the first."""
    case = synth.deformation_case(600_000, 3_000, seed=12)
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    first = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0,
                               want_trace=False)
    for ax in "xyz":
        np.testing.assert_array_equal(first[ax], ref[ax])
    poses = _new_poses(case, 5)
    h.update_poses(poses)
    again = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    info = h.last_step_info()
    assert info["warm"] and info["stale_items"] == 0, info
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], poses, k=4, tol_t=10.0,
                               want_trace=False)
    for ax in "xyz":
        np.testing.assert_array_equal(again[ax], ref[ax])
    h.close()
