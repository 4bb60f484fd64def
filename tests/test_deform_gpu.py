"""GPU parity: pgmo_deform_* (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): vertex->control associations bit-exact in reference order,
FP64-mode positions within 1e-9 m (we assert bit-identical pre-cast doubles and bit-identical
floats), FP32-mode positions within 1e-4 m.
"""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-9  # metres, north_star
FP32_TOL = 1e-4  # metres, north_star


def _set(handle, case):
    handle.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])


def _compare(gpu, ref, exact_positions=True, tol=FP64_TOL):
    assert gpu["rc"] == ref["rc"]
    if ref["rc"] != 0:
        return
    np.testing.assert_array_equal(gpu["processed"], ref["processed"])
    done = ref["processed"].astype(bool)
    np.testing.assert_array_equal(gpu["assoc_cnt"][done].astype(np.int32), ref["assoc_cnt"][done])
    ga = gpu["assoc"].astype(np.int64)
    ga[ga == 0xFFFFFFFF] = -1
    np.testing.assert_array_equal(ga[done], ref["assoc"][done].astype(np.int64))
    if exact_positions:
        np.testing.assert_array_equal(gpu["f64"][done], ref["f64"][done])  # bit-identical doubles
        for ax in "xyz":
            np.testing.assert_array_equal(gpu[ax], ref[ax])
    for ax in "xyz":
        np.testing.assert_allclose(gpu[ax].astype(np.float64), ref[ax].astype(np.float64), rtol=0,
                                   atol=tol)


def _run_both(capi, oracle, case, k, tol_t, precision=0, subset=None, stamped=True, stride=8):
    h = capi.DeformHandle(0)
    _set(h, case)
    st = case["stamps"] if stamped else None
    gpu = h.deform_points(case["x"], case["y"], case["z"], st, k=k, tol_t=tol_t,
                          precision=precision, subset=subset, want_f64=True, want_assoc=True,
                          assoc_stride=stride)
    ref = oracle.deform_points(case["x"], case["y"], case["z"], st, case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=k, tol_t=tol_t,
                               subset=subset, assoc_stride=stride)
    h.close()
    return gpu, ref


def test_reference_unit_test_deform_points(capi):
    # kimera_pgmo/test/test_mesh_deformation.cpp:16-50 through the C ABI
    xs = np.arange(100, dtype=np.float32)
    zeros = np.zeros(100, dtype=np.float32)
    ctrl = np.array([[i, 0, 0] for i in range(0, 100, 10)], dtype=np.float64)
    poses = np.zeros((10, 12))
    poses[:, [0, 4, 8]] = 1.0
    poses[:, 9] = ctrl[:, 0]
    poses[:, 10] = 1.0
    h = capi.DeformHandle(0)
    h.set_controls(ctrl, None, None, poses)
    r = h.deform_points(xs, zeros, zeros, None, k=4)
    assert r["rc"] == 0
    np.testing.assert_allclose(r["x"], xs, atol=1e-9)
    np.testing.assert_allclose(r["y"], 1.0, atol=1e-9)
    np.testing.assert_allclose(r["z"], 0.0, atol=1e-9)


def test_reference_unit_test_time_check(capi, oracle):
    # test_mesh_deformation.cpp:52-112: k = 3, tol 10 s, stamps 0 / 20 s (uint64 wrap quirk)
    xs = np.arange(100, dtype=np.float32)
    zeros = np.zeros(100, dtype=np.float32)
    ctrl = np.array([[i, 0, 0] for i in range(0, 100, 10)], dtype=np.float64)
    poses = np.zeros((10, 12))
    poses[:, [0, 4, 8]] = 1.0
    poses[:, 9] = ctrl[:, 0]
    poses[:, 10] = np.where(ctrl[:, 0] > 50, 1.0, -1.0)
    t20 = capi.stamp_from_sec(20.0)
    cst = np.where(ctrl[:, 0] > 50, t20, 0).astype(np.uint64)
    stamps = np.where(np.arange(100) < 50, 0, t20).astype(np.uint64)
    case = dict(x=xs, y=zeros, z=zeros, stamps=stamps, ctrl_xyz=ctrl, ctrl_stamps=cst,
                valid=np.ones(10, np.uint8), poses=poses)
    gpu, ref = _run_both(capi, oracle, case, 3, 10.0)
    _compare(gpu, ref)
    np.testing.assert_array_equal(gpu["y"][:50], -1.0)
    np.testing.assert_array_equal(gpu["y"][51:], 1.0)
    np.testing.assert_array_equal(gpu["x"], xs)


def test_cube_known_answer(capi):
    # test_deformation_graph.cpp:250-309 second phase: pins the LINEAR weight (1.273)
    cube = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1],
                     [1, 1, 1]], dtype=np.float32)
    ctrl = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    poses = np.zeros((3, 12))
    poses[:, [0, 4, 8]] = 1.0
    poses[:, 9:] = [[-0.5, 0, 0], [1.5, 0, 0], [0.5, 1, 0]]
    h = capi.DeformHandle(0)
    h.set_controls(ctrl, np.zeros(3, np.uint64), None, poses)
    r = h.deform_points(cube[:, 0].copy(), cube[:, 1].copy(), cube[:, 2].copy(),
                        np.zeros(8, np.uint64), k=2, tol_t=10.0)
    assert abs(r["x"][0] + 0.5) < 1e-3
    assert abs(r["x"][1] - 1.273) < 1e-3


@pytest.mark.parametrize("tol_t", [10.0, 20.0])
@pytest.mark.parametrize("invalid", [0.0, 0.01])
def test_config_a_parity(capi, oracle, tol_t, invalid):
    # configs[0]: 200k vertices, 2k control points, k = 4
    case = synth.deformation_case(200_000, 2_000, invalid_fraction=invalid)
    gpu, ref = _run_both(capi, oracle, case, 4, tol_t)
    _compare(gpu, ref)
    assert ref["processed"].all()


@pytest.mark.parametrize("spread", [(6.0, 5.0, 10), (40.0, 2.0, 10), (0.5, 30.0, 10),
                                    (6.0, 5.0, 40), (3.0, 0.0, 10), (6.0, 5.0, 20)])
def test_search_modes_agree(capi, oracle, spread):
    """Bucket-ordered lists (default), grid walk and brute force over the window are
    bit-identical, also when vertices lie far outside / deep inside the control-point cloud,
    when the alive set exceeds the bucket kernel's capacity (40 control points per keyframe:
    ~840 alive: grid walk; 20: ~420 alive: the non-staged bucket kernel) and when all control
    points of a keyframe coincide."""
    case = synth.deformation_case(150_000, 300 * spread[2], ctrl_per_kf=spread[2],
                                  vertex_spread=spread[0], ctrl_spread=spread[1], seed=17)
    outs = []
    for mode in (0, 1, 2):
        h = capi.DeformHandle(0)
        h.set_search_mode(mode)
        _set(h, case)
        outs.append(h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4,
                                    tol_t=10.0, want_f64=True, want_assoc=True))
        h.close()
    for key in ("x", "y", "z", "f64", "assoc", "assoc_cnt", "processed"):
        np.testing.assert_array_equal(outs[0][key], outs[1][key])
        np.testing.assert_array_equal(outs[0][key], outs[2][key])
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0)
    _compare(outs[1], ref)


def test_fp32_mode(capi, oracle):
    case = synth.deformation_case(200_000, 2_000)
    gpu, ref = _run_both(capi, oracle, case, 4, 10.0, precision=1)
    _compare(gpu, ref, exact_positions=False, tol=FP32_TOL)


def test_unsigned_wrap_epoch_zero(capi, oracle):
    # stamps starting at 0: `stamp - tol` wraps (mesh_deformation.h:202)
    case = synth.deformation_case(50_000, 1_000, t0_ns=0)
    gpu, ref = _run_both(capi, oracle, case, 4, 10.0)
    _compare(gpu, ref)


@pytest.mark.parametrize("k", [1, 2, 3, 6, 9])
def test_k_sweep(capi, oracle, k):
    case = synth.deformation_case(20_000, 500, seed=3)
    gpu, ref = _run_both(capi, oracle, case, k, 10.0, stride=12)
    _compare(gpu, ref)


def test_unstamped(capi, oracle):
    case = synth.deformation_case(20_000, 300, seed=5)
    gpu, ref = _run_both(capi, oracle, case, 4, 10.0, stamped=False)
    _compare(gpu, ref)


def test_subset_indices(capi, oracle):
    case = synth.deformation_case(30_000, 600, seed=9)
    rng = np.random.default_rng(1)
    subset = np.sort(rng.choice(30_000, 7_000, replace=False)).astype(np.uint64)
    subset = subset[subset >= 12_000]  # incremental mode: tail of the mesh
    gpu, ref = _run_both(capi, oracle, case, 4, 10.0, subset=subset)
    _compare(gpu, ref)


@pytest.mark.parametrize("ppv", [1, 3])
def test_lattice_ties_follow_pcl_order(capi, oracle, ppv):
    # exact float-distance ties everywhere: associations must follow PCL's traversal order
    case = synth.lattice_case(n_side=12, points_per_voxel=ppv)
    gpu, ref = _run_both(capi, oracle, case, 4, 6.0, stride=12)
    _compare(gpu, ref)
    gpu, ref = _run_both(capi, oracle, case, 2, 100.0, stamped=False)
    _compare(gpu, ref)


def test_many_points_per_voxel_unstable_sort(capi, oracle):
    # > 16 candidates inside one voxel with ties: PCL's std::sort is not stable there
    rng = np.random.default_rng(11)
    n_c = 60
    cpos = np.zeros((n_c, 3))
    cpos[:, 0] = 0.125 * rng.integers(0, 4, n_c)
    cpos[:, 1] = 0.125 * rng.integers(0, 4, n_c)
    cpos[:, 2] = 0.125 * rng.integers(0, 2, n_c)
    cpos[40:] += [3.0, 0, 0]
    poses = np.zeros((n_c, 12))
    poses[:, [0, 4, 8]] = 1.0
    poses[:, 9:] = cpos + rng.uniform(-0.2, 0.2, (n_c, 3))
    g = 0.125 * np.arange(0, 32, dtype=np.float32)
    vx, vy = np.meshgrid(g, g[:8], indexing="ij")
    case = dict(x=vx.ravel().copy(), y=vy.ravel().copy(), z=np.zeros(vx.size, np.float32),
                stamps=np.zeros(vx.size, np.uint64), ctrl_xyz=cpos,
                ctrl_stamps=np.zeros(n_c, np.uint64), valid=np.ones(n_c, np.uint8), poses=poses)
    for k in (4, 12):
        gpu, ref = _run_both(capi, oracle, case, k, 10.0, stamped=False, stride=16)
        _compare(gpu, ref)


def test_too_few_controls_skip_and_k_shrink(capi, oracle):
    # first stretches see < k + 1 voxels: k shrinks (:183), single-point windows skip (:185)
    case = synth.deformation_case(3_000, 30, ctrl_per_kf=1, seed=2)
    case["valid"][1:4] = 0
    gpu, ref = _run_both(capi, oracle, case, 4, 0.4)
    _compare(gpu, ref)


def test_early_returns(capi, oracle):
    case = synth.deformation_case(100, 3, ctrl_per_kf=1)
    gpu, ref = _run_both(capi, oracle, case, 4, 10.0)  # control_points.size() < k
    assert gpu["rc"] == ref["rc"] == 1
    np.testing.assert_array_equal(gpu["x"], case["x"])
    assert not gpu["processed"].any()


def test_shard_invariance_single_gpu(capi):
    """Vertex-range slabs deformed independently give bitwise the same result as one pass."""
    import torch
    case = synth.deformation_case(100_000, 2_000)
    dev = torch.device("cuda:0")
    h = capi.DeformHandle(0)
    _set(h, case)
    t = {k: torch.from_numpy(case[k]).to(dev) for k in ("x", "y", "z")}
    st = torch.from_numpy(case["stamps"].view(np.int64)).to(dev)
    n = case["x"].shape[0]

    def run(n_shards):
        out = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(3)]
        bounds = [n * i // n_shards for i in range(n_shards + 1)]
        stamps, lens = [], []
        for s in range(n_shards):
            a, b = bounds[s], bounds[s + 1]
            ss, ll = h.stamp_stretches_device(st[a:b], b - a)
            stamps.append(ss)
            lens.append(ll)
        h.build_schedule(np.concatenate(stamps), np.concatenate(lens), n, 4, 10.0)
        for s in range(n_shards):
            a, b = bounds[s], bounds[s + 1]
            h.run_device(a, b - a, t["x"][a:b], t["y"][a:b], t["z"][a:b], out[0][a:b],
                         out[1][a:b], out[2][a:b])
        torch.cuda.synchronize()
        return [o.cpu().numpy() for o in out]

    one = run(1)
    for shards in (2, 3, 8):
        many = run(shards)
        for a, b in zip(one, many):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("precision", [0, 1])
def test_config_c_slice_parity(capi, oracle, precision):
    """configs[2] at 1/5 scale (2M vertices, 10k control points, same density as the 10M / 50k
    benchmark workload): full oracle parity through the pipelined host entry point.  The full
    10M-vertex comparison is scripts/full_parity.py (profiles/r01_full_parity.txt)."""
    case = synth.deformation_case(2_000_000, 10_000)
    h = capi.DeformHandle(0)
    _set(h, case)
    gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0,
                          precision=precision)
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0,
                               want_trace=False)
    h.close()
    assert gpu["rc"] == ref["rc"] == 0
    assert gpu["processed"].all()
    for ax in "xyz":
        if precision == 0:
            np.testing.assert_array_equal(gpu[ax], ref[ax])
        else:
            np.testing.assert_allclose(gpu[ax].astype(np.float64), ref[ax].astype(np.float64),
                                       rtol=0, atol=FP32_TOL)


@pytest.mark.parametrize("n_ctrl", [7, 126, 127, 128, 254, 255, 256, 257, 510, 511, 512, 513, 700])
def test_alive_set_sizes_at_kernel_boundaries(capi, oracle, n_ctrl):
    """Unstamped: one run whose alive set is exactly n_ctrl points — the staged / unstaged
    variants of the list kernel (<= 256 / <= 512 points), the grid-walk kernel beyond, and every
    width of the list-position field inside the 32-bit keys (7 to 10 bits)."""
    case = synth.deformation_case(6_000, n_ctrl, ctrl_per_kf=n_ctrl, seed=100 + n_ctrl, ctrl_spread=9.0,
                                  vertex_spread=11.0)
    gpu, ref = _run_both(capi, oracle, case, 4, 10.0, stamped=False)
    _compare(gpu, ref)
    # a stamped sweep whose window holds the same set for every vertex
    gpu, ref = _run_both(capi, oracle, case, 4, 1.0e6)
    _compare(gpu, ref)


def test_appended_controls_extend_the_schedule(capi, oracle):
    """New keyframes since the last deformMesh: control points appended, mesh grown.  The window
    schedule is resumed from the first stretch that can see them (not replayed from the start)
    and the result is the one a fresh handle / the oracle gives."""
    case = synth.deformation_case(60_000, 1200, seed=31, invalid_fraction=0.02)
    n_c_old, n_v_old = 1100, 55_000
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"][:n_c_old], case["ctrl_stamps"][:n_c_old], case["valid"][:n_c_old],
                   case["poses"][:n_c_old])
    first = h.deform_points(case["x"][:n_v_old], case["y"][:n_v_old], case["z"][:n_v_old],
                            case["stamps"][:n_v_old], k=4, tol_t=10.0)
    assert first["rc"] == 0
    builds0 = h.schedule_stats()[0]
    _set(h, case)
    gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0, want_f64=True,
                          want_assoc=True)
    builds, resumed, n_stretches, _ = h.schedule_stats()
    assert builds == builds0 + 1 and resumed > 0.5 * n_stretches, (builds, resumed, n_stretches)
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0)
    _compare(gpu, ref)
    # same table again, different k: nothing to extend from -> full replay, still right
    gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=3, tol_t=10.0, want_f64=True,
                          want_assoc=True)
    assert h.schedule_stats()[1] == 0
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=3, tol_t=10.0)
    _compare(gpu, ref)
    # an old control point moved: not an append -> full replay
    moved = case["ctrl_xyz"].copy()
    moved[5, 0] += 0.25
    h.set_controls(moved, case["ctrl_stamps"], case["valid"], case["poses"])
    gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=3, tol_t=10.0, want_f64=True,
                          want_assoc=True)
    assert h.schedule_stats()[1] == 0
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], moved, case["ctrl_stamps"],
                               case["valid"], case["poses"], k=3, tol_t=10.0)
    _compare(gpu, ref)
    h.close()


@pytest.mark.parametrize("tie_fraction", [1e-3, 1e-2, 0.2])
def test_planted_ties_follow_the_reference_order(capi, oracle, tie_fraction):
    """Exact float-distance ties at realistic and absurd rates (vertices on the midpoint of two
    lattice-snapped control points: bit-identical d^2 for their two nearest neighbours).  Which of
    the two comes first is PCL's visit order (K4); positions must not depend on it being rare.
    The second deformation runs from the association records the resolver left behind."""
    case, tied = synth.deformation_case_with_ties(200_000, 2_000, tie_fraction)
    assert tied.size > 0.5 * tie_fraction * 200_000
    h = capi.DeformHandle(0)
    _set(h, case)
    gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0,
                          want_f64=True, want_assoc=True)
    assert h.last_step_info()["resolved"] >= tied.size * 0.95  # the planted vertices went through K4
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0)
    _compare(gpu, ref)
    # pose-only update: blend from the records (incl. the resolver's), against a fresh oracle run
    poses2 = case["poses"].copy()
    poses2[:, 9:] += np.array([0.25, -0.5, 0.125])
    h.update_poses(poses2)
    gpu2 = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0,
                           want_f64=True, want_assoc=True)
    ref2 = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                                case["ctrl_stamps"], case["valid"], poses2, k=4, tol_t=10.0)
    _compare(gpu2, ref2)
    h.close()
