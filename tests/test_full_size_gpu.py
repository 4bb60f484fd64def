"""Full-size parity on the BASELINE.json configurations themselves (round 1 had them only as
builder-run scripts under scripts/ and profiles/).

configs[2]: the 10M-vertex / 50k-control-point mesh of bench.py through the pipelined host
entry point, against the oracle on (a) the first 1.5M vertices and (b) 40 two-stretch slices
spread over the whole mesh.  A slice is swept by the oracle with the FULL control table from 22
stretches earlier: the reference's add loop (mesh_deformation.h:167-178) inserts every control
point from index 0 in index order, so the octree is anchored and grown exactly as in the full
sweep.  The alive set needs the warm-up: deleteVoxelAtPoint takes every point of the voxel with
it (mesh_deformation.cpp:40), so whether a young control point survives depends on whether it
was inserted before or after an old neighbour in its voxel expired — history that a sweep
started mid-way gets wrong for the control points it bulk-inserts at its first vertex, i.e. for
2 x tol_t = 20 stretches.  After that every alive point was inserted by the running sweep and
the state is the full sweep's.

Plus properties that hold for every vertex without an oracle (identity correction, a common
rigid motion, first call == pose-only call).

configs[1]: 50k-vertex messages, 24 of them, as 1 block and as 64 voxblox-like blocks per
message, horizon pruning / archiving before each message, through VoxelClearing, Octree and
Delta compression: every remap, every output list and the stored mesh against the oracle.
"""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu

N_V, N_C = 10_000_000, 50_000


@pytest.fixture(scope="module")
def mesh10m():
    return synth.deformation_slab(N_V, N_C, 0, N_V)


@pytest.fixture(scope="module")
def deformed10m(capi, mesh10m):
    case = mesh10m
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    first = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    assert first["rc"] == 0 and first["processed"].all()
    assert not h.last_step_info()["warm"]
    # a second solve: every slab of the pipelined call now starts from the association records
    poses2 = case["poses"].copy()
    poses2[:, 9:] += np.random.default_rng(2).normal(0, 0.05, poses2[:, 9:].shape)
    h.update_poses(poses2)
    second = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    info = h.last_step_info()
    assert second["rc"] == 0 and info["warm"] and info["stale_items"] == 0, info
    h.close()
    return first, second, poses2


def _oracle_range(oracle, case, poses, a, b):
    return oracle.deform_points(case["x"][a:b], case["y"][a:b], case["z"][a:b], case["stamps"][a:b],
                                case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], poses, k=4,
                                tol_t=10.0, want_trace=False)


def test_config2_prefix_parity(oracle, mesh10m, deformed10m):
    first, second, poses2 = deformed10m
    n = 1_500_000
    for gpu, poses in ((first, mesh10m["poses"]), (second, poses2)):
        ref = _oracle_range(oracle, mesh10m, poses, 0, n)
        assert ref["rc"] == 0
        for ax in "xyz":
            np.testing.assert_array_equal(gpu[ax][:n], ref[ax])


def test_config2_slices_over_the_whole_mesh(oracle, mesh10m, deformed10m):
    first, second, poses2 = deformed10m
    per_kf = -(-N_V // (N_C // 10))      # vertices per stamp stretch
    n_kf = N_C // 10
    checked = 0
    warm_up = 22
    for n, s in enumerate(np.linspace(800, n_kf - 3, 40).astype(np.int64)):
        a, c, b = (int(s) - warm_up) * per_kf, int(s) * per_kf, min((int(s) + 2) * per_kf, N_V)
        gpu, poses = (first, mesh10m["poses"]) if n % 2 == 0 else (second, poses2)
        ref = _oracle_range(oracle, mesh10m, poses, a, b)
        assert ref["rc"] == 0
        for ax in "xyz":
            np.testing.assert_array_equal(gpu[ax][c:b], ref[ax][c - a:], err_msg=f"stretch {s}")
        checked += b - c
    assert checked > 150_000


# ---- size-independent properties of interpPoint (mesh_deformation.cpp:73-106) on all 10 M
# vertices: no oracle needed, the expected values follow from the formula
#   out = sum_j w_j (R_j (v - g_j) + t_j) / sum_j w_j

def _deform_with(capi, case, poses):
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], poses)
    out = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    assert out["rc"] == 0 and out["processed"].all()
    h.close()
    return out


def test_config2_identity_correction_returns_the_mesh(capi, mesh10m):
    """R_j = I, t_j = g_j: every term is (v - g_j) + g_j = v, the weighted mean of equal doubles is
    within a few 1e-16 relative of v, and v is a float: bit-identical output for every vertex."""
    case = mesh10m
    poses = np.zeros((N_C, 12))
    poses[:, [0, 4, 8]] = 1.0
    poses[:, 9:] = case["ctrl_xyz"]
    out = _deform_with(capi, case, poses)
    for ax in "xyz":
        # (v - g) + g is exact in double unless |v| is some 2^-29 of |g|: such a coordinate may move
        # by a rounding of g's magnitude, nothing else may move at all
        bad = out[ax] != case[ax]
        assert bad.sum() <= 8, int(bad.sum())
        assert (np.abs(case[ax][bad]) < 1e-4).all()
        assert (np.abs(out[ax][bad].astype(np.float64) - case[ax][bad]) < 1e-11).all()


def test_config2_common_rigid_motion(capi, mesh10m):
    """the same rigid motion T = (R, d) on every node: R_j = R, t_j = R g_j + d gives
    R (v - g_j) + R g_j + d = R v + d for every neighbour, so the blend is R v + d whatever the
    neighbour sets and weights are — up to the rounding of the terms (1e-5 m at 1 km)."""
    case = mesh10m
    th = np.deg2rad(7.0)
    R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
    d = np.array([0.5, -0.25, 0.125])
    poses = np.empty((N_C, 12))
    poses[:, :9] = R.reshape(-1)
    poses[:, 9:] = case["ctrl_xyz"] @ R.T + d
    out = _deform_with(capi, case, poses)
    v = np.stack([case["x"], case["y"], case["z"]], axis=1).astype(np.float64)
    want = v @ R.T + d
    got = np.stack([out["x"], out["y"], out["z"]], axis=1).astype(np.float64)
    scale = np.maximum(1.0, np.abs(want).max())
    assert np.abs(got - want).max() <= 2.0 ** -22 * scale  # a float ulp at that magnitude + slack


def test_config2_first_call_equals_pose_only_call(capi, mesh10m):
    """search + blend (first deformMesh) and blend-from-records (after update_poses) are the same
    function of (mesh, control table, poses): bit-identical on all vertices"""
    case = mesh10m
    poses2 = case["poses"].copy()
    poses2[:, 9:] += np.random.default_rng(5).normal(0, 0.05, poses2[:, 9:].shape)
    cold = _deform_with(capi, case, poses2)          # records do not exist yet
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    assert h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)["rc"] == 0
    h.update_poses(poses2)
    warm = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
    info = h.last_step_info()
    assert warm["rc"] == 0 and info["warm"] and info["stale_items"] == 0, info
    h.close()
    for ax in "xyz":
        np.testing.assert_array_equal(warm[ax], cold[ax])


def _vcc_state(gpu, ora):
    g, o = gpu.snapshot(), ora.snapshot()
    for key in ("vertices", "rgba", "stamps", "active", "polygons"):
        np.testing.assert_array_equal(g[key], o[key], err_msg=key)
    np.testing.assert_array_equal(np.sort(g["invalid"]), np.sort(o["invalid"]))


@pytest.mark.parametrize("n_blocks", [1, 64])
def test_config1_voxel_clearing(capi, oracle, n_blocks):
    gpu, ora = capi.VccHandle(0.05), oracle.Compression("voxel_clearing", 0.05)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(
            synth.compression_stream(24, 50_000, n_blocks=n_blocks)):
        gpu.prune(t - 10.0)
        ora.prune(t - 10.0)
        remap, keys = gpu.integrate(blocks, offsets, xyz, rgba, t, want_keys=True)
        idx, _, okeys = ora.integrate(blocks, offsets, xyz, rgba, t)
        np.testing.assert_array_equal(keys, okeys, err_msg=f"message {m}")
        np.testing.assert_array_equal(remap.astype(np.int64), idx, err_msg=f"message {m}")
        if m % 8 == 7:
            _vcc_state(gpu, ora)
    _vcc_state(gpu, ora)
    gpu.close()


@pytest.mark.parametrize("n_blocks", [1, 64])
def test_config1_octree(capi, oracle, n_blocks):
    gpu, ora = capi.OctHandle(0.05, 0), oracle.Compression("octree", 0.05)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(
            synth.compression_stream(24, 50_000, n_blocks=n_blocks)):
        gpu.prune(t - 10.0)
        ora.prune(t - 10.0)
        gi, gk = gpu.integrate(blocks, offsets, xyz, rgba, t)
        oi, ok, _ = ora.integrate(blocks, offsets, xyz, rgba, t)
        np.testing.assert_array_equal(gk, ok, err_msg=f"message {m}: remap kind")
        np.testing.assert_array_equal(gi, oi, err_msg=f"message {m}: remap index")
        go, oo = gpu.last_outputs(), ora.snapshot()
        for key in ("new_indices", "new_vertices", "new_triangles"):
            np.testing.assert_array_equal(go[key], oo[key], err_msg=f"message {m}: {key}")
    gs, os_ = gpu.snapshot(), ora.snapshot()
    for key in ("vertices", "rgba", "stamps", "polygons", "active", "active_xyz", "active_stamps"):
        np.testing.assert_array_equal(gs[key], os_[key], err_msg=key)
    gpu.close()


@pytest.mark.parametrize("n_blocks", [1, 64])
def test_config1_delta(capi, oracle, n_blocks):
    fields = ("vertices", "rgba", "stamps", "faces", "faces_archive", "prev_to_curr", "deleted",
              "observed", "new")
    gpu, ora = capi.DeltaHandle(0.05, 0), oracle.DeltaCompression(0.05)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(
            synth.compression_stream(24, 50_000, n_blocks=n_blocks)):
        cut = int((t - 10.0) * 1e9)
        gpu.archive_by_time(cut)
        ora.archive_by_time(cut)
        full = m % 6 == 5 or m == 23
        gr, gd = gpu.update(blocks, offsets, xyz, rgba, int(t * 1e9), fetch=full)
        orr, od = ora.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        np.testing.assert_array_equal(gr, orr, err_msg=f"message {m}: remap")
        if full:
            for key in ("vertex_start", "face_start", "num_archived_vertices", "active_voxels",
                        "pending_archived_faces"):
                assert gd[key] == od[key], f"message {m}: {key}"
            for key in fields:
                np.testing.assert_array_equal(gd[key], od[key], err_msg=f"message {m}: {key}")
    gpu.close()
