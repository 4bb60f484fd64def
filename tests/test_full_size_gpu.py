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
