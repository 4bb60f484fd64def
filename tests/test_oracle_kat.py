"""CPU: the oracle against the reference's own known-answer tests (oracle/kat_*.cpp replay
kimera_pgmo/test/test_mesh_deformation.cpp, test_deformation_graph.cpp,
test_voxel_clearing_compression.cpp, test_octree_compression.cpp,
test_mesh_frontend_interface.cpp and test_common_*.cpp scenarios with their expected values)."""
import numpy as np


def test_reference_known_answer_tests(oracle):
    rc, out = oracle.run_kat()
    assert rc == 0, out
    assert "0 failures" in out


def test_stamp_from_sec(oracle):
    # test_common_structs.cpp:12-15
    assert oracle.stamp_from_sec(105867.312345678) == 105867312345678


def test_deform_points_python_binding(oracle):
    # test_mesh_deformation.cpp:16-50 through the ctypes binding
    xs = np.arange(100, dtype=np.float32)
    zeros = np.zeros(100, dtype=np.float32)
    ctrl = np.array([[i, 0, 0] for i in range(0, 100, 10)], dtype=np.float64)
    poses = np.zeros((10, 12))
    poses[:, [0, 4, 8]] = 1.0
    poses[:, 9] = ctrl[:, 0]
    poses[:, 10] = 1.0
    r = oracle.deform_points(xs, zeros, zeros, None, ctrl, None, None, poses, k=4)
    assert r["rc"] == 0
    np.testing.assert_allclose(r["x"], xs, atol=1e-9)
    np.testing.assert_allclose(r["y"], 1.0, atol=1e-9)
    np.testing.assert_allclose(r["z"], 0.0, atol=1e-9)
    assert r["processed"].all()
    assert (r["assoc_cnt"] == 5).all()  # k + 1 neighbours are blended


def test_delta_compression_python_binding(oracle):
    """DeltaCompression restatement through ctypes: a small stream with archival; structural
    invariants of MeshDelta (mesh_delta.cpp:203-260 validate / checkFaces)."""
    import numpy as np
    from kimera_pgmo_b200 import synth

    d = oracle.DeltaCompression(0.05)
    total_archived = 0
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(8, 3000, n_blocks=5)):
        if m >= 3:
            d.archive_by_time(int((t - 2.0) * 1e9))
        remap, delta = d.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        assert delta["vertex_start"] == total_archived
        total_archived += delta["num_archived_vertices"]
        n_total = delta["vertex_start"] + delta["vertices"].shape[0]
        assert remap.shape[0] == xyz.shape[0] and (remap < n_total).all()
        assert (remap >= delta["vertex_start"] + delta["num_archived_vertices"]).all()  # active vertices
        assert (delta["faces"] < n_total).all()
        assert (delta["faces_archive"] < total_archived).all()
        assert delta["active_voxels"] == delta["vertices"].shape[0] - delta["num_archived_vertices"]
        assert np.array_equal(np.intersect1d(delta["new"], delta["observed"]), delta["new"])
    assert total_archived > 0
