"""GPU parity: pgmo_oct_* (OctreeCompression through the C ABI) against the CPU oracle.

Bar: bit-exact — the dense remap record (index + kind of container operation) per input vertex,
new_indices / new_vertices / new_triangles of every call, and the stored mesh (vertices, stamps,
polygons, active set) after the stream, including pruneStoredMesh re-anchoring."""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu


def _compare_state(gpu, ora):
    gs, os_ = gpu.snapshot(), ora.snapshot()
    np.testing.assert_array_equal(gs["vertices"], os_["vertices"])
    np.testing.assert_array_equal(gs["rgba"], os_["rgba"])
    np.testing.assert_array_equal(gs["stamps"], os_["stamps"])
    np.testing.assert_array_equal(gs["polygons"], os_["polygons"])
    np.testing.assert_array_equal(gs["active"], os_["active"])
    np.testing.assert_array_equal(gs["active_xyz"], os_["active_xyz"])
    np.testing.assert_array_equal(gs["active_stamps"], os_["active_stamps"])


def _run_stream(capi, oracle, stream, res, horizon=None, check_every=1, kind="octree"):
    gpu = capi.OctHandle(res, 0, block=(kind == "block"))
    ora = oracle.Compression(kind, res)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(stream):
        if horizon is not None:
            gpu.prune(t - horizon)
            ora.prune(t - horizon)
        gi, gk = gpu.integrate(blocks, offsets, xyz, rgba, t)
        oi, ok, _ = ora.integrate(blocks, offsets, xyz, rgba, t)
        np.testing.assert_array_equal(gk, ok, err_msg=f"message {m}: remap kind")
        np.testing.assert_array_equal(gi, oi, err_msg=f"message {m}: remap index")
        go, oo = gpu.last_outputs(), ora.snapshot()
        np.testing.assert_array_equal(go["new_indices"], oo["new_indices"], err_msg=f"message {m}")
        np.testing.assert_array_equal(go["new_vertices"], oo["new_vertices"], err_msg=f"message {m}")
        np.testing.assert_array_equal(go["new_triangles"], oo["new_triangles"], err_msg=f"message {m}")
        if m % check_every == 0:
            _compare_state(gpu, ora)
    _compare_state(gpu, ora)
    gpu.close()


def _mesh(faces, block=(0, 0, 0)):
    xyz = np.array(faces, dtype=np.float32).reshape(-1, 3)
    rgba = (np.uint32(0xFF000000) + np.arange(xyz.shape[0], dtype=np.uint32))
    return (np.array([block], dtype=np.int32), np.array([0, xyz.shape[0]], dtype=np.uint64), xyz, rgba)


def test_small_meshes_reference_scenarios(capi, oracle):
    # the shapes of test_octree_compression.cpp: a few triangles sharing vertices, re-sent
    tri = [[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[1, 0, 0], [1, 1, 0], [0, 1, 0]],
           [[0, 0, 0], [0, 1, 0], [0, 0, 1]]]
    tri2 = [[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[2, 0, 0], [2, 1, 0], [1, 1, 0]],
            [[-3, -3, -3], [-3, -2, -3], [-2, -3, -3]]]
    stream = [(*_mesh(tri), 100.0), (*_mesh(tri), 100.5), (*_mesh(tri2), 101.0),
              (*_mesh(tri), 140.0), (*_mesh(tri2), 141.0)]
    _run_stream(capi, oracle, stream, 0.1)
    _run_stream(capi, oracle, stream, 0.1, horizon=10.0)
    _run_stream(capi, oracle, stream, 1.5)   # vertices closer than the resolution collapse
    _run_stream(capi, oracle, stream, 3.0, horizon=10.0)


def test_empty_and_degenerate_inputs(capi, oracle):
    empty = (np.zeros((1, 3), np.int32), np.array([0, 0], np.uint64), np.zeros((0, 3), np.float32),
             np.zeros(0, np.uint32))
    same = _mesh([[[1, 1, 1], [1, 1, 1], [1, 1, 1]], [[1, 1, 1], [2, 1, 1], [2, 1, 1]]])
    ragged = _mesh([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[5, 5, 5], [6, 5, 5], [5, 6, 5]]])
    ragged = (ragged[0], np.array([0, 5], np.uint64), ragged[2][:5], ragged[3][:5])  # 5 vertices
    stream = [(*empty, 1.0), (*same, 2.0), (*ragged, 3.0), (*empty, 4.0), (*same, 5.0)]
    _run_stream(capi, oracle, stream, 0.5)


@pytest.mark.parametrize("n_blocks", [1, 7])
@pytest.mark.parametrize("res", [0.05, 0.2])
def test_stream_parity(capi, oracle, n_blocks, res):
    stream = list(synth.compression_stream(6, 6_000, n_blocks=n_blocks, seed=3))
    _run_stream(capi, oracle, stream, res)


def test_stream_with_horizon_pruning(capi, oracle):
    # pruneStoredMesh re-anchors the octree at the first surviving vertex (reInitializeStructure)
    stream = list(synth.compression_stream(12, 3_000, n_blocks=3, seed=9, slide=4.0))
    _run_stream(capi, oracle, stream, 0.05, horizon=2.5)
    _run_stream(capi, oracle, stream, 0.3, horizon=1.5)


@pytest.mark.parametrize("res", [0.05, 0.2, 1.0])
def test_block_compression_parity(capi, oracle, res):
    """BlockCompression: same base flow, cell-hash structure tests."""
    stream = list(synth.compression_stream(6, 6_000, n_blocks=4, seed=3))
    _run_stream(capi, oracle, stream, res, kind="block")
    stream = list(synth.compression_stream(10, 3_000, n_blocks=2, seed=9, slide=4.0))
    _run_stream(capi, oracle, stream, res, horizon=2.5, kind="block")


def test_block_compression_reference_scenarios(capi, oracle):
    # kimera_pgmo/test/test_block_compression.cpp: createSimpleMesh(1.0) then (2.0), res 0.1 and 10
    def simple(scale):
        v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1]], dtype=np.float32) * scale
        faces = [[0, 1, 2], [1, 3, 2], [0, 1, 4], [0, 4, 2]]
        return _mesh([[v[i] for i in f] for f in faces])
    for res in (0.1, 10.0):
        stream = [(*simple(1.0), 100.0), (*simple(2.0), 101.0), (*simple(1.0), 102.0)]
        _run_stream(capi, oracle, stream, res, kind="block")
        _run_stream(capi, oracle, stream, res, horizon=0.9, kind="block")
    h = capi.OctHandle(0.1, 0, block=True)
    h.integrate(*simple(1.0), 100.0)
    assert h.sizes()[0] == 5 and h.sizes()[1] == 4          # :52-54
    assert h.last_outputs()["new_indices"].tolist() == [0, 1, 2, 3, 4]
    h.integrate(*simple(2.0), 101.0)
    assert h.last_outputs()["new_indices"].tolist() == [0, 5, 6, 7, 8]  # :79
    h.close()


def test_dense_chains_need_many_rounds(capi, oracle):
    # vertices 0.6 res apart along lines: every vertex depends on its predecessor's fate, the
    # fixed-point iteration needs about as many rounds as the chain is long
    rng = np.random.default_rng(2)
    n_tri = 300
    base = np.cumsum(np.full(n_tri * 3, 0.06, dtype=np.float32))
    xyz = np.stack([base, 0.013 * rng.standard_normal(n_tri * 3).astype(np.float32),
                    np.zeros(n_tri * 3, np.float32)], axis=1).astype(np.float32)
    rgba = np.arange(n_tri * 3, dtype=np.uint32)
    mesh = (np.zeros((1, 3), np.int32), np.array([0, n_tri * 3], np.uint64), xyz, rgba)
    _run_stream(capi, oracle, [(*mesh, 1.0), (*mesh, 2.0)], 0.1)


def test_frozen_octree_golden(capi):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
    oct_ = capi.OctHandle(0.05, 0)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(
            synth.compression_stream(4, 3_000, n_blocks=3, seed=5)):
        oct_.prune(t - 2.0)
        idx, kind = oct_.integrate(blocks, offsets, xyz, rgba, t)
        np.testing.assert_array_equal(idx, g[f"oct_remap_{m}"])
        np.testing.assert_array_equal(kind, g[f"oct_kind_{m}"])
    snap = oct_.snapshot()
    np.testing.assert_array_equal(snap["vertices"], g["oct_vertices"])
    np.testing.assert_array_equal(snap["polygons"], g["oct_polygons"])
