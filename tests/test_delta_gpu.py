"""GPU parity: pgmo_delta_* (DeltaCompression through the C ABI) against the CPU oracle.

Bar: bit-exact — per update the remap of every input vertex and every field of the MeshDelta
(vertex_start, face_start, vertices, colours, stamps, face_updates, face_archive_updates,
prev_to_curr, deleted / observed / new indices), including archiveBlocksByTime /
clearArchivedBlocks in between.  The oracle itself is pinned by every configuration of the
reference's test_delta_compression.cpp (oracle/kat_delta.cpp); the same configurations are
replayed here through the GPU path."""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu

SEC = 1_000_000_000
FIELDS = ("vertices", "rgba", "stamps", "faces", "faces_archive", "prev_to_curr", "deleted", "observed", "new")


def _same(gd, od, where):
    for key in ("vertex_start", "face_start", "num_archived_vertices", "active_voxels", "pending_archived_faces"):
        assert gd[key] == od[key], f"{where}: {key} {gd[key]} vs {od[key]}"
    for key in FIELDS:
        np.testing.assert_array_equal(gd[key], od[key], err_msg=f"{where}: {key}")


B1, B2 = (0, 0, 0), (-1, 0, 0)
F_A = [(0.5, 0.5, 0.5), (0.5, 0.75, 0.75), (0.5, 0.75, 0.5)]
F_B = [(0.0, 0.0, 0.0), (0.0, 0.5, 0.5), (0.0, 0.0, 0.5)]
F_C = [(-0.5, 0.5, 0.5), (-0.5, 0.75, 0.75), (-0.5, 0.75, 0.5)]
F_D = [(0.0, 0.0, 0.0), (0.0, 0.5, 0.5), (0.0, 0.0, -0.5)]
BLOCKS = {"b1_empty": (B1, []), "b1_v1": (B1, [F_A, F_A, F_B]), "b1_v2": (B1, [F_A, F_B]),
          "b2_empty": (B2, []), "b2_v1": (B2, [F_C, F_B]), "b2_v2": (B2, [F_C, F_D])}
# test_delta_compression.cpp:334-435: (prune time [s] | None, stamp [s], blocks)
CONFIGS = {
    "SingleBlockClearing": [(None, 100, ["b1_v1"]), (None, 101, ["b1_empty"])],
    "SingleBlockPrune": [(None, 100, ["b1_v1"]), (101, 102, ["b1_empty"]), (None, 103, ["b1_v1"])],
    "MultiBlockClearing": [(None, 100, ["b1_v1", "b2_v1"]), (None, 101, ["b1_empty", "b2_v1"]),
                           (None, 102, ["b1_v1", "b2_v1"]), (None, 103, ["b1_v1", "b2_empty"])],
    "MultiBlockPrune": [(None, 100, ["b1_v1"]), (None, 102, ["b2_v1"]), (101, 103, ["b1_empty", "b2_empty"]),
                        (None, 104, ["b1_v1"]), (None, 105, ["b1_v1"])],
    "MultiBlockPartialUpdates": [(None, 100, ["b1_v1"]), (None, 102, ["b2_v1"]), (101, 103, ["b1_empty", "b2_v1"]),
                                 (None, 104, ["b2_v1"]), (None, 105, ["b1_v1", "b2_v1"]),
                                 (None, 106, ["b1_empty", "b2_empty"])],
    "MultiBlockPartialArchive": [(None, 100, ["b1_v2"]), (None, 102, ["b2_v2"]), (101, 103, ["b1_empty", "b2_v2"]),
                                 (None, 104, ["b2_empty"])],
    "RepeatedTimestamp": [(None, 100, ["b1_v1"]), (None, 100, ["b2_v1"])],
}
# expected (vertex_start, face_start, num_vertices) per step, same source
EXPECTED_STATE = {
    "SingleBlockClearing": [(0, 0, 6), (0, 0, 0)],
    "SingleBlockPrune": [(0, 0, 6), (0, 0, 6), (6, 2, 6)],
    "MultiBlockClearing": [(0, 0, 9), (0, 0, 6), (0, 0, 9), (0, 0, 6)],
    "MultiBlockPrune": [(0, 0, 6), (0, 0, 9), (0, 0, 6), (6, 2, 6), (6, 2, 6)],
    "MultiBlockPartialUpdates": [(0, 0, 6), (0, 0, 9), (0, 0, 9), (3, 1, 6), (3, 1, 9), (3, 1, 3)],
    "MultiBlockPartialArchive": [(0, 0, 6), (0, 0, 10), (0, 0, 10), (4, 1, 2)],
    "RepeatedTimestamp": [(0, 0, 6), (0, 0, 9)],
}


def _mesh(names, point_index):
    idx, offsets, pts, cols = [], [0], [], []
    for name in names:
        index, faces = BLOCKS[name]
        idx.append(index)
        for f in faces:
            for v in f:
                pts.append(v)
                cols.append(0xFF000000 | (point_index << 16) | (point_index << 8) | point_index)
                point_index += 1
        offsets.append(len(pts))
    return (np.array(idx, np.int32).reshape(-1, 3), np.array(offsets, np.uint64),
            np.array(pts, np.float32).reshape(-1, 3), np.array(cols, np.uint32)), point_index


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_reference_configurations(capi, oracle, name):
    gpu, ora = capi.DeltaHandle(1.0e-3, 0), oracle.DeltaCompression(1.0e-3)
    point_index = 0
    for step, (prune, stamp, blocks) in enumerate(CONFIGS[name]):
        if prune is not None:
            gpu.archive_by_time(prune * SEC)
            ora.archive_by_time(prune * SEC)
        (bi, off, xyz, rgba), point_index = _mesh(blocks, point_index)
        gr, gd = gpu.update(bi, off, xyz, rgba, stamp * SEC)
        orr, od = ora.update(bi, off, xyz, rgba, stamp * SEC)
        np.testing.assert_array_equal(gr, orr, err_msg=f"{name} step {step}: remap")
        _same(gd, od, f"{name} step {step}")
        vs, fs, nv = EXPECTED_STATE[name][step]
        assert (gd["vertex_start"], gd["face_start"], gd["vertices"].shape[0]) == (vs, fs, nv)
    gpu.close()


@pytest.mark.parametrize("n_blocks,horizon", [(1, None), (6, 3.0), (40, 2.0)])
def test_stream_parity(capi, oracle, n_blocks, horizon):
    gpu, ora = capi.DeltaHandle(0.05, 0), oracle.DeltaCompression(0.05)
    seen_archived = seen_pending = 0
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(14, 6000, n_blocks=n_blocks)):
        if horizon is not None and m >= 3:
            cut = int((t - horizon) * 1e9)
            gpu.archive_by_time(cut)
            ora.archive_by_time(cut)
        gr, gd = gpu.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        orr, od = ora.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        np.testing.assert_array_equal(gr, orr, err_msg=f"message {m}: remap")
        _same(gd, od, f"message {m}")
        seen_archived += gd["num_archived_vertices"]
        seen_pending = max(seen_pending, gd["pending_archived_faces"])
    if horizon is not None:
        assert seen_archived > 1000 and seen_pending > 0
    gpu.close()


def test_clear_archived_blocks_and_reobservation(capi, oracle):
    gpu, ora = capi.DeltaHandle(0.05, 0), oracle.DeltaCompression(0.05)
    stream = list(synth.compression_stream(6, 4500, n_blocks=9))
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(stream):
        if m in (2, 4):  # archive half of the blocks of the previous message by index, twice
            prev = stream[m - 1][0][::2]
            gpu.archive_blocks(prev)
            ora.archive_blocks(prev)
            gpu.archive_blocks(prev[:1])  # already gone: no-op
            ora.archive_blocks(prev[:1])
        gr, gd = gpu.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        orr, od = ora.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        np.testing.assert_array_equal(gr, orr, err_msg=f"message {m}: remap")
        _same(gd, od, f"message {m}")
    # the same first message again after everything was archived: all vertices are new again
    gpu.archive_by_time(2**62)
    ora.archive_by_time(2**62)
    blocks, offsets, xyz, rgba, t = stream[0]
    gr, gd = gpu.update(blocks, offsets, xyz, rgba, int(100e9))
    orr, od = ora.update(blocks, offsets, xyz, rgba, int(100e9))
    np.testing.assert_array_equal(gr, orr)
    _same(gd, od, "re-observation")
    assert gd["new"].shape[0] == gd["active_voxels"]
    gpu.close()


def test_empty_and_invalid_inputs(capi, oracle):
    gpu, ora = capi.DeltaHandle(0.1, 0), oracle.DeltaCompression(0.1)
    e = (np.zeros((0, 3), np.int32), np.zeros(1, np.uint64), np.zeros((0, 3), np.float32), np.zeros(0, np.uint32))
    gr, gd = gpu.update(*e, 5)
    orr, od = ora.update(*e, 5)
    _same(gd, od, "empty message")
    gpu.archive_by_time(10)  # nothing to archive
    ora.archive_by_time(10)
    gr, gd = gpu.update(*e, 6)
    orr, od = ora.update(*e, 6)
    _same(gd, od, "empty message after an empty archive")
    # a block whose vertex count is not a multiple of 3: Face(indices, i) throws in the reference
    bad = (np.zeros((1, 3), np.int32), np.array([0, 4], np.uint64), np.zeros((4, 3), np.float32), np.zeros(4, np.uint32))
    with pytest.raises(capi.PgmoError) as err:
        gpu.update(*bad, 7)
    assert err.value.code == capi.PGMO_ERR_OUT_OF_RANGE
    gpu.close()


def test_compaction_of_records_and_pools_keeps_parity(capi, oracle, monkeypatch):
    """Erased voxel records and replaced block segments are garbage-collected (stable compaction,
    table rebuild, voxel lists renumbered); with the slack set to 0 this happens after almost
    every message and must not change a single output."""
    monkeypatch.setenv("PGMO_DELTA_COMPACT_SLACK", "0")
    gpu, ora = capi.DeltaHandle(0.05, 0), oracle.DeltaCompression(0.05)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(16, 6000, n_blocks=5, slide=4.0)):
        if m >= 2:
            cut = int((t - 2.0) * 1e9)
            gpu.archive_by_time(cut)
            ora.archive_by_time(cut)
        gr, gd = gpu.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        orr, od = ora.update(blocks, offsets, xyz, rgba, int(t * 1e9))
        np.testing.assert_array_equal(gr, orr, err_msg=f"message {m}: remap")
        _same(gd, od, f"message {m}")
    assert gpu.last_stats()[3] >= 3  # record compactions happened
    gpu.close()


@pytest.mark.parametrize("n_blocks,horizon", [(1, None), (6, 3.0), (40, 2.0)])
def test_semantic_labels_follow_the_points(capi, oracle, n_blocks, horizon, monkeypatch):
    """MeshInterface::getActiveSemantics -> VertexInfo::label -> MeshDelta::semantic_updates
    (delta_compression.cpp:76, :120, :128, :285, :347): a voxel keeps the label of the vertex that
    set its point in its last observed pass; archived and carried-over vertices keep theirs;
    unlabelled messages in between reset the labels of what they re-observe.  Also with the record
    compaction forced after every message."""
    for compact in (False, True):
        if compact:
            monkeypatch.setenv("PGMO_DELTA_COMPACT_SLACK", "0")
        gpu, ora = capi.DeltaHandle(0.05, 0), oracle.DeltaCompression(0.05)
        rng = np.random.default_rng(4)
        for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(12, 6000, n_blocks=n_blocks)):
            if horizon is not None and m >= 3:
                cut = int((t - horizon) * 1e9)
                gpu.archive_by_time(cut)
                ora.archive_by_time(cut)
            labels = None
            if m % 4 != 2:  # every fourth message carries no semantics
                labels = rng.integers(0, 40, xyz.shape[0]).astype(np.uint32)
                labels[rng.random(xyz.shape[0]) < 0.05] = 0xFFFFFFFF  # std::nullopt for some vertices
            gr, gd = gpu.update(blocks, offsets, xyz, rgba, int(t * 1e9), labels=labels)
            orr, od = ora.update(blocks, offsets, xyz, rgba, int(t * 1e9), labels=labels)
            np.testing.assert_array_equal(gr, orr, err_msg=f"message {m}: remap")
            _same(gd, od, f"message {m}")
            gl, any_label = gpu.labels()
            ol, n_sem = ora.labels()
            assert any_label
            np.testing.assert_array_equal(gl, ol, err_msg=f"message {m}: labels")
            assert int((gl != 0xFFFFFFFF).sum()) == n_sem
        gpu.close()
