"""GPU parity: pgmo_vcc_* (through the C ABI) against the CPU oracle and against the reference's
own unit-test expectations (kimera_pgmo/test/test_voxel_clearing_compression.cpp:106-356).

Bar: voxel keys, remap indices, stored vertices, stamps, active indices and the polygon list
bit-exact; invalid indices equal as a set (the reference's order follows unordered_set
iteration and is unpinned)."""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu

B1 = (0, 0, 0)
B2 = (-1, 0, 0)
BLOCK1 = [[(0.5, 0.5, 0.5), (0.5, 0.75, 0.75), (0.5, 0.75, 0.5)],
          [(0.5, 0.5, 0.5), (0.5, 0.75, 0.75), (0.5, 0.75, 0.5)],
          [(0.0, 0.0, 0.0), (0.0, 0.5, 0.5), (0.0, 0.0, 0.5)]]
BLOCK2 = [[(-0.5, 0.5, 0.5), (-0.5, 0.75, 0.75), (-0.5, 0.75, 0.5)],
          [(0.0, 0.0, 0.0), (0.0, 0.5, 0.5), (0.0, 0.0, 0.5)]]


def mesh(*blocks):
    """blocks: (index, faces) pairs -> (block_idx, offsets, xyz, rgba)"""
    idx, offsets, pts = [], [0], []
    for index, faces in blocks:
        idx.append(index)
        for f in faces:
            pts.extend(f)
        offsets.append(len(pts))
    xyz = np.array(pts, dtype=np.float32).reshape(-1, 3)
    rgba = (np.arange(xyz.shape[0], dtype=np.uint32) | np.uint32(0xFF000000))
    return (np.array(idx, dtype=np.int32).reshape(-1, 3), np.array(offsets, dtype=np.uint64), xyz,
            rgba)


def dedup_triangles(polys):
    out = []
    for p in polys.tolist():
        if p not in out:
            out.append(p)
    return [set(p) for p in out]


def expect_triangles(snap, expected):
    assert dedup_triangles(snap["polygons"]) == [set(e) for e in expected]


def test_clearing_correct_basic(capi):
    c = capi.VccHandle(1.0e-3)
    c.integrate(*mesh((B1, BLOCK1)), 100.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 6 and s["invalid"].size == 0 and s["stamps"].shape[0] == 6
    assert s["active"].tolist() == [0, 1, 2, 3, 4, 5]
    expect_triangles(s, [{0, 1, 2}, {3, 4, 5}])
    c.integrate(*mesh((B1, [])), 100.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 6 and s["invalid"].size == 6
    assert s["polygons"].shape[0] == 0 and s["active"].size == 0


def test_pruning_correct_basic(capi):
    c = capi.VccHandle(1.0e-3)
    c.integrate(*mesh((B1, BLOCK1)), 100.0)
    c.prune(101.0)
    c.integrate(*mesh((B1, [])), 105.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 6 and s["invalid"].size == 0 and s["active"].size == 0
    expect_triangles(s, [{0, 1, 2}, {3, 4, 5}])
    c.integrate(*mesh((B1, BLOCK1)), 105.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 12 and s["invalid"].size == 0
    assert s["active"].tolist() == [6, 7, 8, 9, 10, 11]
    expect_triangles(s, [{0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {9, 10, 11}])


def test_clearing_correct_multiblock(capi):
    c = capi.VccHandle(1.0e-3)
    c.integrate(*mesh((B1, BLOCK1), (B2, BLOCK2)), 100.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 9 and s["invalid"].size == 0
    expect_triangles(s, [{6, 7, 8}, {3, 4, 5}, {0, 1, 2}])
    c.integrate(*mesh((B1, []), (B2, BLOCK2)), 100.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 9 and s["invalid"].size == 3
    expect_triangles(s, [{6, 7, 8}, {3, 4, 5}])
    c.integrate(*mesh((B1, BLOCK1), (B2, BLOCK2)), 100.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 12 and s["invalid"].size == 3
    expect_triangles(s, [{6, 7, 8}, {3, 4, 5}, {9, 10, 11}])
    c.integrate(*mesh((B1, BLOCK1), (B2, [])), 100.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 12 and s["invalid"].size == 6
    expect_triangles(s, [{9, 10, 11}, {3, 4, 5}])


def test_pruning_correct_multiblock(capi):
    c = capi.VccHandle(1.0e-3)
    c.integrate(*mesh((B1, BLOCK1)), 100.0)
    c.integrate(*mesh((B2, BLOCK2)), 102.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 9 and s["invalid"].size == 0
    expect_triangles(s, [{6, 7, 8}, {3, 4, 5}, {0, 1, 2}])
    c.prune(101.0)
    c.integrate(*mesh((B1, []), (B2, [])), 105.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 9 and s["invalid"].size == 3
    expect_triangles(s, [{0, 1, 2}, {3, 4, 5}])
    c.integrate(*mesh((B1, BLOCK1)), 105.0)
    s = c.snapshot()
    assert s["vertices"].shape[0] == 15 and s["invalid"].size == 3 and s["stamps"].shape[0] == 15
    expect_triangles(s, [{0, 1, 2}, {3, 4, 5}, {9, 10, 11}, {12, 13, 14}])


def _compare_state(gpu, ora):
    g, o = gpu.snapshot(), ora.snapshot()
    np.testing.assert_array_equal(g["vertices"], o["vertices"])
    np.testing.assert_array_equal(g["rgba"], o["rgba"])
    np.testing.assert_array_equal(g["stamps"], o["stamps"])
    np.testing.assert_array_equal(g["active"], o["active"])
    np.testing.assert_array_equal(np.sort(g["invalid"]), np.sort(o["invalid"]))
    np.testing.assert_array_equal(g["polygons"], o["polygons"])


@pytest.mark.parametrize("n_blocks", [1, 16])
def test_stream_parity(capi, oracle, n_blocks):
    """configs[1] at reduced size: sliding wall, prune(t - 10) before each message."""
    gpu = capi.VccHandle(0.05)
    ora = oracle.Compression("voxel_clearing", 0.05)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(
            synth.compression_stream(14, 6_000, n_blocks=n_blocks, slide=1.0, wall=4.0)):
        gpu.prune(t - 3.0)
        ora.prune(t - 3.0)
        remap, keys = gpu.integrate(blocks, offsets, xyz, rgba, t, want_keys=True)
        idx, kind, okeys = ora.integrate(blocks, offsets, xyz, rgba, t)
        np.testing.assert_array_equal(keys, okeys)
        np.testing.assert_array_equal(remap.astype(np.int64), idx)
        assert (kind == 1).all()
        if m % 4 == 3:
            _compare_state(gpu, ora)
    _compare_state(gpu, ora)
    assert gpu.sizes()[3] > 0  # invalidated slots were exercised


def test_blocks_sharing_voxels_and_reintegration(capi, oracle):
    """Voxels shared between blocks, erased by one block and re-created by a later block of the
    same message, empty blocks, clearArchivedBlocks."""
    rng = np.random.default_rng(3)
    gpu = capi.VccHandle(0.1)
    ora = oracle.Compression("voxel_clearing", 0.1)
    all_blocks = [(bx, by, 0) for bx in range(3) for by in range(2)]
    for step in range(12):
        chosen = [b for b in all_blocks if rng.random() < 0.7]
        rng.shuffle(chosen)
        parts = []
        for b in chosen:
            n_faces = int(rng.integers(0, 40))
            # vertices near the block, quantised so that many share voxels across blocks
            pts = (np.array(b, dtype=np.float32) * 0.8
                   + np.round(rng.uniform(-0.2, 1.0, (3 * n_faces, 3)) * 8) / 8).astype(np.float32)
            parts.append((b, [tuple(map(tuple, pts[3 * f:3 * f + 3])) for f in range(n_faces)]))
        if not parts:
            continue
        args = mesh(*parts)
        t = 50.0 + step
        if step % 3 == 2:
            gpu.prune(t - 2.0)
            ora.prune(t - 2.0)
        if step == 7:
            gpu.clear_blocks(np.array([all_blocks[0], (9, 9, 9)], dtype=np.int32))
            ora.clear_blocks(np.array([all_blocks[0], (9, 9, 9)], dtype=np.int32))
        remap, keys = gpu.integrate(*args, t, want_keys=True)
        idx, _, okeys = ora.integrate(*args, t)
        np.testing.assert_array_equal(keys, okeys)
        np.testing.assert_array_equal(remap.astype(np.int64), idx)
        _compare_state(gpu, ora)


def test_rounding_half_away_and_negative_keys(capi, oracle):
    # round(p / res) at exact halves and negative coordinates (std::round semantics)
    vals = np.array([0.025, -0.025, 0.075, -0.075, 0.124999, -0.125, 0.0, -0.0, 1.025, -3.475],
                    dtype=np.float32)
    pts = np.stack([vals, vals[::-1], np.roll(vals, 3)], axis=1)
    pts = np.concatenate([pts, pts[:2]])  # 12 vertices = 4 faces
    args = (np.zeros((1, 3), np.int32), np.array([0, 12], np.uint64), pts, None)
    gpu = capi.VccHandle(0.05)
    ora = oracle.Compression("voxel_clearing", 0.05)
    remap, keys = gpu.integrate(*args, 1.0, want_keys=True)
    idx, _, okeys = ora.integrate(*args, 1.0)
    np.testing.assert_array_equal(keys, okeys)
    np.testing.assert_array_equal(remap.astype(np.int64), idx)


def test_key_range_is_reported(capi):
    gpu = capi.VccHandle(1.0e-3)
    pts = np.array([[2000.0, 0, 0], [0, 0, 0], [1, 1, 1]], dtype=np.float32)  # 2e6 > 2^20
    with pytest.raises(capi.PgmoError) as e:
        gpu.integrate(np.zeros((1, 3), np.int32), np.array([0, 3], np.uint64), pts, None, 1.0)
    assert e.value.code == capi.PGMO_ERR_UNSUPPORTED


def test_two_handles_two_threads(capi, oracle):
    """Independent handles driven concurrently (mesh_frontend_interface.cpp:100-110)."""
    import threading
    msgs = list(synth.compression_stream(6, 3_000, wall=3.0))
    results = {}

    def worker(name, res):
        h = capi.VccHandle(res)
        out = []
        for blocks, offsets, xyz, rgba, t in msgs:
            out.append(h.integrate(blocks, offsets, xyz, rgba, t)[0])
        results[name] = out

    ths = [threading.Thread(target=worker, args=("a", 0.05)),
           threading.Thread(target=worker, args=("b", 0.2))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for name, res in (("a", 0.05), ("b", 0.2)):
        ora = oracle.Compression("voxel_clearing", res)
        for m, (blocks, offsets, xyz, rgba, t) in enumerate(msgs):
            idx, _, _ = ora.integrate(blocks, offsets, xyz, rgba, t)
            np.testing.assert_array_equal(results[name][m].astype(np.int64), idx)


def test_device_resident_integrate_on_a_caller_stream(capi):
    """pgmo_vcc_integrate_device queues on the caller's stream, pgmo_vcc_prune on the handle's
    own: calls on one handle take effect in call order (the bench's device-resident loop).  The
    stored state must equal the host-buffer path's on the same messages."""
    import torch
    dev = torch.device("cuda", 0)
    msgs = list(synth.compression_stream(60, 50_000, n_blocks=1))
    host = capi.VccHandle(0.05)
    for blocks, offsets, xyz, rgba, t in msgs:
        host.prune(t - 3.0)
        host.integrate(blocks, offsets, xyz, rgba, t)
    want = host.snapshot()
    host.close()
    side = torch.cuda.Stream(device=dev)
    d_msgs = [(torch.from_numpy(x).to(dev), torch.from_numpy(c.view(np.int32)).to(dev))
              for _, _, x, c, _ in msgs]
    d_remap = torch.empty(50_000, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for rep in range(3):
        gpu = capi.VccHandle(0.05)
        with torch.cuda.stream(side):
            for m, (blocks, offsets, _, _, t) in enumerate(msgs):
                gpu.prune(t - 3.0)
                gpu.integrate_device(blocks, offsets, d_msgs[m][0], d_msgs[m][1], t, d_remap,
                                     side.cuda_stream if m % 5 else 0)
        got = gpu.snapshot()
        gpu.close()
        for key in ("vertices", "stamps", "active", "polygons"):
            np.testing.assert_array_equal(got[key], want[key], err_msg=f"{key}, repetition {rep}")
        np.testing.assert_array_equal(np.sort(got["invalid"]), np.sort(want["invalid"]))
