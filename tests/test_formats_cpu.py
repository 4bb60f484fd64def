"""CPU: disk / wire formats (kimera_pgmo_b200/formats.py) and the oracle's restatement of the
KimeraPgmoMesh conversions, against the reference's own fixtures (frozen in
tests/golden/reference_meshes.npz by make_reference_meshes.py) and the values its tests expect."""
import json
import os

import numpy as np

from kimera_pgmo_b200 import formats

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _gold():
    return np.load(os.path.join(GOLD, "reference_meshes.npz")), json.load(open(os.path.join(GOLD, "reference_mesh_kat.json")))


def test_ply_round_trip_like_io_data(tmp_path):
    g, kat = _gold()
    v, f = g["sphere_vertices"], g["sphere_faces"]
    assert v.shape == (kat["mesh_conversion"]["num_vertices"], 3) and f.shape == (840, 3)
    assert g["cube_vertices"].shape == (8, 3) and g["cube_faces"].shape == (12, 3)
    stamps = np.uint64(1_700_000_000_000_000_000) + np.arange(v.shape[0], dtype=np.uint64) * np.uint64(1_234_567_891)
    rgba = (np.arange(v.shape[0] * 4) % 251).astype(np.uint8).reshape(-1, 4)
    labels = np.arange(v.shape[0], dtype=np.uint32) * 3
    path = tmp_path / "mesh.ply"
    formats.write_ply(path, v, f, rgba=rgba, stamps=stamps, labels=labels)
    head = open(path).read().split("end_header")[0]
    for prop in ("property float x", "property uchar alpha", "property uint secs", "property uint nsecs",
                 "property uint label", "property list uchar uint vertex_indices", "format ascii 1.0"):
        assert prop in head  # mesh_io.cpp:27-62
    back = formats.read_ply(path)
    for key, want in (("vertices", v), ("faces", f), ("rgba", rgba), ("stamps", stamps), ("labels", labels)):
        np.testing.assert_array_equal(back[key], want)
    # colours are all-or-nothing (mesh_io.cpp:68-74), stamps need both columns
    formats.write_ply(path, v[:3], f[:0])
    bare = formats.read_ply(path)
    assert bare["rgba"] is None and bare["stamps"] is None and bare["faces"].shape == (0, 3)


def test_binary_little_endian_ply(tmp_path):
    v = np.array([[0.5, -1.25, 3.0], [1e-3, 2.0, -7.5]], dtype=np.float32)
    path = tmp_path / "b.ply"
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\n"
              "property float z\nproperty uint secs\nproperty uint nsecs\nelement face 1\n"
              "property list uchar uint vertex_indices\nend_header\n").encode()
    rec = np.zeros(2, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("s", "<u4"), ("n", "<u4")])
    rec["x"], rec["y"], rec["z"], rec["s"], rec["n"] = v[:, 0], v[:, 1], v[:, 2], [10, 4_000_000_000], [5, 999_999_999]
    face = bytes([3]) + np.array([0, 1, 0], dtype="<u4").tobytes()
    open(path, "wb").write(header + rec.tobytes() + face)
    m = formats.read_ply(path)
    np.testing.assert_array_equal(m["vertices"], v)
    assert m["stamps"].tolist() == [10 * 10**9 + 5, 4_000_000_000 * 10**9 + 999_999_999]
    assert m["faces"].tolist() == [[0, 1, 0]]


def test_dgrf_control_table(tmp_path):
    g, kat = _gold()
    path = tmp_path / "graph.dgrf"
    open(path, "wb").write(g["dgrf_text"].tobytes())
    xyz, stamps, valid, poses = formats.load_dgrf(path, "v")
    np.testing.assert_array_equal(xyz, np.array(kat["dgrf"]["v_positions"], dtype=np.float64))
    assert valid.tolist() == [1, 1, 1] and stamps.tolist() == [0, 0, 0]
    np.testing.assert_array_equal(poses[:, :9], np.tile(np.eye(3).ravel(), (3, 1)))
    np.testing.assert_array_equal(poses[:, 9:], xyz)  # the saved graph is undeformed
    np.testing.assert_array_equal(xyz, g["dgrf_v_xyz"])
    assert formats.load_dgrf(path, "a")[0].shape[0] == 0  # pose nodes have no VERTEX lines
    # a rotated node: Eigen's quaternion -> matrix, 90 degrees about z
    s = np.sqrt(0.5)
    R = np.array(formats.quat_to_rot(0.0, 0.0, s, s)).reshape(3, 3)
    np.testing.assert_allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)


def test_mesh_conversion_kat_through_the_oracle(oracle):
    """test_mesh_conversion.cpp:24-48: sphere.ply -> KimeraPgmoMesh -> back."""
    g, kat = _gold()
    v = g["sphere_vertices"]
    stamps = np.uint64(1_600_000_000_123_456_789) + np.arange(v.shape[0], dtype=np.uint64) * np.uint64(999_999_937)
    rgba8 = (np.arange(v.shape[0] * 4) * 7 % 256).astype(np.uint8).reshape(-1, 4)
    packed = (rgba8[:, 3].astype(np.uint32) << 24) | (rgba8[:, 0].astype(np.uint32) << 16) | \
             (rgba8[:, 1].astype(np.uint32) << 8) | rgba8[:, 2]
    mv, mc, mt = oracle.cloud_to_msg(v, packed, stamps)
    nv, nc, nt = formats.mesh_to_msg_arrays(v, rgba8, stamps)
    np.testing.assert_array_equal(mv, nv)
    np.testing.assert_array_equal(mc, nc)
    np.testing.assert_array_equal(mt, nt)
    xyz, rgba, ns = oracle.msg_to_cloud(mv, mc, mt)
    k = kat["mesh_conversion"]
    assert xyz.shape[0] == k["num_vertices"]
    for i, comp in k["points"].items():
        for axis, val in comp.items():
            assert xyz[int(i), "xyz".index(axis)] == val
    np.testing.assert_array_equal(xyz, v)
    np.testing.assert_array_equal(ns, stamps)
    # 255.0f * ((1.0f / 255) * byte) truncates back to the byte for every byte value but a few:
    # the reference's round trip is not the identity on colours — the oracle reproduces it
    conv = np.float32(255.0) * ((np.float32(1.0) / np.float32(255.0)) * np.arange(256, dtype=np.float32))
    table = conv.astype(np.uint8)  # truncation
    want = (table[rgba8[:, 3]].astype(np.uint32) << 24) | (table[rgba8[:, 0]].astype(np.uint32) << 16) | \
           (table[rgba8[:, 1]].astype(np.uint32) << 8) | table[rgba8[:, 2]]
    np.testing.assert_array_equal(rgba, want)
    # no colours in the message: PointXYZRGBA's default
    assert (oracle.msg_to_cloud(mv, None, None)[1] == 0xFF000000).all()
