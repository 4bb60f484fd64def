"""CPU: the C++ PLY / .dgrf readers and the prefix tables of include/kimera_pgmo_b200/io.h,
against files written by kimera_pgmo_b200/formats.py from the reference's frozen fixtures, and
back (what the C++ side writes is read by the Python reader)."""
import os
import subprocess

import numpy as np

from kimera_pgmo_b200 import formats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_readers_and_prefix_tables(tmp_path, capi):
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_meshes.npz"))
    v, f = g["sphere_vertices"], g["sphere_faces"]
    n = v.shape[0]
    stamps = np.uint64(1_700_000_000_000_000_000) + np.arange(n, dtype=np.uint64) * np.uint64(1_234_567_891)
    rgba = (np.arange(n * 4) % 251).astype(np.uint8).reshape(-1, 4)
    labels = np.arange(n, dtype=np.uint32) * 3
    ascii_ply, binary_ply, dgrf, out_ply = (str(tmp_path / name) for name in
                                            ("a.ply", "b.ply", "graph.dgrf", "out.ply"))
    formats.write_ply(ascii_ply, v, f, rgba=rgba, stamps=stamps, labels=labels)
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\n"
              "property float z\nproperty uint secs\nproperty uint nsecs\nelement face 1\n"
              "property list uchar uint vertex_indices\nend_header\n").encode()
    rec = np.zeros(2, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("s", "<u4"), ("n", "<u4")])
    rec["x"], rec["y"], rec["z"] = [0.5, 1e-3], [-1.25, 2.0], [3.0, -7.5]
    rec["s"], rec["n"] = [10, 4_000_000_000], [5, 999_999_999]
    open(binary_ply, "wb").write(header + rec.tobytes() + bytes([3]) + np.array([0, 1, 0], dtype="<u4").tobytes())
    open(dgrf, "wb").write(g["dgrf_text"].tobytes())

    exe = os.path.join(ROOT, "build", "tests", "test_io")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    lib_dir = os.path.join(ROOT, "kimera_pgmo_b200", "lib")
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread", "-I",
                        os.path.join(ROOT, "include"), "-o", exe,
                        os.path.join(ROOT, "tests", "cpp", "test_io.cpp"), "-L", lib_dir,
                        "-lpgmo_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe, ascii_ply, binary_ply, dgrf, out_ply], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr

    back = formats.read_ply(out_ply)  # written by WriteMeshWithStampsToPly
    np.testing.assert_array_equal(back["vertices"], v)
    np.testing.assert_array_equal(back["faces"], f)
    np.testing.assert_array_equal(back["rgba"], rgba)
    np.testing.assert_array_equal(back["stamps"], stamps)
    assert back["labels"] is None
