"""Disk / wire formats at either side of the mesh path (SURVEY.md 8f "next" row 4), host side.

* PLY as kimera_pgmo::IOData writes and reads it (src/utils/mesh_io.cpp:14-94, through happly):
  element vertex with float x y z, optional uchar red green blue alpha, optional uint secs nsecs
  (stamp = stampFromSec(secs) + nsecs on load, :78-80), optional uint label; element face with a
  uchar-counted uint list vertex_indices.  The writer emits ASCII like the reference
  (happly::DataFormat::ASCII, :62); the reader also takes binary_little_endian files.
* .dgrf deformation-graph dumps (src/deformation_graph_io.cpp:60-170, :194-349): the lines this
  path needs — `VERTEX key stamp x y z` (vertex_positions_ / vertex_stamps_ of the key's prefix)
  and `NODE key x y z qx qy qz qw` (the optimised values) — become the control table of
  pgmo_deform_set_controls.
* KimeraPgmoMesh message arrays: mesh_to_msg_arrays / msg_arrays_to_mesh restate
  conversions::fillMsg / fillFromMsg in numpy for the tests; the product converts on the GPU
  (pgmo_deform_msg_to_cloud / pgmo_deform_cloud_to_msg).
"""
from __future__ import annotations

import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
              "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2",
              "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def read_ply(path):
    """-> dict(vertices f32 [n,3], faces u32 [m,3], rgba u8 [n,4] | None, stamps u64 [n] | None,
    labels u32 [n] | None)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    header = data[:end].decode("ascii").split("\n")
    fmt = None
    elements = []  # (name, count, [(prop, type) | (prop, ("list", count_type, item_type))])
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append((tok[4], ("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
            else:
                elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
    body = data[end:]
    columns = {}
    if fmt == "ascii":
        tokens = body.split()
        pos = 0
        for name, count, props in elements:
            cols = {p: [] for p, _ in props}
            for _ in range(count):
                for p, t in props:
                    if isinstance(t, tuple):
                        k = int(tokens[pos])
                        cols[p].append([int(v) for v in tokens[pos + 1:pos + 1 + k]])
                        pos += 1 + k
                    else:
                        cols[p].append(float(tokens[pos]) if t[0] == "f" else int(tokens[pos]))
                        pos += 1
            columns[name] = {p: (cols[p] if isinstance(t, tuple) else np.array(cols[p], dtype=t))
                             for p, t in props}
    elif fmt == "binary_little_endian":
        off = 0
        for name, count, props in elements:
            if all(not isinstance(t, tuple) for _, t in props):
                dt = np.dtype([(p, "<" + t) for p, t in props])
                arr = np.frombuffer(body, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                columns[name] = {p: arr[p].copy() for p, _ in props}
            else:
                cols = {p: [] for p, _ in props}
                for _ in range(count):
                    for p, t in props:
                        if isinstance(t, tuple):
                            k = int(np.frombuffer(body, dtype="<" + t[1], count=1, offset=off)[0])
                            off += np.dtype(t[1]).itemsize
                            cols[p].append(np.frombuffer(body, dtype="<" + t[2], count=k, offset=off).tolist())
                            off += np.dtype(t[2]).itemsize * k
                        else:
                            cols[p].append(np.frombuffer(body, dtype="<" + t, count=1, offset=off)[0])
                            off += np.dtype(t).itemsize
                columns[name] = cols
    else:
        raise ValueError(f"unsupported PLY format {fmt!r}")
    v = columns["vertex"]
    n = len(v["x"])
    out = dict(vertices=np.stack([np.asarray(v[a], dtype=np.float32) for a in "xyz"], axis=1)
               if n else np.zeros((0, 3), np.float32), rgba=None, stamps=None, labels=None)
    if all(k in v for k in ("red", "green", "blue", "alpha")):  # mesh_io.cpp:68-74: all four or none
        out["rgba"] = np.stack([np.asarray(v[k], dtype=np.uint8) for k in ("red", "green", "blue", "alpha")], 1)
    if "secs" in v and "nsecs" in v:
        # stampFromSec(double(sec)) + nsec (:78-80); exact for every uint32 second count
        out["stamps"] = (np.asarray(v["secs"], dtype=np.uint64) * np.uint64(1_000_000_000)
                         + np.asarray(v["nsecs"], dtype=np.uint64))
    if "label" in v:
        out["labels"] = np.asarray(v["label"], dtype=np.uint32)
    faces = columns.get("face", {}).get("vertex_indices", [])
    out["faces"] = (np.array(faces, dtype=np.uint32).reshape(-1, 3) if len(faces)
                    else np.zeros((0, 3), np.uint32))
    return out


def write_ply(path, vertices, faces, rgba=None, stamps=None, labels=None):
    """IOData::save (mesh_io.cpp:14-63): ASCII; stamps split into uint secs / nsecs."""
    vertices = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)
    faces = np.asarray(faces, dtype=np.uint32).reshape(-1, 3)
    n = vertices.shape[0]
    lines = ["ply", "format ascii 1.0", f"element vertex {n}", "property float x",
             "property float y", "property float z"]
    cols = [vertices[:, 0], vertices[:, 1], vertices[:, 2]]
    if rgba is not None:
        rgba = np.asarray(rgba, dtype=np.uint8).reshape(-1, 4)
        lines += [f"property uchar {c}" for c in ("red", "green", "blue", "alpha")]
        cols += [rgba[:, 0], rgba[:, 1], rgba[:, 2], rgba[:, 3]]
    if stamps is not None and len(stamps):
        st = np.asarray(stamps, dtype=np.uint64)
        lines += ["property uint secs", "property uint nsecs"]
        cols += [(st // np.uint64(1_000_000_000)).astype(np.uint32),
                 (st % np.uint64(1_000_000_000)).astype(np.uint32)]
    if labels is not None and len(labels):
        lines.append("property uint label")
        cols.append(np.asarray(labels, dtype=np.uint32))
    lines += [f"element face {faces.shape[0]}", "property list uchar uint vertex_indices", "end_header"]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
        for i in range(n):
            f.write(" ".join(repr(float(c[i])) if c.dtype.kind == "f" else str(int(c[i])) for c in cols) + "\n")
        for tri in faces:
            f.write(f"3 {int(tri[0])} {int(tri[1])} {int(tri[2])}\n")


def quat_to_rot(qx, qy, qz, qw):
    """Eigen::Quaterniond::toRotationMatrix (what gtsam::Rot3::Quaternion(w, x, y, z) stores),
    operation for operation; row-major 9 doubles."""
    tx, ty, tz = 2.0 * qx, 2.0 * qy, 2.0 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    return [1.0 - (tyy + tzz), txy - twz, txz + twy,
            txy + twz, 1.0 - (txx + tzz), tyz - twx,
            txz - twy, tyz + twx, 1.0 - (txx + tyy)]


def load_dgrf(path, prefix: str):
    """Control table of `prefix` from a DeformationGraph::save dump: (ctrl_xyz f64 [n,3],
    ctrl_stamps u64 [n], valid u8 [n], poses12 f64 [n,12]).  VERTEX lines give positions / stamps
    in index order (deformation_graph_io.cpp:44-50, :300-320); a NODE line with the same key gives
    the optimised pose (valid = values.exists(Symbol(prefix, j)))."""
    want = ord(prefix)
    verts, nodes = {}, {}
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "VERTEX":
                key = int(tok[1])
                if key >> 56 == want:
                    verts[key & ((1 << 56) - 1)] = (int(tok[2]), [float(t) for t in tok[3:6]])
            elif tok[0] == "NODE":
                key = int(tok[1])
                if key >> 56 == want:
                    nodes[key & ((1 << 56) - 1)] = [float(t) for t in tok[2:9]]
    n = (max(verts) + 1) if verts else 0
    xyz = np.zeros((n, 3), dtype=np.float64)
    stamps = np.zeros(n, dtype=np.uint64)
    valid = np.zeros(n, dtype=np.uint8)
    poses = np.zeros((n, 12), dtype=np.float64)
    for j, (st, p) in verts.items():
        xyz[j] = p
        stamps[j] = st
    for j, q in nodes.items():
        if j < n:
            valid[j] = 1
            poses[j, :9] = quat_to_rot(q[3], q[4], q[5], q[6])
            poses[j, 9:] = q[:3]
    return xyz, stamps, valid, poses


def mesh_to_msg_arrays(vertices, rgba=None, stamps=None):
    """conversions::fillMsg in numpy: (vertices f64 [n,3], colors f32 [n,4] | None,
    stamps {sec, nsec} u32 [n,2] | None)."""
    v = np.asarray(vertices, dtype=np.float32).astype(np.float64)
    c = None
    if rgba is not None:
        c = (np.float32(1.0) / np.float32(255.0)) * np.asarray(rgba, dtype=np.uint8).astype(np.float32)
    t = None
    if stamps is not None:
        st = np.asarray(stamps, dtype=np.uint64)
        t = np.stack([(st // np.uint64(1_000_000_000)).astype(np.uint32),
                      (st % np.uint64(1_000_000_000)).astype(np.uint32)], axis=1)
    return v, c, t
