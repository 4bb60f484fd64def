"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (the CPU restatement of the reference).  May be imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs —
as the checker or the CPU baseline, never as the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
KAT_PATH = os.path.join(_HERE, "oracle_kat")

_vp = C.c_void_p
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-j4"], check=True, capture_output=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_deform_points.restype = C.c_int
    lib.oracle_deform_points.argtypes = [
        C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint32, C.c_double,
        _vp, C.c_uint64, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp]
    lib.oracle_dgraph_create.restype = _vp
    lib.oracle_dgraph_destroy.argtypes = [_vp]
    lib.oracle_dgraph_set_controls.argtypes = [_vp, C.c_char, C.c_uint64, _vp, _vp, _vp, _vp]
    lib.oracle_dgraph_set_recalculate.argtypes = [_vp, C.c_int]
    lib.oracle_dgraph_deform.restype = C.c_int64
    lib.oracle_dgraph_deform.argtypes = [_vp, C.c_char, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint32,
                                         C.c_double, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_uint32,
                                         _vp]
    lib.oracle_comp_create.restype = _vp
    lib.oracle_comp_create.argtypes = [C.c_int, C.c_double]
    lib.oracle_comp_destroy.argtypes = [_vp]
    lib.oracle_comp_integrate.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_double, _vp,
                                          _vp, _vp, _vp]
    lib.oracle_comp_prune.argtypes = [_vp, C.c_double]
    lib.oracle_comp_clear_blocks.argtypes = [_vp, C.c_uint32, _vp]
    lib.oracle_comp_sizes.argtypes = [_vp, _vp]
    lib.oracle_comp_snapshot.argtypes = [_vp] + [_vp] * 11
    lib.oracle_comp_remap_lookup.restype = C.c_int64
    lib.oracle_comp_remap_lookup.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint64]
    lib.oracle_comp_remap_block_size.restype = C.c_int64
    lib.oracle_comp_remap_block_size.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32]
    lib.oracle_delta_create.restype = _vp
    lib.oracle_delta_create.argtypes = [C.c_double]
    lib.oracle_delta_destroy.argtypes = [_vp]
    lib.oracle_delta_update.restype = C.c_int
    lib.oracle_delta_update.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp]
    lib.oracle_delta_update_labeled.restype = C.c_int
    lib.oracle_delta_update_labeled.argtypes = [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp]
    lib.oracle_delta_fetch_labels.restype = C.c_uint64
    lib.oracle_delta_fetch_labels.argtypes = [_vp, _vp]
    lib.oracle_delta_archive_by_time.argtypes = [_vp, C.c_uint64]
    lib.oracle_delta_archive_blocks.argtypes = [_vp, C.c_uint32, _vp]
    lib.oracle_delta_sizes.argtypes = [_vp, _vp]
    lib.oracle_delta_fetch.argtypes = [_vp] * 10
    lib.oracle_msg_to_cloud.restype = None
    lib.oracle_msg_to_cloud.argtypes = [C.c_uint64] + [_vp] * 6
    lib.oracle_cloud_to_msg.restype = None
    lib.oracle_cloud_to_msg.argtypes = [C.c_uint64] + [_vp] * 6
    lib.oracle_voxblox_decode.restype = None
    lib.oracle_voxblox_decode.argtypes = [C.c_uint32] + [_vp] * 8 + [C.c_float, _vp, _vp]
    lib.oracle_m2g_create.restype = _vp
    lib.oracle_m2g_destroy.argtypes = [_vp]
    lib.oracle_m2g_update.restype = None
    lib.oracle_m2g_update.argtypes = [_vp, C.c_uint32] + [_vp] * 6
    lib.oracle_m2g_size.restype = C.c_uint64
    lib.oracle_m2g_size.argtypes = [_vp]
    lib.oracle_m2g_fetch.restype = None
    lib.oracle_m2g_fetch.argtypes = [_vp, _vp, _vp]
    lib.oracle_m2g_block_mapping.restype = C.c_int64
    lib.oracle_m2g_block_mapping.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int, _vp, C.c_uint64]
    lib.oracle_stamp_from_sec.restype = C.c_uint64
    lib.oracle_stamp_from_sec.argtypes = [C.c_double]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_vp)


def stamp_from_sec(sec):
    return int(load().oracle_stamp_from_sec(float(sec)))


def deform_points(x, y, z, stamps_ns, ctrl_xyz, ctrl_stamps, valid, poses12, k=4, tol_t=10.0,
                  subset=None, want_trace=True, assoc_stride=8, out=None):
    """deformation::deformPoints on the CPU restatement.  Same result dict as
    kimera_pgmo_b200.capi.DeformHandle.deform_points plus `elapsed_s`."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    z = np.ascontiguousarray(z, dtype=np.float32)
    n_v = x.shape[0]
    st = None if stamps_ns is None else np.ascontiguousarray(stamps_ns, dtype=np.uint64)
    pos = np.ascontiguousarray(ctrl_xyz, dtype=np.float64).reshape(-1, 3)
    n_c = pos.shape[0]
    cst = None if ctrl_stamps is None else np.ascontiguousarray(ctrl_stamps, dtype=np.uint64)
    va = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    po = np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.uint64)
    n_pos = n_v if sub is None else sub.shape[0]
    if out is None:
        ox, oy, oz = x.copy(), y.copy(), z.copy()
    else:
        ox, oy, oz = out
    f64 = np.full((n_pos, 3), np.nan, dtype=np.float64) if want_trace else None
    assoc = np.full((n_pos, assoc_stride), -1, dtype=np.int32) if want_trace else None
    cnt = np.zeros(n_pos, dtype=np.int32) if want_trace else None
    processed = np.zeros(n_pos, dtype=np.uint8) if want_trace else None
    elapsed = C.c_double(0.0)
    rc = lib.oracle_deform_points(
        n_v, _ptr(x), _ptr(y), _ptr(z), _ptr(st), n_c, _ptr(pos), _ptr(cst), _ptr(va), _ptr(po), k,
        float(tol_t), _ptr(sub), 0 if sub is None else sub.shape[0], 1 if want_trace else 0,
        _ptr(ox), _ptr(oy), _ptr(oz), _ptr(f64), _ptr(assoc), assoc_stride, _ptr(cnt),
        _ptr(processed), C.cast(C.byref(elapsed), _vp))
    return dict(rc=rc, x=ox, y=oy, z=oz, f64=f64, assoc=assoc, assoc_cnt=cnt, processed=processed,
                elapsed_s=elapsed.value)


class GraphPath:
    """DeformationGraph mesh path (start index / cache / predeform), deformation_graph.h:715-870."""

    def __init__(self):
        self._lib = load()
        self._h = _vp(self._lib.oracle_dgraph_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_dgraph_destroy(self._h)
            self._h = None

    def set_controls(self, prefix, ctrl_xyz, ctrl_stamps, valid, poses12):
        pos = np.ascontiguousarray(ctrl_xyz, dtype=np.float64).reshape(-1, 3)
        cst = np.ascontiguousarray(ctrl_stamps, dtype=np.uint64)
        va = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        po = np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
        self._lib.oracle_dgraph_set_controls(self._h, prefix.encode(), pos.shape[0], _ptr(pos),
                                             _ptr(cst), _ptr(va), _ptr(po))

    def set_recalculate(self, v):
        self._lib.oracle_dgraph_set_recalculate(self._h, 1 if v else 0)

    def deform(self, prefix, x, y, z, stamps_ns, k, tol_t, graph_indices=None, hint=-1,
               assoc_stride=8):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.float32)
        st = np.ascontiguousarray(stamps_ns, dtype=np.uint64)
        gi = None if graph_indices is None else np.ascontiguousarray(graph_indices, dtype=np.int32)
        n_v = x.shape[0]
        ox, oy, oz = x.copy(), y.copy(), z.copy()
        assoc = np.full((n_v, assoc_stride), -1, dtype=np.int32)
        cnt = np.zeros(n_v, dtype=np.int32)
        start = self._lib.oracle_dgraph_deform(self._h, prefix.encode(), n_v, _ptr(x), _ptr(y),
                                               _ptr(z), _ptr(st), k, float(tol_t), _ptr(gi), hint,
                                               _ptr(ox), _ptr(oy), _ptr(oz), _ptr(assoc),
                                               assoc_stride, _ptr(cnt))
        return dict(start=int(start), x=ox, y=oy, z=oz, assoc=assoc, assoc_cnt=cnt)


class Compression:
    """MeshCompression family: kind 'voxel_clearing' | 'octree' | 'block'."""

    def __init__(self, kind, resolution):
        self._lib = load()
        self.kind = kind
        self._h = _vp(self._lib.oracle_comp_create({"voxel_clearing": 0, "octree": 1, "block": 2}[kind],
                                                   float(resolution)))
        self.elapsed_s = 0.0

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_comp_destroy(self._h)
            self._h = None

    def integrate(self, block_idx, offsets, xyz, rgba, stamp_s):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        pts = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        col = None if rgba is None else np.ascontiguousarray(rgba, dtype=np.uint32)
        n = pts.shape[0]
        idx = np.full(n, -1, dtype=np.int64)
        kind = np.zeros(n, dtype=np.uint8)
        keys = np.zeros((n, 3), dtype=np.int64)
        elapsed = C.c_double(0.0)
        self._lib.oracle_comp_integrate(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(pts),
                                        _ptr(col), float(stamp_s), _ptr(idx), _ptr(kind),
                                        _ptr(keys), C.cast(C.byref(elapsed), _vp))
        self.elapsed_s = elapsed.value
        return idx, kind, keys

    def prune(self, earliest_s):
        self._lib.oracle_comp_prune(self._h, float(earliest_s))

    def clear_blocks(self, block_idx):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        self._lib.oracle_comp_clear_blocks(self._h, bi.shape[0], _ptr(bi))

    def sizes(self):
        s = np.zeros(8, dtype=np.uint64)
        self._lib.oracle_comp_sizes(self._h, _ptr(s))
        return [int(v) for v in s]

    def snapshot(self):
        s = self.sizes()
        out = dict(
            vertices=np.zeros((s[0], 3), dtype=np.float32), rgba=np.zeros(s[0], dtype=np.uint32),
            stamps=np.zeros(s[0], dtype=np.uint64), polygons=np.zeros((s[1], 3), dtype=np.uint32),
            active=np.zeros(s[2], dtype=np.uint64), invalid=np.zeros(s[3], dtype=np.uint64),
            new_vertices=np.zeros((s[4], 3), dtype=np.float32),
            new_triangles=np.zeros((s[5], 3), dtype=np.uint32),
            new_indices=np.zeros(s[6], dtype=np.uint64),
            active_xyz=np.zeros((s[7], 3), dtype=np.float32),
            active_stamps=np.zeros(s[7], dtype=np.float64))
        self._lib.oracle_comp_snapshot(
            self._h, _ptr(out["vertices"]), _ptr(out["rgba"]), _ptr(out["stamps"]),
            _ptr(out["polygons"]), _ptr(out["active"]), _ptr(out["invalid"]),
            _ptr(out["new_vertices"]), _ptr(out["new_triangles"]), _ptr(out["new_indices"]),
            _ptr(out["active_xyz"]), _ptr(out["active_stamps"]))
        return out

    def remap_lookup(self, block, local):
        return int(self._lib.oracle_comp_remap_lookup(self._h, int(block[0]), int(block[1]),
                                                      int(block[2]), int(local)))

    def remap_block_size(self, block):
        return int(self._lib.oracle_comp_remap_block_size(self._h, int(block[0]), int(block[1]),
                                                          int(block[2])))


def run_kat():
    """Runs the known-answer tests (reference unit-test scenarios); returns (rc, output)."""
    if not os.path.exists(KAT_PATH):
        build()
    p = subprocess.run([KAT_PATH], capture_output=True, text=True)
    return p.returncode, p.stdout + p.stderr


def voxblox_decode(block_idx, offsets, x16, y16, z16, r, g, b, block_edge_length):
    """VoxbloxMsgInterface::getActiveVertex for every vertex of every block -> (xyz f32, rgba)"""
    lib = load()
    bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = int(off[-1]) if off.shape[0] else 0
    arrs = [np.ascontiguousarray(a, dtype=np.uint16) for a in (x16, y16, z16)]
    cols = [np.ascontiguousarray(a, dtype=np.uint8) for a in (r, g, b)]
    xyz = np.zeros((n, 3), dtype=np.float32)
    rgba = np.zeros(n, dtype=np.uint32)
    lib.oracle_voxblox_decode(bi.shape[0], _ptr(bi), _ptr(off), *[_ptr(a) for a in arrs],
                              *[_ptr(a) for a in cols], float(block_edge_length), _ptr(xyz), _ptr(rgba))
    return xyz, rgba


def msg_to_cloud(vertices_xyz, colors_rgba, stamps_sec_nsec):
    """conversions::fillFromMsg per vertex -> (xyz f32, rgba u32, stamps_ns u64 | None)"""
    lib = load()
    v = np.ascontiguousarray(vertices_xyz, dtype=np.float64).reshape(-1, 3)
    c = None if colors_rgba is None else np.ascontiguousarray(colors_rgba, dtype=np.float32).reshape(-1, 4)
    t = None if stamps_sec_nsec is None else np.ascontiguousarray(stamps_sec_nsec, dtype=np.uint32).reshape(-1, 2)
    n = v.shape[0]
    xyz = np.zeros((n, 3), dtype=np.float32)
    rgba = np.zeros(n, dtype=np.uint32)
    ns = np.zeros(n, dtype=np.uint64) if t is not None else None
    lib.oracle_msg_to_cloud(n, _ptr(v), _ptr(c), _ptr(t), _ptr(xyz), _ptr(rgba), _ptr(ns))
    return xyz, rgba, ns


def cloud_to_msg(xyz, rgba, stamps_ns):
    """conversions::fillMsg per vertex -> (vertices f64, colors f32 x 4, stamps {sec, nsec} | None)"""
    lib = load()
    p = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    col = np.ascontiguousarray(rgba, dtype=np.uint32)
    ns = None if stamps_ns is None else np.ascontiguousarray(stamps_ns, dtype=np.uint64)
    n = p.shape[0]
    v = np.zeros((n, 3), dtype=np.float64)
    c = np.zeros((n, 4), dtype=np.float32)
    t = np.zeros((n, 2), dtype=np.uint32) if ns is not None else None
    lib.oracle_cloud_to_msg(n, _ptr(p), _ptr(col), _ptr(ns), _ptr(v), _ptr(c), _ptr(t))
    return v, c, t


class DeltaCompression:
    """kimera_pgmo::DeltaCompression (insertion-ordered containers, see oracle/delta_oracle.h)."""

    def __init__(self, resolution):
        self._lib = load()
        self._h = _vp(self._lib.oracle_delta_create(float(resolution)))
        self.elapsed_s = 0.0

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_delta_destroy(self._h)
            self._h = None

    def labels(self):
        """(label per delta vertex, 0xFFFFFFFF = none; len(semantic_updates))"""
        n = self.sizes()[2]
        out = np.zeros(n, dtype=np.uint32)
        k = self._lib.oracle_delta_fetch_labels(self._h, _ptr(out))
        return out, int(k)

    def update(self, block_idx, offsets, xyz, rgba, stamp_ns, labels=None):
        """-> (remap per input vertex, delta dict)"""
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        pts = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        col = None if rgba is None else np.ascontiguousarray(rgba, dtype=np.uint32)
        remap = np.zeros(pts.shape[0], dtype=np.uint64)
        elapsed = C.c_double(0.0)
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.uint32)
        rc = self._lib.oracle_delta_update_labeled(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(pts),
                                                   _ptr(col), _ptr(lab), int(stamp_ns), _ptr(remap),
                                                   C.cast(C.byref(elapsed), _vp))
        assert rc == 0, "the reference would have thrown"
        self.elapsed_s = elapsed.value
        return remap, self.delta()

    def archive_by_time(self, earliest_ns):
        self._lib.oracle_delta_archive_by_time(self._h, int(earliest_ns))

    def archive_blocks(self, block_idx):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        self._lib.oracle_delta_archive_blocks(self._h, bi.shape[0], _ptr(bi))

    def sizes(self):
        s = np.zeros(12, dtype=np.uint64)
        self._lib.oracle_delta_sizes(self._h, _ptr(s))
        return [int(v) for v in s]

    def delta(self):
        s = self.sizes()
        d = dict(vertex_start=s[0], face_start=s[1], num_archived_vertices=s[3],
                 vertices=np.zeros((s[2], 3), np.float32), rgba=np.zeros(s[2], np.uint32),
                 stamps=np.zeros(s[2], np.uint64), faces=np.zeros((s[4], 3), np.uint32),
                 faces_archive=np.zeros((s[5], 3), np.uint32), prev_to_curr=np.zeros((s[6], 2), np.uint64),
                 deleted=np.zeros(s[7], np.uint64), observed=np.zeros(s[8], np.uint64),
                 new=np.zeros(s[9], np.uint64), active_voxels=s[10], pending_archived_faces=s[11])
        self._lib.oracle_delta_fetch(self._h, _ptr(d["vertices"]), _ptr(d["rgba"]), _ptr(d["stamps"]),
                                     _ptr(d["faces"]), _ptr(d["faces_archive"]), _ptr(d["prev_to_curr"]),
                                     _ptr(d["deleted"]), _ptr(d["observed"]), _ptr(d["new"]))
        return d


class MeshToGraph:
    """mesh -> graph index bookkeeping of MeshFrontendInterface (oracle/frontend_oracle.h)."""

    def __init__(self):
        self._lib = load()
        self._h = _vp(self._lib.oracle_m2g_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_m2g_destroy(self._h)
            self._h = None

    def update(self, block_idx, offsets, mesh_index, mesh_kind, graph_index, graph_kind):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        mi = np.ascontiguousarray(mesh_index, dtype=np.int64)
        gi = np.ascontiguousarray(graph_index, dtype=np.int64)
        mk = None if mesh_kind is None else np.ascontiguousarray(mesh_kind, dtype=np.uint8)
        gk = None if graph_kind is None else np.ascontiguousarray(graph_kind, dtype=np.uint8)
        self._lib.oracle_m2g_update(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(mi), _ptr(mk),
                                    _ptr(gi), _ptr(gk))

    def mesh_to_graph(self):
        """-> (mesh indices ascending, graph indices)"""
        n = int(self._lib.oracle_m2g_size(self._h))
        m, g = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self._lib.oracle_m2g_fetch(self._h, _ptr(m), _ptr(g))
        return m, g

    def block_mapping(self, block, which):
        """dense msg_to_mesh_idx_[block] (which 0) / msg_to_graph_idx_[block] (which 1), -1 = no entry"""
        n = int(self._lib.oracle_m2g_block_mapping(self._h, int(block[0]), int(block[1]), int(block[2]),
                                                   int(which), None, 0))
        if n < 0:
            return None
        out = np.full(n, -1, np.int64)
        self._lib.oracle_m2g_block_mapping(self._h, int(block[0]), int(block[1]), int(block[2]),
                                           int(which), _ptr(out), n)
        return out
