"""ctypes binding of libpgmo_b200.so (include/pgmo_b200.h).

Test / bench plumbing only: the product's host side is the C++ shim under
include/kimera_pgmo_b200/.  Everything here goes through the C ABI; there is no Python or CPU
implementation behind it, and loading fails loudly when the CUDA library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGMO_B200_LIB: developer override to compare builds (e.g. one with -DPGMO_DEBUG_COUNTERS)
LIB_PATH = os.environ.get("PGMO_B200_LIB") or os.path.join(_HERE, "lib", "libpgmo_b200.so")

PGMO_OK = 0
PGMO_NOOP = 1
PGMO_ERR_INVALID = -1
PGMO_ERR_CUDA = -2
PGMO_ERR_NO_DEVICE = -3
PGMO_ERR_UNSUPPORTED = -4
PGMO_ERR_OUT_OF_RANGE = -5
FP64 = 0
FP32 = 1

_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "pgmo_last_error": (C.c_char_p, []),
    "pgmo_version": (C.c_char_p, []),
    "pgmo_device_count": (C.c_int, []),
    "pgmo_stamp_from_sec": (C.c_uint64, [C.c_double]),
    "pgmo_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(_vp)]),
    "pgmo_host_free": (C.c_int, [_vp]),
    "pgmo_deform_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pgmo_deform_destroy": (None, [_vp]),
    "pgmo_deform_set_controls": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp]),
    "pgmo_deform_set_controls_device": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp]),
    "pgmo_deform_update_poses": (C.c_int, [_vp, _vp]),
    "pgmo_deform_update_poses_device": (C.c_int, [_vp, _vp, _vp]),
    "pgmo_deform_set_profiling": (C.c_int, [_vp, C.c_int]),
    "pgmo_deform_set_search_mode": (C.c_int, [_vp, C.c_int]),
    "pgmo_deform_drop_schedule": (C.c_int, [_vp]),
    "pgmo_deform_last_timings": (C.c_int, [_vp, _vp, _u64p]),
    "pgmo_deform_set_assoc_cache": (C.c_int, [_vp, C.c_int]),
    "pgmo_deform_pose_source_create": (C.c_int, [_vp, C.c_uint64, _vp, C.POINTER(_vp)]),
    "pgmo_deform_pose_source_attach": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int]),
    "pgmo_deform_pose_source_write": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "pgmo_deform_update_poses_from_source": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_uint64, _vp]),
    "pgmo_deform_set_scan_overlap": (C.c_int, [_vp, C.c_int]),
    "pgmo_deform_pose_source_publish": (C.c_int, [_vp, C.c_uint64, C.c_uint32, _vp, _vp]),
    "pgmo_deform_update_poses_from_source_seq": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, _vp]),
    "pgmo_deform_pose_source_timeouts": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "pgmo_deform_slab_control_range": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _u64p, _u64p]),
    "pgmo_deform_run_host_slab": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, C.c_int, _vp,
                                            _vp, _vp, _vp]),
    "pgmo_deform_last_step_info": (C.c_int, [_vp, _vp, _vp]),
    "pgmo_deform_cache_records": (C.c_int, [_vp, _vp, C.c_uint64, _u64p]),
    "pgmo_deform_group_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "pgmo_deform_group_destroy": (None, [_vp]),
    "pgmo_deform_group_size": (C.c_int, [_vp]),
    "pgmo_deform_group_set_controls": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp]),
    "pgmo_deform_group_update_poses": (C.c_int, [_vp, _vp]),
    "pgmo_deform_group_points": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint32, C.c_double,
                                           C.c_int, _vp, _vp, _vp, _vp]),
    "pgmo_deform_group_repeat": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "pgmo_deform_points": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, C.c_uint64,
                                     C.c_uint32, C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                     C.c_uint32, _vp, _vp]),
    "pgmo_deform_stamp_stretches_device": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, _vp, _vp,
                                                     _u64p, _vp]),
    "pgmo_deform_build_schedule": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_uint64, C.c_uint32,
                                             C.c_double]),
    "pgmo_deform_run_device": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, C.c_int,
                                         _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "pgmo_deform_schedule_info": (C.c_int, [_vp, _u64p, C.c_uint64, _vp]),
    "pgmo_deform_schedule_stats": (C.c_int, [_vp, _vp]),
    "pgmo_deform_msg_to_cloud": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pgmo_deform_cloud_to_msg": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pgmo_deform_mesh_msg": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_uint32, C.c_double, C.c_int,
                                       _vp, _vp]),
    "pgmo_deform_stamp_check_begin": (C.c_int, [_vp, C.c_uint64, _vp, _vp]),
    "pgmo_deform_stamp_check_finish": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "pgmo_deform_mesh_device": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_uint32, C.c_double,
                                          C.c_int, _vp, _vp, _vp, _vp]),
    "pgmo_vcc_device_view": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), _u64p]),
    "pgmo_oct_device_view": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), _u64p]),
    "pgmo_deform_predeform": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp,
                                        _vp, _vp, _u64p]),
    "pgmo_vcc_create": (C.c_int, [C.c_double, C.c_int, C.POINTER(_vp)]),
    "pgmo_vcc_destroy": (None, [_vp]),
    "pgmo_vcc_integrate": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp]),
    "pgmo_vcc_prune": (C.c_int, [_vp, C.c_double]),
    "pgmo_vcc_clear_blocks": (C.c_int, [_vp, C.c_uint32, _vp]),
    "pgmo_vcc_sizes": (C.c_int, [_vp, _vp]),
    "pgmo_vcc_snapshot": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pgmo_vcc_integrate_device": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_double, _vp,
                                            _vp]),
    "pgmo_oct_create": (C.c_int, [C.c_double, C.c_int, C.POINTER(_vp)]),
    "pgmo_oct_destroy": (None, [_vp]),
    "pgmo_oct_integrate": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp]),
    "pgmo_oct_prune": (C.c_int, [_vp, C.c_double]),
    "pgmo_oct_sizes": (C.c_int, [_vp, _vp]),
    "pgmo_oct_snapshot": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pgmo_oct_last_outputs": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "pgmo_oct_last_stats": (C.c_int, [_vp, _vp]),
    "pgmo_oct_create_block": (C.c_int, [C.c_double, C.c_int, C.POINTER(_vp)]),
    "pgmo_delta_create": (C.c_int, [C.c_double, C.c_int, C.POINTER(_vp)]),
    "pgmo_delta_destroy": (None, [_vp]),
    "pgmo_delta_update": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "pgmo_delta_update_labeled": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "pgmo_delta_fetch_labels": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "pgmo_delta_archive_by_time": (C.c_int, [_vp, C.c_uint64]),
    "pgmo_delta_archive_blocks": (C.c_int, [_vp, C.c_uint32, _vp]),
    "pgmo_delta_sizes": (C.c_int, [_vp, _vp]),
    "pgmo_delta_fetch": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64p, _vp, _u64p, _vp, _u64p,
                                   _vp, _u64p]),
    "pgmo_delta_last_stats": (C.c_int, [_vp, _vp]),
    "pgmo_delta_blocks": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _u64p]),
    "pgmo_m2g_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pgmo_m2g_destroy": (None, [_vp]),
    "pgmo_m2g_update": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pgmo_m2g_size": (C.c_int, [_vp, _u64p]),
    "pgmo_m2g_fetch": (C.c_int, [_vp, _vp, _vp]),
    "pgmo_m2g_block_mapping": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_uint64, _u64p]),
    "pgmo_m2g_device_view": (C.c_int, [_vp, C.POINTER(_vp), _u64p]),
    "pgmo_voxblox_decode": (C.c_int, [C.c_int, C.c_uint32] + [_vp] * 8 + [C.c_float, _vp, _vp]),
    "pgmo_vcc_integrate_voxblox": (C.c_int, [_vp, C.c_uint32] + [_vp] * 8 + [C.c_float, C.c_double,
                                                                          _vp, _vp]),
    "pgmo_oct_integrate_voxblox": (C.c_int, [_vp, C.c_uint32] + [_vp] * 8 + [C.c_float, C.c_double,
                                                                          _vp, _vp]),
}

_lib = None


class PgmoError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"pgmo_b200 error {code}: {message}")
        self.code = code


def load():
    """Loads libpgmo_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PGMO_LIB_PATH", LIB_PATH)  # build variants (scripts/, experiments)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc):
    if rc < 0:
        raise PgmoError(rc, load().pgmo_last_error().decode())
    return rc


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be contiguous"
        return a.ctypes.data_as(_vp)
    if isinstance(a, int):
        return _vp(a)
    if hasattr(a, "data_ptr"):  # torch tensor
        return _vp(a.data_ptr())
    raise TypeError(type(a))


def stamp_from_sec(sec: float) -> int:
    return int(load().pgmo_stamp_from_sec(float(sec)))


class DeformHandle:
    """pgmo_deform_t: deformation::deformPoints for one control table (one prefix)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = _vp()
        _check(self._lib.pgmo_deform_create(device, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgmo_deform_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_controls(self, pos_xyz, stamps_ns, valid, poses12):
        pos = np.ascontiguousarray(pos_xyz, dtype=np.float64).reshape(-1, 3)
        n_c = pos.shape[0]
        st = None if stamps_ns is None else np.ascontiguousarray(stamps_ns, dtype=np.uint64)
        va = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        po = None if poses12 is None else np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
        _check(self._lib.pgmo_deform_set_controls(self._h, n_c, _ptr(pos), _ptr(st), _ptr(va), _ptr(po)))

    def set_controls_device(self, n_c, d_pos, d_stamps, d_valid, d_poses, stream=0):
        _check(self._lib.pgmo_deform_set_controls_device(
            self._h, n_c, _ptr(d_pos), _ptr(d_stamps), _ptr(d_valid), _ptr(d_poses), _vp(stream)))

    def update_poses(self, poses12):
        po = np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
        _check(self._lib.pgmo_deform_update_poses(self._h, _ptr(po)))

    def update_poses_device(self, d_poses, stream=0):
        _check(self._lib.pgmo_deform_update_poses_device(self._h, _ptr(d_poses), _vp(stream)))

    def drop_schedule(self):
        _check(self._lib.pgmo_deform_drop_schedule(self._h))

    def set_search_mode(self, mode):
        _check(self._lib.pgmo_deform_set_search_mode(self._h, int(mode)))

    def set_profiling(self, enabled=True):
        _check(self._lib.pgmo_deform_set_profiling(self._h, 1 if enabled else 0))

    # --- pose source (multi-GPU pose replication by peer pull) -----------------------------------
    def pose_source_create(self, n_c):
        """-> (64-byte CUDA IPC handle, device base pointer)"""
        handle = np.zeros(64, dtype=np.uint8)
        base = _vp()
        _check(self._lib.pgmo_deform_pose_source_create(self._h, n_c, _ptr(handle), C.byref(base)))
        return handle, int(base.value)

    def pose_source_attach(self, n_c, ipc_handle=None, d_base=None, owner_device=0):
        _check(self._lib.pgmo_deform_pose_source_attach(
            self._h, n_c, None if ipc_handle is None else _ptr(np.ascontiguousarray(ipc_handle, dtype=np.uint8)),
            None if d_base is None else _vp(d_base), owner_device))

    def pose_source_write(self, buffer, poses12, stream=0):
        _check(self._lib.pgmo_deform_pose_source_write(self._h, buffer, _ptr(poses12), _vp(stream)))

    def pose_source_publish(self, seq, n_readers, poses12, stream=0):
        _check(self._lib.pgmo_deform_pose_source_publish(self._h, seq, n_readers, _ptr(poses12), _vp(stream)))

    def update_poses_from_source_seq(self, seq, reader, row_begin, row_end, stream=0):
        _check(self._lib.pgmo_deform_update_poses_from_source_seq(self._h, seq, reader, row_begin, row_end,
                                                                  _vp(stream)))

    def pose_source_timeouts(self):
        n = C.c_uint64(0)
        _check(self._lib.pgmo_deform_pose_source_timeouts(self._h, C.byref(n)))
        return int(n.value)

    def update_poses_from_source(self, buffer, row_begin, row_end, stream=0):
        _check(self._lib.pgmo_deform_update_poses_from_source(self._h, buffer, row_begin, row_end, _vp(stream)))

    def slab_control_range(self, first_position, n_local):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib.pgmo_deform_slab_control_range(self._h, first_position, n_local, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def run_host_slab(self, first_position, n_local, x, y, z, out, precision=FP64, processed=None):
        """pinned / pageable host arrays (numpy or CPU tensors) in and out, current schedule"""
        ox, oy, oz = out
        return _check(self._lib.pgmo_deform_run_host_slab(
            self._h, first_position, n_local, _ptr(x), _ptr(y), _ptr(z), precision, _ptr(ox), _ptr(oy),
            _ptr(oz), _ptr(processed)))

    def set_scan_overlap(self, enabled=True):
        _check(self._lib.pgmo_deform_set_scan_overlap(self._h, 1 if enabled else 0))

    def set_assoc_cache(self, enabled=True):
        _check(self._lib.pgmo_deform_set_assoc_cache(self._h, 1 if enabled else 0))

    def cache_records(self):
        """(n_slots, 8) uint16: the association records in search order (diagnostics)."""
        n = C.c_uint64(0)
        _check(self._lib.pgmo_deform_cache_records(self._h, None, 0, C.byref(n)))
        rec = np.zeros((int(n.value), 8), dtype=np.uint16)
        _check(self._lib.pgmo_deform_cache_records(self._h, _ptr(rec), rec.shape[0], C.byref(n)))
        return rec

    def last_step_info(self):
        """dict(search_ms, resolver_ms, blend_ms, total_ms, items, stale_items, resolved, warm)."""
        ms = np.zeros(4, dtype=np.float32)
        cnt = np.zeros(4, dtype=np.uint64)
        _check(self._lib.pgmo_deform_last_step_info(self._h, _ptr(ms), _ptr(cnt)))
        return dict(search_ms=float(ms[0]), resolver_ms=float(ms[1]), blend_ms=float(ms[2]),
                    total_ms=float(ms[3]), items=int(cnt[0]), stale_items=int(cnt[1]),
                    resolved=int(cnt[2]), warm=bool(cnt[3]))

    def last_timings(self):
        """(knn_blend_ms, resolver_ms, schedule_builds) of the last run_device."""
        ms = np.zeros(2, dtype=np.float32)
        builds = C.c_uint64(0)
        _check(self._lib.pgmo_deform_last_timings(self._h, _ptr(ms), C.byref(builds)))
        return float(ms[0]), float(ms[1]), int(builds.value)

    def deform_points(self, x, y, z, stamps_ns, k=4, tol_t=10.0, precision=FP64, subset=None,
                      out=None, want_f64=False, want_assoc=False, assoc_stride=8):
        """Host-buffer call.  Returns dict(rc, x, y, z, f64, assoc, assoc_cnt, processed)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.float32)
        n_v = x.shape[0]
        st = None if stamps_ns is None else np.ascontiguousarray(stamps_ns, dtype=np.uint64)
        sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.uint64)
        n_pos = n_v if sub is None else sub.shape[0]
        if out is None:
            ox, oy, oz = x.copy(), y.copy(), z.copy()
        else:
            ox, oy, oz = out
        f64 = np.full((n_pos, 3), np.nan, dtype=np.float64) if want_f64 else None
        assoc = np.full((n_pos, assoc_stride), 0xFFFFFFFF, dtype=np.uint32) if want_assoc else None
        cnt = np.zeros(n_pos, dtype=np.uint8) if want_assoc else None
        processed = np.zeros(n_pos, dtype=np.uint8)
        rc = _check(self._lib.pgmo_deform_points(
            self._h, n_v, _ptr(x), _ptr(y), _ptr(z), _ptr(st), _ptr(sub),
            0 if sub is None else sub.shape[0], k, float(tol_t), precision, _ptr(ox), _ptr(oy),
            _ptr(oz), _ptr(f64), _ptr(assoc), assoc_stride, _ptr(cnt), _ptr(processed)))
        return dict(rc=rc, x=ox, y=oy, z=oz, f64=f64, assoc=assoc, assoc_cnt=cnt,
                    processed=processed)

    # --- device-resident path -------------------------------------------------------------
    def stamp_stretches_device(self, d_stamps, n_local, stream=0):
        n = C.c_uint64(0)
        cap = 1 << 16
        while True:
            stamps = np.zeros(cap, dtype=np.uint64)
            lens = np.zeros(cap, dtype=np.uint64)
            _check(self._lib.pgmo_deform_stamp_stretches_device(
                self._h, n_local, _ptr(d_stamps), cap, _ptr(stamps), _ptr(lens), C.byref(n),
                _vp(stream)))
            if n.value <= cap:
                return stamps[: n.value].copy(), lens[: n.value].copy()
            cap = int(n.value)

    def stamp_check_begin(self, d_stamps, n_local, stream=0):
        _check(self._lib.pgmo_deform_stamp_check_begin(self._h, n_local, _ptr(d_stamps), _vp(stream)))

    def stamp_check_finish(self):
        changed = C.c_int(1)
        _check(self._lib.pgmo_deform_stamp_check_finish(self._h, C.byref(changed)))
        return bool(changed.value)

    def build_schedule(self, stretch_stamp, stretch_len, total_positions, k, tol_t):
        if stretch_stamp is None:
            return _check(self._lib.pgmo_deform_build_schedule(
                self._h, 0, None, None, total_positions, k, float(tol_t)))
        ss = np.ascontiguousarray(stretch_stamp, dtype=np.uint64)
        sl = np.ascontiguousarray(stretch_len, dtype=np.uint64)
        return _check(self._lib.pgmo_deform_build_schedule(
            self._h, ss.shape[0], _ptr(ss), _ptr(sl), total_positions, k, float(tol_t)))

    def schedule_stats(self):
        s = np.zeros(4, dtype=np.uint64)
        _check(self._lib.pgmo_deform_schedule_stats(self._h, _ptr(s)))
        return [int(v) for v in s]

    def schedule_info(self):
        n = C.c_uint64(0)
        _check(self._lib.pgmo_deform_schedule_info(self._h, C.byref(n), 0, None))
        runs = np.zeros((max(int(n.value), 1), 5), dtype=np.uint64)
        _check(self._lib.pgmo_deform_schedule_info(self._h, C.byref(n), runs.shape[0], _ptr(runs)))
        return runs[: n.value]

    def run_device(self, first_position, n_local, d_x, d_y, d_z, d_ox, d_oy, d_oz,
                   precision=FP64, d_subset=None, d_f64=None, d_assoc=None, assoc_stride=0,
                   d_assoc_cnt=None, d_processed=None, want_resolved=False, stream=0):
        resolved = C.c_uint64(0)
        _check(self._lib.pgmo_deform_run_device(
            self._h, first_position, n_local, _ptr(d_x), _ptr(d_y), _ptr(d_z), _ptr(d_subset),
            precision, _ptr(d_ox), _ptr(d_oy), _ptr(d_oz), _ptr(d_f64), _ptr(d_assoc),
            assoc_stride, _ptr(d_assoc_cnt), _ptr(d_processed),
            C.cast(C.byref(resolved), _vp) if want_resolved else None, _vp(stream)))
        return int(resolved.value)

    def deform_mesh_device(self, n_v, d_vertices_xyzw, d_stamps, d_out_xyzw, k=4, tol_t=10.0,
                           precision=FP64, d_processed=None, want_resolved=False, stream=0):
        """deformMesh on a device-resident {x, y, z, rgba} cloud (raw device pointers / tensors'
        data_ptr()); returns (rc, n_resolved)."""
        resolved = C.c_uint64(0)
        rc = _check(self._lib.pgmo_deform_mesh_device(
            self._h, n_v, _ptr(d_vertices_xyzw), _ptr(d_stamps), k, tol_t, precision,
            _ptr(d_out_xyzw), _ptr(d_processed),
            C.cast(C.byref(resolved), _vp) if want_resolved else None, _vp(stream)))
        return rc, int(resolved.value)

    def msg_to_cloud(self, vertices_xyz, colors_rgba, stamps_sec_nsec, d_cloud, d_stamps, stream=0):
        v = np.ascontiguousarray(vertices_xyz, dtype=np.float64).reshape(-1, 3)
        c = None if colors_rgba is None else np.ascontiguousarray(colors_rgba, dtype=np.float32).reshape(-1, 4)
        t = None if stamps_sec_nsec is None else np.ascontiguousarray(stamps_sec_nsec, dtype=np.uint32).reshape(-1, 2)
        _check(self._lib.pgmo_deform_msg_to_cloud(self._h, v.shape[0], _ptr(v), _ptr(c), _ptr(t),
                                                  _ptr(d_cloud), _ptr(d_stamps), _vp(stream)))
        return v, c, t  # keep the host arrays alive until the stream has consumed them

    def cloud_to_msg(self, n_v, d_cloud, d_stamps, want_colors=True, stream=0):
        v = np.zeros((n_v, 3), dtype=np.float64)
        c = np.zeros((n_v, 4), dtype=np.float32) if want_colors else None
        t = np.zeros((n_v, 2), dtype=np.uint32) if d_stamps is not None else None
        _check(self._lib.pgmo_deform_cloud_to_msg(self._h, n_v, _ptr(d_cloud), _ptr(d_stamps), _ptr(v),
                                                  _ptr(c), _ptr(t), _vp(stream)))
        return v, c, t

    def deform_mesh_msg(self, vertices_xyz, stamps_sec_nsec, k=4, tol_t=10.0, precision=FP64):
        v = np.ascontiguousarray(vertices_xyz, dtype=np.float64).reshape(-1, 3)
        t = None if stamps_sec_nsec is None else np.ascontiguousarray(stamps_sec_nsec, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros_like(v)
        processed = np.zeros(v.shape[0], dtype=np.uint8)
        rc = _check(self._lib.pgmo_deform_mesh_msg(self._h, v.shape[0], _ptr(v), _ptr(t), k, float(tol_t),
                                                   precision, _ptr(out), _ptr(processed)))
        return dict(rc=rc, vertices=out, processed=processed)

    def predeform(self, x, y, z, graph_indices, start_idx, out):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.float32)
        gi = np.ascontiguousarray(graph_indices, dtype=np.int32)
        n_v = x.shape[0]
        to_deform = np.zeros(max(n_v - start_idx, 1), dtype=np.uint64)
        n_to = C.c_uint64(0)
        ox, oy, oz = out
        _check(self._lib.pgmo_deform_predeform(self._h, n_v, _ptr(x), _ptr(y), _ptr(z), _ptr(gi),
                                               start_idx, _ptr(ox), _ptr(oy), _ptr(oz),
                                               _ptr(to_deform), C.byref(n_to)))
        return to_deform[: n_to.value].copy()


class DeformGroup:
    """pgmo_deform_group_t: the deformation path over several GPUs of one box from one call."""

    def __init__(self, devices=None):
        self._lib = load()
        g = _vp()
        if devices is None:
            _check(self._lib.pgmo_deform_group_create(0, None, C.byref(g)))
        else:
            d = np.ascontiguousarray(devices, dtype=np.int32)
            _check(self._lib.pgmo_deform_group_create(d.shape[0], _ptr(d), C.byref(g)))
        self._g = g

    def close(self):
        if getattr(self, "_g", None):
            self._lib.pgmo_deform_group_destroy(self._g)
            self._g = None

    def __del__(self):
        self.close()

    def size(self):
        return int(self._lib.pgmo_deform_group_size(self._g))

    def set_controls(self, pos_xyz, stamps_ns, valid, poses12):
        pos = np.ascontiguousarray(pos_xyz, dtype=np.float64).reshape(-1, 3)
        st = None if stamps_ns is None else np.ascontiguousarray(stamps_ns, dtype=np.uint64)
        va = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        po = np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
        _check(self._lib.pgmo_deform_group_set_controls(self._g, pos.shape[0], _ptr(pos), _ptr(st), _ptr(va),
                                                        _ptr(po)))

    def update_poses(self, poses12):
        po = np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
        _check(self._lib.pgmo_deform_group_update_poses(self._g, _ptr(po)))

    def deform_points(self, x, y, z, stamps_ns, k=4, tol_t=10.0, precision=FP64, out=None):
        n_v = x.shape[0]
        ox, oy, oz = out if out is not None else (np.array(x, copy=True), np.array(y, copy=True), np.array(z, copy=True))
        processed = np.zeros(n_v, dtype=np.uint8)
        rc = _check(self._lib.pgmo_deform_group_points(
            self._g, n_v, _ptr(x), _ptr(y), _ptr(z), _ptr(stamps_ns), k, float(tol_t), precision, _ptr(ox),
            _ptr(oy), _ptr(oz), _ptr(processed)))
        return dict(rc=rc, x=ox, y=oy, z=oz, processed=processed)

    def repeat(self, x, y, z, precision=FP64, out=None):
        n_v = x.shape[0]
        ox, oy, oz = out if out is not None else (np.array(x, copy=True), np.array(y, copy=True), np.array(z, copy=True))
        processed = np.zeros(n_v, dtype=np.uint8)
        rc = _check(self._lib.pgmo_deform_group_repeat(self._g, n_v, _ptr(x), _ptr(y), _ptr(z), precision,
                                                       _ptr(ox), _ptr(oy), _ptr(oz), _ptr(processed)))
        return dict(rc=rc, x=ox, y=oy, z=oz, processed=processed)


def _flatten_blocks(block_idx, offsets, xyz, rgba):
    bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    pts = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    col = None if rgba is None else np.ascontiguousarray(rgba, dtype=np.uint32)
    assert off.shape[0] == bi.shape[0] + 1 and int(off[-1]) == pts.shape[0]
    return bi, off, pts, col


def _voxblox_args(block_idx, offsets, x16, y16, z16, r, g, b):
    bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    packed = [np.ascontiguousarray(a, dtype=np.uint16) for a in (x16, y16, z16)]
    packed += [np.ascontiguousarray(a, dtype=np.uint8) for a in (r, g, b)]
    return bi, off, packed


def voxblox_decode(block_idx, offsets, x16, y16, z16, r, g, b, edge, device=0):
    """VoxbloxMsgInterface::getActiveVertex for a whole message on the GPU -> (xyz f32, rgba)"""
    bi, off, packed = _voxblox_args(block_idx, offsets, x16, y16, z16, r, g, b)
    n = int(off[-1]) if off.shape[0] else 0
    xyz = np.zeros((n, 3), dtype=np.float32)
    rgba = np.zeros(n, dtype=np.uint32)
    _check(load().pgmo_voxblox_decode(device, bi.shape[0], _ptr(bi), _ptr(off),
                                      *[_ptr(a) for a in packed], float(edge), _ptr(xyz), _ptr(rgba)))
    return xyz, rgba


class VccHandle:
    """pgmo_vcc_t: VoxelClearingCompression."""

    def __init__(self, resolution: float, device: int = 0):
        self._lib = load()
        h = _vp()
        _check(self._lib.pgmo_vcc_create(float(resolution), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgmo_vcc_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def integrate(self, block_idx, offsets, xyz, rgba, stamp_s, want_keys=False):
        bi, off, pts, col = _flatten_blocks(block_idx, offsets, xyz, rgba)
        n = pts.shape[0]
        remap = np.zeros(n, dtype=np.uint64)
        keys = np.zeros((n, 3), dtype=np.int64) if want_keys else None
        _check(self._lib.pgmo_vcc_integrate(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(pts),
                                            _ptr(col), float(stamp_s), _ptr(remap), _ptr(keys)))
        return remap, keys

    def integrate_voxblox(self, block_idx, offsets, x16, y16, z16, r, g, b, edge, stamp_s,
                          want_keys=False):
        bi, off, packed = _voxblox_args(block_idx, offsets, x16, y16, z16, r, g, b)
        n = int(off[-1]) if off.shape[0] else 0
        remap = np.zeros(n, dtype=np.uint64)
        keys = np.zeros((n, 3), dtype=np.int64) if want_keys else None
        _check(self._lib.pgmo_vcc_integrate_voxblox(
            self._h, bi.shape[0], _ptr(bi), _ptr(off), *[_ptr(a) for a in packed], float(edge),
            float(stamp_s), _ptr(remap), _ptr(keys)))
        return (remap, keys) if want_keys else remap

    def integrate_device(self, block_idx, offsets, d_xyz, d_rgba, stamp_s, d_remap, stream=0):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(self._lib.pgmo_vcc_integrate_device(self._h, bi.shape[0], _ptr(bi), _ptr(off),
                                                   _ptr(d_xyz), _ptr(d_rgba), float(stamp_s),
                                                   _ptr(d_remap), _vp(stream)))

    def prune(self, earliest_s):
        _check(self._lib.pgmo_vcc_prune(self._h, float(earliest_s)))

    def clear_blocks(self, block_idx):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        _check(self._lib.pgmo_vcc_clear_blocks(self._h, bi.shape[0], _ptr(bi)))

    def device_view(self):
        """(device pointer to the {x, y, z, rgba} records, device pointer to the u64 stamps, n)"""
        v, st, n = _vp(), _vp(), C.c_uint64(0)
        _check(self._lib.pgmo_vcc_device_view(self._h, C.byref(v), C.byref(st), C.byref(n)))
        return v.value, st.value, int(n.value)

    def sizes(self):
        s = np.zeros(4, dtype=np.uint64)
        _check(self._lib.pgmo_vcc_sizes(self._h, _ptr(s)))
        return [int(v) for v in s]

    def snapshot(self):
        nv, npoly, nact, ninv = self.sizes()
        out = dict(
            vertices=np.zeros((nv, 3), dtype=np.float32), rgba=np.zeros(nv, dtype=np.uint32),
            stamps=np.zeros(nv, dtype=np.uint64), polygons=np.zeros((npoly, 3), dtype=np.uint32),
            active=np.zeros(nact, dtype=np.uint64), invalid=np.zeros(ninv, dtype=np.uint64))
        _check(self._lib.pgmo_vcc_snapshot(self._h, _ptr(out["vertices"]), _ptr(out["rgba"]),
                                           _ptr(out["stamps"]), _ptr(out["polygons"]),
                                           _ptr(out["active"]), _ptr(out["invalid"])))
        return out


class OctHandle:
    """pgmo_oct_t: OctreeCompression, or BlockCompression with block=True (same base flow, cell-hash
    structure tests)."""

    def __init__(self, resolution: float, device: int = 0, block: bool = False):
        self._lib = load()
        h = _vp()
        create = self._lib.pgmo_oct_create_block if block else self._lib.pgmo_oct_create
        _check(create(float(resolution), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgmo_oct_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def integrate(self, block_idx, offsets, xyz, rgba, stamp_s):
        bi, off, pts, col = _flatten_blocks(block_idx, offsets, xyz, rgba)
        n = pts.shape[0]
        idx = np.full(n, -1, dtype=np.int64)
        kind = np.zeros(n, dtype=np.uint8)
        _check(self._lib.pgmo_oct_integrate(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(pts),
                                            _ptr(col), float(stamp_s), _ptr(idx), _ptr(kind)))
        return idx, kind

    def integrate_voxblox(self, block_idx, offsets, x16, y16, z16, r, g, b, edge, stamp_s):
        bi, off, packed = _voxblox_args(block_idx, offsets, x16, y16, z16, r, g, b)
        n = int(off[-1]) if off.shape[0] else 0
        idx = np.full(n, -1, dtype=np.int64)
        kind = np.zeros(n, dtype=np.uint8)
        _check(self._lib.pgmo_oct_integrate_voxblox(
            self._h, bi.shape[0], _ptr(bi), _ptr(off), *[_ptr(a) for a in packed], float(edge),
            float(stamp_s), _ptr(idx), _ptr(kind)))
        return idx, kind

    def prune(self, earliest_s):
        _check(self._lib.pgmo_oct_prune(self._h, float(earliest_s)))

    def device_view(self):
        """(device pointer to the {x, y, z, rgba} records, device pointer to the u64 stamps, n)"""
        v, st, n = _vp(), _vp(), C.c_uint64(0)
        _check(self._lib.pgmo_oct_device_view(self._h, C.byref(v), C.byref(st), C.byref(n)))
        return v.value, st.value, int(n.value)

    def sizes(self):
        s = np.zeros(8, dtype=np.uint64)
        _check(self._lib.pgmo_oct_sizes(self._h, _ptr(s)))
        return [int(v) for v in s]

    def snapshot(self):
        s = self.sizes()
        out = dict(
            vertices=np.zeros((s[0], 3), dtype=np.float32), rgba=np.zeros(s[0], dtype=np.uint32),
            stamps=np.zeros(s[0], dtype=np.uint64), polygons=np.zeros((s[1], 3), dtype=np.uint32),
            active=np.zeros(s[2], dtype=np.uint64), active_xyz=np.zeros((s[2], 3), dtype=np.float32),
            active_stamps=np.zeros(s[2], dtype=np.float64))
        _check(self._lib.pgmo_oct_snapshot(self._h, _ptr(out["vertices"]), _ptr(out["rgba"]),
                                           _ptr(out["stamps"]), _ptr(out["polygons"]),
                                           _ptr(out["active"]), _ptr(out["active_xyz"]),
                                           _ptr(out["active_stamps"])))
        return out

    def last_stats(self):
        """[fixed-point rounds, candidates, temp vertices, accepted] of the last integrate."""
        st = np.zeros(4, dtype=np.uint64)
        _check(self._lib.pgmo_oct_last_stats(self._h, _ptr(st)))
        return [int(v) for v in st]

    def last_outputs(self):
        s = self.sizes()
        out = dict(new_vertices=np.zeros((s[4], 3), dtype=np.float32),
                   new_rgba=np.zeros(s[4], dtype=np.uint32),
                   new_triangles=np.zeros((s[5], 3), dtype=np.uint32),
                   new_indices=np.zeros(s[6], dtype=np.uint64))
        _check(self._lib.pgmo_oct_last_outputs(self._h, _ptr(out["new_vertices"]),
                                               _ptr(out["new_rgba"]), _ptr(out["new_triangles"]),
                                               _ptr(out["new_indices"])))
        return out


class DeltaHandle:
    """kimera_pgmo::DeltaCompression through the C ABI; same result dicts as
    oracle.ffi.DeltaCompression."""

    def __init__(self, resolution: float, device: int = 0):
        self._lib = load()
        h = _vp()
        _check(self._lib.pgmo_delta_create(float(resolution), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgmo_delta_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def update(self, block_idx, offsets, xyz, rgba, stamp_ns, fetch=True, labels=None):
        bi, off, pts, col = _flatten_blocks(block_idx, offsets, xyz, rgba)
        remap = np.zeros(pts.shape[0], dtype=np.uint64)
        if labels is None:
            _check(self._lib.pgmo_delta_update(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(pts), _ptr(col),
                                               int(stamp_ns), _ptr(remap)))
        else:
            lab = np.ascontiguousarray(labels, dtype=np.uint32)
            assert lab.shape[0] == pts.shape[0]
            _check(self._lib.pgmo_delta_update_labeled(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(pts),
                                                       _ptr(col), _ptr(lab), int(stamp_ns), _ptr(remap)))
        return remap, (self.delta() if fetch else None)

    def labels(self):
        """semantic_updates of the last delta (0xFFFFFFFF = none), and whether labels were ever given"""
        n = self.sizes()[2]
        out = np.zeros(n, dtype=np.uint32)
        flag = C.c_int(0)
        _check(self._lib.pgmo_delta_fetch_labels(self._h, _ptr(out), C.byref(flag)))
        return out, bool(flag.value)

    def archive_by_time(self, earliest_ns):
        _check(self._lib.pgmo_delta_archive_by_time(self._h, int(earliest_ns)))

    def archive_blocks(self, block_idx):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        _check(self._lib.pgmo_delta_archive_blocks(self._h, bi.shape[0], _ptr(bi)))

    def sizes(self):
        s = np.zeros(12, dtype=np.uint64)
        _check(self._lib.pgmo_delta_sizes(self._h, _ptr(s)))
        return [int(v) for v in s]

    def last_stats(self):
        s = np.zeros(4, dtype=np.uint64)
        _check(self._lib.pgmo_delta_last_stats(self._h, _ptr(s)))
        return [int(v) for v in s]

    def blocks(self):
        """block_info_map_ in iteration order: (index [n,3], update_time_ns [n], sequence [n])"""
        n = C.c_uint64(0)
        _check(self._lib.pgmo_delta_blocks(self._h, 0, None, None, None, C.byref(n)))
        idx = np.zeros((n.value, 3), np.int32)
        tm = np.zeros(n.value, np.uint64)
        sq = np.zeros(n.value, np.uint32)
        _check(self._lib.pgmo_delta_blocks(self._h, n.value, _ptr(idx), _ptr(tm), _ptr(sq), C.byref(n)))
        return idx, tm, sq

    def delta(self):
        s = self.sizes()
        d = dict(vertex_start=s[0], face_start=s[1], num_archived_vertices=s[3],
                 vertices=np.zeros((s[2], 3), np.float32), rgba=np.zeros(s[2], np.uint32),
                 stamps=np.zeros(s[2], np.uint64), faces=np.zeros((s[4], 3), np.uint32),
                 faces_archive=np.zeros((s[5], 3), np.uint32), prev_to_curr=np.zeros((s[6], 2), np.uint64),
                 deleted=np.zeros(s[7], np.uint64), observed=np.zeros(s[8], np.uint64),
                 new=np.zeros(s[9], np.uint64), active_voxels=s[10], pending_archived_faces=s[11])
        n = [C.c_uint64(0) for _ in range(4)]
        _check(self._lib.pgmo_delta_fetch(
            self._h, _ptr(d["vertices"]), _ptr(d["rgba"]), _ptr(d["stamps"]), _ptr(d["faces"]),
            _ptr(d["faces_archive"]), _ptr(d["prev_to_curr"]), C.byref(n[0]), _ptr(d["deleted"]),
            C.byref(n[1]), _ptr(d["observed"]), C.byref(n[2]), _ptr(d["new"]), C.byref(n[3])))
        d["prev_to_curr"] = d["prev_to_curr"][: n[0].value]
        d["deleted"] = d["deleted"][: n[1].value]
        d["observed"] = d["observed"][: n[2].value]
        d["new"] = d["new"][: n[3].value]
        return d


class MeshToGraphHandle:
    """pgmo_m2g_t: msg_to_mesh_idx_ / msg_to_graph_idx_ / mesh_to_graph_idx_ of
    MeshFrontendInterface, in HBM; same methods as oracle.ffi.MeshToGraph."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = _vp()
        _check(self._lib.pgmo_m2g_create(device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgmo_m2g_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def update(self, block_idx, offsets, mesh_index, mesh_kind, graph_index, graph_kind):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        mi = np.ascontiguousarray(mesh_index).astype(np.int64, copy=False)
        gi = np.ascontiguousarray(graph_index).astype(np.int64, copy=False)
        mk = None if mesh_kind is None else np.ascontiguousarray(mesh_kind, dtype=np.uint8)
        gk = None if graph_kind is None else np.ascontiguousarray(graph_kind, dtype=np.uint8)
        _check(self._lib.pgmo_m2g_update(self._h, bi.shape[0], _ptr(bi), _ptr(off), _ptr(mi), _ptr(mk),
                                         _ptr(gi), _ptr(gk)))

    def mesh_to_graph(self):
        """-> (mesh indices ascending, graph indices)"""
        n = C.c_uint64(0)
        _check(self._lib.pgmo_m2g_size(self._h, C.byref(n)))
        m, g = np.zeros(n.value, np.uint64), np.zeros(n.value, np.uint64)
        if n.value:
            _check(self._lib.pgmo_m2g_fetch(self._h, _ptr(m), _ptr(g)))
        return m, g

    def block_mapping(self, block, which):
        bi = np.ascontiguousarray(block, dtype=np.int32)
        n = C.c_uint64(0)
        _check(self._lib.pgmo_m2g_block_mapping(self._h, _ptr(bi), int(which), None, 0, C.byref(n)))
        out = np.full(n.value, -1, np.int64)
        if n.value:
            _check(self._lib.pgmo_m2g_block_mapping(self._h, _ptr(bi), int(which), _ptr(out), n.value,
                                                    C.byref(n)))
        return out

    def device_view(self):
        """-> (device pointer of the dense mesh -> graph array, its length)"""
        p, n = _vp(), C.c_uint64(0)
        _check(self._lib.pgmo_m2g_device_view(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value
