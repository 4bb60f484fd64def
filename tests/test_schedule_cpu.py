"""CPU: the product's host-side window schedule (deform_schedule.cpp) against the oracle.

A numpy brute force over the scheduled alive sets must reproduce the oracle's processed mask,
neighbour counts and — wherever no exact distance tie is involved — the neighbour sets.  This
exercises leaf counting, whole-voxel deletion, k mutation and skips without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from kimera_pgmo_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build", "tests")


@pytest.fixture(scope="module")
def sched_lib():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libschedule_test.so")
    srcs = [os.path.join(ROOT, "tests", "cpp", "schedule_capi.cpp"),
            os.path.join(ROOT, "kimera_pgmo_b200", "csrc", "deform_schedule.cpp")]
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "kimera_pgmo_b200", "csrc"), "-o", so] + srcs,
                   check=True)
    lib = C.CDLL(so)
    lib.schedule_build.restype = C.c_long
    lib.schedule_build.argtypes = [C.c_ulong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong,
                                   C.c_void_p, C.c_void_p, C.c_ulong, C.c_uint, C.c_double,
                                   C.c_int, C.c_ulong, C.c_void_p, C.c_void_p]
    return lib


def _stretches(stamps):
    b = np.flatnonzero(np.concatenate([[True], stamps[1:] != stamps[:-1]]))
    lens = np.diff(np.concatenate([b, [stamps.shape[0]]]))
    return stamps[b].astype(np.uint64), lens.astype(np.uint64)


def _schedule(lib, case, k, tol_t, stamped=True):
    n_c = case["ctrl_xyz"].shape[0]
    pos = np.ascontiguousarray(case["ctrl_xyz"], dtype=np.float64)
    cst = np.ascontiguousarray(case["ctrl_stamps"], dtype=np.uint64)
    valid = np.ascontiguousarray(case["valid"], dtype=np.uint8)
    ss, sl = _stretches(case["stamps"])
    cap = 2 * ss.shape[0] + 2 * n_c + 16
    runs = np.zeros((cap, 6), dtype=np.uint64)
    del_run = np.zeros(n_c, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib.schedule_build(n_c, p(pos), p(cst), p(valid), ss.shape[0], p(ss), p(sl),
                           case["x"].shape[0], k, tol_t, 1 if stamped else 0, cap, p(runs),
                           p(del_run))
    return n, runs[: max(n, 0)], del_run


def _dist2(c, q):
    d = c.astype(np.float32) - q.astype(np.float32)
    sq = d * d
    return sq[:, 0] + (sq[:, 1] + sq[:, 2])


def _check(lib, oracle, case, k, tol_t, stamped=True):
    ref = oracle.deform_points(case["x"], case["y"], case["z"],
                               case["stamps"] if stamped else None, case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=k, tol_t=tol_t,
                               assoc_stride=16)
    n, runs, del_run = _schedule(lib, case, k, tol_t, stamped)
    if ref["rc"] == 1:
        assert n == -1
        return 0
    assert n > 0
    n_v = case["x"].shape[0]
    starts = runs[:, 0].astype(np.int64)
    cf = case["ctrl_xyz"].astype(np.float32)
    compared = 0
    rng = np.random.default_rng(0)
    # every run head (where the state just changed) plus a random sample
    sample = set(int(s) for s in starts) | set(int(s) + 1 for s in starts if s + 1 < n_v)
    sample |= set(int(v) for v in rng.choice(n_v, min(n_v, 1500), replace=False))
    for v in sorted(sample):
        r = int(np.searchsorted(starts, v, side="right") - 1)
        lo, hi, kk, skip = (int(runs[r, 1]), int(runs[r, 2]), int(runs[r, 3]), int(runs[r, 4]))
        assert bool(ref["processed"][v]) == (skip == 0), (v, r)
        if skip:
            continue
        j = np.arange(lo, hi)
        alive = j[del_run[lo:hi] > r]
        K = kk + 2
        n_nbr = min(K, alive.shape[0])
        assert ref["assoc_cnt"][v] == n_nbr - 1, (v, r, kk, alive.shape[0])
        d2 = _dist2(cf[alive], np.array([case["x"][v], case["y"][v], case["z"][v]]))
        order = np.lexsort((alive, d2))
        top = order[: min(K + 1, alive.shape[0])]
        if np.unique(d2[top]).shape[0] != top.shape[0]:
            continue  # exact tie: order is the resolver's job (GPU test)
        np.testing.assert_array_equal(alive[order[: n_nbr - 1]], ref["assoc"][v][: n_nbr - 1])
        compared += 1
    return compared


def test_config_a_small(sched_lib, oracle):
    case = synth.deformation_case(20_000, 400)
    assert _check(sched_lib, oracle, case, 4, 10.0) > 1000
    assert _check(sched_lib, oracle, case, 4, 20.0) > 1000


def test_invalid_controls_and_epoch_zero(sched_lib, oracle):
    case = synth.deformation_case(20_000, 400, invalid_fraction=0.05, t0_ns=0)
    assert _check(sched_lib, oracle, case, 4, 10.0) > 1000


def test_dense_voxels(sched_lib, oracle):
    # many control points per 1 m voxel: leaf count != point count, voxel-wide deletion
    case = synth.deformation_case(10_000, 1_000, ctrl_per_kf=50, ctrl_spread=1.2, vertex_spread=3.0)
    assert _check(sched_lib, oracle, case, 4, 3.0) > 500
    assert _check(sched_lib, oracle, case, 9, 1.0) > 500


def test_sparse_controls_k_shrink_and_skip(sched_lib, oracle):
    case = synth.deformation_case(3_000, 30, ctrl_per_kf=1, seed=2)
    case["valid"][1:4] = 0
    _check(sched_lib, oracle, case, 4, 0.4)
    _check(sched_lib, oracle, case, 2, 0.0)
    case["valid"][:] = 0
    case["valid"][7] = 1
    _check(sched_lib, oracle, case, 1, 0.4)


def test_unstamped(sched_lib, oracle):
    case = synth.deformation_case(5_000, 200, seed=5)
    assert _check(sched_lib, oracle, case, 4, 10.0, stamped=False) > 500


def test_reference_time_check_scenario(sched_lib, oracle):
    xs = np.arange(100, dtype=np.float32)
    zeros = np.zeros(100, dtype=np.float32)
    ctrl = np.array([[i, 0, 0] for i in range(0, 100, 10)], dtype=np.float64)
    poses = np.zeros((10, 12))
    poses[:, [0, 4, 8]] = 1.0
    t20 = oracle.stamp_from_sec(20.0)
    case = dict(x=xs, y=zeros, z=zeros,
                stamps=np.where(np.arange(100) < 50, 0, t20).astype(np.uint64), ctrl_xyz=ctrl,
                ctrl_stamps=np.where(ctrl[:, 0] > 50, t20, 0).astype(np.uint64),
                valid=np.ones(10, np.uint8), poses=poses)
    _check(sched_lib, oracle, case, 3, 10.0)
    n, runs, del_run = _schedule(sched_lib, case, 3, 10.0)
    # first vertex sees 6 controls, then the wrap removes two voxels; at stamp 20 s four more
    # arrive and the four stamp-0 survivors go
    assert runs[0, 2] == 6 and runs[1, 1] == 2
    assert list(del_run[:6]) == [1, 1, 3, 3, 3, 3] or del_run[0] == 1


def _extend_check(lib, case, n_old, v_old, k=4, tol_t=10.0):
    """old = first n_old controls + first v_old vertices; new = everything"""
    lib.schedule_extend_check.restype = C.c_long
    lib.schedule_extend_check.argtypes = [C.c_ulong, C.c_ulong, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_ulong, C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p,
                                          C.c_void_p, C.c_uint, C.c_double, C.c_void_p]
    pos = np.ascontiguousarray(case["ctrl_xyz"], dtype=np.float64)
    cst = np.ascontiguousarray(case["ctrl_stamps"], dtype=np.uint64)
    valid = np.ascontiguousarray(case["valid"], dtype=np.uint8)
    so, lo = _stretches(case["stamps"][:v_old])
    sn, ln = _stretches(case["stamps"])
    resumed = C.c_ulong(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.schedule_extend_check(n_old, pos.shape[0], p(pos), p(cst), p(valid), so.shape[0], p(so), p(lo),
                                   sn.shape[0], p(sn), p(ln), k, tol_t, C.byref(resumed))
    return rc, int(resumed.value), sn.shape[0]


def test_incremental_schedule_equals_full_rebuild(sched_lib):
    """Control points appended (new keyframes since the last deformMesh) and / or the mesh grown:
    extendStampedSchedule must give exactly the from-scratch schedule, and resume late."""
    resumed_late = 0
    for seed, (n_v, n_c, kw) in enumerate([
            (40_000, 800, {}), (40_000, 800, dict(invalid_fraction=0.05)), (30_000, 600, dict(ctrl_spread=1.0)),
            (30_000, 300, dict(ctrl_per_kf=3)), (20_000, 400, dict(t0_ns=0))]):
        case = synth.deformation_case(n_v, n_c, seed=20 + seed, **kw)
        n_kf = n_c // kw.get("ctrl_per_kf", 10)
        for frac_c, frac_v in ((0.9, 0.9), (0.99, 1.0), (1.0, 0.95), (0.5, 0.5), (0.97, 0.6), (0.3, 1.0), (0.999, 0.999)):
            n_old = max(int(n_c * frac_c), 1)
            v_old = max(int(n_v * frac_v), 1)
            for k, tol in ((4, 10.0), (2, 3.0), (4, 1.0e6)):
                rc, resumed, n_stretches = _extend_check(sched_lib, case, n_old, v_old, k, tol)
                assert rc == 0, (seed, frac_c, frac_v, k, tol, rc)
                if tol == 10.0 and frac_c >= 0.9 and frac_v >= 0.9:
                    assert resumed > 0.5 * n_stretches, (seed, frac_c, frac_v, resumed, n_stretches, n_kf)
                    resumed_late += 1
    assert resumed_late >= 10
    # nothing appended at all: still identical
    case = synth.deformation_case(10_000, 200, seed=3)
    assert _extend_check(sched_lib, case, 200, 10_000)[0] == 0
    # fewer controls than k before the extension: the old schedule was a no-op
    assert _extend_check(sched_lib, case, 3, 10_000)[0] == 0
