"""GPU: the CUDA path (through the C ABI) against the committed golden vectors — the expected
values of the reference's own unit tests (reference_kat.json) and the frozen oracle outputs
(oracle_vectors.npz).  No oracle is executed here."""
import json
import os

import numpy as np
import pytest

from kimera_pgmo_b200 import synth
from test_golden_cpu import GOLDEN, kat_deform_case, kat_vcc_mesh

pytestmark = pytest.mark.gpu


def _kat():
    return json.load(open(os.path.join(GOLDEN, "reference_kat.json")))


def _vec():
    return np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))


def test_kat_deform_points(capi):
    k = _kat()["deform_points"]
    xs, zeros, ctrl, poses = kat_deform_case(k)
    poses[:, 9:] = ctrl + np.array(k["pose_translation_offset"])
    h = capi.DeformHandle(0)
    h.set_controls(ctrl, None, None, poses)
    r = h.deform_points(xs, zeros, zeros, None, k=k["k"])
    assert r["rc"] == 0
    np.testing.assert_allclose(r["x"], xs, atol=k["expect"]["tol"])
    np.testing.assert_allclose(r["y"], k["expect"]["y"], atol=k["expect"]["tol"])
    np.testing.assert_allclose(r["z"], k["expect"]["z"], atol=k["expect"]["tol"])


def test_kat_time_check(capi):
    k = _kat()["deform_points_time_check"]
    xs, zeros, ctrl, poses = kat_deform_case(k)
    poses[:, 9] = ctrl[:, 0]
    poses[:, 10] = k["pose_y"]
    cst = np.array([capi.stamp_from_sec(s) for s in k["control_stamp_s"]], dtype=np.uint64)
    vst = np.array([capi.stamp_from_sec(s) for s in k["vertex_stamp_s"]], dtype=np.uint64)
    h = capi.DeformHandle(0)
    h.set_controls(ctrl, cst, None, poses)
    r = h.deform_points(xs, zeros, zeros, vst, k=k["k"], tol_t=k["tol_t"])
    e = k["expect"]
    np.testing.assert_array_equal(r["y"][:50], e["y_first_half"])
    np.testing.assert_array_equal(r["y"][e["second_half_from"]:], e["y_second_half"])


def test_kat_stamp_from_sec(capi):
    k = _kat()["stamp_from_sec"]
    assert capi.stamp_from_sec(k["sec"]) == k["expect_ns"]


def test_kat_voxel_clearing_basic(capi):
    k = _kat()["voxel_clearing_basic"]
    vcc = capi.VccHandle(k["resolution"], 0)
    blocks, offsets, xyz, rgba = kat_vcc_mesh(k)
    vcc.integrate(blocks, offsets, xyz, rgba, 100.0)
    snap = vcc.snapshot()
    assert snap["vertices"].shape[0] == k["expect"]["num_vertices"]
    assert snap["invalid"].shape[0] == k["expect"]["num_invalid"]
    assert sorted(snap["active"].tolist()) == k["expect"]["active_indices"]


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frozen_deformation(capi, mode):
    g = _vec()
    case = synth.deformation_case(4_000, 200, seed=123, invalid_fraction=0.01)
    h = capi.DeformHandle(0)
    h.set_search_mode(mode)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    r = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0,
                        want_f64=True, want_assoc=True)
    ga = r["assoc"].astype(np.int64)
    ga[ga == 0xFFFFFFFF] = -1
    np.testing.assert_array_equal(ga, g["deform_assoc"].astype(np.int64))
    np.testing.assert_array_equal(r["f64"], g["deform_f64"])
    np.testing.assert_array_equal(np.stack([r["x"], r["y"], r["z"]], 1), g["deform_xyz"])
    np.testing.assert_array_equal(r["processed"], g["deform_processed"])


def test_frozen_lattice_ties(capi):
    g = _vec()
    lat = synth.lattice_case(n_side=8, points_per_voxel=2)
    h = capi.DeformHandle(0)
    h.set_controls(lat["ctrl_xyz"], lat["ctrl_stamps"], lat["valid"], lat["poses"])
    r = h.deform_points(lat["x"], lat["y"], lat["z"], lat["stamps"], k=4, tol_t=6.0,
                        want_f64=True, want_assoc=True)
    assert r["rc"] == int(g["lattice_rc"][0])
    done = g["lattice_processed"].astype(bool)
    ga = r["assoc"].astype(np.int64)
    ga[ga == 0xFFFFFFFF] = -1
    np.testing.assert_array_equal(ga[done], g["lattice_assoc"].astype(np.int64)[done])
    np.testing.assert_array_equal(r["f64"][done], g["lattice_f64"][done])


def test_frozen_voxel_clearing(capi):
    g = _vec()
    vcc = capi.VccHandle(0.05, 0)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(
            synth.compression_stream(4, 3_000, n_blocks=3, seed=5)):
        vcc.prune(t - 2.0)
        remap, keys = vcc.integrate(blocks, offsets, xyz, rgba, t, want_keys=True)
        np.testing.assert_array_equal(remap.astype(np.int64), g[f"vcc_remap_{m}"])
        np.testing.assert_array_equal(keys, g[f"vcc_keys_{m}"])
    snap = vcc.snapshot()
    np.testing.assert_array_equal(snap["vertices"], g["vcc_vertices"])
    np.testing.assert_array_equal(snap["stamps"], g["vcc_stamps"])
    np.testing.assert_array_equal(np.sort(snap["active"]), g["vcc_active"])
