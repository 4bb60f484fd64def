"""GPU: KimeraPgmoMesh wire format on the device (SURVEY 8f "next" row 4).

Bar: pgmo_deform_msg_to_cloud / _cloud_to_msg bit-identical to the oracle's restatement of
conversions::fillFromMsg / fillMsg, including the reference test's known answers on sphere.ply
(test_mesh_conversion.cpp:24-48); pgmo_deform_mesh_msg (message in, message out) bit-identical
to the host-buffer path run on the float-narrowed vertices, widened back to double."""
import json
import os

import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_sphere_round_trip_matches_reference_test_and_oracle(capi, oracle):
    import torch

    g = np.load(os.path.join(GOLD, "reference_meshes.npz"))
    kat = json.load(open(os.path.join(GOLD, "reference_mesh_kat.json")))["mesh_conversion"]
    v = g["sphere_vertices"]
    n = v.shape[0]
    rng = np.random.default_rng(2)
    msg_v = v.astype(np.float64) + rng.uniform(-1e-9, 1e-9, v.shape)  # doubles that are not floats
    msg_c = rng.uniform(0.0, 1.0, (n, 4)).astype(np.float32)
    msg_c[:5] = [[0, 0, 0, 0], [1, 1, 1, 1], [0.5, 0.25, 0.999, 1e-4], [1.0 / 255, 2.0 / 255, 254.0 / 255, 0.99609375],
                 [0.2, 0.4, 0.6, 0.8]]
    msg_t = np.stack([rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32),
                      rng.integers(0, 10**9, n, dtype=np.uint64).astype(np.uint32)], axis=1)
    h = capi.DeformHandle(0)
    cloud = torch.zeros((n, 4), dtype=torch.float32, device="cuda:0")
    ns = torch.zeros(n, dtype=torch.int64, device="cuda:0")
    keep = h.msg_to_cloud(msg_v, msg_c, msg_t, cloud, ns)
    torch.cuda.synchronize()
    oxyz, orgba, ons = oracle.msg_to_cloud(msg_v, msg_c, msg_t)
    got = cloud.cpu().numpy()
    assert np.array_equal(got[:, :3].view(np.uint32), oxyz.view(np.uint32))
    assert np.array_equal(got[:, 3].view(np.uint32), orgba)
    assert np.array_equal(ns.cpu().numpy().view(np.uint64), ons)
    for i, comp in kat["points"].items():  # the reference test's expectations
        for axis, val in comp.items():
            assert got[int(i), "xyz".index(axis)] == val
    assert n == kat["num_vertices"]
    # and back
    bv, bc, bt = h.cloud_to_msg(n, cloud, ns)
    ov, oc, ot = oracle.cloud_to_msg(oxyz, orgba, ons)
    assert np.array_equal(bv, ov) and np.array_equal(bc, oc) and np.array_equal(bt, ot)
    assert np.array_equal(bt, msg_t)
    # without colours / stamps
    h.msg_to_cloud(msg_v, None, None, cloud, None)
    torch.cuda.synchronize()
    assert (cloud.cpu().numpy()[:, 3].view(np.uint32) == 0xFF000000).all()
    del keep
    h.close()


@pytest.mark.parametrize("precision", [0, 1])
def test_deform_mesh_msg_equals_host_path(capi, oracle, precision):
    case = synth.deformation_case(60_000, 600)
    rng = np.random.default_rng(9)
    msg_v = np.stack([case["x"], case["y"], case["z"]], axis=1).astype(np.float64)
    msg_v += rng.uniform(-1e-8, 1e-8, msg_v.shape)  # what a publisher with double geometry sends
    st = case["stamps"]
    msg_t = np.stack([(st // np.uint64(10**9)).astype(np.uint32), (st % np.uint64(10**9)).astype(np.uint32)], 1)
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    out = h.deform_mesh_msg(msg_v, msg_t, k=4, tol_t=10.0, precision=precision)
    assert out["rc"] == 0
    again = h.deform_mesh_msg(msg_v, msg_t, k=4, tol_t=10.0, precision=precision)  # cached schedule
    assert np.array_equal(out["vertices"], again["vertices"])
    narrowed = msg_v.astype(np.float32)
    h2 = capi.DeformHandle(0)
    h2.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    ref = h2.deform_points(narrowed[:, 0].copy(), narrowed[:, 1].copy(), narrowed[:, 2].copy(), st, k=4,
                           tol_t=10.0, precision=precision)
    assert np.array_equal(out["processed"], ref["processed"])
    for a, ax in enumerate("xyz"):
        assert np.array_equal(out["vertices"][:, a], ref[ax].astype(np.float64)), ax
    if precision == 0:
        o = oracle.deform_points(narrowed[:, 0].copy(), narrowed[:, 1].copy(), narrowed[:, 2].copy(), st,
                                 case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"], k=4,
                                 tol_t=10.0)
        for a, ax in enumerate("xyz"):
            assert np.array_equal(out["vertices"][:, a], o[ax].astype(np.float64)), ax
    # unstamped message (no vertex_stamps): the unstamped overload
    u = h.deform_mesh_msg(msg_v, None, k=4, precision=precision)
    ru = h2.deform_points(narrowed[:, 0].copy(), narrowed[:, 1].copy(), narrowed[:, 2].copy(), None, k=4,
                          precision=precision)
    for a, ax in enumerate("xyz"):
        assert np.array_equal(u["vertices"][:, a], ru[ax].astype(np.float64)), ax
    # empty message
    assert h.deform_mesh_msg(np.zeros((0, 3)), None)["rc"] == 1
    h.close()
    h2.close()
