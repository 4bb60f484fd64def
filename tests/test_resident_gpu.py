"""GPU: the device-resident mesh path (SURVEY 8f "next" row 3).

pgmo_{vcc,oct}_device_view hand the compressor's all_vertices_ / all_vertex_stamps_ to
pgmo_deform_mesh_device without a host round trip.  Bar: bit-identical to the host-buffer path
(pgmo_deform_points on the snapshot of the same compressor) and to the CPU oracle run on that
snapshot; rgba untouched; unprocessed vertices keep their input record."""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu


def _controls(snap, every, rng):
    """control points = a subsample of the stored vertices (float-representable), stamps ascending"""
    v, st = snap["vertices"], snap["stamps"]
    pick = np.sort(rng.choice(v.shape[0], max(v.shape[0] // every, 8), replace=False))
    order = np.argsort(st[pick], kind="stable")
    pick = pick[order]
    cp = v[pick].astype(np.float64)
    cs = st[pick].astype(np.uint64)
    ramp = np.linspace(0.0, 1.0, cp.shape[0])
    th = np.deg2rad(4.0) * ramp
    poses = np.zeros((cp.shape[0], 12))
    poses[:, 0], poses[:, 1], poses[:, 3], poses[:, 4], poses[:, 8] = np.cos(th), -np.sin(th), np.sin(th), np.cos(th), 1.0
    poses[:, 9:] = cp + np.stack([0.5 * ramp, -0.3 * ramp, 0.1 * ramp], 1)
    return cp, cs, poses


def _fill(handle, n_msgs, verts, n_blocks):
    for blocks, offsets, xyz, rgba, t in synth.compression_stream(n_msgs, verts, n_blocks=n_blocks):
        handle.integrate(blocks, offsets, xyz, rgba, t)


@pytest.mark.parametrize("kind", ["vcc", "oct"])
@pytest.mark.parametrize("precision", [0, 1])
def test_resident_equals_host_path_and_oracle(capi, oracle, kind, precision):
    import torch

    comp = capi.VccHandle(0.05, 0) if kind == "vcc" else capi.OctHandle(0.05, 0)
    _fill(comp, 12, 6000, 4)
    snap = comp.snapshot()
    n = snap["vertices"].shape[0]
    d_v, d_st, n_view = comp.device_view()
    assert n_view == n and n > 10_000

    rng = np.random.default_rng(7)
    cp, cs, poses = _controls(snap, 400, rng)
    h = capi.DeformHandle(0)
    h.set_controls(cp, cs, None, poses)
    x, y, z = (np.ascontiguousarray(snap["vertices"][:, a]) for a in range(3))
    host = h.deform_points(x, y, z, snap["stamps"], k=4, tol_t=2.0, precision=precision)
    assert host["rc"] == 0

    out = torch.full((n, 4), float("nan"), dtype=torch.float32, device="cuda:0")
    processed = torch.zeros(n, dtype=torch.uint8, device="cuda:0")
    rc, _ = h.deform_mesh_device(n, d_v, d_st, out, k=4, tol_t=2.0, precision=precision,
                                 d_processed=processed)
    torch.cuda.synchronize()
    assert rc == 0
    res = out.cpu().numpy()
    done = processed.cpu().numpy() == 1
    assert np.array_equal(done, host["processed"] == 1)
    for a, ax in enumerate("xyz"):
        assert np.array_equal(res[done, a], host[ax][done]), ax
        # unprocessed vertices: the input record
        assert np.array_equal(res[~done, a], snap["vertices"][~done, a])
    assert np.array_equal(res[:, 3].view(np.uint32), snap["rgba"])

    if precision == 0:
        ref = oracle.deform_points(x, y, z, snap["stamps"], cp, cs, np.ones(cp.shape[0], np.uint8),
                                   poses, k=4, tol_t=2.0)
        assert ref["rc"] == 0
        ref_done = ref["processed"] == 1
        assert np.array_equal(done, ref_done)
        for a, ax in enumerate("xyz"):
            assert np.array_equal(res[done, a], ref[ax][done]), ax
    h.close()
    comp.close()


def test_resident_in_place_and_view_tracks_the_compressor(capi):
    import torch

    vcc = capi.VccHandle(0.05, 0)
    _fill(vcc, 4, 6000, 1)
    _, _, n0 = vcc.device_view()
    _fill(vcc, 4, 6000, 1)  # same stream again: re-observations only
    d_v, d_st, n1 = vcc.device_view()
    assert n1 == vcc.sizes()[0] and n1 >= n0
    snap = vcc.snapshot()
    cp, cs, poses = _controls(snap, 300, np.random.default_rng(3))
    h = capi.DeformHandle(0)
    h.set_controls(cp, cs, None, poses)
    a = torch.empty((n1, 4), dtype=torch.float32, device="cuda:0")
    rc, _ = h.deform_mesh_device(n1, d_v, d_st, a, k=4, tol_t=2.0)
    assert rc == 0
    # in place on a private copy of the cloud
    cloud = torch.from_numpy(np.concatenate([snap["vertices"], snap["rgba"].view(np.float32)[:, None]], 1)).cuda()
    rc, _ = h.deform_mesh_device(n1, cloud, d_st, cloud, k=4, tol_t=2.0)
    torch.cuda.synchronize()
    assert rc == 0
    assert torch.equal(a.view(torch.int32), cloud.view(torch.int32))
    # the compressor's own cloud was not modified
    assert np.array_equal(vcc.snapshot()["vertices"], snap["vertices"])
    h.close()
    vcc.close()


def test_speculative_rerun_detects_changed_stamps(capi):
    """A second call on the same mesh deforms on the cached schedule while the stamp scan is in
    flight; a mesh whose stamps changed (same size) must fall back to the full path."""
    import torch

    vcc = capi.VccHandle(0.05, 0)
    _fill(vcc, 10, 6000, 1)
    snap = vcc.snapshot()
    n = snap["vertices"].shape[0]
    cp, cs, poses = _controls(snap, 300, np.random.default_rng(11))
    h = capi.DeformHandle(0)
    h.set_controls(cp, cs, None, poses)
    cloud = torch.from_numpy(np.concatenate([snap["vertices"], snap["rgba"].view(np.float32)[:, None]], 1)).cuda()
    x, y, z = (np.ascontiguousarray(snap["vertices"][:, a]) for a in range(3))

    def resident(stamps_np):
        d_st = torch.from_numpy(stamps_np.view(np.int64)).cuda()
        out = torch.empty_like(cloud)
        rc, _ = h.deform_mesh_device(n, cloud, d_st, out, k=4, tol_t=2.0)
        torch.cuda.synchronize()
        assert rc == 0
        return out.cpu().numpy()

    first = resident(snap["stamps"])
    builds0 = h.last_timings()[2]
    again = resident(snap["stamps"])  # speculative: cached schedule, fingerprint matches
    assert h.last_timings()[2] == builds0
    assert np.array_equal(first.view(np.uint32), again.view(np.uint32))

    # move one stretch boundary: same vertex count, different stretch list
    st2 = snap["stamps"].copy()
    b = np.flatnonzero(st2[1:] != st2[:-1])
    assert b.shape[0] > 2
    cut = int(b[b.shape[0] // 2])
    st2[cut - 40:cut + 1] = st2[cut + 1]
    changed = resident(st2)
    assert h.last_timings()[2] == builds0 + 1
    h2 = capi.DeformHandle(0)
    h2.set_controls(cp, cs, None, poses)
    ref = h2.deform_points(x, y, z, st2, k=4, tol_t=2.0)
    for a, ax in enumerate("xyz"):
        assert np.array_equal(changed[:, a], ref[ax]), ax
    # and the explicit check API
    d_st2 = torch.from_numpy(st2.view(np.int64)).cuda()
    h.stamp_check_begin(d_st2, n)
    assert h.stamp_check_finish() is False
    d_st1 = torch.from_numpy(snap["stamps"].view(np.int64)).cuda()
    h.stamp_check_begin(d_st1, n)
    assert h.stamp_check_finish() is True
    h.close()
    h2.close()
    vcc.close()
