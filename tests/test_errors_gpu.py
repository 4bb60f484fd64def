"""GPU: error conventions of the C ABI (include/pgmo_b200.h): return codes instead of
exceptions, PGMO_NOOP for the reference's silent early returns (outputs untouched),
PGMO_ERR_OUT_OF_RANGE where the reference's .at() would throw, PGMO_ERR_INVALID for bad arguments."""
import ctypes as C

import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu


def _handle(capi, case):
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
    return h


def test_zero_vertices_is_noop(capi):
    case = synth.deformation_case(100, 50)
    h = _handle(capi, case)
    e = np.zeros(0, np.float32)
    r = h.deform_points(e, e, e, np.zeros(0, np.uint64), k=4)
    assert r["rc"] == capi.PGMO_NOOP


def test_fewer_controls_than_k_leaves_outputs_untouched(capi):
    # mesh_deformation.h:149-152
    case = synth.deformation_case(1000, 3, ctrl_per_kf=1)
    h = _handle(capi, case)
    out = (np.full(1000, 7.0, np.float32), np.full(1000, 8.0, np.float32), np.full(1000, 9.0, np.float32))
    r = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, out=out)
    assert r["rc"] == capi.PGMO_NOOP
    assert (out[0] == 7.0).all() and (out[1] == 8.0).all() and (out[2] == 9.0).all()
    assert not r["processed"].any()


def test_subset_index_out_of_range(capi):
    # indices->at(i) / cloud.at(): std::out_of_range in the reference
    case = synth.deformation_case(1000, 50)
    h = _handle(capi, case)
    with pytest.raises(capi.PgmoError) as e:
        h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4,
                        subset=np.array([1, 5, 1000], dtype=np.uint64))
    assert e.value.code == capi.PGMO_ERR_OUT_OF_RANGE


def test_stamped_call_without_control_stamps_is_invalid(capi):
    case = synth.deformation_case(1000, 50)
    h = capi.DeformHandle(0)
    h.set_controls(case["ctrl_xyz"], None, None, case["poses"])
    with pytest.raises(capi.PgmoError) as e:
        h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4)
    assert e.value.code == capi.PGMO_ERR_INVALID


def test_run_device_without_schedule_is_invalid(capi):
    import torch
    h = capi.DeformHandle(0)
    t = torch.zeros(16, dtype=torch.float32, device="cuda:0")
    with pytest.raises(capi.PgmoError) as e:
        h.run_device(0, 16, t, t, t, t.clone(), t.clone(), t.clone())
    assert e.value.code == capi.PGMO_ERR_INVALID


def test_null_handle_and_bad_arguments(capi):
    lib = capi.load()
    assert lib.pgmo_deform_points(None, 0, None, None, None, None, None, 0, 4, 10.0, 0, None, None,
                                  None, None, None, 0, None, None) == capi.PGMO_ERR_INVALID
    h = C.c_void_p()
    assert lib.pgmo_vcc_create(0.0, 0, C.byref(h)) == capi.PGMO_ERR_INVALID   # resolution must be > 0
    assert lib.pgmo_oct_create(-1.0, 0, C.byref(h)) == capi.PGMO_ERR_INVALID
    assert lib.pgmo_deform_create(999, C.byref(h)) < 0                        # no such device
    assert b"" != lib.pgmo_last_error()


def test_large_k_goes_through_the_resolver(capi, oracle):
    # k + 3 > 16 register slots: every vertex is resolved by K4; still exact
    case = synth.deformation_case(3_000, 300, seed=4)
    h = _handle(capi, case)
    gpu = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=15, tol_t=10.0,
                          want_f64=True, want_assoc=True, assoc_stride=20)
    ref = oracle.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                               case["ctrl_stamps"], case["valid"], case["poses"], k=15, tol_t=10.0,
                               assoc_stride=20)
    assert gpu["rc"] == ref["rc"] == 0
    done = ref["processed"].astype(bool)
    ga = gpu["assoc"].astype(np.int64)
    ga[ga == 0xFFFFFFFF] = -1
    np.testing.assert_array_equal(ga[done], ref["assoc"][done].astype(np.int64))
    np.testing.assert_array_equal(gpu["f64"][done], ref["f64"][done])


def test_duplicate_block_index_in_one_vcc_message_is_unsupported(capi):
    vcc = capi.VccHandle(0.05, 0)
    blocks = np.zeros((2, 3), np.int32)
    offsets = np.array([0, 3, 6], np.uint64)
    xyz = np.random.default_rng(0).random((6, 3)).astype(np.float32)
    with pytest.raises(capi.PgmoError) as e:
        vcc.integrate(blocks, offsets, xyz, np.zeros(6, np.uint32), 1.0)
    assert e.value.code == capi.PGMO_ERR_UNSUPPORTED


def test_duplicate_block_in_one_message_is_rejected(capi):
    """The per-block bookkeeping of every compressor and of the mesh -> graph maps assumes that a
    block index appears once per message (the reference resolves duplicates through a block map);
    a duplicate is refused before anything changes."""
    blocks = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 0]], dtype=np.int32)
    offsets = np.array([0, 3, 6, 9], dtype=np.uint64)
    xyz = np.random.default_rng(0).uniform(0, 1, (9, 3)).astype(np.float32)
    rgba = np.zeros(9, dtype=np.uint32)
    for make in (lambda: capi.OctHandle(0.05, 0), lambda: capi.VccHandle(0.05), lambda: capi.DeltaHandle(0.05, 0)):
        h = make()
        with pytest.raises(capi.PgmoError) as e:
            if isinstance(h, capi.DeltaHandle):
                h.update(blocks, offsets, xyz, rgba, 1_000_000_000)
            else:
                h.integrate(blocks, offsets, xyz, rgba, 1.0)
        assert e.value.code == capi.PGMO_ERR_UNSUPPORTED
        # nothing changed: the same handle takes a well-formed message
        if isinstance(h, capi.DeltaHandle):
            h.update(blocks[:2], offsets[:3], xyz[:6], rgba[:6], 1_000_000_000)
        else:
            h.integrate(blocks[:2], offsets[:3], xyz[:6], rgba[:6], 1.0)
        h.close()
    m2g = capi.MeshToGraphHandle(0)
    idx = np.arange(9, dtype=np.int64)
    with pytest.raises(capi.PgmoError) as e:
        m2g.update(blocks, offsets, idx, None, idx, None)
    assert e.value.code == capi.PGMO_ERR_UNSUPPORTED
    m2g.update(blocks[:2], offsets[:3], idx[:6], None, idx[:6], None)
    m2g.close()


def test_device_side_error_poisons_the_compressor(capi):
    """A range error found by the kernels is reported after reference counts / the voxel table
    have changed: the handle refuses further work instead of continuing on that state."""
    vcc = capi.VccHandle(1.0e-3)
    far = np.array([[2000.0, 0, 0], [0, 0, 0], [1, 1, 1]], dtype=np.float32)  # key 2e6 > 2^20
    one_block = (np.zeros((1, 3), np.int32), np.array([0, 3], np.uint64))
    with pytest.raises(capi.PgmoError) as e:
        vcc.integrate(*one_block, far, None, 1.0)
    assert e.value.code == capi.PGMO_ERR_UNSUPPORTED
    near = np.array([[0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1]], dtype=np.float32)
    with pytest.raises(capi.PgmoError) as e:
        vcc.integrate(*one_block, near, None, 2.0)
    assert e.value.code == capi.PGMO_ERR_INVALID and "unusable" in str(e.value)
    with pytest.raises(capi.PgmoError):
        vcc.prune(0.0)
    vcc.close()
    # the octree compressor: a vertex farther than 2^18 voxels from the first one
    oc = capi.OctHandle(1.0e-3, 0)
    oc.integrate(*one_block, near, None, 1.0)
    with pytest.raises(capi.PgmoError) as e:
        oc.integrate(*one_block, near + np.float32(500.0), None, 2.0)
    assert e.value.code in (capi.PGMO_ERR_UNSUPPORTED, capi.PGMO_ERR_CUDA)
    with pytest.raises(capi.PgmoError) as e:
        oc.integrate(*one_block, near, None, 3.0)
    assert e.value.code == capi.PGMO_ERR_INVALID and "unusable" in str(e.value)
    oc.close()
