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
