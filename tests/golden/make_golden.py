#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/.

Two kinds of content:
  * reference_kat.json  — inputs and EXPECTED VALUES transcribed from the reference's own unit
    tests (file:line cited per entry).  They are independent of the oracle: the CPU suite checks
    the oracle against them, the GPU suite checks the CUDA path against them.
  * oracle_vectors.npz  — outputs of the pinned oracle (oracle/liboracle.so) on small seeded
    cases, so that the GPU parity tests can also run against frozen vectors and any later drift
    of the oracle itself is caught by the CPU suite.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from kimera_pgmo_b200 import synth  # noqa: E402
from oracle import ffi  # noqa: E402


def reference_kats():
    kats = {}
    # kimera_pgmo/test/test_mesh_deformation.cpp:16-50 (deformPoints, unstamped, k = 4)
    kats["deform_points"] = {
        "cite": "kimera_pgmo/test/test_mesh_deformation.cpp:16-50",
        "vertices_x": list(range(100)), "control_x": list(range(0, 100, 10)),
        "pose_translation_offset": [0.0, 1.0, 0.0], "k": 4,
        "expect": {"x": "same", "y": 1.0, "z": 0.0, "tol": 1e-9},
    }
    # test_mesh_deformation.cpp:52-112 (deformPointsWithTimeCheck, k = 3, tol 10 s)
    kats["deform_points_time_check"] = {
        "cite": "kimera_pgmo/test/test_mesh_deformation.cpp:52-112",
        "vertices_x": list(range(100)), "control_x": list(range(0, 100, 10)),
        "control_stamp_s": [0.0 if x <= 50 else 20.0 for x in range(0, 100, 10)],
        "vertex_stamp_s": [0.0 if i < 50 else 20.0 for i in range(100)],
        "pose_y": [-1.0 if x <= 50 else 1.0 for x in range(0, 100, 10)], "k": 3, "tol_t": 10.0,
        "expect": {"y_first_half": -1.0, "y_second_half": 1.0, "second_half_from": 51},
    }
    # test_deformation_graph.cpp:250-309 (cube, second phase: linear weights -> 1.273)
    kats["cube_linear_weight"] = {
        "cite": "kimera_pgmo/test/test_deformation_graph.cpp:250-309",
        "vertices": [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1],
                     [1, 1, 1]],
        "controls": [[0, 0, 0], [1, 0, 0], [0, 1, 0]],
        "pose_translations": [[-0.5, 0, 0], [1.5, 0, 0], [0.5, 1, 0]], "k": 2, "tol_t": 10.0,
        "expect": {"x0": -0.5, "x1": 1.273, "tol": 1e-3},
    }
    # test_common_structs.cpp:12-15
    kats["stamp_from_sec"] = {"cite": "kimera_pgmo/test/test_common_structs.cpp:12-15",
                              "sec": 105867.312345678, "expect_ns": 105867312345678}
    # test_voxel_clearing_compression.cpp:115-135 (clearingCorrectBasic, compression_factor
    # 1e-3 at :103): the full scenarios with their per-position triangle checks live in
    # oracle/kat_compression.cpp; here: the expectations of the first integrate.
    kats["voxel_clearing_basic"] = {
        "cite": "kimera_pgmo/test/test_voxel_clearing_compression.cpp:115-135",
        "resolution": 1.0e-3,
        "block1": {"index": [0, 0, 0],
                   "faces": [[[0.5, 0.5, 0.5], [0.5, 0.75, 0.75], [0.5, 0.75, 0.5]],
                             [[0.5, 0.5, 0.5], [0.5, 0.75, 0.75], [0.5, 0.75, 0.5]],
                             [[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.0, 0.0, 0.5]]]},
        "expect": {"num_vertices": 6, "num_invalid": 0, "active_indices": [0, 1, 2, 3, 4, 5]},
    }
    return kats


def oracle_vectors():
    out = {}
    # deformation: 4000 vertices, 200 control points, k = 4, tol 10 s; 1 % invalid
    case = synth.deformation_case(4_000, 200, seed=123, invalid_fraction=0.01)
    r = ffi.deform_points(case["x"], case["y"], case["z"], case["stamps"], case["ctrl_xyz"],
                          case["ctrl_stamps"], case["valid"], case["poses"], k=4, tol_t=10.0)
    assert r["rc"] == 0
    out["deform_assoc"] = r["assoc"]
    out["deform_f64"] = r["f64"]
    out["deform_xyz"] = np.stack([r["x"], r["y"], r["z"]], axis=1)
    out["deform_processed"] = r["processed"]
    # lattice (ties everywhere): associations follow the PCL traversal order
    lat = synth.lattice_case(n_side=8, points_per_voxel=2)
    r = ffi.deform_points(lat["x"], lat["y"], lat["z"], lat["stamps"], lat["ctrl_xyz"],
                          lat["ctrl_stamps"], lat["valid"], lat["poses"], k=4, tol_t=6.0,
                          assoc_stride=8)
    out["lattice_rc"] = np.array([r["rc"]])
    out["lattice_assoc"] = r["assoc"]
    out["lattice_f64"] = r["f64"]
    out["lattice_processed"] = r["processed"]
    # voxel clearing: 4 messages x 3000 vertices, 3 blocks, prune(t - 2)
    for kind, tag in (("voxel_clearing", "vcc"), ("octree", "oct")):
        comp = ffi.Compression(kind, 0.05)
        for m, (blocks, offsets, xyz, rgba, t) in enumerate(
                synth.compression_stream(4, 3_000, n_blocks=3, seed=5)):
            comp.prune(t - 2.0)
            idx, kinds, keys = comp.integrate(blocks, offsets, xyz, rgba, t)
            out[f"{tag}_remap_{m}"] = idx
            out[f"{tag}_kind_{m}"] = kinds
            if tag == "vcc":
                out[f"{tag}_keys_{m}"] = keys
        snap = comp.snapshot()
        out[f"{tag}_vertices"] = snap["vertices"]
        out[f"{tag}_stamps"] = snap["stamps"]
        out[f"{tag}_active"] = np.sort(snap["active"])
        out[f"{tag}_polygons"] = snap["polygons"]
    return out


def main():
    with open(os.path.join(HERE, "reference_kat.json"), "w") as f:
        json.dump(reference_kats(), f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **oracle_vectors())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
