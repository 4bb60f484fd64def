#!/usr/bin/env python
"""Freezes the reference's own fixture files as arrays (they are not available on the GPU box):
kimera_pgmo/test/data/{cube.ply, sphere.ply, graph.dgrf}, read with kimera_pgmo_b200/formats.py,
plus the values the reference's tests expect of them
(kimera_pgmo_ros/test/test_mesh_conversion.cpp:24-48, kimera_pgmo/test/test_deformation_graph.cpp
:824-860).  Run here, where /root/reference exists:  python tests/golden/make_reference_meshes.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kimera_pgmo_b200 import formats  # noqa: E402

DATA = "/root/reference/kimera_pgmo/test/data"


def main():
    cube = formats.read_ply(os.path.join(DATA, "cube.ply"))
    sphere = formats.read_ply(os.path.join(DATA, "sphere.ply"))
    out = {"cube_vertices": cube["vertices"], "cube_faces": cube["faces"],
           "sphere_vertices": sphere["vertices"], "sphere_faces": sphere["faces"]}
    if cube["rgba"] is not None:
        out["cube_rgba"] = cube["rgba"]
    for prefix in ("a", "v"):
        xyz, stamps, valid, poses = formats.load_dgrf(os.path.join(DATA, "graph.dgrf"), prefix)
        out[f"dgrf_{prefix}_xyz"], out[f"dgrf_{prefix}_stamps"] = xyz, stamps
        out[f"dgrf_{prefix}_valid"], out[f"dgrf_{prefix}_poses"] = valid, poses
    with open(os.path.join(DATA, "graph.dgrf")) as f:
        out["dgrf_text"] = np.frombuffer(f.read().encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_meshes.npz"), **out)
    kat = {
        "mesh_conversion": {
            "cite": "kimera_pgmo_ros/test/test_mesh_conversion.cpp:24-48 (PCL -> KimeraPgmoMesh -> PCL on sphere.ply)",
            "num_vertices": 422, "points": {"0": {"z": -127}, "1": {"x": 5}, "420": {"x": -1}, "421": {"z": 127}},
            "faces_kept": [[0, 2], [100, 1], [839, 0]],
        },
        "cube": {"cite": "kimera_pgmo/test/data/cube.ply (test_deformation_graph.cpp:252)", "num_vertices": 8,
                 "num_faces": 12},
        "dgrf": {"cite": "kimera_pgmo/test/data/graph.dgrf (test_deformation_graph.cpp:824-860): 3 mesh "
                         "vertices with prefix 'v' at (0,0,0), (1,0,0), (0,1,0), 3 pose nodes with prefix 'a'",
                 "v_positions": [[0, 0, 0], [1, 0, 0], [0, 1, 0]], "a_nodes": 3},
    }
    with open(os.path.join(HERE, "reference_mesh_kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)
    print("written")


if __name__ == "__main__":
    main()
