"""BASELINE configs[4] as a function (used by bench.py's `config4` key and scripts/config5.py):
OctreeCompression with horizon eviction plus full re-deformation per loop closure over a synthetic
keyframe trajectory.

Per keyframe: one 50k-vertex mesh message -> pgmo_oct_prune(t - horizon) + pgmo_oct_integrate.
Every `closure_every` keyframes the whole accumulated (compressed) mesh is deformed from scratch
(start index 0) with the control points / poses known so far: pgmo_deform_set_controls +
pgmo_deform_points with host buffers (window schedule rebuilt because the control table grew), the
same call again after a pose-only update, and the same re-deformation on the compressor's
device-resident cloud (pgmo_oct_device_view + pgmo_deform_mesh_device: no vertex upload, the
deformed cloud stays in HBM)."""
import time

import numpy as np

from . import capi, synth


def run_config4(n_kf=1000, closure_every=50, device=0, horizon=10.0, res=0.05, verts=50_000,
                ctrl_per_kf=10):
    import torch
    oc = capi.OctHandle(res, device)
    dh = capi.DeformHandle(device)
    rng = np.random.default_rng(0)
    lat_ms, closures, rounds = [], [], []
    ctrl_xyz, ctrl_stamps = [], []
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(n_kf, verts, n_blocks=1)):
        t0 = time.perf_counter()
        oc.prune(t - horizon)
        oc.integrate(blocks, offsets, xyz, rgba, t)
        lat_ms.append(1e3 * (time.perf_counter() - t0))
        rounds.append(oc.last_stats()[0])
        # control points of this keyframe: a subsample of the message's vertices (float-representable)
        pick = rng.choice(xyz.shape[0], ctrl_per_kf, replace=False)
        ctrl_xyz.append(xyz[pick].astype(np.float64))
        ctrl_stamps.append(np.full(ctrl_per_kf, capi.stamp_from_sec(t), dtype=np.uint64))
        if (m + 1) % closure_every == 0:
            snap = oc.snapshot()
            v = snap["vertices"]
            order = np.argsort(snap["stamps"], kind="stable")  # mesh vertices in stamp order
            x, y, z = (np.ascontiguousarray(v[order, a]) for a in range(3))
            st = np.ascontiguousarray(snap["stamps"][order])
            cp = np.concatenate(ctrl_xyz)
            cs = np.concatenate(ctrl_stamps)
            ramp = np.linspace(0.0, 1.0, cp.shape[0])
            poses = np.zeros((cp.shape[0], 12))
            poses[:, [0, 4, 8]] = 1.0
            poses[:, 9:] = cp + np.stack([0.6 * ramp, 0.8 * ramp, 0 * ramp], 1)
            t1 = time.perf_counter()
            dh.set_controls(cp, cs, None, poses)
            r = dh.deform_points(x, y, z, st, k=4, tol_t=10.0)
            t_first = 1e3 * (time.perf_counter() - t1)
            t2 = time.perf_counter()
            dh.update_poses(poses)
            dh.deform_points(x, y, z, st, k=4, tol_t=10.0)
            t_again = 1e3 * (time.perf_counter() - t2)
            d_v, d_st, n_res = oc.device_view()
            d_out = torch.empty((n_res, 4), dtype=torch.float32, device=f"cuda:{device}")
            dh.deform_mesh_device(n_res, d_v, d_st, d_out, k=4, tol_t=10.0)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            dh.update_poses(poses)
            dh.deform_mesh_device(n_res, d_v, d_st, d_out, k=4, tol_t=10.0)
            torch.cuda.synchronize()
            t_res = 1e3 * (time.perf_counter() - t3)
            closures.append({"keyframe": m + 1, "mesh_vertices": int(x.shape[0]), "controls": int(cp.shape[0]),
                             "deform_ms_new_controls": t_first, "deform_ms_same_controls": t_again,
                             "deform_ms_resident_cloud": t_res, "rc": int(r["rc"])})
    oc.close()
    dh.close()
    lat = np.array(lat_ms[5:]) if len(lat_ms) > 5 else np.array(lat_ms)
    return {
        "scenario": f"configs[4]: {n_kf} keyframes x {verts}-vertex messages, OctreeCompression res {res} m, "
                    f"horizon {horizon} s, closure every {closure_every}",
        "compression_ms_per_message": {"mean": float(lat.mean()), "p50": float(np.percentile(lat, 50)),
                                       "p99": float(np.percentile(lat, 99))},
        "compressed_vertices_per_s": verts / (lat.mean() * 1e-3),
        "fixed_point_rounds": {"mean": float(np.mean(rounds)), "max": int(np.max(rounds))},
        "slowest_messages": [(int(i), round(float(lat_ms[i]), 2), int(rounds[i])) for i in np.argsort(lat_ms)[-5:]],
        "final_mesh_vertices": closures[-1]["mesh_vertices"] if closures else 0,
        "closures": closures[:3] + closures[-3:] if len(closures) > 6 else closures,
    }
