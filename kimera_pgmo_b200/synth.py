"""Synthetic workloads of SURVEY.md §8(d) / BASELINE.json `configs` (numpy, seeded).

Everything is generated once on the host and fed unchanged to both the CUDA path and the
oracle.  Control-point doubles are float-representable (the reference derives them from float
mesh vertices, common_structs.cpp:169).
"""
from __future__ import annotations

import numpy as np

T0_NS = 1_700_000_000_000_000_000  # realistic epoch: no unsigned wrap in stamp - tol


def _trajectory(n_kf: int) -> np.ndarray:
    """Closed 100 m x 50 m rectangle walked at 1 m/s, one keyframe per second."""
    per = 300.0
    s = np.arange(n_kf, dtype=np.float64) % per
    x = np.where(s < 100, s, np.where(s < 150, 100.0, np.where(s < 250, 250.0 - s, 0.0)))
    y = np.where(s < 100, 0.0, np.where(s < 150, s - 100.0, np.where(s < 250, 50.0, 300.0 - s)))
    return np.stack([x, y, np.zeros_like(x)], axis=1)


def _rot_zy(theta, phi):
    cz, sz, cy, sy = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    # Rz(theta) @ Ry(phi), row-major
    R = np.empty(theta.shape + (9,), dtype=np.float64)
    R[..., 0] = cz * cy
    R[..., 1] = -sz
    R[..., 2] = cz * sy
    R[..., 3] = sz * cy
    R[..., 4] = cz
    R[..., 5] = sz * sy
    R[..., 6] = -sy
    R[..., 7] = 0.0
    R[..., 8] = cy
    return R


def deformation_case(n_vertices: int, n_controls: int, ctrl_per_kf: int = 10, seed: int = 42,
                     t0_ns: int = T0_NS, invalid_fraction: float = 0.0,
                     vertex_spread: float = 6.0, ctrl_spread: float = 5.0):
    """configs[0]/[2]: n_kf = n_controls / ctrl_per_kf keyframes; each owns ctrl_per_kf control
    points within +-ctrl_spread m and n_vertices / n_kf vertices within +-vertex_spread m,
    z in [0, 3]; every vertex / control point carries its keyframe's stamp (1 s apart).
    Poses: a smooth loop-closure-like correction ramping along the trajectory to 5 deg yaw,
    1 deg pitch, 1 m translation."""
    rng = np.random.default_rng(seed)
    n_kf = max(n_controls // ctrl_per_kf, 1)
    kf = _trajectory(n_kf)
    kf_stamp = (np.uint64(t0_ns) + np.arange(n_kf, dtype=np.uint64) * np.uint64(1_000_000_000))

    c_kf = np.minimum(np.arange(n_controls) // ctrl_per_kf, n_kf - 1)
    cpos = kf[c_kf] + np.stack([
        rng.uniform(-ctrl_spread, ctrl_spread, n_controls),
        rng.uniform(-ctrl_spread, ctrl_spread, n_controls),
        rng.uniform(0.0, 3.0, n_controls)], axis=1)
    cpos = cpos.astype(np.float32).astype(np.float64)
    cstamps = kf_stamp[c_kf]

    ramp = (np.arange(n_controls, dtype=np.float64) / max(n_controls - 1, 1))
    theta = np.deg2rad(5.0) * ramp
    phi = np.deg2rad(1.0) * ramp
    delta = np.stack([0.6 * ramp, 0.8 * ramp, np.zeros_like(ramp)], axis=1)
    poses = np.empty((n_controls, 12), dtype=np.float64)
    poses[:, :9] = _rot_zy(theta, phi)
    poses[:, 9:] = cpos + delta
    valid = np.ones(n_controls, dtype=np.uint8)
    if invalid_fraction > 0:
        valid[rng.random(n_controls) < invalid_fraction] = 0

    per_kf = -(-n_vertices // n_kf)
    v_kf = np.minimum(np.arange(n_vertices) // per_kf, n_kf - 1)
    base = kf[v_kf].astype(np.float32)
    x = base[:, 0] + rng.uniform(-vertex_spread, vertex_spread, n_vertices).astype(np.float32)
    y = base[:, 1] + rng.uniform(-vertex_spread, vertex_spread, n_vertices).astype(np.float32)
    z = rng.uniform(0.0, 3.0, n_vertices).astype(np.float32)
    stamps = kf_stamp[v_kf]
    return dict(x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), z=np.ascontiguousarray(z),
                stamps=np.ascontiguousarray(stamps), ctrl_xyz=cpos, ctrl_stamps=cstamps,
                valid=valid, poses=poses)


def deformation_case_with_ties(n_vertices: int, n_controls: int, tie_fraction: float,
                               ctrl_per_kf: int = 10, seed: int = 42, lattice: float = 2.0 ** -5):
    """deformation_case with exact float-distance ties planted at a chosen rate (VERDICT r1 #8:
    voxblox vertices sit on a lattice, nobody knows what ties cost).  Control points are snapped
    to a `lattice` grid (0.03125 m: sums of two coordinates stay exact in float); a fraction
    `tie_fraction` of the vertices is moved to the exact midpoint of a control point of its own
    keyframe and that point's nearest control point among the keyframes within +-5 s (both alive
    at the vertex's stamp): dx, dy, dz to the two are equal and opposite, so the two float d^2
    are bit-identical and they are the vertex's nearest two.  Returns (case, tied vertex indices)."""
    from scipy.spatial import cKDTree
    case = deformation_case(n_vertices, n_controls, ctrl_per_kf, seed)
    cpos = np.round(case["ctrl_xyz"] / lattice) * lattice
    case["poses"][:, 9:] += cpos - case["ctrl_xyz"]
    case["ctrl_xyz"] = cpos
    n_kf = max(n_controls // ctrl_per_kf, 1)
    c_kf = np.minimum(np.arange(n_controls) // ctrl_per_kf, n_kf - 1)
    _, nbr = cKDTree(cpos).query(cpos, k=9)
    partner = np.full(n_controls, -1, dtype=np.int64)
    for j in range(1, 9):  # nearest other control point within +-5 keyframes, not coincident
        cand = nbr[:, j]
        ok = (partner < 0) & (np.abs(c_kf[cand] - c_kf) <= 5) & (cpos[cand] != cpos).any(axis=1)
        partner[ok] = cand[ok]
    rng = np.random.default_rng([seed, 977])
    per_kf = -(-n_vertices // n_kf)
    tied = np.nonzero(rng.random(n_vertices) < tie_fraction)[0]
    v_kf = np.minimum(tied // per_kf, n_kf - 1)
    c1 = np.minimum(v_kf * ctrl_per_kf + rng.integers(0, ctrl_per_kf, tied.shape[0]), n_controls - 1)
    keep = partner[c1] >= 0
    tied, c1 = tied[keep], c1[keep]
    mid = (0.5 * (cpos[c1] + cpos[partner[c1]])).astype(np.float32)
    case["x"][tied], case["y"][tied], case["z"][tied] = mid[:, 0], mid[:, 1], mid[:, 2]
    return case, tied


def deformation_slab(n_vertices_total: int, n_controls: int, first: int, count: int,
                     ctrl_per_kf: int = 10, seed: int = 42, t0_ns: int = T0_NS,
                     vertex_spread: float = 6.0, ctrl_spread: float = 5.0):
    """Same workload family as deformation_case, but only vertices [first, first + count) are
    materialised (multi-GPU ranks generate their own slab); the control table is global."""
    ctrl = deformation_case(0, n_controls, ctrl_per_kf, seed, t0_ns, 0.0, vertex_spread,
                            ctrl_spread)
    n_kf = max(n_controls // ctrl_per_kf, 1)
    kf = _trajectory(n_kf)
    kf_stamp = (np.uint64(t0_ns) + np.arange(n_kf, dtype=np.uint64) * np.uint64(1_000_000_000))
    per_kf = -(-n_vertices_total // n_kf)
    idx = np.arange(first, first + count, dtype=np.int64)
    v_kf = np.minimum(idx // per_kf, n_kf - 1)
    rng = np.random.default_rng([seed, first, count])
    base = kf[v_kf].astype(np.float32)
    x = base[:, 0] + rng.uniform(-vertex_spread, vertex_spread, count).astype(np.float32)
    y = base[:, 1] + rng.uniform(-vertex_spread, vertex_spread, count).astype(np.float32)
    z = rng.uniform(0.0, 3.0, count).astype(np.float32)
    ctrl.update(x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), z=np.ascontiguousarray(z),
                stamps=np.ascontiguousarray(kf_stamp[v_kf]))
    return ctrl


def deformation_range(n_vertices_total: int, n_controls: int, first: int, count: int,
                      ctrl_per_kf: int = 10, seed: int = 42, t0_ns: int = T0_NS,
                      vertex_spread: float = 6.0, ctrl_spread: float = 5.0,
                      chunk: int = 1 << 20, with_controls: bool = True):
    """Vertices [first, first + count) of a mesh whose every vertex is a function of (seed, its
    global index) only: generated in fixed chunks of `chunk` vertices, each from its own stream,
    so that any partition of the index range into slabs materialises the identical mesh (shard
    invariance tests, configs[3]).  Same workload family as deformation_case."""
    ctrl = (deformation_case(0, n_controls, ctrl_per_kf, seed, t0_ns, 0.0, vertex_spread,
                             ctrl_spread) if with_controls else {})
    n_kf = max(n_controls // ctrl_per_kf, 1)
    kf = _trajectory(n_kf).astype(np.float32)
    kf_stamp = (np.uint64(t0_ns) + np.arange(n_kf, dtype=np.uint64) * np.uint64(1_000_000_000))
    per_kf = -(-n_vertices_total // n_kf)
    x = np.empty(count, dtype=np.float32)
    y = np.empty(count, dtype=np.float32)
    z = np.empty(count, dtype=np.float32)
    last = first + count
    c = first // chunk
    while c * chunk < last:
        lo, hi = c * chunk, min((c + 1) * chunk, n_vertices_total)
        rng = np.random.default_rng([seed, 7919, c])
        n = hi - lo
        dx = rng.uniform(-vertex_spread, vertex_spread, n).astype(np.float32)
        dy = rng.uniform(-vertex_spread, vertex_spread, n).astype(np.float32)
        dz = rng.uniform(0.0, 3.0, n).astype(np.float32)
        a, b = max(lo, first), min(hi, last)
        if b > a:
            v_kf = np.minimum(np.arange(a, b, dtype=np.int64) // per_kf, n_kf - 1)
            x[a - first:b - first] = kf[v_kf, 0] + dx[a - lo:b - lo]
            y[a - first:b - first] = kf[v_kf, 1] + dy[a - lo:b - lo]
            z[a - first:b - first] = dz[a - lo:b - lo]
        c += 1
    v_kf = np.minimum(np.arange(first, last, dtype=np.int64) // per_kf, n_kf - 1)
    ctrl.update(x=x, y=y, z=z, stamps=np.ascontiguousarray(kf_stamp[v_kf]))
    return ctrl


# kimera_pgmo/include/kimera_pgmo/utils/common_functions.h:19-61
ROBOT_ID_TO_PREFIX = {i: chr(ord("a") + i) for i in range(8)}
ROBOT_ID_TO_VERTEX_PREFIX = {i: chr(ord("s") + i) for i in range(8)}


def symbol_key(prefix: str, index: int) -> int:
    """gtsam::Symbol(chr, idx).key() = (chr << 56) | idx (test/data/graph.dgrf: 'a' ->
    6989586621679009792)."""
    return (ord(prefix) << 56) | int(index)


def robot_range(robot_id: int, n_vertices: int, n_controls: int, first: int, count: int, **kw):
    """configs[3]: robot `robot_id` of a multi-robot mesh — its own trajectory phase, stamps and
    control table (seed and epoch offset by the robot id)."""
    return deformation_range(n_vertices, n_controls, first, count, seed=42 + 1000 * robot_id,
                             t0_ns=T0_NS + robot_id * 3_600_000_000_000, **kw)


def lattice_case(n_side: int = 12, seed: int = 7, t0_ns: int = T0_NS, n_groups: int = 4,
                 points_per_voxel: int = 1):
    """Tie-heavy case: vertices and control points on a 0.5 m lattice, so exact float-distance
    ties are the norm and the PCL traversal order decides associations."""
    rng = np.random.default_rng(seed)
    g = np.arange(n_side, dtype=np.float32) * np.float32(0.5)
    vx, vy, vz = np.meshgrid(g, g, g[: max(n_side // 3, 2)], indexing="ij")
    x, y, z = vx.ravel(), vy.ravel(), vz.ravel()
    n_v = x.shape[0]
    grp = (np.arange(n_v) * n_groups) // n_v
    stamps = np.uint64(t0_ns) + grp.astype(np.uint64) * np.uint64(4_000_000_000)
    # control points: every second lattice node (+ optional extra points inside the same voxel)
    c = []
    for i in range(0, n_side, 2):
        for j in range(0, n_side, 2):
            for r in range(points_per_voxel):
                c.append((g[i] + 0.125 * r, g[j], 0.5 * (i % 3) + 0.125 * r))
    cpos = np.array(c, dtype=np.float32).astype(np.float64)
    n_c = cpos.shape[0]
    cgrp = (np.arange(n_c) * n_groups) // n_c
    cstamps = np.uint64(t0_ns) + cgrp.astype(np.uint64) * np.uint64(4_000_000_000)
    ang = rng.uniform(-0.1, 0.1, n_c)
    poses = np.empty((n_c, 12), dtype=np.float64)
    poses[:, :9] = _rot_zy(ang, 0.5 * ang)
    poses[:, 9:] = cpos + rng.uniform(-0.3, 0.3, (n_c, 3))
    return dict(x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), z=np.ascontiguousarray(z),
                stamps=np.ascontiguousarray(stamps), ctrl_xyz=cpos, ctrl_stamps=cstamps,
                valid=np.ones(n_c, dtype=np.uint8), poses=poses)


def compression_stream(n_messages: int, verts_per_message: int, n_blocks: int = 1,
                       seed: int = 42, slide: float = 1.0, wall: float = 20.0):
    """configs[1]: a stream of mesh messages.  Each message is `verts_per_message` unindexed
    vertices (3 consecutive = 1 small triangle, edge 0.03-0.15 m) on a `wall` x `wall` m patch
    sliding `slide` m per message, so most voxels are re-observed.  Triangles are drawn from a
    fixed per-cell pool so that re-observed places repeat the same geometry.  Yields
    (block_idx[n_blocks,3], offsets[n_blocks+1], xyz[N,3], rgba[N], stamp_s)."""
    rng = np.random.default_rng(seed)
    n_tri = verts_per_message // 3
    for m in range(n_messages):
        # triangle anchors: deterministic function of (global cell) so overlap repeats exactly
        u = rng.uniform(0.0, wall, n_tri) + slide * m
        v = rng.uniform(0.0, wall, n_tri)
        cell_u = np.floor(u / 0.1)
        cell_v = np.floor(v / 0.1)
        hu = np.modf(np.sin(cell_u * 12.9898 + cell_v * 78.233) * 43758.5453)[0]
        hv = np.modf(np.sin(cell_u * 39.3468 + cell_v * 11.135) * 24634.6345)[0]
        ax = (cell_u * 0.1 + 0.05 * np.abs(hu)).astype(np.float32)
        ay = (cell_v * 0.1 + 0.05 * np.abs(hv)).astype(np.float32)
        e = (0.03 + 0.12 * np.abs(hu)).astype(np.float32)
        tri = np.empty((n_tri, 3, 3), dtype=np.float32)
        tri[:, 0, 0] = ax
        tri[:, 0, 1] = ay
        tri[:, 1, 0] = ax + e
        tri[:, 1, 1] = ay
        tri[:, 2, 0] = ax
        tri[:, 2, 1] = ay + e
        tri[:, :, 2] = (0.5 * np.abs(hv)).astype(np.float32)[:, None]
        if n_blocks > 1:
            # voxblox-like blocks: group triangles by 0.8 m block of their first vertex
            bu = np.floor(tri[:, 0, 0] / 0.8).astype(np.int64)
            bv = np.floor(tri[:, 0, 1] / 0.8).astype(np.int64)
            key = bu * 100003 + bv
            order = np.argsort(key, kind="stable")
            tri = tri[order]
            key = key[order]
            uniq, first = np.unique(key, return_index=True)
            sel = np.argsort(first)[:n_blocks]  # keep block order of first appearance... capped
            bounds = np.sort(first)
            # merge everything beyond n_blocks into the last block
            starts = list(bounds[:n_blocks])
            offsets = np.array([3 * s for s in starts] + [3 * n_tri], dtype=np.uint64)
            blocks = np.stack([bu[order][bounds[:n_blocks]], bv[order][bounds[:n_blocks]],
                               np.zeros(len(starts), dtype=np.int64)], axis=1).astype(np.int32)
            del sel, uniq
        else:
            offsets = np.array([0, 3 * n_tri], dtype=np.uint64)
            blocks = np.zeros((1, 3), dtype=np.int32)
        xyz = np.ascontiguousarray(tri.reshape(-1, 3))
        rgba = (np.uint32(0xFF000000) | (np.arange(xyz.shape[0], dtype=np.uint32) & np.uint32(0xFFFFFF)))
        yield blocks, offsets, xyz, np.ascontiguousarray(rgba), 100.0 + 1.0 * m
