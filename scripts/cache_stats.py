"""Shared-memory wavefront model of the blend kernel on the real association records
(diagnostics).  An LDS.128 is served per quarter warp (8 consecutive records); its cost is the
largest number of DIFFERENT addresses falling into one 16-byte bank group.  A neighbour field is
rel | copy << 11 | class << 13: the address is (rel, copy), the bank group of chunk c is
c ^ class.  Run with PGMO_BLEND_COPIES=1 to see the single-placement figure."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n_v = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_c = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000
case = synth.deformation_case(n_v, n_c)
h = capi.DeformHandle(0)
h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
rec = h.cache_records().astype(np.int64)
n = rec.shape[0] // 32 * 32
nbr = rec[:n, 1:6]
addr = nbr & 0x1FFF            # (rel, copy)
cls = nbr >> 13
copy = (nbr >> 11) & 3
q = n // 8
A = addr.reshape(q, 8, 5)
C = cls.reshape(q, 8, 5)
print("copies used:", {int(m): float((copy == m).mean()) for m in range(4)})
tot = 0.0
for j in range(5):
    aj, cj = A[:, :, j], C[:, :, j]
    srt = np.sort(aj, axis=1)
    d = 1 + (srt[:, 1:] != srt[:, :-1]).sum(axis=1)
    worst = np.zeros(q, dtype=np.int64)
    for g in range(8):
        key = np.where(cj == g, aj, -1)
        s2 = np.sort(key, axis=1)
        first = np.ones((q, 8), dtype=bool)
        first[:, 1:] = s2[:, 1:] != s2[:, :-1]
        worst = np.maximum(worst, (first & (s2 >= 0)).sum(axis=1))
    tot += worst.mean()
    print(f"slot {j}: distinct rows per quarter warp {d.mean():.2f}, wavefronts per quarter warp {worst.mean():.3f}")
print(f"mean wavefronts per quarter warp and LDS.128: {tot / 5:.3f}")
u = np.sort((nbr & 0x7FF).reshape(n // 32, 160), axis=1)
print("distinct rows (union over slots) per group of 32 records: mean %.2f"
      % (1 + (u[:, 1:] != u[:, :-1]).sum(axis=1)).mean())


# --- would another order of the 32 records of a group help?  (the search kernel is free to
# choose it.)  Needs the vertex of every record: synth.deformation_case gives every keyframe
# per_kf vertices = one head item of 1 vertex + one item with the rest.
n_kf = max(n_c // 10, 1)
per_kf = -(-n_v // n_kf)
if per_kf <= 2049:
    slot = np.arange(n)
    within = slot % per_kf
    item_start = slot - within + (within > 0)
    vert = item_start + rec[:n, 0]
    vx, vy, vz = case["x"][vert], case["y"][vert], case["z"][vert]
    rank_in_item = slot - item_start
    grp = (item_start.astype(np.int64) << 8) + rank_in_item // 32      # group id
    full = np.bincount(np.unique(grp, return_inverse=True)[1])          # group sizes
    gid = np.unique(grp, return_inverse=True)[1]
    keep = full[gid] == 32
    R5 = (nbr & 0x7FF)[keep]
    g_of = gid[keep]
    def morton(cell):
        qx, qy, qz = (np.floor(c / cell).astype(np.int64) & 1023 for c in (vx[keep], vy[keep], vz[keep]))
        m = np.zeros_like(qx)
        for b in range(10):
            m |= ((qx >> b) & 1) << (3 * b) | ((qy >> b) & 1) << (3 * b + 1) | ((qz >> b) & 1) << (3 * b + 2)
        return m

    for label, key in (("as recorded", rank_in_item[keep] % 32), ("by z", vz[keep]),
                       ("by x", vx[keep]), ("by y", vy[keep]), ("by Morton code, 0.75 m cells", morton(0.75)),
                       ("by Morton code, 0.4 m cells", morton(0.4))):
        o = np.lexsort((key, g_of))
        Rr = R5[o].reshape(-1, 8, 5)
        srt = np.sort(Rr, axis=1)
        d = 1 + (srt[:, 1:, :] != srt[:, :-1, :]).sum(axis=1)
        print(f"records of a group ordered {label}: distinct rows per quarter warp and slot {d.mean():.2f}")
