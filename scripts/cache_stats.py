"""Bank-conflict model of the blend kernel on the real association records (diagnostics):
per quarter warp (8 consecutive records) and neighbour slot, the number of distinct pose rows
and the largest number of distinct rows sharing a bank class (= shared-memory wavefronts)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n_v, n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 5_000
if len(sys.argv) > 2:
    n_c = int(sys.argv[2])
case = synth.deformation_case(n_v, n_c)
h = capi.DeformHandle(0)
h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
r = h.deform_points(case["x"], case["y"], case["z"], case["stamps"], k=4, tol_t=10.0)
rec = h.cache_records().astype(np.int64)
n = rec.shape[0] // 8 * 8
rec = rec[:n]
nbr = rec[:, 1:6]
row, cls = nbr & 0x1FFF, nbr >> 13
q = n // 8
row = row.reshape(q, 8, 5)
cls = cls.reshape(q, 8, 5)
for j in range(5):
    d = np.zeros(q, dtype=np.int64)
    worst = np.zeros(q, dtype=np.int64)
    worst_idx = np.zeros(q, dtype=np.int64)
    worst_rand = np.zeros(q, dtype=np.int64)
    rj, cj = row[:, :, j], cls[:, :, j]
    srt = np.sort(rj, axis=1)
    distinct = np.ones((q, 8), dtype=bool)
    distinct[:, 1:] = srt[:, 1:] != srt[:, :-1]
    d = distinct.sum(axis=1)
    # per class: distinct rows
    for c in range(8):
        key = np.where(cj == c, rj, -1)
        s2 = np.sort(key, axis=1)
        dd = np.ones((q, 8), dtype=bool)
        dd[:, 1:] = s2[:, 1:] != s2[:, :-1]
        cnt = (dd & (s2 >= 0)).sum(axis=1)
        worst = np.maximum(worst, cnt)
        key = np.where((rj & 7) == c, rj, -1)
        s2 = np.sort(key, axis=1)
        dd = np.ones((q, 8), dtype=bool)
        dd[:, 1:] = s2[:, 1:] != s2[:, :-1]
        cnt = (dd & (s2 >= 0)).sum(axis=1)
        worst_idx = np.maximum(worst_idx, cnt)
    print(f"slot {j}: distinct rows per quarter warp {d.mean():.2f}; wavefronts with classes {worst.mean():.3f}, "
          f"with row & 7 {worst_idx.mean():.3f}")


def stats(row, cls, label):
    q = row.shape[0] // 8
    row = row[: q * 8].reshape(q, 8, 5)
    cls = cls[: q * 8].reshape(q, 8, 5)
    out = []
    for j in range(5):
        rj, cj = row[:, :, j], cls[:, :, j]
        srt = np.sort(rj, axis=1)
        distinct = np.ones((q, 8), dtype=bool)
        distinct[:, 1:] = srt[:, 1:] != srt[:, :-1]
        d = distinct.sum(axis=1)
        worst = np.zeros(q, dtype=np.int64)
        for c in range(8):
            key = np.where(cj == c, rj, -1)
            s2 = np.sort(key, axis=1)
            dd = np.ones((q, 8), dtype=bool)
            dd[:, 1:] = s2[:, 1:] != s2[:, :-1]
            worst = np.maximum(worst, (dd & (s2 >= 0)).sum(axis=1))
        out.append((d.mean(), worst.mean()))
    print(label, " ".join(f"d={a:.2f}/w={b:.3f}" for a, b in out), "mean w %.3f" % np.mean([b for _, b in out]))


per_kf = -(-n_v // (n_c // 10))
nb = rec[:, 1:6]
R, Cc = nb & 0x1FFF, nb >> 13
stats(R, Cc, "as recorded      ")
# lexicographic by neighbour list inside each keyframe block (~ work item)
order = []
for b0 in range(0, n, per_kf):
    blk = R[b0:b0 + per_kf]
    o = np.lexsort((blk[:, 4], blk[:, 3], blk[:, 2], blk[:, 1], blk[:, 0]))
    order.append(b0 + o)
order = np.concatenate(order)
stats(R[order], Cc[order], "lexicographic    ")
# by sorted neighbour SET
order = []
for b0 in range(0, n, per_kf):
    blk = np.sort(R[b0:b0 + per_kf], axis=1)
    o = np.lexsort((blk[:, 4], blk[:, 3], blk[:, 2], blk[:, 1], blk[:, 0]))
    order.append(b0 + o)
order = np.concatenate(order)
stats(R[order], Cc[order], "by neighbour set ")
