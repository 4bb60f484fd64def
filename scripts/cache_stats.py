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


# --- what the shared-memory pipe sees: distinct pose rows per quarter warp (8 consecutive
# records) and neighbour slot.  LDS.128 with <= 2 distinct addresses per quarter warp costs
# ~2.5 cycles per warp instruction, more than that 4 (scripts/microbench/smem_fp64.cu).
def quarter_histogram(row, label):
    q = row.shape[0] // 8
    r = np.sort(row[: q * 8].reshape(q, 8, 5), axis=1)
    d = 1 + (r[:, 1:, :] != r[:, :-1, :]).sum(axis=1)          # (q, 5)
    hist = np.bincount(d.ravel(), minlength=9)[1:] / d.size
    print(label, "distinct rows per quarter warp and slot: " +
          " ".join(f"{i + 1}:{h:.3f}" for i, h in enumerate(hist[:6])),
          f"| <=2: {hist[:2].sum():.3f}  mean {d.mean():.2f}")


quarter_histogram(R, "as recorded          ")
g = n // 32
R32 = R[: g * 32].reshape(g, 32, 5)
# sorted inside every group of 32 records by the whole ordered list
key = np.zeros((g, 32), dtype=np.int64)
for j in range(5):
    key = key * 8192 + R32[:, :, j]
o = np.argsort(key, axis=1, kind="stable")
quarter_histogram(np.take_along_axis(R32, o[:, :, None], axis=1).reshape(-1, 5), "sorted in 32-groups  ")
# ... by a 32-bit key: first two neighbours + 6 hashed bits of the rest
h6 = (R32[:, :, 2] * 7 + R32[:, :, 3] * 13 + R32[:, :, 4] * 29) & 63
key32 = (R32[:, :, 0] << 19) | (R32[:, :, 1] << 6) | h6
o = np.argsort(key32, axis=1, kind="stable")
quarter_histogram(np.take_along_axis(R32, o[:, :, None], axis=1).reshape(-1, 5), "32-groups, 32-bit key")
# sorted inside groups of 256 (one slab)
g2 = n // 256
R256 = R[: g2 * 256].reshape(g2, 256, 5)
key = np.zeros((g2, 256), dtype=np.int64)
for j in range(5):
    key = key * 8192 + R256[:, :, j]
o = np.argsort(key, axis=1, kind="stable")
quarter_histogram(np.take_along_axis(R256, o[:, :, None], axis=1).reshape(-1, 5), "sorted in 256-groups ")
# distinct ordered lists / distinct rows (any slot) per group of 32
lists = np.sort(key.reshape(-1, 32), axis=1) if False else None
k32 = np.zeros((g, 32), dtype=np.int64)
for j in range(5):
    k32 = k32 * 8192 + R32[:, :, j]
ks = np.sort(k32, axis=1)
print("distinct ordered lists per 32-group: mean %.2f" % (1 + (ks[:, 1:] != ks[:, :-1]).sum(axis=1)).mean())
u = np.sort(R32.reshape(g, 160), axis=1)
print("distinct rows (union over slots) per 32-group: mean %.2f" % (1 + (u[:, 1:] != u[:, :-1]).sum(axis=1)).mean())


# --- two placements per row (copy A: class of the colouring, copy B: another class): a greedy
# pass over the 8 records of a quarter warp picks, per distinct row, a placement whose bank group
# is still free.  Result: mean shared-memory wavefronts per quarter warp and slot.
def two_choice(row, cls, label):
    q = row.shape[0] // 8
    r = row[: q * 8].reshape(q, 8, 5)
    a = cls[: q * 8].reshape(q, 8, 5)
    b = (a + 1 + r % 7) & 7
    tot_one, tot_two = 0.0, 0.0
    for j in range(5):
        rj, aj, bj = r[:, :, j], a[:, :, j], b[:, :, j]
        used = np.zeros(q, dtype=np.int64)          # bank groups taken
        load = np.zeros((q, 8), dtype=np.int64)     # distinct rows per bank group
        load_one = np.zeros((q, 8), dtype=np.int64)
        for i in range(8):
            first = np.ones(q, dtype=bool)
            for p in range(i):
                first &= rj[:, p] != rj[:, i]
            ca, cb = aj[:, i], bj[:, i]
            free_a = ((used >> ca) & 1) == 0
            free_b = ((used >> cb) & 1) == 0
            pick = np.where(free_a, ca, np.where(free_b, cb, ca))
            used = np.where(first, used | (1 << pick), used)
            np.add.at(load, (np.nonzero(first)[0], pick[first]), 1)
            np.add.at(load_one, (np.nonzero(first)[0], ca[first]), 1)
        tot_one += load_one.max(axis=1).mean()
        tot_two += load.max(axis=1).mean()
    print(label, "wavefronts per quarter warp and slot: one placement %.3f, two placements (greedy) %.3f"
          % (tot_one / 5, tot_two / 5))


two_choice(R, Cc, "as recorded")
