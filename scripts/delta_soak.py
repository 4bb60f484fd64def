"""Long DeltaCompression stream against the oracle: `n` 50k-vertex messages in 64 voxblox-like
blocks, archiveBlocksByTime(t - 10 s) before each; every `check_every`-th message all outputs are
compared bit for bit with the CPU oracle (which is fed every message).  Prints a JSON line.
  python scripts/delta_soak.py [n=300] [check_every=10]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402
from oracle import ffi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
check_every = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gpu, ora = capi.DeltaHandle(0.05, 0), ffi.DeltaCompression(0.05)
lat, cpu, mismatches, checked = [], [], 0, 0
archived = 0
for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(n, 50_000, n_blocks=64)):
    cut = int((t - 10.0) * 1e9)
    t0 = time.perf_counter()
    gpu.archive_by_time(cut)
    gr, _ = gpu.update(blocks, offsets, xyz, rgba, int(t * 1e9), fetch=False)
    lat.append(1e3 * (time.perf_counter() - t0))
    t0 = time.perf_counter()
    ora.archive_by_time(cut)
    orr, od = ora.update(blocks, offsets, xyz, rgba, int(t * 1e9))
    cpu.append(1e3 * (time.perf_counter() - t0))
    archived += od["num_archived_vertices"]
    bad = not np.array_equal(gr, orr)
    if m % check_every == 0 or m == n - 1:
        gd = gpu.delta()
        checked += 1
        for key in ("vertex_start", "face_start", "num_archived_vertices", "pending_archived_faces"):
            bad = bad or gd[key] != od[key]
        for key in ("vertices", "rgba", "stamps", "faces", "faces_archive", "prev_to_curr", "deleted", "observed", "new"):
            bad = bad or not np.array_equal(gd[key], od[key])
    mismatches += int(bad)
lat, cpu = np.array(lat[5:]), np.array(cpu[5:])
stats = gpu.last_stats()
print(json.dumps({
    "scenario": f"DeltaCompression, {n} messages x 50000 vertices in 64 blocks, res 0.05 m, archiveBlocksByTime(t - 10 s)",
    "messages_with_any_difference": mismatches, "full_delta_comparisons": checked,
    "gpu_ms_per_message": {"mean": float(lat.mean()), "p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))},
    "oracle_ms_per_message": {"mean": float(cpu.mean())},
    "archived_vertices_total": int(archived), "active_vertices_at_end": int(od["active_voxels"]),
    "voxel_records_at_end": stats[1], "blocks_at_end": stats[2], "record_compactions": stats[3]}))
sys.exit(0 if mismatches == 0 else 1)
