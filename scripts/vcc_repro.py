"""Repeats the host-buffer VoxelClearing stream of bench.py on fresh handles and reports where a
device-side error shows up (diagnostics for a flaky 'previous voxel missing from the table')."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
msgs = list(synth.compression_stream(n, 50_000, n_blocks=1))
for rep in range(reps):
    vcc = capi.VccHandle(0.05, 0)
    try:
        for m, (blocks, offsets, xyz, rgba, t) in enumerate(msgs):
            vcc.prune(t - 10.0)
            vcc.integrate(blocks, offsets, xyz, rgba, t)
        print("rep", rep, "ok", vcc.sizes())
    except capi.PgmoError as e:
        print("rep", rep, "FAILED at message", m, str(e)[:120])
    vcc.close()
