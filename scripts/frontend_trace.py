"""Per-message wall time of the MeshFrontendInterface stream of bench.py (VoxelClearing 0.05 m +
Octree 0.5 m + mesh -> graph bookkeeping, 16 blocks per message) with allocation tracing
(PGMO_TRACE_ALLOC=1): finds what is behind a slow message."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
msgs = list(synth.compression_stream(n, 50_000, n_blocks=16))
full, graph, m2g = capi.VccHandle(0.05, 0), capi.OctHandle(0.5, 0), capi.MeshToGraphHandle(0)
for m, (blocks, offsets, xyz, rgba, t) in enumerate(msgs):
    sys.stderr.write(f"--- message {m}\n")
    sys.stderr.flush()
    ts = [time.perf_counter()]
    full.prune(t - 10.0)
    ts.append(time.perf_counter())
    graph.prune(t - 10.0)
    ts.append(time.perf_counter())
    remap, _ = full.integrate(blocks, offsets, xyz, rgba, t)
    ts.append(time.perf_counter())
    gi, gk = graph.integrate(blocks, offsets, xyz, rgba, t)
    ts.append(time.perf_counter())
    m2g.update(blocks, offsets, remap.view(np.int64), None, gi, gk)
    ts.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(ts[:-1], ts[1:])]
    sys.stderr.write(f"message {m}: vcc prune {d[0]:.3f}, oct prune {d[1]:.3f}, vcc integrate {d[2]:.3f}, "
                     f"oct integrate {d[3]:.3f}, m2g {d[4]:.3f}, total {sum(d):.3f} ms\n")
