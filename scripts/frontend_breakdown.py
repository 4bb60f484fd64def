"""Per-call wall time of the MeshFrontendInterface path (VoxelClearing full mesh + Octree graph
compressor + mesh -> graph bookkeeping) for 1 and 16 blocks per message; second pass reported
(the first process on a fresh box is cold).  python scripts/frontend_breakdown.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kimera_pgmo_b200 import capi, synth
for rep in range(2):
  for nb in (1, 16):
    msgs = list(synth.compression_stream(40, 50000, n_blocks=nb))
    full, graph, m2g = capi.VccHandle(0.05, 0), capi.OctHandle(0.5, 0), capi.MeshToGraphHandle(0)
    rows = []
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(msgs):
        ts = [time.perf_counter()]
        full.prune(t - 10.0); ts.append(time.perf_counter())
        graph.prune(t - 10.0); ts.append(time.perf_counter())
        remap, _ = full.integrate(blocks, offsets, xyz, rgba, t); ts.append(time.perf_counter())
        gi, gk = graph.integrate(blocks, offsets, xyz, rgba, t); ts.append(time.perf_counter())
        m2g.update(blocks, offsets, remap.view(np.int64), None, gi, gk); ts.append(time.perf_counter())
        if m >= 5:
            rows.append(1e3 * np.diff(ts))
    rows = np.array(rows)
    if rep:
        print(nb, 'blocks: vcc.prune, oct.prune, vcc.integrate, oct.integrate, m2g.update; p50', np.round(np.median(rows, 0), 3),
              'mean', np.round(rows.mean(0), 3), 'max', np.round(rows.max(0), 2), 'n_blocks/msg', len(msgs[-1][0]))
