import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kimera_pgmo_b200 import capi, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
msgs = list(synth.compression_stream(40, 50000, n_blocks=nb))
for rep in range(2):
    dc = capi.DeltaHandle(0.05, 0)
    lat = []
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(msgs):
        t0 = time.perf_counter(); dc.archive_by_time(int((t - 10.0) * 1e9)); t1 = time.perf_counter()
        dc.update(blocks, offsets, xyz, rgba, int(t * 1e9), fetch=False); t2 = time.perf_counter()
        lat.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), dc.last_stats()[0]))
    lat = np.array(lat[5:])
    print("blocks", nb, "archive ms", lat[:, 0].mean(), "update ms", lat[:, 1].mean(), "p50", np.percentile(lat[:, 1], 50), "rounds", lat[:, 2].mean(), dc.sizes()[10], dc.last_stats())
    dc.close()
print(np.round(lat[:, 1], 1).tolist())
