import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kimera_pgmo_b200 import capi, synth
from oracle import ffi
for res, nb, verts in ((0.05, 1, 50000), (0.2, 7, 6000)):
    gpu = capi.OctHandle(res, 0); ora = ffi.Compression("octree", res)
    for m, (blocks, offsets, xyz, rgba, t) in enumerate(synth.compression_stream(8, verts, n_blocks=nb, seed=3)):
        gpu.prune(t - 3.0); ora.prune(t - 3.0)
        t0 = time.perf_counter(); gi, gk = gpu.integrate(blocks, offsets, xyz, rgba, t); t1 = time.perf_counter()
        oi, ok, _ = ora.integrate(blocks, offsets, xyz, rgba, t); t2 = time.perf_counter()
        print(res, m, 'kinds', np.bincount(gk, minlength=3), 'sizes', gpu.sizes(), 'match', np.array_equal(gi, oi) and np.array_equal(gk, ok), f'gpu {1e3*(t1-t0):.2f} ms cpu {1e3*(t2-t1):.2f} ms')
