"""Per-message wall time of the DeltaCompression stream of bench.py with allocation tracing
(PGMO_TRACE_ALLOC=1 in the environment makes the library report every allocation on stderr):
finds what is behind a slow message."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import capi, synth  # noqa: E402

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
msgs = list(synth.compression_stream(n, 50_000, n_blocks=n_blocks))
dc = capi.DeltaHandle(0.05, 0)
for m, (blocks, offsets, xyz, rgba, t) in enumerate(msgs):
    sys.stderr.write(f"--- message {m}\n")
    sys.stderr.flush()
    t0 = time.perf_counter()
    dc.archive_by_time(int((t - 10.0) * 1e9))
    t1 = time.perf_counter()
    dc.update(blocks, offsets, xyz, rgba, int(t * 1e9), fetch=False)
    t2 = time.perf_counter()
    sys.stderr.write(f"message {m}: archive {1e3 * (t1 - t0):.3f} ms, update {1e3 * (t2 - t1):.3f} ms\n")
dc.close()
