"""BASELINE configs[4]: OctreeCompression with horizon eviction plus full re-deformation per loop
closure over a synthetic keyframe trajectory.

Per keyframe: one 50k-vertex mesh message -> pgmo_oct_prune(t - horizon) + pgmo_oct_integrate.
Every `closure_every` keyframes: the whole accumulated (compressed) mesh is deformed from
scratch (start index 0) with the control points / poses known so far —
pgmo_deform_set_controls + pgmo_deform_points with host buffers, window schedule rebuilt because
the control table grew.  Prints one JSON line: per-message compression latency (mean / p50 / p99)
and, per closure, mesh size, control count and deformation time.
  python scripts/config5.py [n_keyframes=1000] [closure_every=50]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kimera_pgmo_b200 import scenarios  # noqa: E402

n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
closure_every = int(sys.argv[2]) if len(sys.argv) > 2 else 50
print(json.dumps(scenarios.run_config4(n_kf, closure_every)))
