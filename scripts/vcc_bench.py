"""Compression leg of bench.py alone (VoxelClearing + Octree on configs[1]); used for profiling."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
print(json.dumps(bench.bench_compression(0, n_messages=n, with_cpu=False)))
