"""A/B of blend-kernel build variants on configs[2] (10 M vertices, 50 k controls): first deform
and pose-only step per variant library, outputs compared bit for bit across variants.
    python scripts/blend_variants.py build/variants/*.so"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import hashlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from kimera_pgmo_b200 import capi, synth
n_v, n_c = 10_000_000, 50_000
case = synth.deformation_case(n_v, n_c)
dev = torch.device("cuda", 0)
d = {k: torch.from_numpy(case[k]).to(dev) for k in "xyz"}
d_st = torch.from_numpy(case["stamps"].view(np.int64)).to(dev)
out = [torch.empty(n_v, dtype=torch.float32, device=dev) for _ in range(3)]
h = capi.DeformHandle(0)
h.set_controls(case["ctrl_xyz"], case["ctrl_stamps"], case["valid"], case["poses"])
h.set_profiling(True)
stream = torch.cuda.current_stream().cuda_stream
ss, sl = h.stamp_stretches_device(d_st, n_v, stream)
assert h.build_schedule(ss, sl, n_v, 4, 10.0) == 0
poses2 = case["poses"].copy(); poses2[:, 9:] += 0.25
res = {}
def run():
    h.run_device(0, n_v, d["x"], d["y"], d["z"], out[0], out[1], out[2], precision=0, stream=stream)
    torch.cuda.synchronize()
    return h.last_step_info()
first = run()
res["first"] = first
warm = []
for i in range(8):
    h.update_poses(poses2 if i %% 2 == 0 else case["poses"])
    warm.append(run())
res["warm_blend_ms"] = float(np.median([w["blend_ms"] for w in warm]))
res["warm_total_ms"] = float(np.median([w["total_ms"] for w in warm]))
res["warm"] = all(w["warm"] for w in warm)
h.update_poses(poses2)
run()
m = hashlib.sha256()
for o in out:
    m.update(o.cpu().numpy().tobytes())
res["sha"] = m.hexdigest()[:16]
print("RESULT " + json.dumps(res))
''' % ROOT

rows = []
for lib in sys.argv[1:]:
    for sort in ("1", "0"):
        env = dict(os.environ, PGMO_LIB_PATH=os.path.abspath(lib), PGMO_SORT_RECORDS=sort)
        p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        if not line:
            print(lib, "sort", sort, "FAILED", p.stderr[-1500:])
            continue
        r = json.loads(line[0][7:])
        f = r["first"]
        print(f"{os.path.basename(lib):28s} sort={sort}: first deform search {f['search_ms']:.3f} + resolver/sort {f['resolver_ms']:.3f} "
              f"+ blend {f['blend_ms']:.3f} = {f['total_ms']:.3f} ms | pose-only blend {r['warm_blend_ms']:.3f} total {r['warm_total_ms']:.3f} ms "
              f"warm={r['warm']} sha={r['sha']}", flush=True)
