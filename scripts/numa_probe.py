"""Where the GPU sits (NUMA node, PCIe) and what pinning the process to each NUMA node does to
the page-locked copy rates (H2D + D2H at once, 120 MB each)."""
import glob
import os
import subprocess
import sys
import time

import torch

print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'; nvidia-smi topo -m; "
                     "nvidia-smi --query-gpu=index,pci.bus_id --format=csv", shell=True,
                     capture_output=True, text=True).stdout)
p = torch.cuda.get_device_properties(0)
bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
try:
    print("gpu0", bdf, "numa_node file:", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
except Exception as e:  # noqa: BLE001
    print("numa_node unreadable:", e)
print("allowed cpus:", sorted(os.sched_getaffinity(0))[:8], "... n =", len(os.sched_getaffinity(0)))
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes:", nodes)
allowed0 = os.sched_getaffinity(0)


def cpus_of(node_dir):
    cpus = []
    for part in open(node_dir + "/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return [c for c in cpus if c in allowed0]


n = 120_000_000
dev = torch.device("cuda", 0)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for node_dir in nodes + [None]:
    cpus = cpus_of(node_dir) if node_dir else sorted(allowed0)
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    both()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        both()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 10
    print(f"process on {os.path.basename(node_dir) if node_dir else 'all allowed cpus'} ({len(cpus)} cpus): "
          f"up + down at once {1e3 * t:.3f} ms = {2 * n / t / 1e9:.1f} GB/s", flush=True)
    del h_in, h_out
