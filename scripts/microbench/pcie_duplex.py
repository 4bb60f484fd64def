"""PCIe copy rates of the box (pinned host memory): H2D alone, D2H alone, both at once — the floor
of the end-to-end deformation step (12 B/vertex up, 12 B/vertex down)."""
import time

import torch

n = 120_000_000
dev = torch.device("cuda", 0)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def up():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    up()
    down()


for name, fn, nbytes in (("H2D alone", up, n), ("D2H alone", down, n), ("H2D + D2H at once", both, 2 * n)):
    t = timed(fn)
    print(f"{name}: {1e3 * t:.3f} ms per {nbytes / 1e6:.0f} MB = {nbytes / t / 1e9:.1f} GB/s")
