"""Host-side plumbing of the vertex-range sharded deformation.

One process per GPU.  Nothing here touches vertex data: ranks own contiguous slabs of the
processing order, the control table / poses are replicated by a broadcast, and the only exchange
is the list of equal-stamp stretches of every slab (a few KB), all-gathered and concatenated in
rank order so that every rank replays the identical window state machine
(pgmo_deform_build_schedule).  torch.distributed is plumbing: NCCL on the GPU box, gloo in the
CPU tests (tests/test_multi_rank_cpu.py).
"""
from __future__ import annotations

import numpy as np


def slab_bounds(n_total: int, world: int, rank: int):
    """Contiguous vertex range [first, first + count) of `rank`."""
    first = n_total * rank // world
    last = n_total * (rank + 1) // world
    return first, last - first


def stretches_of(stamps: np.ndarray):
    """(stamp, length) of the maximal equal-stamp stretches — what K1
    (pgmo_deform_stamp_stretches_device) computes on the GPU."""
    stamps = np.ascontiguousarray(stamps, dtype=np.uint64)
    if stamps.shape[0] == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    b = np.flatnonzero(np.concatenate([[True], stamps[1:] != stamps[:-1]]))
    lens = np.diff(np.concatenate([b, [stamps.shape[0]]]))
    return stamps[b].copy(), lens.astype(np.uint64)


def merge_stretch_lists(lists):
    """Concatenates per-rank (stamps, lens) in rank order, fusing a stretch that continues across
    a slab boundary (the same stamp at the end of rank r and the start of rank r + 1), so that
    the result equals the stretch list of the unsharded stamp stream."""
    ss = np.concatenate([np.asarray(a, dtype=np.uint64) for a, _ in lists]) if lists else np.zeros(0, np.uint64)
    sl = np.concatenate([np.asarray(b, dtype=np.uint64) for _, b in lists]) if lists else np.zeros(0, np.uint64)
    keep = sl > 0
    ss, sl = ss[keep], sl[keep]
    if ss.shape[0] == 0:
        return ss, sl
    head = np.concatenate([[True], ss[1:] != ss[:-1]])
    starts = np.flatnonzero(head)
    return ss[starts].copy(), np.add.reduceat(sl, starts).astype(np.uint64)


class StretchExchange:
    """all_gather of variable-length stretch lists through one fixed-size int64 tensor
    (count, stamps[cap], lens[cap]) per rank; works on any torch.distributed backend.  The
    result is cached: when no rank's local list changed since the last call (one 1-element
    all_reduce), the cached global list is returned without moving anything."""

    def __init__(self, world: int, device, cap: int = 1 << 13):
        import torch
        self.world = world
        self.cap = cap
        self.device = device
        self._torch = torch
        self._buf = None
        self._last_local = None
        self._last_global = None
        self._flag = torch.zeros(1, dtype=torch.int64, device=device) if world > 1 else None

    def __call__(self, ss: np.ndarray, sl: np.ndarray):
        if self.world == 1:
            return ss, sl
        import torch.distributed as dist
        torch = self._torch
        n = ss.shape[0]
        same = (self._last_local is not None and np.array_equal(self._last_local[0], ss)
                and np.array_equal(self._last_local[1], sl))
        # value = 0 if unchanged, else the capacity this rank needs
        self._flag[0] = 0 if same else max(n, 1)
        dist.all_reduce(self._flag, op=dist.ReduceOp.MAX)
        need = int(self._flag.item())
        if need == 0:
            return self._last_global
        while self.cap < need:
            self.cap *= 2
            self._buf = None
        if self._buf is None:
            self._buf = torch.zeros(self.world * (2 * self.cap + 1), dtype=torch.int64,
                                    device=self.device)
        mine = torch.zeros(2 * self.cap + 1, dtype=torch.int64)
        mine[0] = n
        mine[1:1 + n] = torch.from_numpy(ss.view(np.int64))
        mine[1 + self.cap:1 + self.cap + n] = torch.from_numpy(sl.view(np.int64))
        dist.all_gather_into_tensor(self._buf, mine.to(self.device))
        g = self._buf.cpu().numpy().reshape(self.world, 2 * self.cap + 1)
        lists = []
        for r in range(self.world):
            m = int(g[r, 0])
            lists.append((g[r, 1:1 + m].view(np.uint64),
                          g[r, 1 + self.cap:1 + self.cap + m].view(np.uint64)))
        self._last_local = (ss.copy(), sl.copy())
        self._last_global = merge_stretch_lists(lists)
        return self._last_global


def plan_segments(counts, world: int):
    """configs[3] / SURVEY 8e: a multi-robot mesh is one independent deformation problem per
    robot prefix.  The robots' vertices are laid end to end (in the order given) and cut into
    `world` contiguous slabs of equal size; a slab that straddles a robot boundary becomes two
    segments.  Returns, per rank, a list of (robot, first, count) — `first` local to the robot.
    Every vertex of every robot is owned by exactly one rank."""
    counts = [int(c) for c in counts]
    total = sum(counts)
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    plan = []
    for rank in range(world):
        lo, n = slab_bounds(total, world, rank)
        hi = lo + n
        segs = []
        for robot, c in enumerate(counts):
            a, b = max(lo, int(starts[robot])), min(hi, int(starts[robot]) + c)
            if b > a:
                segs.append((robot, a - int(starts[robot]), b - a))
        plan.append(segs)
    return plan


def robot_readers(plan, n_robots: int):
    """Per robot, the ranks that own a segment of it, in rank order: the readers of that robot's
    pose source (a rank's reader id = its position in the list, pgmo_deform_pose_source_publish
    waits for that many acknowledgements)."""
    return {r: [rank for rank, segs in enumerate(plan) if any(seg[0] == r for seg in segs)]
            for r in range(n_robots)}


class MultiRobotExchange:
    """One StretchExchange per robot: ranks that own no vertex of a robot contribute an empty
    list (the collectives stay world-wide, a few KB each, and only run when a list changed)."""

    def __init__(self, n_robots: int, world: int, device):
        self._ex = [StretchExchange(world, device) for _ in range(n_robots)]

    def __call__(self, robot: int, ss: np.ndarray, sl: np.ndarray):
        return self._ex[robot](np.ascontiguousarray(ss, dtype=np.uint64),
                               np.ascontiguousarray(sl, dtype=np.uint64))


class PoseSource:
    """Replication of a solve result by peer pull (pgmo_deform_pose_source_*): the owner rank
    writes the new poses once into its device's source table (two buffers, used alternately);
    after one stream-ordered synchronisation (any collective the caller already runs, e.g. a
    1-element all_reduce) every rank's table-build kernel reads the rows its slab references
    straight from the owner's memory over NVLink (CUDA IPC mapping).  `handle` is this rank's
    DeformHandle for the control table in question."""

    def __init__(self, handle, n_c: int, rank: int, world: int, device, owner: int = 0):
        import torch
        self.h, self.n_c, self.rank, self.world, self.owner = handle, n_c, rank, world, owner
        self.step = 0
        ipc = torch.zeros(64, dtype=torch.uint8, device=device)
        if rank == owner:
            handle_bytes, _ = handle.pose_source_create(n_c)
            ipc.copy_(torch.from_numpy(handle_bytes))
        if world > 1:
            import torch.distributed as dist
            dist.broadcast(ipc, src=owner)
            if rank != owner:
                handle.pose_source_attach(n_c, ipc_handle=ipc.cpu().numpy())

    def write(self, poses, stream=0):
        """owner: poses (device tensor, or pinned host tensor / array) -> the buffer of this step"""
        if self.rank == self.owner:
            self.h.pose_source_write(self.step % 2, poses, stream)

    def pull(self, row_begin: int, row_end: int, stream=0):
        """every rank, after the synchronisation that follows write()"""
        self.h.update_poses_from_source(self.step % 2, row_begin, row_end, stream)
        self.step += 1

    # The same exchange without a collective: sequence numbers in the source's own memory order
    # the owner's write and everybody's pull (pgmo_deform_pose_source_publish /
    # pgmo_deform_update_poses_from_source_seq); `seq` counts solves from 1.
    def publish_and_pull(self, poses, row_begin: int, row_end: int, stream=0):
        self.seq = getattr(self, "seq", 0) + 1
        if self.rank == self.owner:
            self.h.pose_source_publish(self.seq, self.world, poses, stream)
        self.h.update_poses_from_source_seq(self.seq, self.rank, row_begin, row_end, stream)

    def pull_again(self, row_begin: int, row_end: int, stream=0):
        """other rows of the solve published last (e.g. after the slab's control range changed)"""
        self.h.update_poses_from_source_seq(self.seq, self.rank, row_begin, row_end, stream)


def gpu_numa_affinity(device_index: int):
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (so that the
    staging buffers it allocates next are local to the GPU's PCIe root).  Returns the node id, or
    None when the topology cannot be read."""
    import os
    try:
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = device_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if device_index < len(ids) and ids[device_index].strip().isdigit():
                    phys = int(ids[device_index])
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            dom, rest = bus.split(":", 1)
            bdf = f"{int(dom, 16):04x}:{rest.lower()}"
        except Exception:
            import torch
            p = torch.cuda.get_device_properties(device_index)
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def broadcast_controls(tensors, src: int = 0):
    """Replicates the control table / poses (device or CPU tensors) from `src` to every rank."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
