"""Frame-range sharding of a batch over ranks (one process per GPU).

Frames decode independently (AlacFile.cs:430-435 resets all state per frame), so the
global frame list (track-major) is cut into `world` contiguous ranges balanced by
compressed bytes -- the same plan alacgpu_plan_partition gives a multi-device context --
and each rank stages and decodes only its own range.  No data-path collective: the only
communication is the timing/throughput reduction at the end of a bench step.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .decoder import plan_partition


@dataclass
class TrackSlice:
    track: int          # index into the caller's track list
    frame_lo: int       # frames [frame_lo, frame_hi) of that track
    frame_hi: int
    byte_lo: int        # byte range of those frames inside the track's mdat payload
    byte_hi: int


def rank_slices(stsz_per_track, world: int, rank: int) -> list[TrackSlice]:
    """The pieces of each track that `rank` owns under the byte-balanced contiguous plan."""
    sizes = [np.ascontiguousarray(s, dtype=np.uint32) for s in stsz_per_track]
    flat = np.concatenate(sizes) if sizes else np.zeros(0, dtype=np.uint32)
    cut = plan_partition(flat, world)
    lo, hi = int(cut[rank]), int(cut[rank + 1])
    out, first = [], 0
    for t, s in enumerate(sizes):
        a, b = max(lo, first), min(hi, first + s.size)
        if a < b:
            offs = np.concatenate([[0], np.cumsum(s.astype(np.int64))])
            out.append(TrackSlice(t, a - first, b - first, int(offs[a - first]), int(offs[b - first])))
        first += s.size
    return out


def reduce_step(dist, device, ms_local: float, units_local: float):
    """-> (max over ranks of ms, sum over ranks of units).  `dist` is torch.distributed
    (initialised) or None for a single process."""
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPU cores of the NUMA node the GPU hangs off, BEFORE any pinned
    buffer is allocated (first touch then places the staging buffers next to the GPU's PCIe root).
    With one rank per GPU the end-to-end path is bound by host<->device copies; on a two-socket box
    half of the ranks otherwise push every byte over the socket interconnect.  Best effort: returns
    what it did, never raises."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info = {"numa_node": node, "cpus": len(use)}
    except Exception:
        pass
    return info
