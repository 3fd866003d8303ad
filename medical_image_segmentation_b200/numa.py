"""Host-side placement for the H2D leg of the hot path.

The end-to-end step is bound by the PCIe copy of the raw slices (512 KB each).  On a two-socket 8-GPU box a rank whose
pinned staging buffers live on the other socket's memory copies across the inter-socket link: round 1 measured
52 GB/s per GPU alone but 23 GB/s per GPU with 4-8 ranks.  ``bind_to_gpu_numa_node`` pins the calling process to the
CPUs of the NUMA node its GPU hangs off BEFORE the pinned buffers are allocated (first-touch then places them there).
Pure host plumbing (sysfs + sched_setaffinity); it does nothing, and says so, where the topology is not exposed.
"""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> int | None:
    """NUMA node of a CUDA device from sysfs, or None when unknown (-1, containers without sysfs, ...)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Restrict the calling process to the CPUs of the GPU's NUMA node.  Returns what was done (for logs / bench JSON)."""
    info = {"numa_node": None, "bound": False, "cpus": None}
    node = gpu_numa_node(device_index)
    info["numa_node"] = node
    if node is None or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
            info["cpus"] = len(cpus)
    except Exception as e:          # never fail a training job over placement
        info["error"] = f"{type(e).__name__}: {e}"
    return info
