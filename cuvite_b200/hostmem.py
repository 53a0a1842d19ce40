"""Placement of pinned host buffers: on a multi-socket host a PCIe transfer from the far socket's memory
crosses the inter-socket link first.  ``near_gpu(i)`` restricts the calling thread to the CPUs of GPU i's NUMA
node while the body allocates (and first touches) its pinned buffers, then restores the affinity — the
buffers stay where they were placed.  Best effort: every failure (no sysfs entry, container cpuset, no GPU)
leaves the affinity alone."""
from __future__ import annotations

import contextlib
import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(device_index: int) -> set[int] | None:
    """CPUs of the NUMA node the GPU hangs off, or None when the platform does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        name = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{name}/numa_node").read())
        if node < 0:
            return None
        return _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except Exception:
        return None


@contextlib.contextmanager
def near_gpu(device_index: int):
    old = None
    try:
        cpus = gpu_numa_cpus(device_index)
        if cpus and hasattr(os, "sched_getaffinity"):
            cur = os.sched_getaffinity(0)
            want = cur & cpus
            if want and want != cur:
                os.sched_setaffinity(0, want)
                old = cur
    except Exception:
        old = None
    try:
        yield old is not None
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass
