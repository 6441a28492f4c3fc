"""Keep a rank's host threads -- and so, by first touch, its pinned staging buffers -- on the NUMA node of its GPU.

The end-to-end path is a host-to-device stream at PCIe speed per GPU.  With one process per GPU on a two-socket box, a
rank whose pinned buffers live on the other socket pulls its whole GAF across the socket interconnect.  Linux places
pages on the node of the thread that first touches them (`cudaHostAlloc` touches them while pinning), so binding the
process to the CPUs next to its GPU before the first allocation is all that is needed.  Best effort: a box that does
not expose NUMA topology (a single node, a VM) is left alone.
"""
from __future__ import annotations

import os


def _read(path: str):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text: str) -> set[int]:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11}"""
    cpus: set[int] = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device: int, sysfs: str = "/sys") -> int | None:
    """NUMA node of CUDA device `device` (None: unknown / not exposed)."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device)
        bus_id = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None
    text = _read(os.path.join(sysfs, "bus", "pci", "devices", bus_id, "numa_node"))
    if text is None:
        return None
    try:
        node = int(text)
    except ValueError:
        return None
    return node if node >= 0 else None


def bind_to_gpu_node(device: int, sysfs: str = "/sys", node: int | None = None) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (`node`: skip the look-up, tests).  -> what was done, for
    the logs / the bench line."""
    info = {"device": int(device), "node": None, "cpus": None, "bound": False}
    nodes = _read(os.path.join(sysfs, "devices", "system", "node", "online"))
    if nodes is None or len(parse_cpulist(nodes)) < 2:
        info["why"] = "one NUMA node"
        return info
    if node is None:
        node = gpu_numa_node(device, sysfs)
    if node is None:
        info["why"] = "GPU's node not exposed"
        return info
    cpus = parse_cpulist(_read(os.path.join(sysfs, "devices", "system", "node", f"node{node}", "cpulist")))
    try:
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if not want:
            info["why"] = "no allowed CPU on the GPU's node"
            return info
        os.sched_setaffinity(0, want)
    except (AttributeError, OSError) as e:
        info["why"] = f"sched_setaffinity: {e}"
        return info
    info.update(node=node, cpus=len(want), bound=True)
    return info
