"""CPU / memory locality of a GPU, for the host-buffer (`e2e`) path with one process per GPU.

With N ranks streaming pinned host buffers to N GPUs at once, a rank whose pinned pages live on the
other socket sends every byte over the inter-socket link.  `bind_to_gpu` pins the calling process to the
CPUs that are local to its GPU (sysfs `local_cpulist` of the PCI device) BEFORE the pinned buffers are
allocated, so first touch places them on the GPU's own NUMA node.  Best effort: on a box that does not
expose the topology (containers often report numa_node = -1) it changes nothing and says so.
"""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def _pci_bus_id(device_index: int) -> str | None:
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        return bus.lower()
    except Exception:
        return None


def gpu_locality(device_index: int) -> dict:
    """{'pci_bus_id', 'numa_node', 'local_cpus'} of a CUDA device (None / -1 / '' where unknown)."""
    bus = _pci_bus_id(device_index)
    out = {"pci_bus_id": bus, "numa_node": -1, "local_cpus": ""}
    if not bus:
        return out
    dom, rest = bus.split(":", 1)
    path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
    try:
        with open(os.path.join(path, "numa_node")) as f:
            out["numa_node"] = int(f.read().strip())
        with open(os.path.join(path, "local_cpulist")) as f:
            out["local_cpus"] = f.read().strip()
    except OSError:
        pass
    return out


def bind_to_gpu(device_index: int) -> bool:
    """Restrict this process to the CPUs local to the GPU (intersection with the current affinity).
    Returns True when the affinity was narrowed; call before allocating pinned host buffers."""
    if os.environ.get("LLFE_NO_NUMA_BIND"):
        return False
    loc = gpu_locality(device_index)
    try:
        local = _parse_cpulist(loc["local_cpus"]) if loc["local_cpus"] else set()
        cur = os.sched_getaffinity(0)
        want = local & cur
        if not want or want == cur:
            return False
        os.sched_setaffinity(0, want)
        return True
    except (OSError, ValueError, AttributeError):
        return False
