"""Host-side placement for the host-buffer pipelines: run the rank's threads on -- and therefore allocate its pinned
staging buffers from -- the NUMA node its GPU hangs off.

One process per GPU is the deployment model.  With N ranks started by torchrun every process inherits
the same CPU mask; pinned buffers (cudaHostAlloc follows the first-touch / local allocation policy of the calling
thread) then pile up on whatever node the launcher ran on, and half of the GPUs pull their frames across the
socket interconnect.  `bind_to_device` narrows the process to the CPUs sysfs lists as local to the GPU's PCI device
before any buffer is allocated.  Nothing here touches the data path.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch


def _pci_sysfs_dir(index: int) -> Optional[str]:
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    d = os.path.join("/sys/bus/pci/devices", bdf)
    return d if os.path.isdir(d) else None


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def device_locality(index: int) -> Dict[str, object]:
    """NUMA node and local CPUs of cuda:<index> as sysfs reports them (node -1 / empty set when unknown)."""
    d = _pci_sysfs_dir(index)
    node, cpus = -1, set()
    if d:
        try:
            node = int(open(os.path.join(d, "numa_node")).read().strip())
        except Exception:
            node = -1
        try:
            cpus = _parse_cpulist(open(os.path.join(d, "local_cpulist")).read())
        except Exception:
            cpus = set()
    return {"node": node, "cpus": cpus, "sysfs": d}


def bind_to_device(index: int) -> Dict[str, object]:
    """Restrict this process to the CPUs local to cuda:<index>.  Call before allocating pinned buffers.
    Returns what was done (for logs / bench records); never raises on hosts without the sysfs information."""
    loc = device_locality(index)
    allowed = os.sched_getaffinity(0)
    target = loc["cpus"] & allowed if loc["cpus"] else set()
    done = False
    if target and target != allowed:
        try:
            os.sched_setaffinity(0, target)
            done = True
        except OSError:
            done = False
    return {"numa_node": loc["node"], "local_cpus": len(loc["cpus"]), "bound": done,
            "cpus_now": len(os.sched_getaffinity(0))}
