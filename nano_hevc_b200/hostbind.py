"""Host-side placement for the host-buffer (e2e) path: keep a rank's pinned buffers and widening threads
on the NUMA node its GPU hangs off.

The host pass of ``nh_host_pipeline_dcplanar`` is bound by host DRAM bandwidth; on a two-socket box a rank
whose staging buffers or worker threads sit on the other socket pays the inter-socket link for every byte.
``bind_to_gpu_numa_node`` restricts the calling process (and therefore the threads and first-touch page
placement that follow) to the CPUs of the GPU's node.  It must run BEFORE pinned memory is allocated and
before the library's host thread pool starts.  On single-node boxes and VMs that hide the topology
(``numa_node`` = -1) it changes nothing and says so.
"""
from __future__ import annotations

import os
import subprocess


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index: int) -> int:
    """NUMA node of GPU ``index`` (as numbered by CUDA_VISIBLE_DEVICES order of nvidia-smi), -1 if unknown."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip().lower()
        # nvidia-smi prints an 8-digit domain (00000000:1B:00.0); sysfs uses 4 digits
        dom, rest = out.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return -1


def bind_to_gpu_numa_node(index: int) -> dict:
    """Restrict this process to the CPUs of GPU ``index``'s NUMA node.  Returns what was done."""
    info = {"gpu": index, "numa_node": -1, "bound": False}
    try:
        before = os.sched_getaffinity(0)
    except AttributeError:
        return info
    info["cpus_before"] = len(before)
    node = gpu_numa_node(index)
    info["numa_node"] = node
    if node < 0:
        return info
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & before
    except OSError:
        return info
    if not cpus or cpus == before:
        info["cpus_after"] = len(before)
        return info
    os.sched_setaffinity(0, cpus)
    info.update(bound=True, cpus_after=len(cpus))
    return info
