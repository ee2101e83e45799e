"""Host staging memory for the host-buffer entry points (strotss_eval_host*): pinned buffers placed on the NUMA node the
GPU hangs off.

With one process per GPU on a two-socket host, pinned pages allocated by a thread running on the far socket make every
host<->device copy cross the inter-socket link.  `numa_local(gpu)` binds the calling thread to the CPUs next to the GPU
(sysfs: /sys/bus/pci/devices/<bus id>/local_cpulist, intersected with the CPUs this process may use) and prefers that
node's memory (set_mempolicy(MPOL_PREFERRED)) while the staging buffers are allocated and first touched; the previous
affinity and policy are restored afterwards.  Everything degrades to a no-op (and says so in the returned record) when
sysfs, the bus id or the syscall is unavailable -- placement is an optimisation, never a requirement.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Dict, Optional

import torch

_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1
_SYS_SET_MEMPOLICY = 238          # x86_64


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


def _bus_id(index: int) -> Optional[str]:
    try:
        p = torch.cuda.get_device_properties(index)
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def gpu_numa(index: int) -> Dict:
    """-> {"bus_id", "numa_node", "local_cpus"} of CUDA device `index` (None entries when unknown)."""
    info = {"bus_id": _bus_id(index), "numa_node": None, "local_cpus": None}
    if info["bus_id"] is None:
        return info
    base = os.path.join("/sys/bus/pci/devices", info["bus_id"])
    try:
        with open(os.path.join(base, "numa_node")) as f:
            info["numa_node"] = int(f.read().strip())
        with open(os.path.join(base, "local_cpulist")) as f:
            info["local_cpus"] = sorted(_parse_cpulist(f.read()))
    except (OSError, ValueError):
        pass
    return info


def _set_mempolicy(mode: int, node: Optional[int]) -> int:
    libc = ctypes.CDLL(None, use_errno=True)
    if node is None or node < 0:
        rc = libc.syscall(_SYS_SET_MEMPOLICY, mode, None, 0)
    else:
        nbits = 1024
        mask = (ctypes.c_ulong * (nbits // (8 * ctypes.sizeof(ctypes.c_ulong))))()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] |= 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        rc = libc.syscall(_SYS_SET_MEMPOLICY, mode, mask, nbits + 1)
    return 0 if rc == 0 else ctypes.get_errno()


@contextlib.contextmanager
def numa_local(index: int, record: Optional[Dict] = None):
    """Run the body on the CPUs / memory node next to CUDA device `index`; `record` (if given) receives what was done."""
    rec = record if record is not None else {}
    info = gpu_numa(index)
    allowed = os.sched_getaffinity(0)
    rec.update({"gpu": index, "bus_id": info["bus_id"], "gpu_numa_node": info["numa_node"], "allowed_cpus": len(allowed),
                "cpus_bound": None, "mempolicy": "unchanged"})
    bound = False
    if info["local_cpus"]:
        near = allowed & set(info["local_cpus"])
        if near:
            try:
                os.sched_setaffinity(0, near)
                bound = True
                rec["cpus_bound"] = f"{min(near)}-{max(near)} ({len(near)} of the {len(info['local_cpus'])} CPUs of the GPU's node)"
            except OSError as e:
                rec["cpus_bound"] = f"sched_setaffinity failed: {e}"
        else:
            rec["cpus_bound"] = "none of the GPU's local CPUs is in this process's cpuset"
    policy_set = False
    if info["numa_node"] is not None and info["numa_node"] >= 0:
        err = _set_mempolicy(_MPOL_PREFERRED, info["numa_node"])
        policy_set = err == 0
        rec["mempolicy"] = f"MPOL_PREFERRED node {info['numa_node']}" if err == 0 else f"set_mempolicy errno {err}"
    try:
        yield rec
    finally:
        if policy_set:
            _set_mempolicy(_MPOL_DEFAULT, None)
        if bound:
            try:
                os.sched_setaffinity(0, allowed)
            except OSError:
                pass


def pinned_empty(shape, index: Optional[int] = None, dtype=torch.float32, record: Optional[Dict] = None) -> torch.Tensor:
    """Pinned host tensor whose pages are allocated and first touched next to CUDA device `index`."""
    if index is None:
        index = torch.cuda.current_device()
    with numa_local(index, record):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.zero_()                    # first touch under the policy (cudaHostAlloc populates the pages, this is belt and braces)
    return t
