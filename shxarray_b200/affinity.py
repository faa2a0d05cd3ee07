"""CPU / memory locality for multi-GPU hosts: keep a rank's host threads (and therefore the page-locked buffers they first
touch) on the NUMA node its GPU hangs off.

Why: the host entry points (``shb200_*_host``) move 1.67 GB per cfg3 step over PCIe through page-locked bounce and result
buffers.  With eight ranks on a two-socket box, a rank whose threads run on the other socket sends every byte across the
inter-socket link twice (r01 SCALE: e2e efficiency 0.25 at 8 GPUs, 132 GB/s aggregate).  Binding must happen BEFORE the
first host call: the library's copy threads inherit the caller's affinity and ``cudaHostAlloc`` pages land where the calling
thread runs.

Nothing here touches the device path; without sysfs / NVML information the call is a no-op and says so.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_local_cpus(device=0):
    """CPUs local to CUDA device ``device`` (sysfs ``local_cpulist`` of its PCI function), or None when unknown."""
    bus_id = None
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        if hasattr(p, "pci_bus_id") and hasattr(p, "pci_device_id"):
            bus_id = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        bus_id = None
    if bus_id is None:
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
            bus_id = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            if isinstance(bus_id, bytes):
                bus_id = bus_id.decode()
            bus_id = bus_id.lower()[-12:]
        except Exception:
            return None
    for cand in (bus_id, bus_id.lower()):
        path = f"/sys/bus/pci/devices/{cand}/local_cpulist"
        try:
            cpus = _parse_cpulist(open(path).read())
            if cpus:
                return cpus
        except OSError:
            continue
    return None


def bind_to_device_numa(device=0):
    """Restrict this process to the CPUs local to ``device`` (intersected with the CPUs it may already use).

    Returns a dict describing what happened: {"bound": bool, "cpus": n, "of": n_allowed, "reason": ...}."""
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return {"bound": False, "reason": "no sched_getaffinity on this platform"}
    local = device_local_cpus(device)
    if not local:
        return {"bound": False, "cpus": len(allowed), "of": len(allowed), "reason": "device-local CPU list unavailable"}
    target = allowed & local
    if not target or target == allowed:
        return {"bound": False, "cpus": len(allowed), "of": len(allowed),
                "reason": "all allowed CPUs are already device-local" if target else "no allowed CPU is device-local"}
    os.sched_setaffinity(0, target)
    return {"bound": True, "cpus": len(target), "of": len(allowed), "reason": "sysfs local_cpulist"}
