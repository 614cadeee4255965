"""Host-side placement for the streamed (host -> GPU -> host) path.

A pinned staging buffer is only as fast as the path between its DRAM and the
GPU's PCIe root port: on a two-socket box, pages that sit on the other socket
cross the inter-socket link in both directions and several ranks then share it.
``bind_host_to_device`` pins the calling process to the CPUs that are local to
the GPU (sysfs ``local_cpulist`` of the GPU's PCI function) BEFORE it allocates
pinned memory, so that first-touch places the pages on the GPU's own NUMA node.
It only narrows the affinity the process already has (a container's cpuset is
respected) and never fails: if sysfs / the device properties are unavailable
the process is left as it is and the reason is returned.
"""
import os


def _parse_cpulist(text):
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


def device_pci_address(device):
    import torch

    pr = torch.cuda.get_device_properties(device)
    return "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)


def device_numa_info(device):
    """``{"pci", "numa_node", "local_cpus"}`` of a CUDA device from sysfs (values None when unknown)."""
    info = {"pci": None, "numa_node": None, "local_cpus": None}
    try:
        addr = device_pci_address(device)
        info["pci"] = addr
        base = "/sys/bus/pci/devices/%s/" % addr
        with open(base + "numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "local_cpulist") as f:
            info["local_cpus"] = sorted(_parse_cpulist(f.read()))
    except Exception as exc:   # no sysfs in this container / old torch without pci ids
        info["error"] = repr(exc)[:120]
    return info


def bind_host_to_device(device):
    """Restrict this process to the CPUs local to ``device`` (intersection with the current affinity).
    Returns a dict describing what was done; never raises."""
    info = device_numa_info(device)
    info["bound"] = False
    try:
        allowed = os.sched_getaffinity(0)
        info["allowed_before"] = len(allowed)
        local = set(info["local_cpus"] or ())
        target = allowed & local
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info["bound"] = True
        info["allowed_after"] = len(os.sched_getaffinity(0))
    except Exception as exc:
        info["bind_error"] = repr(exc)[:120]
    info.pop("local_cpus", None)
    return info
