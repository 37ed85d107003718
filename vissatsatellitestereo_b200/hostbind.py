"""Host-side placement for the host-buffer entry points: run the calling process on the CPUs of the NUMA node its GPU
hangs off, so that pinned staging buffers (first touched after the call) and the PCIe copies stay on that socket.
Matters when several ranks stream depth maps and DSMs over PCIe at once (one process per GPU).  Best effort: any
missing piece of information (no NVML, no sysfs NUMA data, a container that hides it) leaves the affinity unchanged."""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_bus_id(device_index):
    """'dddd:bb:dd.f' of CUDA device `device_index` in sysfs spelling, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        if all(hasattr(p, a) for a in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            return '{:04x}:{:02x}:{:02x}.0'.format(int(p.pci_domain_id), int(p.pci_bus_id), int(p.pci_device_id))
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = device_index
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        if visible:
            ids = [v.strip() for v in visible.split(',') if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else str(bus)).lower()
        return bus[4:] if len(bus.split(':')[0]) == 8 else bus     # NVML prints an 8-digit PCI domain, sysfs uses 4
    except Exception:
        return None


def gpu_numa_node(device_index, sysfs='/sys'):
    """NUMA node of CUDA device `device_index` (-1 if unknown)."""
    bus = gpu_pci_bus_id(device_index)
    if bus is None:
        return -1
    try:
        with open(os.path.join(sysfs, 'bus/pci/devices', bus, 'numa_node')) as fp:
            return int(fp.read().strip())
    except Exception:
        return -1


def bind_to_gpu_numa_node(device_index, sysfs='/sys'):
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns the node (or -1 if nothing was changed)."""
    node = gpu_numa_node(device_index, sysfs)
    if node < 0 or not hasattr(os, 'sched_setaffinity'):
        return -1
    try:
        with open(os.path.join(sysfs, 'devices/system/node/node{}/cpulist'.format(node))) as fp:
            cpus = _parse_cpulist(fp.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return -1 if not cpus else node
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return -1
