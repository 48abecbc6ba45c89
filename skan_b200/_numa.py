"""Host-side placement for the end-to-end path: run the calling thread on the CPUs next to a GPU, so that the pinned
host buffers it allocates afterwards (first touch) live on that GPU's NUMA node and host->device copies of all ranks
of a node do not share one socket's memory controllers and inter-socket link.

The reference has no counterpart (it is a single-process CPU library); this belongs to the multi-GPU plumbing of
SURVEY.md section 8(e)."""
from __future__ import annotations

import os


def bind_thread_to_gpu(device) -> dict:
    """Set the CPU affinity of the calling thread to the CPUs NVML reports as local to `device` (a torch.device or an
    index into the visible CUDA devices).  Returns {'cpus': n, 'numa_node': k or None, 'bound': bool}; never raises:
    where NVML, the sysfs entry or the permission is missing the thread stays where it is and 'bound' is False."""
    info = {'bound': False, 'cpus': None, 'numa_node': None}
    try:
        import torch
        import pynvml
        idx = device.index if hasattr(device, 'index') and device.index is not None else int(device)
        props = torch.cuda.get_device_properties(idx)
        pynvml.nvmlInit()
        uuid = 'GPU-' + str(props.uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = None
        for cand in (bus.lower(), bus.lower()[4:] if len(bus) > 12 else bus.lower()):
            path = f'/sys/bus/pci/devices/{cand}/numa_node'
            if os.path.exists(path):
                node = int(open(path).read().strip())
                break
        info['numa_node'] = node if node is not None and node >= 0 else None
        cpus = None
        if info['numa_node'] is not None:
            path = f'/sys/devices/system/node/node{node}/cpulist'
            if os.path.exists(path):
                cpus = set()
                for part in open(path).read().strip().split(','):
                    a, _, b = part.partition('-')
                    cpus.update(range(int(a), int(b or a) + 1))
        if not cpus:  # NVML's own idea of the GPU's CPUs
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = set(cpus) & set(allowed)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info['bound'] = True
            info['cpus'] = len(cpus)
    except Exception as e:  # placement is an optimisation, never a requirement
        info['error'] = f'{type(e).__name__}: {e}'[:120]
    return info
