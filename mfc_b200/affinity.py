"""Host-side placement of a rank next to its GPU (one process per GPU).

``bind_to_gpu(i)`` restricts the calling process to the CPU cores NVML reports as local to GPU ``i`` -- the cores of the
NUMA node the GPU's PCIe root port hangs off -- so that the pinned host buffers the rank allocates afterwards (first touch)
and its DMA descriptors live on that node.  The reference gets the same placement from the MPI launcher
(``mpirun --bind-to``/``--map-by`` in its run templates, reference ``toolchain/templates/*.mako``); under ``torchrun`` nothing
binds the ranks, and eight ranks staging their state through one node's memory is what the round-1 end-to-end number at
N = 8 paid for.  No-ops (returns a note) where NVML is missing, the mask is empty, or the machine
exposes a single node anyway.
"""
import os


def gpu_cpu_mask(index):
    """CPU ids local to GPU `index` according to NVML, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        return [c for c in cpus if c < ncpu] or None
    except Exception:
        return None


def numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return [0]


def bind_to_gpu(index, local_world=1):
    """Pin the process to the cores local to GPU `index`; with several ranks on one node each rank takes its own slice of
    them (so the ranks' copy threads do not share cores).  Returns a dict describing what was done."""
    info = {"gpu": index, "numa_nodes": len(numa_nodes()), "bound": False}
    if not hasattr(os, "sched_setaffinity"):
        info["note"] = "no sched_setaffinity"
        return info
    allowed = sorted(os.sched_getaffinity(0))
    mask = gpu_cpu_mask(index)
    if mask:
        mask = [c for c in mask if c in allowed]
    if not mask:
        info["note"] = "NVML reports no CPU affinity for this GPU"
        return info
    if local_world > 1 and len(mask) >= 2 * local_world:
        # ranks whose GPUs share a node split its cores evenly
        per = len(mask) // local_world
        mine = mask[(index % local_world) * per:(index % local_world + 1) * per]
        if mine:
            mask = mine
    try:
        os.sched_setaffinity(0, mask)
        info.update(bound=True, cpus=f"{mask[0]}-{mask[-1]}", ncpus=len(mask))
    except OSError as e:
        info["note"] = f"sched_setaffinity failed: {e}"
    return info
