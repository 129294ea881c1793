"""Multi-GPU plumbing: environments shard by index, one process per GPU, no exchange inside step.

The reference has no distributed path at all (one engine per process, engine/engine_utils.py:8-15); its only
parallelism is N independent env processes.  Here rank r of W owns the contiguous block of environments
[r*E, (r+1)*E) and the only collective is a sum of a handful of episode statistics at logging cadence.
"""
import numpy as np

STAT_NAMES = ["episodes", "return_sum", "length_sum", "cost_sum", "success", "crash", "out_of_road", "max_step"]


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated (first
    touch puts the staging buffers on that node), so that with one rank per GPU the ranks do not all stream their D2H
    copies into one socket's memory.  Reads the GPU's PCI address from nvidia-smi and the node from sysfs; a box with a
    single node (or without the sysfs entries) is left alone.  Returns a small dict for the bench line."""
    import os
    import subprocess
    info = {"bound": False}
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                       # sysfs uses a 4-digit PCI domain
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info.update(gpu_pci=bus, node=node, nodes=len(nodes))
        if node < 0 or len(nodes) < 2:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except Exception as e:  # no nvidia-smi / sysfs: nothing to bind
        info["error"] = type(e).__name__
    return info


def shard_scenarios(n_library, envs_per_rank, rank):
    """Library indices of the scenarios rank `rank` hosts (weak scaling: per-rank work is fixed)."""
    return [(rank * envs_per_rank + e) % n_library for e in range(envs_per_rank)]


def episode_stats(terminated, truncated, info_flags, info_f, cost_total=None):
    """8 floats from one step's outputs (torch tensors on any device): see STAT_NAMES.
    info_flags bits: include/md_layout.h FL_*; info_f columns: sim.INFO_F_NAMES."""
    import torch
    done = (terminated.bool() | truncated.bool())
    f = info_flags
    crash = (f & 0x1f) != 0
    z = lambda m: (m & done).sum().to(torch.float64)
    return torch.stack([
        done.sum().to(torch.float64), (info_f[:, 6] * done).sum().to(torch.float64),
        (info_f[:, 7] * done).sum().to(torch.float64),
        (cost_total * done).sum().to(torch.float64) if cost_total is not None else torch.zeros((), dtype=torch.float64, device=f.device),
        z((f & 0x800) != 0), z(crash), z((f & 0x400) != 0), z((f & 0x1000) != 0)])


def all_reduce_stats(stats):
    """Sum the statistics vector over ranks (NCCL on GPUs, gloo in the CPU tests). No-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
