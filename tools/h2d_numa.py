#!/usr/bin/env python
"""Host-to-device ceiling of the box with and without NUMA-local pinned buffers.  Run under torchrun (one rank per
GPU): every rank times a cudaMemcpyAsync of a 1 GiB pinned buffer, all ranks at once, first with the buffer wherever
the allocator put it, then with the process bound to the CPUs of its GPU's NUMA node before the buffer is allocated
and first touched.  Prints one JSON line per rank."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def gpu_numa_node(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bdf, bytes):
            bdf = bdf.decode()
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        return node, bdf
    except Exception as e:  # noqa
        return -1, str(e)


def node_cpus(node):
    txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    cpus = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def timed_copy(dev, xh, xd, reps, world):
    st = torch.cuda.current_stream()
    for _ in range(2):
        xd.copy_(xh, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        xd.copy_(xh, non_blocking=True)
    b.record(st)
    torch.cuda.synchronize()
    return xh.numel() * xh.element_size() * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 27                                               # complex64: 1 GiB
    xd = torch.empty(n, dtype=torch.complex64, device=dev)
    out = {"rank": rank, "aff_before": len(os.sched_getaffinity(0)), "nodes": sorted(os.listdir("/sys/devices/system/node"))[:6]}
    xh = torch.empty(n, dtype=torch.complex64, pin_memory=True)
    xh.view(torch.float32).fill_(1.0)
    out["default_gbs"] = timed_copy(dev, xh, xd, 8, world)
    del xh
    node, bdf = gpu_numa_node(local)
    out["numa_node"], out["bdf"] = node, bdf
    if node >= 0:
        cpus = node_cpus(node)
        os.sched_setaffinity(0, cpus)
        out["aff_after"] = len(cpus)
    xh = torch.empty(n, dtype=torch.complex64, pin_memory=True)   # allocated and first touched from the local node
    xh.view(torch.float32).fill_(1.0)
    out["local_gbs"] = timed_copy(dev, xh, xd, 8, world)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
