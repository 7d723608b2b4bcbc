#!/usr/bin/env python3
"""Host <-> device copy floor of the end-to-end path when N GPUs copy AT THE SAME TIME (one process per GPU, like bench.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py

Every rank moves what one C2 step moves (24.9 MB down, 6.2 MB up, pinned memory) K times between barriers; the table gives the
per-rank and the aggregate rate for: downloads only, uploads only, both directions at once -- with the default CPU affinity
and again with each rank (and its pinned buffers, first-touched after the move) bound to the NUMA node of its GPU, when the
box exposes one.  The aggregate download rate x 1 / 24.9 MB is the ceiling of the multi-GPU end-to-end frame rate."""
import os
import re
import time

import torch
import torch.distributed as dist

D2H, H2D, K = 24883200, 6220800, 40


def numa_of_gpu(local):
    """(PCI address, NUMA node or None) of CUDA device `local` (honours CUDA_VISIBLE_DEVICES through torch)"""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:                                               # noqa: BLE001
        return None, None
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception:                                               # noqa: BLE001
        return bdf, None
    return bdf, node


def cpus_of_node(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    except Exception:                                               # noqa: BLE001
        return None
    cpus = set()
    for part in txt.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def measure(tag, rank, world, local):
    h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory(); d_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")
    h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory(); d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
    h_out.fill_(1); h_in.fill_(2)                                   # first touch on this rank's (possibly re-bound) cores
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    rows = []
    for name, down, up in (("D2H 24.9 MB", True, False), ("H2D 6.2 MB", False, True), ("both", True, True)):
        for _ in range(3):
            if down:
                with torch.cuda.stream(s1): h_out.copy_(d_out, non_blocking=True)
            if up:
                with torch.cuda.stream(s2): d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            if down:
                with torch.cuda.stream(s1): h_out.copy_(d_out, non_blocking=True)
            if up:
                with torch.cuda.stream(s2): d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        nbytes = (D2H if down else 0) + (H2D if up else 0)
        rows.append((name, float(tmax.item()), float(tsum.item()) / world, nbytes))
    if rank == 0:
        for name, tmax, tavg, nbytes in rows:
            print("%-28s N=%d  %-12s  slowest rank %.3f ms (mean %.3f)  per-rank %.1f GB/s  aggregate %.1f GB/s%s"
                  % (tag, world, name, tmax * 1e3, tavg * 1e3, nbytes / tmax / 1e9, nbytes * world / tmax / 1e9,
                     ("  -> e2e ceiling %.1f k output MP/s" % (8.2944 * world / tmax / 1e3) if name == "both" else "")), flush=True)


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bdf, node = numa_of_gpu(local)
    info = "rank %d gpu %s numa %s affinity %d cpus" % (rank, bdf, node, len(os.sched_getaffinity(0)))
    gathered = [None] * world
    dist.all_gather_object(gathered, info)
    if rank == 0:
        print("host: %d cpus online; %s" % (os.cpu_count(), "; ".join(gathered)), flush=True)
        try:
            nodes = sorted(int(re.findall(r"\d+", d)[0]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
            print("NUMA nodes:", {n: len(cpus_of_node(n) or ()) for n in nodes}, flush=True)
        except Exception as e:                                      # noqa: BLE001
            print("NUMA nodes: unavailable (%r)" % e, flush=True)
    measure("default affinity", rank, world, local)
    cpus = cpus_of_node(node) if node is not None and node >= 0 else None
    allowed = os.sched_getaffinity(0)
    if cpus and (cpus & allowed):
        os.sched_setaffinity(0, cpus & allowed)
        measure("bound to the GPU's NUMA node", rank, world, local)
    elif rank == 0:
        print("no per-GPU NUMA node exposed (numa_node = %s): the binding variant is not applicable on this box" % node, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
