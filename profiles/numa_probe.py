#!/usr/bin/env python
"""Host<->device bandwidth with every rank copying at once, before and after binding each rank to its GPU's NUMA node
(torchrun, one rank per GPU):   torchrun --nproc-per-node N profiles/numa_probe.py
Prints one JSON line per rank + the aggregate."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
    import super_resolution_system_b200 as pkg
    p = torch.cuda.get_device_properties(local)
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    info = {"rank": rank, "bdf": bdf, "affinity_before": len(os.sched_getaffinity(0))}
    for k in ("numa_node", "local_cpulist"):
        try:
            info[k] = open(f"/sys/bus/pci/devices/{bdf}/{k}").read().strip()
        except OSError as e:
            info[k] = repr(e)
    try:
        info["nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except OSError:
        info["nodes"] = None
    nbytes = 512 << 20
    d_buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_buf2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(device=dev)

    def measure(tag):
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h.fill_(1)
        h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h2.fill_(2)
        res = {}
        for mode in ("h2d", "d2h", "duplex"):
            for _ in range(2):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(4):
                    if mode in ("h2d", "duplex"):
                        d_buf.copy_(h, non_blocking=True)
                    if mode in ("d2h", "duplex"):
                        with torch.cuda.stream(s2):
                            h2.copy_(d_buf2, non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            res[mode] = round(4 * nbytes * (2 if mode == "duplex" else 1) / dt / 1e9, 1)
        info[tag] = res

    measure("all_ranks_default")
    info["bind"] = pkg.distributed.bind_host_to_device_node(local)
    info["affinity_after"] = len(os.sched_getaffinity(0))
    measure("all_ranks_bound")
    # one rank at a time (bound): the per-GPU link by itself
    for r in range(world):
        dist.barrier()
        if r == rank:
            h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            h.fill_(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                d_buf.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            info["alone_h2d"] = round(4 * nbytes / (time.perf_counter() - t0) / 1e9, 1)
            del h
        dist.barrier()
    every = [None] * world
    dist.all_gather_object(every, info)
    if rank == 0:
        for e in every:
            print(json.dumps(e))
        agg = {t: {m: round(sum(e[t][m] for e in every), 1) for m in ("h2d", "d2h", "duplex")} for t in ("all_ranks_default", "all_ranks_bound")}
        print(json.dumps({"world": world, "aggregate_GBps": agg}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
