#!/usr/bin/env python
"""What bounds the end-to-end (host-buffer) figure when all GPUs of the box copy at once?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank copies a 256 MB pinned buffer H2D, D2H and both directions at once (two streams), first ALONE (the
other ranks wait at a barrier) and then ALL RANKS TOGETHER; rank 0 prints one JSON line with the per-rank and the
aggregate GB/s of each case plus the host's topology (CPUs visible, NUMA nodes, which NUMA node each GPU hangs
off).  If the aggregate of "together" stops growing with N while "alone" holds, the limiter is on the host side
(memory bandwidth / root complex), not in the engine's pipeline."""
import json
import os
import time

import torch
import torch.distributed as dist


def bw(fn, nbytes, reps=8):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nbytes = 256 << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = {}
    # alone: one rank at a time
    for r in range(world):
        barrier()
        if r == rank:
            res["alone"] = {"h2d": bw(h2d, nbytes), "d2h": bw(d2h, nbytes), "both": bw(both, 2 * nbytes)}
    # together
    tog = {}
    for name, fn, nb in (("h2d", h2d, nbytes), ("d2h", d2h, nbytes), ("both", both, 2 * nbytes)):
        barrier()
        tog[name] = bw(fn, nb)
    res["together"] = tog
    allres = [None] * world
    if world > 1:
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        topo = {"cpus_visible": len(os.sched_getaffinity(0)), "numa_nodes": [], "gpu_numa": []}
        try:
            topo["numa_nodes"] = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
        except Exception:
            pass
        try:
            import pynvml
            pynvml.nvmlInit()
            for i in range(world):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                bus = pynvml.nvmlDeviceGetPciInfo(h).busId
                bus = bus.decode() if isinstance(bus, bytes) else bus
                path = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
                topo["gpu_numa"].append(open(path).read().strip() if os.path.exists(path) else "?")
        except Exception:
            pass
        out = {"n_gpus": world, "bytes": nbytes, "topology": topo,
               "alone_per_rank": [r["alone"] for r in allres],
               "together_per_rank": [r["together"] for r in allres],
               "together_aggregate": {k: sum(r["together"][k] for r in allres) for k in ("h2d", "d2h", "both")}}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
