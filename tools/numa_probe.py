#!/usr/bin/env python
"""Where do the D2H landing buffers live?  Prints the GPU <-> NUMA topology of the box and, under torchrun, the
concurrent per-rank D2H / H2D copy rate with the rank's threads (a) left where the launcher put them, (b) bound to
the cores of the GPU's own NUMA node, (c) bound to a remote node.  Pinned pages are placed first-touch, so the
binding that is in force when the buffer is allocated decides which socket's memory the PCIe writes land in.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/numa_probe.py
"""
import glob, os, subprocess, sys, time

import torch
import torch.distributed as dist


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def nodes():
    d = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        d[int(p.rsplit("node", 1)[1])] = cpulist(open(p + "/cpulist").read())
    return d


def gpu_node(i):
    pr = torch.cuda.get_device_properties(i)
    bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    try:
        return bdf, int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except OSError as e:
        return bdf, "?(%s)" % e


def rate(dev, label, world):
    n = 44 * 1024 * 1024
    d = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    h = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for x in h:
        x.zero_()
    s = torch.cuda.Stream()
    res = []
    for direction in ("d2h", "h2d"):
        with torch.cuda.stream(s):
            for i in range(8):
                (h[i % 2].copy_(d[i % 2], non_blocking=True) if direction == "d2h" else d[i % 2].copy_(h[i % 2], non_blocking=True))
            s.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            reps = 96
            for i in range(reps):
                (h[i % 2].copy_(d[i % 2], non_blocking=True) if direction == "d2h" else d[i % 2].copy_(h[i % 2], non_blocking=True))
            s.synchronize()
            dt = time.perf_counter() - t0
        res.append(reps * n / dt / 1e9)
    print("rank %d %-28s d2h %.1f GB/s  h2d %.1f GB/s  (cpus now: %d)" % (
        int(os.environ.get("RANK", 0)), label, res[0], res[1], len(os.sched_getaffinity(0))), flush=True)


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nd = nodes()
    if rank == 0:
        print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "nodes", {k: len(v) for k, v in nd.items()})
        for k, v in nd.items():
            print(" node", k, "cpus", v[:4], "...", v[-2:])
            try:
                print("   ", open("/sys/devices/system/node/node%d/meminfo" % k).read().splitlines()[0])
            except OSError:
                pass
        for i in range(torch.cuda.device_count()):
            print(" gpu", i, gpu_node(i))
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
        except Exception as e:
            print("topo failed", e)
    full = os.sched_getaffinity(0)
    bdf, node = gpu_node(local)
    rate(dev, "launcher affinity", world)
    if isinstance(node, int) and node in nd and len(nd) > 1:
        own = set(nd[node]) & full
        other = set(c for k, v in nd.items() if k != node for c in v) & full
        if own:
            os.sched_setaffinity(0, own)
            rate(dev, "bound to own node %d" % node, world)
        if other:
            os.sched_setaffinity(0, other)
            rate(dev, "bound to remote node", world)
        os.sched_setaffinity(0, full)
    else:
        print("rank", rank, "gpu node", node, "- single node box or unknown, nothing to bind")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
