"""Where do the 8 GPUs' D2H copies lose bandwidth?  Prints what the box exposes about NUMA / affinity and measures the
D2H rate of every GPU into pinned buffers first-touched under different CPU affinities, alone and all at once.

    gpurun --gpus 8 -- python tools/numa_probe.py
"""
import glob
import os
import subprocess
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %r" % e


def main():
    print("cpus allowed:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
    print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
    print(sh("nvidia-smi topo -m"))
    for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
        print(p, open(p).read().strip())
    print(sh("cat /proc/self/status | grep -i -E 'cpus_allowed_list|mems_allowed_list'"))
    n = torch.cuda.device_count()
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
        cpus = set()
        for part in open(p).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        nodes[os.path.basename(os.path.dirname(p))] = cpus
    all_cpus = os.sched_getaffinity(0)
    if not nodes:
        half = sorted(all_cpus)
        nodes = {"lowhalf": set(half[:len(half) // 2]), "highhalf": set(half[len(half) // 2:])}
    chunk = 256 << 20
    bufs = {}
    for name, cpus in nodes.items():
        use = cpus & all_cpus
        if not use:
            continue
        try:
            os.sched_setaffinity(0, use)
        except OSError as e:
            print("setaffinity failed", e)
        for g in range(n):
            torch.cuda.set_device(g)
            h = torch.empty(chunk, dtype=torch.uint8, pin_memory=True)
            h.fill_(1)
            bufs[(name, g)] = h
    os.sched_setaffinity(0, all_cpus)
    dsrc = [torch.ones(chunk, dtype=torch.uint8, device="cuda:%d" % g) for g in range(n)]
    streams = [torch.cuda.Stream("cuda:%d" % g) for g in range(n)]

    def run(gs, name, reps=8):
        for g in gs:
            with torch.cuda.stream(streams[g]):
                bufs[(name[g] if isinstance(name, dict) else name, g)].copy_(dsrc[g], non_blocking=True)
        for g in gs:
            streams[g].synchronize()
        t0 = time.perf_counter()
        done = {}
        for _ in range(reps):
            for g in gs:
                with torch.cuda.stream(streams[g]):
                    bufs[(name[g] if isinstance(name, dict) else name, g)].copy_(dsrc[g], non_blocking=True)
        for g in gs:
            streams[g].synchronize()
            done[g] = reps * chunk / (time.perf_counter() - t0) / 1e9
        return done

    names = [k for k in nodes if (k, 0) in bufs]
    for name in names:
        print("alone, buffers touched on", name, {g: round(run([g], name)[g], 1) for g in range(n)})
    for name in names:
        r = run(list(range(n)), name)
        print("all at once, buffers touched on", name, {g: round(v, 1) for g, v in r.items()}, "sum", round(sum(r.values()), 1))
    if len(names) >= 2 and n >= 2:
        for flip in (False, True):
            m = {g: names[(g * 2 // n) ^ flip] if len(names) == 2 else names[g % len(names)] for g in range(n)}
            r = run(list(range(n)), m)
            print("all at once, split", m, {g: round(v, 1) for g, v in r.items()}, "sum", round(sum(r.values()), 1))
    for k in (2, 4):
        if n >= k:
            r = run(list(range(k)), names[0])
            print("first %d at once" % k, {g: round(v, 1) for g, v in r.items()})
            r = run(list(range(n - k, n)), names[0])
            print("last %d at once" % k, {g: round(v, 1) for g, v in r.items()})


if __name__ == "__main__":
    main()
