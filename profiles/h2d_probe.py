#!/usr/bin/env python
"""profiles/h2d_probe.py -- platform probe behind DESIGN.md section 7: pinned host <-> device copy bandwidth per GPU when
1, 2, 4, 8 ranks copy at once, with and without NUMA-local pinned memory.  No codec involved.

    torchrun --nproc-per-node N profiles/h2d_probe.py [--mb 1024]

Prints one JSON line on rank 0: per-rank GB/s for H2D and D2H, all ranks copying simultaneously, (a) pinned memory
allocated wherever the process happens to run, (b) after binding the process to the cores of the GPU's own NUMA node
(first-touch then places the pinned pages there) -- plus the same with only rank 0 copying (the uncontended figure).
"""
import argparse
import json
import os
import re
import subprocess

import torch
import torch.distributed as dist


def gpu_numa(local):
    """NUMA node and CPU list of GPU `local` (sysfs via nvidia-smi's PCI bus id)"""
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip()
        dev = "/sys/bus/pci/devices/" + bus.lower().replace("00000000:", "0000:")
        node = int(open(dev + "/numa_node").read())
        cpus = open(dev + "/local_cpulist").read().strip()
        return node, cpus
    except Exception as e:  # noqa: BLE001
        return None, str(e)


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def measure(host, devbuf, reps, active):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    dist.barrier()
    if not active:
        dist.barrier()
        return None, None
    ev[0].record()
    for _ in range(reps):
        devbuf.copy_(host, non_blocking=True)
    ev[1].record()
    for _ in range(reps):
        host.copy_(devbuf, non_blocking=True)
    ev[2].record()
    torch.cuda.synchronize()
    dist.barrier()
    gb = host.numel() * reps / 1e9
    return gb / (ev[0].elapsed_time(ev[1]) / 1e3), gb / (ev[1].elapsed_time(ev[2]) / 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = args.mb << 20
    devbuf = torch.empty(n, dtype=torch.uint8, device="cuda")
    node, cpus = gpu_numa(local)
    aff0 = sorted(os.sched_getaffinity(0))
    res = {"rank": rank, "gpu_numa_node": node, "gpu_local_cpus": cpus, "affinity_before": "%d cpus, %d..%d" % (len(aff0), aff0[0], aff0[-1])}
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    res["default_all"] = measure(host, devbuf, args.reps, True)
    res["default_alone"] = measure(host, devbuf, args.reps, rank == 0)
    del host
    bound = False
    if node is not None and node >= 0:
        try:
            want = set(parse_cpulist(cpus)) & set(aff0)
            if want:
                os.sched_setaffinity(0, want)
                bound = True
        except Exception:  # noqa: BLE001
            pass
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)   # allocated and first touched on the GPU's own node
    host.fill_(1)
    res["numa_bound"] = bound
    res["numa_all"] = measure(host, devbuf, args.reps, True)
    res["numa_alone"] = measure(host, devbuf, args.reps, rank == 0)
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        topo = ""
        try:
            topo = subprocess.check_output(["nvidia-smi", "topo", "-m"], text=True)
        except Exception:  # noqa: BLE001
            pass
        rnd = lambda v: None if v is None or v[0] is None else [round(v[0], 1), round(v[1], 1)]
        print(json.dumps({"what": "pinned host<->device copies, %d MiB x %d, [H2D GB/s, D2H GB/s] per rank" % (args.mb, args.reps),
                          "n_gpus": world,
                          "ranks": [{k: (rnd(v) if k.endswith(("_all", "_alone")) else v) for k, v in r.items()} for r in allres],
                          "topo": re.sub(r"\x1b\[[0-9;]*m", "", topo)[:6000]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
