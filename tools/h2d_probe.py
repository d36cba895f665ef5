"""Why does host->device bandwidth per GPU collapse when 8 ranks copy at once (round 1: 55 GB/s alone, 4 GB/s each at
N=8)?  Run under torchrun with N ranks.  Prints the box's topology (GPU PCI address -> NUMA node, CPUs per node) and
the per-rank GB/s of pinned-host -> device copies with k = 1, 2, 4, N ranks active, for two placements of the pinned
buffer: wherever the process happens to run ("default"), and with the process first bound to the CPUs of its GPU's
NUMA node ("local").
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py"""
import glob
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_numa(index):
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
        return bus, int(node)
    except Exception as e:
        return "?", -1


def node_cpus(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    except Exception:
        return []
    cpus = []
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def measure(buf_dev, buf_host, active, reps=4):
    torch.cuda.synchronize()
    dist.barrier()
    gbs = 0.0
    if active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        buf_dev.copy_(buf_host, non_blocking=True)
        torch.cuda.synchronize()
    dist.barrier()
    if active:
        e0.record()
        for _ in range(reps):
            buf_dev.copy_(buf_host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = buf_host.numel() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    dist.barrier()
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, round(gbs, 1))
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bus, node = gpu_numa(local)
    info = [None] * world
    dist.all_gather_object(info, {"gpu": local, "bus": bus, "numa": node, "cpus_allowed": len(os.sched_getaffinity(0))})
    nbytes = 2 << 30
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res = {}
    for placement in ("default", "local"):
        if placement == "local" and node >= 0:
            cpus = node_cpus(node)
            if cpus:
                try:
                    os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
                except OSError:
                    pass
        host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        host.fill_(rank + 1)                       # first touch on the CPUs the process is bound to
        for k in sorted({1, 2, 4, world} & set(range(1, world + 1))):
            res["%s_k%d" % (placement, k)] = measure(dev, host, rank < k)
        del host
    if rank == 0:
        nodes = sorted(int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        print(json.dumps({"ranks": info, "numa_nodes": {n: len(node_cpus(n)) for n in nodes}, "h2d_gbs": res}))
        print(topo)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
