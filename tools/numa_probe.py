"""Where should a rank's pinned host buffer live?  sysfs gives no NUMA node for the GPUs inside the container, so measure:
per rank (one at a time) the H2D rate from a pinned buffer allocated while bound to each NUMA node's CPUs; then ALL ranks at
once, default placement vs the best node of each rank.   torchrun --nproc-per-node N tools/numa_probe.py"""
import glob
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
all_cpus = os.sched_getaffinity(0)
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    cpus = set()
    for part in open(p + "/cpulist").read().strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    if cpus & all_cpus:
        nodes[int(p.rsplit("node", 1)[1])] = cpus & all_cpus


def barrier():
    if world > 1:
        dist.barrier()


def h2d_rate(buf, dst, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dst.copy_(buf, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return buf.numel() * buf.element_size() / best / 1e9


SZ = 256 * 1024 * 1024
dst = torch.empty(SZ, dtype=torch.uint8, device=dev)
rates = {}
for r in range(world):
    if r == rank:
        for node, cpus in nodes.items():
            os.sched_setaffinity(0, cpus)
            b = torch.empty(SZ, dtype=torch.uint8).pin_memory()
            b.fill_(1)
            rates[node] = h2d_rate(b, dst)
            del b
        os.sched_setaffinity(0, all_cpus)
    barrier()
best = max(rates, key=rates.get) if rates else None
print(f"rank {rank}: cpus allowed {len(all_cpus)}, nodes {sorted(nodes)}, solo H2D GB/s per node {dict((k, round(v, 1)) for k, v in rates.items())}, best {best}", flush=True)
BIG = 1024 ** 3
dstb = torch.empty(BIG, dtype=torch.uint8, device=dev)
bufs = {}
bufs["default"] = torch.empty(BIG, dtype=torch.uint8).pin_memory(); bufs["default"].fill_(1)
if best is not None:
    os.sched_setaffinity(0, nodes[best])
    bufs["best node"] = torch.empty(BIG, dtype=torch.uint8).pin_memory(); bufs["best node"].fill_(1)
    os.sched_setaffinity(0, all_cpus)
for name, b in bufs.items():
    barrier()
    rate = h2d_rate(b, dstb, reps=4)
    t = torch.tensor([rate], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    if rank == 0:
        print(f"all {world} ranks at once, {name}: {float(t.item()):.1f} GB/s in total ({float(t.item()) / world:.1f} per GPU)", flush=True)
    barrier()
if world > 1:
    dist.destroy_process_group()
