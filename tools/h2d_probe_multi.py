"""Concurrent H2D ceiling of the box: every rank copies a 3 GB pinned buffer to its own GPU at the same time (torchrun, one
process per GPU) — what the e2e leg of bench.py --gpus N can reach at most per GPU when all ranks stream together.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/h2d_probe_multi.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
N = 3 * 1024 ** 3
host = torch.empty(N, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(N, dtype=torch.uint8, device=f"cuda:{local}")
for concurrent in (False, True):
    best = 0.0
    for rep in range(4):
        for turn in range(world if not concurrent else 1):
            dist.barrier()
            if concurrent or turn == rank:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                dev.copy_(host, non_blocking=True)
                torch.cuda.synchronize()
                best = max(best, N / (time.perf_counter() - t0) / 1e9)
        dist.barrier()
    rates = [None] * world
    dist.all_gather_object(rates, round(best, 1))
    if rank == 0:
        print(f"{'all ranks at once' if concurrent else 'one rank at a time'}: GB/s per GPU {rates}, sum {sum(rates):.0f}", flush=True)
dist.destroy_process_group()
