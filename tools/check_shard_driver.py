"""torchrun check of the sharded encode driver (sharding.generate_codes_sharded): the N-GPU semantic ids —
Sinkhorn re-encode rounds and global suffix included — equal the single-GPU generate_codes of the concatenated
catalogue, and how long both take.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/check_shard_driver.py [items_total] [c2_slice,c1_slice]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import sharding, synth   # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden                        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    ok_all = True
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c2_slice", "c1_slice"]
    for name in names:
        g, cfg, cbs = load_golden(name)
        model = build_model(cfg, cbs, device=dev)
        lo, hi = sharding.shard_range(n_total, rank, world)
        x = torch.from_numpy(synth.synth_items(2024, lo, hi - lo, cfg["in_dim"], 1_000_000)).to(dev)
        for rep in range(2):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            mine, stats = sharding.generate_codes_sharded(model, x, dist.group.WORLD)
            torch.cuda.synchronize(); dist.barrier()
            t_sharded = time.perf_counter() - t0
        sizes = [sharding.shard_range(n_total, r, world) for r in range(world)]
        pad = max(b - a for a, b in sizes)
        buf = torch.full((pad, mine.shape[1]), -7, dtype=torch.int64, device=dev)
        buf[:mine.shape[0]] = mine
        allb = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        xs = [torch.empty((pad, x.shape[1]), dtype=torch.float32, device=dev) for _ in range(world)]
        xb = torch.zeros((pad, x.shape[1]), dtype=torch.float32, device=dev)
        xb[:x.shape[0]] = x
        dist.all_gather(xs, xb)
        if rank == 0:
            full = torch.cat([allb[r][:b - a] for r, (a, b) in enumerate(sizes)])
            x_all = torch.cat([xs[r][:b - a] for r, (a, b) in enumerate(sizes)])
            single_model = build_model(cfg, cbs, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref, rstats = rq.generate_codes(single_model, x_all, fast=False)
            torch.cuda.synchronize()
            t_single = time.perf_counter() - t0
            ok = torch.equal(ref, full) and rstats["rounds"] == stats["rounds"] and rstats["distinct"] == stats["distinct"]
            ok_all = ok_all and ok
            print(f"{name}: items={n_total} ranks={world} rounds={stats['rounds']} distinct={stats['distinct']} "
                  f"max_conflicts={stats['max_conflicts']} equal_single_gpu={ok} sharded={t_sharded * 1e3:.1f} ms "
                  f"single_gpu={t_single * 1e3:.1f} ms", flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    assert int(flag.item()) == 1


if __name__ == "__main__":
    main()
