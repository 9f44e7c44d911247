"""torchrun check of the peer-memory sharded dedup (rqb200_shard_*): N-GPU ids == single-GPU dedup of the
concatenated catalogue, on adversarial inputs (ragged shards, an empty rank, one giant collision group).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_shard_p2p.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import sharding    # noqa: E402
from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden                        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    g, cfg, cbs = load_golden("c2_slice")
    model = build_model(cfg, cbs, device=dev)
    cap = 300_000
    pd = sharding.PeerShardDedup(model, dist.group.WORLD, max_local_items=cap, max_recv_items=world * cap)
    ops = sharding.CudaShardOps(model)
    cases = []
    for case, (Ks, sizes, spread) in enumerate([
            ([256, 256, 256], [200_000 + 1234 * r for r in range(world)], 256),     # few collisions
            ([256, 256, 256], [150_000] * world, 6),                                 # heavy collisions (216 codes)
            ([8, 8, 8], [0 if r == 0 else 70_001 for r in range(world)], 8),         # rank 0 empty
            ([1024, 1024, 1024, 1024], [99_999] * world, 1),                         # one giant group
            ([256, 256, 256], [5 + r for r in range(world)], 2)]):                   # tiny
        gen = torch.Generator(device=dev).manual_seed(100 * case + rank)
        n = sizes[rank]
        codes = torch.stack([torch.randint(0, min(k, spread), (n,), generator=gen, device=dev) for k in Ks], 1) \
            if n else torch.zeros((0, len(Ks)), dtype=torch.int64, device=dev)
        for rep in range(2):
            mine = pd(codes, Ks)
        via_nccl = sharding.global_suffix(codes, Ks, ops, dist.group.WORLD)
        # gather (ragged) to rank 0
        pad = max(sizes)
        buf = torch.full((pad, len(Ks) + 1), -7, dtype=torch.int64, device=dev)
        buf[:n] = mine
        allb = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        ok = torch.equal(mine, via_nccl)
        if rank == 0:
            full = torch.cat([allb[r][:sizes[r]] for r in range(world)])
            ref, _ = rq.suffix_dedup(model, full[:, :-1].contiguous()) if full.shape[0] else (full, None)
            ok = ok and torch.equal(ref, full)
            print(f"case {case}: sizes={sizes} Ks={Ks} max_suffix={int(full[:, -1].max()) if full.shape[0] else 0} "
                  f"equal_single_gpu={ok}", flush=True)
        cases.append(ok)
    # timing of the two routes at 1M items per rank
    n = cap
    Ks = [256, 256, 256]
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    codes = torch.stack([torch.randint(0, 256, (n,), generator=gen, device=dev) for _ in Ks], 1)
    for name, fn in (("peer_memory", lambda: pd(codes, Ks)), ("nccl_all_to_all", lambda: sharding.global_suffix(codes, Ks, ops, dist.group.WORLD))):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            print(f"{name}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per call at {n} items/rank x {world} ranks", flush=True)
    assert all(cases)
    pd.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
