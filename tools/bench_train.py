"""BASELINE configs[3]: k-means codebook init + one RQ-VAE training epoch, per-cluster sum/count and gradient
all-reduce over NVLink when launched with torchrun.  Prints one JSON line (rank 0).

    python tools/bench_train.py [--items 1000000] [--batch 4096] [--levels 4 --codes 256 --e-dim 64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_train.py --items 1250000
--items is PER GPU (weak scaling: every rank trains on its own shard, gradients averaged every step).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200 import _cabi      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--in-dim", type=int, default=768)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--codes", type=int, default=256)
    ap.add_argument("--e-dim", type=int, default=64)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--kmeans-iters", type=int, default=10)
    ap.add_argument("--max-steps", type=int, default=0, help="stop the epoch early (0 = whole epoch)")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        group = dist.group.WORLD
    lib = _cabi.lib()
    x = torch.empty((a.items, a.in_dim), dtype=torch.float32, device=dev)
    _cabi.check(lib.rqb200_synth_items(2024, rank * a.items, a.items, a.in_dim, world * a.items, _cabi.ptr(x),
                                       _cabi.stream_ptr(x.device)))
    torch.manual_seed(2024)
    model = rq.RQVAE(in_dim=a.in_dim, num_emb_list=[a.codes] * a.levels, e_dim=a.e_dim, layers=[256, 128],
                     dropout_prob=a.dropout, quant_loss_weight=0.1, beta=0.25, kmeans_init=True, kmeans_iters=a.kmeans_iters,
                     sk_epsilons=[0.01] * a.levels, sk_iters=50)
    loader = rq.DeviceBatches(x, a.batch, shuffle=True, seed=2024 + rank, drop_last=True)
    if a.max_steps:
        full = loader

        class _Head:
            def __len__(self): return min(len(full), a.max_steps)
            def __iter__(self):
                for i, b in enumerate(full):
                    if i >= a.max_steps:
                        return
                    yield b
        loader = _Head()
    params = dict(lr=1e-3, learner="AdamW", lr_scheduler_type="linear", weight_decay=1e-4, epochs=1, warmup_epochs=0,
                  save_limit=1, eval_step=1, device=dev, ckpt_dir=os.path.join(ROOT, "gpurun_out", "bench_ckpt"))
    trainer = rq.Trainer(params, model, len(loader), group=group)
    trainer.slice_batches = False                      # every rank iterates over its own shard
    # first batch: k-means init of every level (timed separately), then the epoch
    first = next(iter(loader))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    model.train()
    with torch.no_grad():
        model(first)
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    l0 = lib.rqb200_launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss, recon = trainer._train_epoch(loader, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_epoch = time.perf_counter() - t0
    steps = trainer.last_epoch_steps
    tt = torch.tensor([t_epoch, t_init], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        items = steps * a.batch * world
        print(json.dumps({
            "metric": "rqvae_train_items_per_s", "value": items / float(tt[0]), "unit": "items/s", "n_gpus": world,
            "steps": steps, "ms_per_step": 1e3 * float(tt[0]) / max(steps, 1), "kmeans_init_s": float(tt[1]),
            "mean_loss": loss / max(steps, 1), "mean_recon": recon / max(steps, 1),
            "launches_per_step": (lib.rqb200_launch_count() - l0) / max(steps, 1),
            "config": {"workload": "BASELINE configs[3] shapes: k-means init + one training epoch", "items_per_gpu": a.items,
                       "batch_per_gpu": a.batch, "in_dim": a.in_dim, "levels": a.levels, "codes": a.codes, "e_dim": a.e_dim,
                       "dropout": a.dropout, "sinkhorn": "eps 0.01 x 50 iters on every level (main.py:27-28)",
                       "optimizer": "fused clip 1.0 + AdamW", "timing": "host wall clock around the epoch, max over ranks"}}),
              flush=True)
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        it = iter(loader)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                data = next(it)
                trainer.optimizer.zero_grad()
                out, ql, _ = model(data)
                lo, _ = model.compute_loss(out, ql, xs=data)
                lo.backward()
                trainer.optimizer.step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
