"""k-means codebook init (reference layers.py:69-82 = scikit-learn KMeans on the CPU) on the GPU kernels.

    python tools/bench_kmeans.py                       # 1 GPU: time seeding + Lloyd iterations, inertia vs scikit-learn
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_kmeans.py                          # sharded samples: centres equal the single-GPU run
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ai_education_generative_recommendation_b200 as rq           # noqa: E402
from ai_education_generative_recommendation_b200.kmeans_gpu import kmeans_fit   # noqa: E402


def make(n, e, K, seed):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn((K, e), generator=g)
    lab = torch.randint(0, K, (n,), generator=g)
    return (centres[lab] + 0.3 * torch.randn((n, e), generator=g)).float()


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        group = dist.group.WORLD
    n, e, K, iters = int(os.environ.get("KM_N", 1_000_000)), 64, 256, 10
    x_all = make(n, e, K, 7)
    lo, hi = rank * n // world, (rank + 1) * n // world
    x = x_all[lo:hi].to(dev)
    out = {"n": n, "e": e, "K": K, "iters": iters, "world": world}
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        centers, info = kmeans_fit(x, K, iters, seed=3, group=group, tol=0.0, return_info=True)
        torch.cuda.synchronize()
        out["seconds_total"] = time.perf_counter() - t0
    out["inertia"] = info["inertia"]
    init = centers.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c2, info2 = kmeans_fit(x, K, iters, init=init, group=group, tol=0.0, return_info=True)
    torch.cuda.synchronize()
    out["seconds_lloyd_only"] = time.perf_counter() - t0
    if world > 1:
        # the sharded run must reproduce the single-GPU run on the concatenated samples (same seed ⇒ same seeding draws)
        if rank == 0:
            ref, rinfo = kmeans_fit(x_all.to(dev), K, iters, seed=3, group=None, tol=0.0, return_info=True)
            out["max_abs_centre_diff_vs_single_gpu"] = float((ref - centers).abs().max())
            out["inertia_single_gpu"] = rinfo["inertia"]
    elif os.environ.get("KM_SKLEARN", "1") == "1":
        from sklearn.cluster import KMeans
        ns = min(n, 200_000)
        t0 = time.perf_counter()
        km = KMeans(n_clusters=K, max_iter=iters, n_init=1, random_state=0).fit(x_all[:ns].numpy())
        out["sklearn_seconds_on_%d_rows" % ns] = time.perf_counter() - t0
        d = torch.cdist(x_all[:ns].to(dev), centers).min(1).values
        out["inertia_gpu_centres_on_those_rows"] = float((d.double() ** 2).sum())
        out["inertia_sklearn_on_those_rows"] = float(km.inertia_)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
