"""GPU k-means for codebook initialisation — replaces reference RQ-VAE/models/layers.py:69-82.

Seeding is k-means++ (D² sampling, single trial) driven by a torch Generator; Lloyd iterations run in
the C-ABI kernels (rqb200_kmeans_assign / _accumulate / _update).  With a process group the samples are
one shard per rank: the per-cluster sums[K,e], counts[K] and inertia are all-reduced (NCCL over
NVLink on the GPU box, gloo in the CPU tests of the host logic) so every rank applies the same update.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


class CudaKMeansOps:
    """Local (per-rank) pieces of one Lloyd iteration, on the CUDA kernels."""

    def assign(self, x, centers):
        n, e = x.shape
        K = centers.shape[0]
        cn = torch.empty((K,), dtype=torch.float32, device=x.device)
        a = torch.empty((n,), dtype=torch.int64, device=x.device)
        check(_cabi.lib().rqb200_kmeans_assign(ptr(x), n, e, ptr(centers), K, ptr(cn), ptr(a), stream_ptr(x.device)))
        return a

    def accumulate(self, x, assign, centers):
        n, e = x.shape
        K = centers.shape[0]
        sums = torch.zeros((K, e), dtype=torch.float64, device=x.device)
        counts = torch.zeros((K,), dtype=torch.int64, device=x.device)
        inertia = torch.zeros((1,), dtype=torch.float64, device=x.device)
        check(_cabi.lib().rqb200_kmeans_accumulate(ptr(x), n, e, ptr(assign), ptr(centers), K, ptr(sums),
                                                   ptr(counts), ptr(inertia), stream_ptr(x.device)))
        return sums, counts, inertia

    def update(self, centers, sums, counts):
        K, e = centers.shape
        shift = torch.zeros((1,), dtype=torch.float64, device=centers.device)
        check(_cabi.lib().rqb200_kmeans_update(ptr(centers), K, e, ptr(sums), ptr(counts), ptr(shift),
                                               stream_ptr(centers.device)))
        return shift

    def sqdist(self, x, cands):
        """[n, T] squared distances to T candidate centres (seeding) — the quantizer's distance kernel."""
        n, e = x.shape
        T = cands.shape[0]
        cn = torch.empty((T,), dtype=torch.float32, device=x.device)
        d = torch.empty((n, T), dtype=torch.float32, device=x.device)
        if n:
            check(_cabi.lib().rqb200_kmeans_distances(ptr(x), n, e, ptr(cands.contiguous()), T, ptr(cn), ptr(d),
                                                      stream_ptr(x.device)))
        return d.clamp_min_(0.0)


def _all_reduce(t, group):
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def kmeans_pp_seed(x: torch.Tensor, K: int, gen: torch.Generator, ops, group=None) -> torch.Tensor:
    """Greedy k-means++ seeding (like scikit-learn: 2 + log K candidates per step, keep the one that lowers the
    potential most).  Sharded mode: every rank draws the same uniform numbers; the rank that owns a selected
    global position contributes the row through an all-reduce, potentials are all-reduced.
    The whole loop stays on the device: every decision (degenerate weights, which rank owns a draw, the best candidate)
    is a tensor expression, so a step is a fixed sequence of launches with no host read — the K steps queue up behind each
    other instead of paying a device round trip per centre."""
    import math
    n, e = x.shape
    dev = x.device
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    trials = 2 + int(math.log(K)) if K > 1 else 1
    uniforms = torch.rand((K, trials), generator=gen, dtype=torch.float64).to(dev)        # the same numbers on every rank
    centers = torch.empty((K, e), dtype=torch.float32, device=dev)
    ones = torch.ones((n,), dtype=torch.float64, device=dev)
    counts = torch.zeros(world, dtype=torch.float64, device=dev)
    counts[rank] = float(n)
    _all_reduce(counts, group)
    ranks = torch.arange(world, device=dev)
    mind = None                                   # current min squared distance of every local sample
    for k in range(K):
        T = 1 if k == 0 else trials
        u = uniforms[k, :T]
        w_local = ones if mind is None else mind.to(torch.float64)
        tot = torch.zeros(world, dtype=torch.float64, device=dev)
        tot[rank] = w_local.sum()
        _all_reduce(tot, group)
        degenerate = ~(tot.sum() > 0)             # all remaining samples coincide with a centre: draw uniformly
        w_local = torch.where(degenerate, ones, w_local)
        tot = torch.where(degenerate, counts, tot)
        total = tot.sum()
        target = u * total                        # [T] positions in the global cumulative weight
        before = (torch.cumsum(tot, 0) - tot)[rank]
        mine = tot[rank]
        last_nonempty = (ranks * (tot > 0)).max()
        cands = torch.zeros((T, e), dtype=torch.float32, device=dev)
        if n > 0:
            cs = torch.cumsum(w_local, 0)
            local_t = target - before
            own = ((local_t >= 0) & (local_t < mine)) | ((last_nonempty == rank) & (local_t >= mine))
            own = own & (mine > 0)
            j = torch.searchsorted(cs, local_t.clamp(min=0)).clamp(0, n - 1)
            cands = torch.where(own[:, None], x[j], cands)
        _all_reduce(cands, group)
        d = ops.sqdist(x, cands)                  # [n, T]
        if mind is not None:
            d = torch.minimum(d, mind[:, None])
        pot = d.to(torch.float64).sum(0)
        _all_reduce(pot, group)
        best = torch.argmin(pot).reshape(1)       # stays on the device
        centers[k] = cands.index_select(0, best)[0]
        mind = d.index_select(1, best).reshape(-1).contiguous()
    return centers


def _relocate_empty(x, assign, centers, sums, counts, group):
    """scikit-learn's `_relocate_empty_clusters_dense` (what `KMeans.fit` at layers.py:77 does when a cluster loses all
    its samples): the n_empty samples farthest from their current centre become the centres of the empty clusters and
    leave their old cluster.  Rare path (host logic over the statistics; the far samples of all shards meet through an
    all-gather).  Farthest sample → lowest empty cluster id; scikit-learn hands them out in argpartition order, so the
    SET of centres is the same, their order among the relocated ids may differ."""
    empty = torch.nonzero(counts == 0).flatten()
    n_empty, (n, e) = int(empty.numel()), x.shape
    k = min(n_empty, n)
    dist = ((x - centers[assign]) ** 2).sum(1) if n else x.new_zeros((0,))
    vals = torch.full((n_empty,), -1.0, dtype=torch.float32, device=x.device)
    rows = torch.zeros((n_empty, e), dtype=torch.float32, device=x.device)
    labs = torch.zeros((n_empty,), dtype=torch.int64, device=x.device)
    if k:
        v, i = torch.topk(dist, k)
        vals[:k], rows[:k], labs[:k] = v, x[i], assign[i]
    if group is not None:
        import torch.distributed as dist_
        world = dist_.get_world_size(group)
        gv = [torch.empty_like(vals) for _ in range(world)]
        gr = [torch.empty_like(rows) for _ in range(world)]
        gl = [torch.empty_like(labs) for _ in range(world)]
        dist_.all_gather(gv, vals, group=group)
        dist_.all_gather(gr, rows, group=group)
        dist_.all_gather(gl, labs, group=group)
        vals, rows, labs = torch.cat(gv), torch.cat(gr), torch.cat(gl)
    order = torch.argsort(vals, descending=True, stable=True)[:n_empty]
    if not bool(vals[order[0]] > 0):
        return sums, counts                     # more clusters than distinct samples: relocating is pointless
    sums, counts = sums.clone(), counts.clone()
    for new_id, j in zip(empty.tolist(), order.tolist()):
        if float(vals[j]) < 0:
            break
        old_id = int(labs[j])
        row = rows[j].to(torch.float64)
        sums[old_id] -= row
        sums[new_id] = row
        counts[new_id] = 1
        counts[old_id] -= 1
    return sums, counts


def kmeans_fit(samples: torch.Tensor, num_clusters: int, num_iters: int = 10, seed: Optional[int] = None,
               init: Optional[torch.Tensor] = None, group=None, tol: float = 1e-4, ops=None,
               return_info: bool = False):
    ops = ops or CudaKMeansOps()
    x = samples.detach().reshape(-1, samples.shape[-1]).to(torch.float32).contiguous()
    n, e = x.shape
    n_total = torch.tensor([float(n)], dtype=torch.float64, device=x.device)
    _all_reduce(n_total, group)
    if int(n_total.item()) < num_clusters:
        # same failure as scikit-learn's KMeans inside the reference (layers.py:77)
        raise ValueError(f"n_samples={int(n_total.item())} should be >= n_clusters={num_clusters}.")
    if isinstance(ops, CudaKMeansOps) and not x.is_cuda:
        raise RuntimeError("kmeans: CUDA tensor required (no CPU fallback)")
    if init is not None:
        centers = init.detach().to(torch.float32).to(x.device).contiguous().clone()
    else:
        gen = torch.Generator()
        gen.manual_seed(2024 if seed is None else int(seed))
        centers = kmeans_pp_seed(x, num_clusters, gen, ops, group)
    # tolerance like scikit-learn: tol * mean feature variance
    s1 = torch.cat([x.to(torch.float64).sum(0), (x.to(torch.float64) ** 2).sum(0)])
    _all_reduce(s1, group)
    nt = float(n_total.item())
    var = (s1[e:] / nt - (s1[:e] / nt) ** 2).clamp_min(0).mean().item()
    info = {"iters": 0, "inertia": None}
    for it in range(int(num_iters)):
        a = ops.assign(x, centers)
        sums, counts, inertia = ops.accumulate(x, a, centers)
        packed = torch.cat([sums.reshape(-1), counts.to(torch.float64), inertia])
        _all_reduce(packed, group)
        K = centers.shape[0]
        sums = packed[:K * e].reshape(K, e).contiguous()
        counts = packed[K * e:K * e + K].round().to(torch.int64).contiguous()
        info["inertia"] = float(packed[-1].item())
        if bool((counts == 0).any()):
            sums, counts = _relocate_empty(x, a, centers, sums, counts, group)
        shift = ops.update(centers, sums, counts)
        info["iters"] = it + 1
        if tol > 0 and float(shift.item()) <= tol * var:
            break
    return (centers, info) if return_info else centers
