"""CPU, world_size 2 over gloo: the multi-GPU host logic (global suffix dedup exchange, sharded k-means
statistics all-reduce).  The per-rank kernels are replaced by oracle-backed numpy twins — only here, in tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ai_education_generative_recommendation_b200 import sharding
from ai_education_generative_recommendation_b200.kmeans_gpu import kmeans_fit


class NumpyShardOps:
    def pack_keys(self, codes, num_emb_list):
        bits = sharding.key_bits(num_emb_list)
        keys = torch.zeros(codes.shape[0], dtype=torch.int64)
        shift = 0
        for l in range(len(num_emb_list) - 1, -1, -1):
            keys |= codes[:, l] << shift
            shift += bits[l]
        return keys

    def bucket_by_owner(self, owner, world):
        order = torch.sort(owner, stable=True).indices
        return order, torch.bincount(owner, minlength=world).to(torch.int64)

    def rank_among_equal(self, keys, bits):
        from oracle import oracle as O
        if keys.numel() == 0:
            return torch.zeros(0, dtype=torch.int64)
        return torch.from_numpy(O.suffix_dedup(keys.numpy()[:, None])[:, 1].copy())


class NumpyDriverOps(NumpyShardOps):
    """Oracle-backed twin of the per-rank pieces CudaShardOps provides to generate_codes_sharded (tests only)."""

    def __init__(self, enc_w, enc_b, cbs, eps, iters):
        self.enc_w, self.enc_b, self.cbs, self.eps, self.iters = enc_w, enc_b, cbs, eps, iters

    def encode_codes(self, data):
        from oracle import oracle as O
        z = O.mlp(data.numpy(), self.enc_w, self.enc_b, threads=1)
        return torch.from_numpy(O.quantize(z, self.cbs, want_xq=False, threads=1)[0])

    def last_level_sinkhorn(self):
        return self.eps > 0

    def resolve_local(self, codes, data, max_rounds):
        raise AssertionError("world == 1 is the single-process driver; not exercised here")

    def groups(self, codes):
        from oracle import oracle as O
        n = codes.shape[0]
        sizes = torch.ones((n,), dtype=torch.int64)
        groups = O.collision_groups(codes.numpy()) if n else []
        groups.sort(key=lambda g: tuple(codes[int(g[0])].tolist()))          # any order: groups are disjoint
        items = np.concatenate(groups).astype(np.int64) if groups else np.zeros(0, dtype=np.int64)
        offsets = np.cumsum([0] + [len(g) for g in groups]).astype(np.int64)
        for g in groups:
            sizes[torch.from_numpy(np.asarray(g))] = len(g)
        return torch.from_numpy(items), torch.from_numpy(offsets), sizes

    def reencode_members(self, data, members, msize, codes):
        from oracle import oracle as O
        out = torch.zeros((members.numel(), self.cbs[0].shape[1]), dtype=torch.float32)
        x = data.numpy()
        for M in sorted(set(msize.tolist())):
            sel = torch.nonzero(msize == M).flatten()
            rows = members[sel].numpy()
            pc, r = O.reencode_prefix(x[rows], self.enc_w, self.enc_b, self.cbs, batch_size=int(M))
            codes[members[sel], :len(self.cbs) - 1] = torch.from_numpy(pc)
            out[sel] = torch.from_numpy(r)
        return out

    def sinkhorn_groups(self, codes, items, offsets, residual_rows):
        from oracle import oracle as O
        items, offsets, res = items.numpy(), offsets.numpy(), residual_rows.numpy()
        for a, b in zip(offsets[:-1], offsets[1:]):
            g = items[a:b]
            codes[torch.from_numpy(g), -1] = torch.from_numpy(O.sinkhorn_last_level(res[g], self.cbs[-1], self.eps, self.iters))


class _Cfg:
    def __init__(self, Ks):
        self.num_emb_list = Ks


def _driver_job(rank, world, x_all, enc_w, enc_b, cbs, eps, iters):
    lo, hi = sharding.shard_range(x_all.shape[0], rank, world)
    ops = NumpyDriverOps(enc_w, enc_b, cbs, eps, iters)
    return sharding.generate_codes_sharded(_Cfg([c.shape[0] for c in cbs]), x_all[lo:hi], dist.group.WORLD, ops=ops)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_driver_matches_single_process_driver(tmp_path, oracle, world):
    """Sinkhorn re-encode rounds + suffix over a sharded catalogue == the single-process driver (infer.py:88-177)."""
    from conftest import load_golden, synth_weights
    from ai_education_generative_recommendation_b200 import synth
    g, cfg, cbs = load_golden("c1_slice")
    _, (ew, eb), _ = synth_weights(cfg)
    n = 600
    x = synth.synth_items(2024, 0, n, cfg["in_dim"], 1_000_000)
    x[300:340] = x[10:50]                                   # exact duplicates across the shard boundary
    ref, rstats = oracle.generate_codes(x, ew, eb, cbs, cfg["sk_epsilons"], cfg["sk_iters"], group_order=True)
    assert rstats["rounds"] >= 1
    res = run_world(_driver_job, (torch.from_numpy(x), ew, eb, cbs, float(cfg["sk_epsilons"][-1]), cfg["sk_iters"]),
                    tmp_path, world=world)
    got = torch.cat([r[0] for r in res]).numpy()
    assert np.array_equal(got, ref)
    assert all(r[1] == res[0][1] for r in res)
    assert res[0][1]["rounds"] == rstats["rounds"]
    assert res[0][1]["distinct"] == len(np.unique(ref[:, :-1], axis=0))


class NumpyKMeansOps:
    def assign(self, x, centers):
        d = ((x[:, None, :].double() - centers[None].double()) ** 2).sum(-1)
        return d.argmin(1)

    def accumulate(self, x, assign, centers):
        K, e = centers.shape
        sums = torch.zeros((K, e), dtype=torch.float64).index_add_(0, assign, x.double())
        counts = torch.bincount(assign, minlength=K).to(torch.int64)
        inertia = ((x.double() - centers[assign].double()) ** 2).sum().reshape(1)
        return sums, counts, inertia

    def update(self, centers, sums, counts):
        new = centers.clone()
        m = counts > 0
        new[m] = (sums[m] / counts[m][:, None].double()).float()
        shift = ((new.double() - centers.double()) ** 2).sum().reshape(1)
        centers.copy_(new)
        return shift

    def sqdist(self, x, cands):
        return ((x[:, None, :].double() - cands[None].double()) ** 2).sum(-1).float()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, args, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = fn(rank, world, *args)
        torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def run_world(fn, args, tmp_path, world=2):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, fn, args, str(tmp_path)), nprocs=world, join=True)
    return [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]


def _suffix_job(rank, world, codes_all, Ks):
    lo, hi = sharding.shard_range(codes_all.shape[0], rank, world)
    out = sharding.global_suffix(codes_all[lo:hi].clone(), Ks, NumpyShardOps(), dist.group.WORLD)
    stats = sharding.global_stats(out, dist.group.WORLD)
    return out, stats


@pytest.mark.parametrize("n,Ks", [(5000, [8, 8, 8]), (20001, [256, 256, 256]), (3, [4, 4]), (4097, [1024] * 4)])
def test_global_suffix_matches_single_process(tmp_path, oracle, n, Ks):
    rng = np.random.default_rng(n)
    codes = np.stack([rng.integers(0, min(k, 6), size=n) for k in Ks], axis=1).astype(np.int64)
    codes[rng.integers(0, n, size=n // 2)] = codes[rng.integers(0, n, size=n // 2)]
    res = run_world(_suffix_job, (torch.from_numpy(codes), Ks), tmp_path)
    got = torch.cat([r[0] for r in res]).numpy()
    ref = oracle.suffix_dedup(codes)
    assert np.array_equal(got, ref)                       # identical to the single-GPU / reference rule
    assert res[0][1] == res[1][1]
    assert res[0][1]["distinct"] == len(np.unique(codes, axis=0))
    assert res[0][1]["max_conflicts"] == int(ref[:, -1].max()) + 1


def _kmeans_job(rank, world, x_all, init):
    lo, hi = sharding.shard_range(x_all.shape[0], rank, world)
    c, info = kmeans_fit(x_all[lo:hi], init.shape[0], 8, init=init, group=dist.group.WORLD, ops=NumpyKMeansOps(),
                         return_info=True, tol=0.0)
    seeded = kmeans_fit(x_all[lo:hi], init.shape[0], 3, seed=7, group=dist.group.WORLD, ops=NumpyKMeansOps())
    return c, info, seeded


def test_sharded_kmeans_equals_single_process(tmp_path, oracle):
    rng = np.random.default_rng(11)
    centres = rng.standard_normal((8, 16)).astype(np.float32)
    x = (centres[rng.integers(0, 8, size=4001)] + 0.05 * rng.standard_normal((4001, 16))).astype(np.float32)
    init = x[rng.choice(4001, size=8, replace=False)].copy()
    res = run_world(_kmeans_job, (torch.from_numpy(x), torch.from_numpy(init)), tmp_path)
    assert torch.equal(res[0][0], res[1][0])              # every rank applies the same update
    single = kmeans_fit(torch.from_numpy(x), 8, 8, init=torch.from_numpy(init), ops=NumpyKMeansOps(), tol=0.0)
    assert torch.allclose(res[0][0], single, rtol=0, atol=1e-6)
    ref = oracle.kmeans_lloyd(x, init, 8, tol=0.0)
    assert np.allclose(res[0][0].numpy(), ref, atol=1e-5)
    assert torch.equal(res[0][2], res[1][2])              # sharded k-means++ seeding is rank-consistent
    with pytest.raises(ValueError, match="should be >= n_clusters"):
        kmeans_fit(torch.from_numpy(x[:4]), 8, 2, ops=NumpyKMeansOps())


def test_shard_ranges_cover_the_catalogue():
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))


class _FakeModel:
    """Stands in for the RQVAE in Trainer._valid_epoch: codes are a fixed function of the rows (CPU)."""
    num_emb_list = [4, 4, 4]

    def eval(self):
        return self

    def get_indices(self, data):
        return (data[:, :3].abs() * 3.999).long().clamp_(0, 3)


def _valid_epoch_job(rank, world, x_all, batch):
    import types
    from ai_education_generative_recommendation_b200.trainer import Trainer
    me = types.SimpleNamespace(model=_FakeModel(), device=torch.device("cpu"), world=world, rank=rank, slice_batches=True,
                               group=dist.group.WORLD, shard_ops=NumpyShardOps())
    loader = [x_all[i:i + batch] for i in range(0, x_all.shape[0], batch)]
    return Trainer._valid_epoch(me, loader)


def test_data_parallel_collision_rate_counts_every_item_once(tmp_path, oracle):
    """Trainer._valid_epoch with sliced batches (every rank iterates over the SAME batches, train.py:126-151): the global
    collision rate equals the single-process value — every item enters the statistics exactly once."""
    torch.manual_seed(3)
    x = torch.rand(1003, 8)
    codes = _FakeModel().get_indices(x).numpy()
    want = 1.0 - len(np.unique(codes, axis=0)) / len(codes)
    res = run_world(_valid_epoch_job, (x, 64), tmp_path, world=2)
    assert abs(res[0] - want) < 1e-12 and abs(res[1] - want) < 1e-12


def test_checkpoint_keeper_follows_the_reference_rotation_rule(monkeypatch):
    """_CheckpointKeeper against the heap / queue rule of train.py:232-248 replayed verbatim on random histories."""
    import heapq
    import random
    import ai_education_generative_recommendation_b200.trainer as T
    gone = []
    monkeypatch.setattr(T, "_delete_file", lambda p, rank=0: gone.append(p))
    random.seed(1)
    for _ in range(100):
        limit = random.choice([1, 2, 3, 5])
        keeper = T._CheckpointKeeper(limit)
        gone.clear()
        heap, queue, ref_gone = [], [], []
        for e in range(30):
            rate = random.choice([0.1, 0.2, 0.3, 0.05, 0.5])
            path = f"epoch_{e}_collision_{rate:.4f}_model.pth"
            keeper.file(rate, path)
            now = (-rate, path)
            if len(queue) < limit:
                queue.append(now)
                heapq.heappush(heap, now)
            else:
                old = queue.pop(0)
                queue.append(now)
                if rate < -heap[0][0]:
                    bad = heapq.heappop(heap)
                    heapq.heappush(heap, now)
                    if bad not in queue:
                        ref_gone.append(bad[1])
                if old not in heap:
                    ref_gone.append(old[1])
        assert sorted(ref_gone) == sorted(gone)
