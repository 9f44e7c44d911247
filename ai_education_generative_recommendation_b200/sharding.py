"""Multi-GPU host logic: one process per GPU, item catalogue sharded by contiguous ranges.

The reference is single-process; this is new design for BASELINE configs 3-5:
  * encode shards independently (weights / codebooks replicated, no data-path collective);
  * global collision handling needs ONE exchange step: every item's packed code is routed to the rank
    that owns the key (all-to-all), the owner ranks equal keys in ascending GLOBAL item order, and the
    ranks travel back (all-to-all).  The result is bit-identical to the single-GPU suffix column;
  * the Sinkhorn re-encode rounds (reference infer.py:109-130): per round the full codes travel to the key's owner,
    which finds the groups; every member's home rank — which holds the item's embedding — re-encodes it as the reference
    does (whole model, in the arithmetic of a batch of the group's size), the residuals entering the last level meet
    at the owner for the group's Sinkhorn pass, and the new last-level codes travel home (`generate_codes_sharded`).
Collectives go through torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests); the
per-rank work is behind a small `ops` object — `CudaShardOps` (C-ABI kernels) in production, an
oracle-backed numpy twin only inside tests/.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi
from ._cabi import check, ptr, stream_ptr


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous item range [lo, hi) of `rank` ([r·N/G, (r+1)·N/G))."""
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def key_bits(num_emb_list) -> List[int]:
    return [max(1, (int(k) - 1).bit_length()) for k in num_emb_list]


class CudaShardOps:
    """Per-rank pieces on the C-ABI kernels."""

    def __init__(self, model):
        self.model = model

    def pack_keys(self, codes: torch.Tensor, num_emb_list) -> torch.Tensor:
        n, Lv = codes.shape
        keys = torch.empty((n,), dtype=torch.int64, device=codes.device)
        check(_cabi.lib().rqb200_pack_keys(ptr(codes.contiguous()), n, Lv, _cabi.int_array(num_emb_list), ptr(keys),
                                           stream_ptr(codes.device)))
        return keys

    def bucket_by_owner(self, owner: torch.Tensor, world: int):
        """Stable bucketing by destination rank: one radix pass of the library's sort (keys = owner)."""
        n = owner.numel()
        self.model._ensure_handle()
        sk = owner.clone()
        pos = torch.arange(n, dtype=torch.int64, device=owner.device)
        if n:
            check(_cabi.lib().rqb200_sort_pairs(self.model._handle, ptr(sk), ptr(pos), n,
                                                max(1, (world - 1).bit_length()), stream_ptr(owner.device)))
        bounds = torch.searchsorted(sk, torch.arange(world + 1, dtype=torch.int64, device=owner.device))
        return pos, (bounds[1:] - bounds[:-1]).to(torch.int64)

    def rank_among_equal(self, keys: torch.Tensor, bits: int) -> torch.Tensor:
        """out[i] = #{j < i : keys[j] == keys[i]} — stable radix sort + segmented rank."""
        n = keys.numel()
        out = torch.empty((n,), dtype=torch.int64, device=keys.device)
        if n == 0:
            return out
        self.model._ensure_handle()
        sk = keys.clone()
        pos = torch.arange(n, dtype=torch.int64, device=keys.device)
        lib = _cabi.lib()
        s = stream_ptr(keys.device)
        check(lib.rqb200_sort_pairs(self.model._handle, ptr(sk), ptr(pos), n, int(bits), s))
        rk = torch.empty((n,), dtype=torch.int64, device=keys.device)
        check(lib.rqb200_segment_rank(self.model._handle, ptr(sk), n, ptr(rk), s))
        out[pos] = rk
        return out


    # ---- the complete sharded driver (generate_codes_sharded) ----
    def encode_codes(self, data):
        """Pass 1 on this rank's shard: codes[n, L] (tensor-core route where the shapes allow it — every row certified by
        the margin gate or recomputed exactly — else the exact route)."""
        from .generate_code import encode_codes_exact, encode_codes_fast
        fast = self.model.fast_route_supported()
        return encode_codes_fast(self.model, data) if fast else encode_codes_exact(self.model, data)

    def last_level_sinkhorn(self) -> bool:
        last = self.model.rq.vq_layers[-1]
        return last.sk_epsilon is not None and last.sk_epsilon > 0

    def resolve_local(self, codes, data, max_rounds):
        from .generate_code import resolve_rounds
        return resolve_rounds(self.model, codes, data, max_rounds=max_rounds)

    def groups(self, codes: torch.Tensor):
        """Collision groups among `codes` (rows in arrival order) → (items, offsets, sizes[n] with 1 = in no group)."""
        from .generate_code import collision_groups
        n = codes.shape[0]
        sizes = torch.ones((n,), dtype=torch.int64, device=codes.device)
        if n == 0:
            return codes.new_zeros((0,)), codes.new_zeros((1,)), sizes
        items, offsets, _ = collision_groups(self.model, codes.contiguous())
        if items.numel():
            sizes[items] = torch.repeat_interleave(offsets[1:] - offsets[:-1], offsets[1:] - offsets[:-1])
        return items, offsets, sizes

    def reencode_members(self, data, members: torch.Tensor, msize: torch.Tensor, codes: torch.Tensor) -> torch.Tensor:
        """Re-encode the listed local items as members of groups of the given sizes (infer.py:120-122 up to the last
        level): codes[members, :L-1] are overwritten, the residuals entering the last level are returned [m, e]."""
        from .generate_code import _as_rows
        m = self.model
        dev = codes.device
        k = members.numel()
        out = torch.empty((k, m.e_dim), dtype=torch.float32, device=dev)
        if k == 0:
            return out
        data = _as_rows(data)
        residual = torch.empty((codes.shape[0], m.e_dim), dtype=torch.float32, device=dev)
        if data.is_cuda:
            x, gathered = data, 0
        else:
            x, gathered = data[members.cpu()].contiguous().to(dev), 1
        m._sync()
        check(_cabi.lib().rqb200_reencode_rows(m._handle, ptr(x), gathered, ptr(members.contiguous()),
                                               ptr(msize.to(torch.int32).contiguous()), k, ptr(codes), ptr(residual),
                                               stream_ptr(dev)))
        return residual[members]

    def sinkhorn_groups(self, codes: torch.Tensor, items, offsets, residual_rows: torch.Tensor):
        """Last level of the re-encode for the groups found by `groups` (vq.py:74-83 over each group's distance matrix):
        codes[item, L-1] overwritten in place.  residual_rows[n, e]: rows of the members are filled."""
        from .generate_code import _regroup_oversized
        m = self.model
        lib = _cabi.lib()
        last = m.rq.vq_layers[-1]
        n_groups = offsets.numel() - 1
        if n_groups <= 0:
            return
        max_group = int((offsets[1:] - offsets[:-1]).max().item())
        cap = lib.rqb200_sinkhorn_group_cap(m._handle)
        m._sync()
        check(lib.rqb200_sinkhorn_regroup(m._handle, ptr(residual_rows), ptr(items), ptr(offsets), n_groups, min(max_group, cap),
                                          float(last.sk_epsilon), int(last.sk_iters), ptr(codes), stream_ptr(codes.device)))
        if max_group > cap:
            _regroup_oversized(m, residual_rows, items, offsets, cap, codes)


class PeerShardDedup:
    """Global suffix column over NVLink peer memory (`rqb200_shard_*`, csrc/dedup.cu): the production path on
    the GPU box.  torch.distributed is used once, to pass the cudaIpc handles around; the data path is the
    library's own kernels storing into peer memory — no NCCL call per step.

    `max_local_items` bounds n of every later call on this rank; an owner may receive up to `max_recv_items`
    keys (default 2 * max_local_items + 65536; RQB200 raises MemoryError on every rank beyond that)."""

    def __init__(self, model, group=None, max_local_items: int = 0, max_recv_items: int = 0):
        self.model = model
        self.group = group
        model._ensure_handle()
        lib = _cabi.lib()
        if group is not None and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        caps = [int(max_local_items)]
        if self.world > 1:
            got = [None] * self.world
            dist.all_gather_object(got, int(max_local_items), group=group)
            caps = got
        self.max_local = max(caps)            # symmetric layout: identical sizes on every rank
        self._h = ctypes.c_void_p()
        check(lib.rqb200_shard_create(ctypes.byref(self._h), model._handle, self.rank, self.world, self.max_local,
                                      int(max_recv_items)))
        if self.world > 1:
            nb = lib.rqb200_shard_handle_bytes()
            mine = ctypes.create_string_buffer(nb)
            check(lib.rqb200_shard_get_handle(self._h, ctypes.cast(mine, ctypes.c_void_p)))
            got = [None] * self.world
            dist.all_gather_object(got, bytes(mine.raw), group=group)
            blob = ctypes.create_string_buffer(b"".join(got), nb * self.world)
            check(lib.rqb200_shard_connect(self._h, ctypes.cast(blob, ctypes.c_void_p)))
            dist.barrier(group=group)

    def __call__(self, codes: torch.Tensor, num_emb_list) -> torch.Tensor:
        """[n_local, L] int64 codes of this rank's contiguous shard → [n_local, L+1] with the global suffix."""
        if not codes.is_cuda or codes.dtype != torch.int64:
            raise RuntimeError("PeerShardDedup needs int64 CUDA codes (there is no CPU fallback)")
        codes = codes.contiguous()
        n, Lv = codes.shape
        out = torch.empty((n, Lv + 1), dtype=torch.int64, device=codes.device)
        check(_cabi.lib().rqb200_shard_suffix_dedup(self._h, ptr(codes), n, Lv, _cabi.int_array(num_emb_list), ptr(out),
                                                    stream_ptr(codes.device)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _cabi.lib().rqb200_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def owner_of(keys: torch.Tensor, world: int) -> torch.Tensor:
    """Rank that owns a key: a multiplicative hash so clustered codes spread evenly."""
    h = keys * -7046029254386353131          # 0x9E3779B97F4A7C15 as int64 (wraps)
    h = (h >> 33) & 0x3FFFFFFF
    return h % world


def _exchange_counts(send_counts: torch.Tensor, group) -> Tuple[List[int], List[int]]:
    """One small all-to-all of the per-destination counts; returns (send, recv) split sizes as host lists."""
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    both = torch.stack([send_counts, recv_counts]).cpu().tolist()      # single device→host sync
    return both[0], both[1]


def _all_to_all(send: torch.Tensor, sc: List[int], rc: List[int], group) -> torch.Tensor:
    recv = torch.empty((sum(rc),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=group)
    return recv


def global_suffix(codes: torch.Tensor, num_emb_list, ops, group=None) -> torch.Tensor:
    """[n_local, L] codes of this rank's contiguous shard → [n_local, L+1] with the GLOBAL suffix column
    (identical to running the single-GPU dedup on the concatenated catalogue)."""
    bits = sum(key_bits(num_emb_list))
    if bits > 63:
        raise ValueError("packed code needs more than 63 bits")
    keys = ops.pack_keys(codes, num_emb_list)
    if group is None or dist.get_world_size(group) == 1:
        suffix = ops.rank_among_equal(keys, bits)
        return torch.cat([codes, suffix[:, None]], dim=1)
    world = dist.get_world_size(group)
    owner = owner_of(keys, world)
    order, send_counts = ops.bucket_by_owner(owner, world)    # stable: ascending item index inside a bucket
    sc, rc = _exchange_counts(send_counts, group)
    recv_keys = _all_to_all(keys[order], sc, rc, group)
    # arrivals are ordered (source rank, local index) == ascending global item index
    rk = ops.rank_among_equal(recv_keys, bits)
    back = _all_to_all(rk, rc, sc, group)                     # same split sizes, reversed roles
    out = torch.empty((codes.shape[0], codes.shape[1] + 1), dtype=torch.int64, device=codes.device)
    out[:, :-1] = codes
    out[order, -1] = back
    return out


def _segment_counts(mask: torch.Tensor, splits: List[int]) -> List[int]:
    """How many True entries of `mask` fall into each consecutive segment of the given lengths (one host read)."""
    cs = torch.cat([mask.new_zeros((1,), dtype=torch.int64), torch.cumsum(mask.to(torch.int64), 0)])
    bounds = torch.tensor([0] + list(_accumulate(splits)), dtype=torch.int64, device=mask.device)
    return (cs[bounds[1:]] - cs[bounds[:-1]]).cpu().tolist()


def _accumulate(xs):
    t = 0
    for v in xs:
        t += int(v)
        yield t


def generate_codes_sharded(model, data_shard, group=None, max_rounds: int = 30, ops=None) -> Tuple[torch.Tensor, dict]:
    """The whole encode driver (reference infer.py:88-177) over a catalogue sharded by contiguous item ranges:
    `data_shard` = this rank's rows; returns this rank's [n_local, L+1] semantic ids, identical to the rows a
    single-GPU `generate_codes` of the concatenated catalogue produces.

    A round of the re-encode loop (infer.py:116-129), sharded:
      1. every item's full code travels to owner = hash(key) mod G (all-to-all; arrival order = ascending global item
         index), the owner finds the groups and answers every item with the size of its group (1 = no collision);
      2. the HOME rank of a member — it holds the embedding — re-encodes it as the reference does: encoder and arg-min
         levels in the arithmetic of a batch of that size (csrc/small_batch.cu); the first L-1 codes are final;
      3. the members' residuals entering the last level travel to the owner (4e bytes per member), which runs the
         Sinkhorn pass over each group's distance matrix, and the new last-level codes travel home.
    Four small exchanges per round, all sized by the number of colliding items except the first; the loop ends as soon
    as no rank owns a group.  The suffix column is one more exchange (`global_suffix`)."""
    ops = ops if ops is not None else CudaShardOps(model)
    Ks = list(model.num_emb_list)
    Lv = len(Ks)
    codes = ops.encode_codes(data_shard)
    world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
    rounds = 0
    if world == 1:
        codes, rounds = ops.resolve_local(codes, data_shard, max_rounds)
        out = global_suffix(codes, Ks, ops, None)
        stats = global_stats(out, None)
        stats["rounds"] = rounds
        return out, stats
    if ops.last_level_sinkhorn():
        n = codes.shape[0]
        dev = codes.device
        while rounds < max_rounds:
            keys = ops.pack_keys(codes, Ks)
            owner = owner_of(keys, world)
            order, send_counts = ops.bucket_by_owner(owner, world)        # stable: ascending item index inside a bucket
            sc, rc = _exchange_counts(send_counts, group)
            recv_codes = _all_to_all(codes[order], sc, rc, group)         # arrival order = (source rank, local index)
            items, offsets, sizes = ops.groups(recv_codes)
            busy = torch.tensor([int(items.numel() > 0)], dtype=torch.int64, device=dev)
            dist.all_reduce(busy, op=dist.ReduceOp.MAX, group=group)
            if int(busy.item()) == 0:
                break
            size_back = _all_to_all(sizes, rc, sc, group)
            msize = torch.ones((n,), dtype=torch.int64, device=dev)
            msize[order] = size_back
            members = torch.nonzero(msize > 1).flatten()
            res_members = ops.reencode_members(data_shard, members, msize[members], codes)
            residual = torch.zeros((n, res_members.shape[1]), dtype=torch.float32, device=dev)
            residual[members] = res_members
            sel = msize[order] > 1                                        # members, in send order
            sc_m = _segment_counts(sel, sc)
            mem_recv = sizes > 1                                          # members, in arrival order
            rc_m = _segment_counts(mem_recv, rc)
            res_recv = _all_to_all(residual[order[sel]], sc_m, rc_m, group)
            res_rows = torch.zeros((recv_codes.shape[0], residual.shape[1]), dtype=torch.float32, device=dev)
            pos = torch.nonzero(mem_recv).flatten()
            res_rows[pos] = res_recv
            ops.sinkhorn_groups(recv_codes, items, offsets, res_rows)
            last_back = _all_to_all(recv_codes[pos, Lv - 1].contiguous(), rc_m, sc_m, group)
            codes[order[sel], Lv - 1] = last_back
            rounds += 1
    out = global_suffix(codes, Ks, ops, group)
    stats = global_stats(out, group)
    stats["rounds"] = rounds
    return out, stats


def global_stats(codes_with_suffix: torch.Tensor, group=None) -> dict:
    """Collision statistics of infer.py:132-137 from the suffix column: an item is 'distinct' iff its
    suffix is 0; the largest group is max suffix + 1."""
    suf = codes_with_suffix[:, -1]
    t = torch.stack([(suf == 0).sum(), torch.tensor(suf.numel(), device=suf.device)]).to(torch.int64)
    mx = suf.max().reshape(1) + 1 if suf.numel() else torch.zeros(1, dtype=torch.int64, device=suf.device)
    if group is not None and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    n = int(t[1].item())
    distinct = int(t[0].item())
    return {"distinct": distinct, "max_conflicts": int(mx.item()), "collision_rate": (n - distinct) / n if n else 0.0}
