"""Multi-GPU host logic: one process per GPU, item catalogue sharded by contiguous ranges.

The reference is single-process; this is new design for BASELINE configs 3-5:
  * encode shards independently (weights / codebooks replicated, no data-path collective);
  * global collision handling needs ONE exchange step: every item's packed code is routed to the rank
    that owns the key (all-to-all), the owner ranks equal keys in ascending GLOBAL item order, and the
    ranks travel back (all-to-all).  The result is bit-identical to the single-GPU suffix column;
  * the Sinkhorn re-encode rounds (reference infer.py:109-130) only ever touch the last level, so items are
    routed once by the hash of their PREFIX codes together with the residual entering the last level; the
    owner runs all rounds and the suffix ranking locally (`generate_codes_sharded`).
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
    def encode_shard(self, data):
        """Pass 1 on this rank's shard: (codes[n, L], residual entering the last level [n, e])."""
        from .generate_code import encode_codes_and_residual
        return encode_codes_and_residual(self.model, data)

    def resolve_owned(self, codes: torch.Tensor, residual: torch.Tensor, max_rounds: int):
        """Passes 2-3 over the items this rank owns (they arrive in ascending global item order)."""
        from .generate_code import resolve_rounds, suffix_dedup
        if codes.shape[0] == 0:
            return torch.empty((0, codes.shape[1] + 1), dtype=torch.int64, device=codes.device), 0
        codes, rounds = resolve_rounds(self.model, codes.contiguous(), residual.contiguous(), max_rounds=max_rounds)
        out, _ = suffix_dedup(self.model, codes)
        return out, rounds


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


def generate_codes_sharded(model, data_shard, group=None, max_rounds: int = 30, ops=None) -> Tuple[torch.Tensor, dict]:
    """The whole encode driver (reference infer.py:88-177) over a catalogue sharded by contiguous item ranges:
    `data_shard` = this rank's rows; returns this rank's [n_local, L+1] semantic ids, identical to the rows a
    single-GPU `generate_codes` of the concatenated catalogue produces.

    Levels < L-1 never change during the Sinkhorn rounds (infer.py:109-110), so every collision group — in any
    round — lives inside one PREFIX (first L-1 codes) class.  Items are therefore routed ONCE to
    owner = hash(prefix) mod G together with the residual entering the last level (all-to-all: 8(L+1) + 4e bytes
    per item); the owner holds them in ascending global item order, runs the ≤30 re-encode rounds and the suffix
    ranking locally with the single-GPU kernels, and the finished rows travel back (all-to-all, 8(L+1) bytes per
    item).  Two exchanges in total, none per round."""
    ops = ops if ops is not None else CudaShardOps(model)
    Ks = list(model.num_emb_list)
    Lv = len(Ks)
    codes, residual = ops.encode_shard(data_shard)
    world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
    if world == 1:
        out, rounds = ops.resolve_owned(codes, residual, max_rounds)
        stats = global_stats(out, None)
        stats["rounds"] = rounds
        return out, stats
    if Lv > 1:
        prefix = ops.pack_keys(codes[:, :Lv - 1].contiguous(), Ks[:Lv - 1])
    else:
        prefix = torch.zeros((codes.shape[0],), dtype=torch.int64, device=codes.device)   # one class: a single owner
    owner = owner_of(prefix, world)
    order, send_counts = ops.bucket_by_owner(owner, world)        # stable: ascending item index inside a bucket
    sc, rc = _exchange_counts(send_counts, group)
    recv_codes = _all_to_all(codes[order], sc, rc, group)         # arrival order = (source rank, local index)
    recv_res = _all_to_all(residual[order], sc, rc, group)        #               = ascending global item index
    owned, rounds = ops.resolve_owned(recv_codes, recv_res, max_rounds)
    back = _all_to_all(owned, rc, sc, group)
    out = torch.empty((codes.shape[0], Lv + 1), dtype=torch.int64, device=codes.device)
    out[order] = back
    r = torch.tensor([rounds], dtype=torch.int64, device=codes.device)
    dist.all_reduce(r, op=dist.ReduceOp.MAX, group=group)
    stats = global_stats(out, group)
    stats["rounds"] = int(r.item())
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
