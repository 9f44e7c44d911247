"""The encode driver — drop-in for reference RQ-VAE/infer.py:44-184 (≡ RQ-VAE/generate_code.py:44-178).

Same three passes, all on the device:
  pass 1  every item → codes with use_sk=False                      (infer.py:93-103)
  pass 2  ≤30 rounds: items sharing a full code are re-encoded group by group with Sinkhorn on the
          LAST level only (infer.py:109-130).  As in the reference, the whole model runs again on each
          group's rows ALONE — in the summation order the reference's CPU GEMM uses for a batch of that
          size (csrc/small_batch.cu), which differs from the catalogue pass for groups of 2..15 rows.
  pass 3  suffix column: out[i, L] = #{j < i : codes[j] == codes[i]} (infer.py:152-163), np.save of the
          [N, L+1] int64 array (infer.py:174-177) and the `_mapping.json` side file (infer.py:180-184).
The reference's per-item string building / dict grouping is replaced by integer sort kernels.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr
from .rqvae import RQVAE

MAX_LEVELS_OF_REFERENCE_DRIVER = 5     # prefix list ["<a_{}>",…,"<e_{}>"] has 5 entries (infer.py:90)


def _chunks(n: int, chunk_rows: int):
    """[lo, hi) row ranges of about chunk_rows rows.  Pass 1 treats the catalogue as ONE batch (the reference with a
    catalogue-sized batch); a tail of fewer than 16 rows is merged into the previous chunk, because a call with < 16 rows
    would be computed in the reference's small-batch summation order (csrc/small_batch.cu)."""
    bounds = list(range(0, n, chunk_rows)) + [n]
    if len(bounds) > 2 and bounds[-1] - bounds[-2] < 16:
        del bounds[-2]
    return list(zip(bounds[:-1], bounds[1:]))


def _as_rows(data):
    if isinstance(data, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    return data


@torch.no_grad()
def encode_latents(model: RQVAE, data, chunk_rows: int = 262144) -> torch.Tensor:
    """Encoder MLP over the whole catalogue → z[N, e] on the model's device.  `data` may be a CUDA tensor,
    a CPU tensor or a numpy array (host data is streamed chunk by chunk through pinned staging)."""
    data = _as_rows(data)
    dev = model._device()
    n = data.shape[0]
    z = torch.empty((n, model.e_dim), dtype=torch.float32, device=dev)
    model._sync()
    L = _cabi.lib()
    for r0, r1 in _chunks(n, chunk_rows):
        chunk = data[r0:r1]
        if not chunk.is_cuda:
            chunk = chunk.contiguous().to(dev, non_blocking=True)
        chunk = chunk.contiguous()
        check(L.rqb200_mlp_exact(model._handle, 0, ptr(chunk), 0, r1 - r0, ptr(z[r0:r1]), stream_ptr(dev)))
    return z


@torch.no_grad()
def collision_groups(model: RQVAE, codes: torch.Tensor):
    """get_collision_item (infer.py:29-42) on device → (items[n_items], offsets[n_groups+1], max_group)."""
    n, Lv = codes.shape
    items = torch.empty((max(n, 1),), dtype=torch.int64, device=codes.device)
    offsets = torch.empty((n + 1,), dtype=torch.int64, device=codes.device)
    ng, ni, mg = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    check(_cabi.lib().rqb200_collision_groups(model._handle, ptr(codes), n, Lv, _cabi.int_array(model.num_emb_list),
                                              ptr(items), ptr(offsets), ctypes.byref(ng), ctypes.byref(ni),
                                              ctypes.byref(mg), stream_ptr(codes.device)))
    return items[:ni.value], offsets[:ng.value + 1], int(mg.value)


@torch.no_grad()
def suffix_dedup(model: Optional[RQVAE], codes: torch.Tensor, num_emb_list=None, want_stats: bool = True
                 ) -> Tuple[torch.Tensor, Optional[dict]]:
    """Suffix column (infer.py:152-163) → ([N, L+1] int64, stats).

    want_stats=False (and the code ranges known, i.e. a model or num_emb_list given): the call only enqueues — the collision
    statistics are the one thing it would otherwise wait for the device to read back — and returns (ids, None)."""
    if not codes.is_cuda:
        raise RuntimeError("suffix_dedup: CUDA tensor required (no CPU fallback)")
    codes = codes.contiguous()
    n, Lv = codes.shape
    out = torch.empty((n, Lv + 1), dtype=torch.int64, device=codes.device)
    nd, mg = ctypes.c_int64(0), ctypes.c_int64(0)
    Ks = num_emb_list if num_emb_list is not None else (model.num_emb_list if model is not None else None)
    handle = model._handle if model is not None else _scratch_handle(codes.device)
    check(_cabi.lib().rqb200_suffix_dedup(handle, ptr(codes), n, Lv, _cabi.int_array(Ks) if Ks else None, ptr(out),
                                          ctypes.byref(nd) if want_stats else None, ctypes.byref(mg) if want_stats else None,
                                          stream_ptr(codes.device)))
    if not want_stats:
        return out, None
    return out, {"distinct": int(nd.value), "max_conflicts": int(mg.value),
                 "collision_rate": (n - int(nd.value)) / n if n else 0.0}


_SCRATCH = {}


def _scratch_handle(device):
    """A minimal model handle that only owns sort workspace (for suffix_dedup without an RQVAE)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SCRATCH:
        h = ctypes.c_void_p(None)
        check(_cabi.lib().rqb200_model_create(ctypes.byref(h), idx, 1, _cabi.int_array([8, 8]), 1, _cabi.int_array([2])))
        _SCRATCH[idx] = h
    return _SCRATCH[idx]


@torch.no_grad()
def encode_codes_fast(model: RQVAE, data, chunk_rows: int = 262144) -> torch.Tensor:
    """Pass 1 on the tensor-core route: codes[N, L] (bit-identical to the exact route)."""
    data = _as_rows(data)
    dev = model._device()
    n, Lv = data.shape[0], len(model.num_emb_list)
    codes = torch.empty((n, Lv), dtype=torch.int64, device=dev)
    model._sync()
    L = _cabi.lib()
    for r0, r1 in _chunks(n, chunk_rows):
        chunk = data[r0:r1]
        if not chunk.is_cuda:
            chunk = chunk.contiguous().to(dev, non_blocking=True)
        chunk = chunk.contiguous()
        check(L.rqb200_get_indices(model._handle, _cabi.ENCODE_FAST, ptr(chunk), r1 - r0, ptr(codes[r0:r1]), 0, None,
                                   stream_ptr(dev)))
    return codes


class _ReencodeMemo:
    """Device memo of the per-item part of the group re-encode (csrc/small_batch.cu, rqb200_reencode_groups_memo): the
    codes of the first L-1 levels and the residual entering the last level are a pure function of (item, size class of its
    group), so rounds after the first answer almost every member from here instead of running the model again."""

    def __init__(self, model: RQVAE, n: int, device):
        self.classes = int(_cabi.lib().rqb200_reencode_classes(model._handle))
        Lv = len(model.num_emb_list)
        self.have = torch.zeros((max(n, 1),), dtype=torch.int32, device=device)
        self.res = torch.empty((self.classes, max(n, 1), model.e_dim), dtype=torch.float32, device=device)
        self.codes = torch.empty((self.classes, max(n, 1), max(Lv - 1, 1)), dtype=torch.int32, device=device)

    @staticmethod
    def try_create(model: RQVAE, n: int, device, data):
        """None when the catalogue is not on the device or the memo does not fit next to it (the rounds then recompute)."""
        if not (torch.is_tensor(data) and data.is_cuda):
            return None
        classes = int(_cabi.lib().rqb200_reencode_classes(model._handle))
        need = classes * n * 4 * (model.e_dim + len(model.num_emb_list)) + 4 * n
        free, _total = torch.cuda.mem_get_info(device)
        if need > 0.6 * free:
            return None
        try:
            return _ReencodeMemo(model, n, device)
        except torch.cuda.OutOfMemoryError:
            return None


class _GroupRecords:
    """Per-item record of the group an item was last re-encoded in (csrc/dedup.cu: group_unchanged_kernel): lets the
    rounds after the first skip every group whose member set did not change — re-encoding it is a no-op by construction."""

    def __init__(self, n: int, device):
        self.first = torch.full((max(n, 1),), -1, dtype=torch.int64, device=device)
        self.meta = torch.full((max(n, 1),), -1, dtype=torch.int64, device=device)
        self.round = 0


@torch.no_grad()
def reencode_round(model: RQVAE, codes: torch.Tensor, data, residual: Optional[torch.Tensor] = None,
                   verbose_round: Optional[int] = None, records: Optional[_GroupRecords] = None,
                   memo: Optional[_ReencodeMemo] = None) -> Tuple[int, int]:
    """ONE round of the loop at infer.py:116-129, in place on `codes`: every group of items sharing a full code goes
    through `model.get_indices(data[group], use_sk=True)` — as in the reference the WHOLE model runs again on the group's
    rows alone (encoder, arg-min levels, Sinkhorn on the last level), in the arithmetic the reference uses for a batch of
    that size (csrc/small_batch.cu), and all L codes of the members are overwritten.  Groups of one round are disjoint
    and are all found before anything is rewritten, so the round is a pure function of the codes it starts from.
    `records` (rounds of one driver run): groups that are member-for-member a group of the previous round are skipped —
    they are fixed points.  `memo` (same): members already re-encoded in a group of the same size class are answered
    from it.  Returns (groups found, groups re-encoded); (0, 0): nothing collides."""
    lib = _cabi.lib()
    dev = codes.device
    data = _as_rows(data)
    last = model.rq.vq_layers[-1]
    if last.sk_epsilon is None or last.sk_epsilon <= 0:
        return 0, 0
    model._sync()
    n, Lv = codes.shape
    if records is None:
        items, offsets, max_group = collision_groups(model, codes)
        n_total = offsets.numel() - 1
    else:
        items = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
        offsets = torch.empty((n + 1,), dtype=torch.int64, device=dev)
        ng, ni, mg, nt = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib.rqb200_collision_groups_changed(model._handle, ptr(codes), n, Lv, _cabi.int_array(model.num_emb_list),
                                                  ptr(records.first), ptr(records.meta), records.round, ptr(items),
                                                  ptr(offsets), ctypes.byref(ng), ctypes.byref(ni), ctypes.byref(mg),
                                                  ctypes.byref(nt), stream_ptr(dev)))
        records.round += 1
        items, offsets, max_group, n_total = items[:ni.value], offsets[:ng.value + 1], int(mg.value), int(nt.value)
    n_groups = offsets.numel() - 1
    if verbose_round is not None and n_total > 0:
        print(f"Iteration {verbose_round}: Found {n_total} collision groups")
    if n_groups <= 0:
        return n_total, 0
    if residual is None:
        residual = torch.empty((n, model.e_dim), dtype=torch.float32, device=dev)
    if data.is_cuda:
        x, gathered = data, 0
    else:
        x, gathered = data[items.cpu()].contiguous().to(dev), 1
    cap = lib.rqb200_sinkhorn_group_cap(model._handle)
    if memo is not None and not gathered:
        check(lib.rqb200_reencode_groups_memo(model._handle, ptr(x), ptr(items), ptr(offsets), n_groups, items.numel(),
                                              ptr(codes), ptr(residual), n, ptr(memo.have), ptr(memo.res), ptr(memo.codes),
                                              stream_ptr(dev)))
    else:
        check(lib.rqb200_reencode_groups(model._handle, ptr(x), gathered, ptr(items), ptr(offsets), n_groups, items.numel(),
                                         ptr(codes), ptr(residual), stream_ptr(dev)))
    check(lib.rqb200_sinkhorn_regroup(model._handle, ptr(residual), ptr(items), ptr(offsets), n_groups,
                                      min(max_group, cap), float(last.sk_epsilon), int(last.sk_iters), ptr(codes),
                                      stream_ptr(dev)))
    if max_group > cap:
        _regroup_oversized(model, residual, items, offsets, cap, codes)
    return n_total, n_groups


@torch.no_grad()
def resolve_rounds(model: RQVAE, codes: torch.Tensor, data, max_rounds: int = 30, verbose: bool = False,
                   stats: Optional[dict] = None) -> Tuple[torch.Tensor, int]:
    """Pass 2 (infer.py:109-130): ≤ max_rounds rounds of `reencode_round`, until no two items share a full code.
    `codes` is updated in place.  `data`: the catalogue rows — a CUDA tensor (members gathered on the device) or a host
    tensor / array (members gathered on the host and uploaded per round).  Once every remaining group is a fixed point
    the remaining rounds of the reference are no-ops; they are counted, not run.  Returns (codes, rounds)."""
    # infer.py:109-110 — only the last level keeps its Sinkhorn epsilon
    for vq in model.rq.vq_layers[:-1]:
        vq.sk_epsilon = 0.0
    last = model.rq.vq_layers[-1]
    rounds = 0
    work = []
    if last.sk_epsilon is not None and last.sk_epsilon > 0:
        residual = torch.empty((codes.shape[0], model.e_dim), dtype=torch.float32, device=codes.device)
        records = _GroupRecords(codes.shape[0], codes.device)
        data = _as_rows(data)
        memo = _ReencodeMemo.try_create(model, codes.shape[0], codes.device, data)
        while rounds < max_rounds:
            found, done = reencode_round(model, codes, data, residual, rounds if verbose else None, records, memo)
            if found == 0:
                break
            work.append(done)
            rounds += 1
            if done == 0:                      # only fixed points are left: the reference spins until tt == 30 without effect
                if verbose:
                    for t in range(rounds, max_rounds):
                        print(f"Iteration {t}: Found {found} collision groups")
                rounds = max_rounds
                break
    if stats is not None:
        stats["groups_reencoded_per_round"] = work
    return codes, rounds


@torch.no_grad()
def encode_codes_exact(model: RQVAE, data, chunk_rows: int = 262144) -> torch.Tensor:
    """Pass 1 on the exact route: codes[N, L]."""
    dev = model._device()
    z = encode_latents(model, data, chunk_rows)
    n, Lv = z.shape[0], len(model.num_emb_list)
    codes = torch.empty((n, Lv), dtype=torch.int64, device=dev)
    model._sync()
    check(_cabi.lib().rqb200_quantize(model._handle, ptr(z), n, ptr(codes), 0, 0, 0, 0, stream_ptr(dev)))
    return codes


@torch.no_grad()
def generate_codes(model: RQVAE, data, max_rounds: int = 30, chunk_rows: int = 262144, verbose: bool = False,
                   fast: Optional[bool] = None, batch_size: Optional[int] = None) -> Tuple[torch.Tensor, dict]:
    """Passes 1–3 for one catalogue on one GPU.  Returns ([N, L+1] int64 CUDA tensor, stats).

    fast (default: whenever the model's shapes allow it): pass 1 runs on the tensor-core route (every row certified by
    the margin gate or recomputed by the exact kernels); the re-encode rounds always run on the exact kernels.
    batch_size (the reference's `params["batch_size"]`, infer.py:85): the reference encodes the catalogue in DataLoader
    batches; full batches of 16 or more rows all share the catalogue arithmetic, but a LAST batch of 2..15 rows is
    computed in its small-batch order (csrc/small_batch.cu) — given batch_size, those rows are encoded that way too.
    (A last batch of exactly one row takes the reference's matrix-vector kernel, which is not restated.)"""
    Lv = len(model.num_emb_list)
    if Lv > MAX_LEVELS_OF_REFERENCE_DRIVER:
        raise IndexError("list index out of range")        # what prefix[i] raises in the reference
    if fast is None:
        fast = model.fast_route_supported()
    was_training = model.training
    model.eval()
    try:
        codes = encode_codes_fast(model, data, chunk_rows) if fast else encode_codes_exact(model, data, chunk_rows)
        tail = codes.shape[0] % int(batch_size) if batch_size else 0
        if 2 <= tail <= 15 and codes.shape[0] > tail:
            rows = _as_rows(data)[codes.shape[0] - tail:]
            rows = (rows if rows.is_cuda else rows.contiguous().to(model._device())).contiguous()
            model._sync()
            check(_cabi.lib().rqb200_get_indices(model._handle, _cabi.ENCODE_EXACT, ptr(rows), tail,
                                                 ptr(codes[codes.shape[0] - tail:]), 0, None, stream_ptr(model._device())))
        rstats = {}
        codes, rounds = resolve_rounds(model, codes, data, max_rounds=max_rounds, verbose=verbose, stats=rstats)
        out, stats = suffix_dedup(model, codes)
        stats.update(rstats)
        stats["rounds"] = rounds
        stats["pass1_route"] = "tensor-core" if fast else "exact"
        return out, stats
    finally:
        if was_training:
            model.train()


def _regroup_oversized(model, residual, items, offsets, cap, codes, scratch_doubles: int = 1 << 27):
    """Groups too large for the shared-memory kernel: one batched launch per wave of groups whose fp64 matrices fit in
    a bounded global scratch (1 GiB by default); a single group beyond that budget gets a scratch of its own."""
    last = model.rq.vq_layers[-1]
    K = model.num_emb_list[-1]
    dev = codes.device
    sizes = (offsets[1:] - offsets[:-1])
    big = torch.nonzero(sizes > cap).flatten()
    if big.numel() == 0:
        return
    big_sizes = sizes[big].cpu().tolist()
    big_ids = big.cpu().tolist()
    lib = _cabi.lib()
    wave_ids, wave_off, used = [], [], 0

    def flush():
        nonlocal wave_ids, wave_off, used
        if not wave_ids:
            return
        scratch = torch.empty((used,), dtype=torch.float64, device=dev)
        gid = torch.tensor(wave_ids, dtype=torch.int64, device=dev)
        off = torch.tensor(wave_off, dtype=torch.int64, device=dev)
        check(lib.rqb200_sinkhorn_regroup_large(model._handle, ptr(residual), ptr(items), ptr(offsets), ptr(gid), ptr(off),
                                                len(wave_ids), ptr(scratch), float(last.sk_epsilon), int(last.sk_iters),
                                                ptr(codes), stream_ptr(dev)))
        wave_ids, wave_off, used = [], [], 0

    for g, b in zip(big_ids, big_sizes):
        need = b * K + (b + 1) // 2 + 1
        if used and used + need > scratch_doubles:
            flush()
        wave_ids.append(g)
        wave_off.append(used)
        used += need
    flush()


def infer(params):
    """Same contract as reference infer.py:44-184: params dict keys of main.py:6-36, writes
    `semantic_id_file` (np.save [N, L+1] int64) and `<…>_mapping.json`."""
    from .dataset import EmbDataset
    h5_path = params["data_path"]
    ckpt_path = os.path.join(params["ckpt_dir"], "best_collision_model.pth")
    output_file = params["semantic_id_file"]
    device = params["device"]
    data = EmbDataset(h5_path)
    model = RQVAE(in_dim=data.dim, num_emb_list=params["num_emb_list"], e_dim=params["e_dim"],
                  layers=params["layers"], dropout_prob=params["dropout"], bn=params["batch_normalize"],
                  loss_type=params["loss_type"], quant_loss_weight=params["quant_loss_weight"],
                  kmeans_init=params["kmeans_init"], kmeans_iters=params["kmeans_iters"],
                  sk_epsilons=params["sk_epsilons"], sk_iters=params["sk_iters"])
    if os.path.exists(ckpt_path):
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        if "state_dict" in ckpt:
            model.load_state_dict(ckpt["state_dict"])
            print(f"Loaded checkpoint from {ckpt_path}")
        else:
            model.load_state_dict(ckpt)
            print(f"Loaded state dict from {ckpt_path}")
    else:
        print(f"Warning: No checkpoint found at {ckpt_path}, using randomly initialized model")
    model = model.to(device)
    model.eval()
    print("Generating codes...")
    codes, stats = generate_codes(model, data.embeddings, verbose=True, batch_size=params.get("batch_size"))
    codes_array = codes.cpu().numpy()
    print("All indices number: ", len(codes_array))
    print("Max number of conflicts: ", stats["max_conflicts"])
    print("Collision Rate", stats["collision_rate"])
    if len(np.unique(codes_array, axis=0)) != len(codes_array):     # infer.py:165-171 re-check
        print("There still have duplicates")
    else:
        print("There are no duplicates in the codes after resolution.")
    os.makedirs(os.path.dirname(output_file) or ".", exist_ok=True)
    print(f"Saving codes to {output_file}")
    print(f"the first 5 codes: {codes_array[:5]}")
    np.save(output_file, codes_array)
    mapping_file = output_file.replace(".npy", "_mapping.json")
    with open(mapping_file, "w") as f:
        json.dump({i: code.tolist() for i, code in enumerate(codes_array)}, f, indent=2)
    print(f"Saved index-to-code mapping to {mapping_file}")
    return codes_array
