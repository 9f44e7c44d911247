"""CPU oracle for the RQ-VAE semantic-ID encode path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker.  The product package never
imports it.

Parity status: PINNED — checked bit-for-bit against the reference's PyTorch CPU modules
(oracle/make_golden.py, run in the build container where /root/reference exists) and
against the committed golden vectors in tests/golden/.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librqvae_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/rqvae_oracle.c with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "rqvae_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librqvae_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        c_f = ctypes.POINTER(ctypes.c_float)
        c_i = ctypes.POINTER(ctypes.c_int)
        c_l = ctypes.POINTER(ctypes.c_int64)
        c_d = ctypes.POINTER(ctypes.c_double)
        L.rq_oracle_linear.argtypes = [c_f, ctypes.c_int64, ctypes.c_int, c_f, c_f, ctypes.c_int,
                                       ctypes.c_int, c_i, ctypes.c_int, c_f, ctypes.c_int]
        L.rq_oracle_linear.restype = ctypes.c_int
        L.rq_oracle_sumsq.argtypes = [c_f, ctypes.c_int64, ctypes.c_int, c_f]
        L.rq_oracle_sumsq.restype = None
        L.rq_oracle_quantize.argtypes = [c_f, ctypes.c_int64, ctypes.c_int, c_f, c_i, ctypes.c_int,
                                         c_l, c_f, c_d, c_f, ctypes.c_int, ctypes.c_int]
        L.rq_oracle_quantize.restype = ctypes.c_int
        L.rq_oracle_quantize_ex.argtypes = L.rq_oracle_quantize.argtypes + [ctypes.c_int]
        L.rq_oracle_quantize_ex.restype = ctypes.c_int
        L.rq_oracle_linear_small.argtypes = [c_f, ctypes.c_int64, ctypes.c_int, c_f, c_f, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, c_f]
        L.rq_oracle_linear_small.restype = ctypes.c_int
        L.rq_oracle_suffix.argtypes = [c_l, ctypes.c_int64, ctypes.c_int, c_l]
        L.rq_oracle_suffix.restype = ctypes.c_int
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _threads(threads: Optional[int]) -> int:
    return int(threads) if threads else (os.cpu_count() or 1)


# --------------------------------------------------------------------------- Linear / MLP

def mkl_kblocks(in_dim: int, out_dim: int = 0) -> List[int]:
    """K-blocking of the reference's CPU GEMM for ``nn.Linear`` (layers.py:23) as observed on
    the reference run: one chain if K ≤ 384, two equal halves for
    384 < K ≤ 768, otherwise blocks of 384 with the remainder last."""
    K = int(in_dim)
    if K <= 384:
        return [K]
    if K <= 768:
        return [K - K // 2, K // 2] if K % 2 else [K // 2, K // 2]
    out = []
    while K > 0:
        out.append(min(384, K))
        K -= out[-1]
    return out


def linear(x: np.ndarray, W: np.ndarray, b: Optional[np.ndarray], relu: bool,
           kblocks: Optional[Sequence[int]] = None, threads: Optional[int] = None) -> np.ndarray:
    """``nn.Linear`` (+ReLU) exactly as the reference computes it on CPU (layers.py:23,28-30)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    n, k = x.shape
    o = W.shape[0]
    assert W.shape[1] == k
    kb = np.asarray(kblocks if kblocks is not None else mkl_kblocks(k, o), dtype=np.int32)
    y = np.empty((n, o), dtype=np.float32)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    rc = lib().rq_oracle_linear(_fp(x), n, k, _fp(W), _fp(bb) if bb is not None else None, o,
                                int(bool(relu)), kb.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                len(kb), _fp(y), _threads(threads))
    if rc != 0:
        raise ValueError(f"rq_oracle_linear failed rc={rc}")
    return y


# Small batches.  The reference re-runs the whole model on every collision group (infer.py:120-122), i.e. with
# 2 … a few dozen rows, and its CPU GEMM switches kernels there.  Measured on the live reference stack with
# oracle/probe_sum_order.py (torch 2.11 + MKL; 1 and 8 threads agree): an [M,K]·[K,N] product uses
#   2 <= M <= 15 and 24·M <= K  → "lane16": 16 interleaved fma chains folded ((p0+p1)+p2)+p3, (s0+s1)+(s2+s3), bias last
#   otherwise                    → the catalogue order (mkl_kblocks).
# Not restated (see the probe's docstring): 1024 → 256 with 16 ≤ M < 176, where the reference's own result depends on
# its thread count, and N ≤ 128 with K in {384, 640}.
LANE16, PAIR4 = 1, 2


def small_batch_plan(M: int, K: int, N: int) -> int:
    """0 = catalogue order, LANE16 — the order the reference uses for an [M,K]·[K,N] product."""
    if 2 <= M <= 15 and 24 * M <= K:
        if N <= 128 and K in (384, 640):
            raise KeyError(f"[{M},{K}]·[{K},{N}]: an exception of the reference's GEMM dispatch that is not restated")
        return LANE16
    if K == 1024 and N == 256 and 16 <= M < 176:
        raise KeyError("1024 -> 256 with 16..175 rows: the reference's result depends on its thread count; not restated")
    return 0


def linear_group(x: np.ndarray, W: np.ndarray, b: Optional[np.ndarray], relu: bool,
                 batch_size: Optional[int] = None) -> np.ndarray:
    """``nn.Linear`` (+ReLU) on ONE small batch, in the order the reference's CPU GEMM uses for that batch size
    (batch_size: the rows given are only SOME rows of a batch of that size — every row is computed independently)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    n, k = x.shape
    kind = small_batch_plan(n if batch_size is None else int(batch_size), k, W.shape[0])
    if kind == 0:
        return linear(x, W, b, relu, threads=1)
    y = np.empty((n, W.shape[0]), dtype=np.float32)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    rc = lib().rq_oracle_linear_small(_fp(x), n, k, _fp(W), _fp(bb) if bb is not None else None, W.shape[0],
                                      int(bool(relu)), kind, _fp(y))
    if rc != 0:
        raise ValueError(f"rq_oracle_linear_small failed rc={rc}")
    return y


def mlp_group(x: np.ndarray, weights, biases, batch_size: Optional[int] = None) -> np.ndarray:
    """``MLPLayers.forward`` on one small batch (a collision group), layer by layer in the batch-size-dependent order."""
    h = x
    for i, (W, b) in enumerate(zip(weights, biases)):
        h = linear_group(h, W, b, relu=(i != len(weights) - 1), batch_size=batch_size)
    return h


def reencode_prefix(xg: np.ndarray, enc_w, enc_b, codebooks, batch_size: Optional[int] = None):
    """The part of ``get_indices(data[group], use_sk=True)`` before the last level, for rows of a batch of
    ``batch_size`` rows (default: the rows given ARE the batch): (codes[m, L-1], residual entering the last level)."""
    M = xg.shape[0] if batch_size is None else int(batch_size)
    r = mlp_group(np.ascontiguousarray(xg, dtype=np.float32), enc_w, enc_b, batch_size=M).copy()
    out = []
    for cb in codebooks[:-1]:
        cb = np.ascontiguousarray(cb, dtype=np.float32)
        idx = quantize(r, [cb], want_xq=False, threads=1, dot_kind=small_batch_plan(M, r.shape[1], cb.shape[0]))[0][:, 0]
        q = cb[idx]
        xres = r + (q - r)
        r = r - xres
        out.append(idx)
    codes = np.stack(out, axis=-1).astype(np.int64) if out else np.zeros((xg.shape[0], 0), dtype=np.int64)
    return codes, r


def sinkhorn_last_level(residual: np.ndarray, cb_last: np.ndarray, eps: float, iters: int) -> np.ndarray:
    """Last level of the re-encode of ONE group: distances of the group's residual rows (in the order the reference's
    matmul uses for that many rows), centred over the group, fp64 Sinkhorn, arg-max (vq.py:71-83)."""
    cb = np.ascontiguousarray(cb_last, dtype=np.float32)
    r = np.ascontiguousarray(residual, dtype=np.float32)
    d = quantize(r, [cb], want_xq=False, dist_level=0, threads=1, dot_kind=small_batch_plan(r.shape[0], r.shape[1], cb.shape[0]))[3]
    return sinkhorn_assign(d, eps, iters)


def mlp(x: np.ndarray, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray],
        kblocks: Optional[Dict[int, Sequence[int]]] = None, threads: Optional[int] = None) -> np.ndarray:
    """``MLPLayers.forward`` in eval mode (layers.py:18-32,42-43): Dropout is the identity, ReLU
    after every Linear but the last, no BatchNorm (fold it into W/b first if present)."""
    h = x
    for i, (W, b) in enumerate(zip(weights, biases)):
        kb = None if kblocks is None else kblocks.get(i)
        h = linear(h, W, b, relu=(i != len(weights) - 1), kblocks=kb, threads=threads)
    return h


# --------------------------------------------------------------------------- quantizer

def sumsq(v: np.ndarray) -> np.ndarray:
    """``torch.sum(v**2, dim=1)`` (vq.py:71-72) with ATen's fp32 summation order."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.empty(v.shape[0], dtype=np.float32)
    lib().rq_oracle_sumsq(_fp(v), v.shape[0], v.shape[1], _fp(out))
    return out


def quantize(z: np.ndarray, codebooks: Sequence[np.ndarray], want_xq: bool = True,
             dist_level: int = -1, threads: Optional[int] = None, dot_kind: int = 0
             ) -> Tuple[np.ndarray, Optional[np.ndarray], np.ndarray, Optional[np.ndarray]]:
    """``ResidualVectorQuantizer.forward`` with ``use_sk=False`` (rq.py:39-56, vq.py:63-99).

    Returns (indices[n,L] int64, x_q[n,e] f32 | None, sum_sq[L] f64, distances | None) where
    sum_sq[l] = Σ (q - r)² over the batch at level l (the mse numerator of vq.py:90-91)."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    n, e = z.shape
    L = len(codebooks)
    K = np.asarray([c.shape[0] for c in codebooks], dtype=np.int32)
    cat = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32) for c in codebooks], 0))
    idx = np.empty((n, L), dtype=np.int64)
    xq = np.empty((n, e), dtype=np.float32) if want_xq else None
    loss = np.zeros(L, dtype=np.float64)
    dist = np.empty((n, int(K[dist_level])), dtype=np.float32) if dist_level >= 0 else None
    rc = lib().rq_oracle_quantize_ex(
        _fp(z), n, e, _fp(cat), K.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), L,
        idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        _fp(xq) if xq is not None else None,
        loss.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        _fp(dist) if dist is not None else None, dist_level, _threads(threads), int(dot_kind))
    if rc != 0:
        raise ValueError(f"rq_oracle_quantize failed rc={rc}")
    return idx, xq, loss, dist


def rq_loss(sum_sq: np.ndarray, n: int, e: int, beta: float) -> float:
    """mean over levels of ``codebook_loss + beta*commitment_loss`` (vq.py:90-92, rq.py:53)."""
    mse = sum_sq / float(n * e)
    return float(np.mean(mse + beta * mse))


def get_indices(x: np.ndarray, enc_w, enc_b, codebooks, threads: Optional[int] = None) -> np.ndarray:
    """``RQVAE.get_indices(xs, use_sk=False)`` (rqvae.py:67-71)."""
    z = mlp(x, enc_w, enc_b, threads=threads)
    return quantize(z, codebooks, want_xq=False, threads=threads)[0]


def forward(x, enc_w, enc_b, codebooks, dec_w, dec_b, beta=0.25, threads=None):
    """``RQVAE.forward(x, use_sk=False)`` + ``compute_loss`` pieces (rqvae.py:60-65,73-84).
    Returns (out, rq_loss, indices, recon_mse)."""
    z = mlp(x, enc_w, enc_b, threads=threads)
    idx, xq, ssq, _ = quantize(z, codebooks, threads=threads)
    out = mlp(xq, dec_w, dec_b, threads=threads)
    recon = float(np.mean((out.astype(np.float64) - np.asarray(x, dtype=np.float64)) ** 2))
    return out, rq_loss(ssq, x.shape[0], z.shape[1], beta), idx, recon


# --------------------------------------------------------------------------- Sinkhorn

def center_distance(d: np.ndarray) -> np.ndarray:
    """``VectorQuantizer.center_distance_for_constraint`` (vq.py:51-61) in fp32."""
    d = np.asarray(d, dtype=np.float32)
    mx, mn = d.max(), d.min()
    middle = np.float32(mx + mn) / np.float32(2)
    amplitude = np.float32(np.float32(mx - middle) + np.float32(1e-5))
    assert amplitude > 0
    return ((d - middle) / amplitude).astype(np.float32)


def sinkhorn(d64: np.ndarray, epsilon: float, iters: int) -> np.ndarray:
    """``sinkhorn_algorithm`` (layers.py:85-108) in fp64."""
    Q = np.exp(-np.asarray(d64, dtype=np.float64) / epsilon)
    B, K = Q.shape
    Q /= Q.sum(-1, keepdims=True).sum(-2, keepdims=True)
    for _ in range(iters):
        Q /= Q.sum(axis=1, keepdims=True)
        Q /= B
        Q /= Q.sum(axis=0, keepdims=True)
        Q /= K
    Q *= B
    return Q


def sinkhorn_assign(d: np.ndarray, epsilon: float, iters: int) -> np.ndarray:
    """The ``use_sk`` branch of ``VectorQuantizer.forward`` (vq.py:74-83): centre, fp64 Sinkhorn, argmax."""
    Q = sinkhorn(center_distance(d).astype(np.float64), epsilon, iters)
    return np.argmax(Q, axis=-1).astype(np.int64)


def top2_gap(Q: np.ndarray) -> np.ndarray:
    """Relative gap between the two largest entries of every row of Q — 0 where the arg-max is an exact tie."""
    if Q.shape[1] < 2:
        return np.ones(Q.shape[0])
    top = np.partition(Q, Q.shape[1] - 2, axis=1)[:, -2:]
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(top[:, 1] > 0, (top[:, 1] - top[:, 0]) / top[:, 1], 0.0)


def quantize_sk(z: np.ndarray, codebooks, sk_epsilons, sk_iters: int, group_order: bool = False,
                gaps: Optional[list] = None) -> np.ndarray:
    """``ResidualVectorQuantizer.forward(use_sk=True)`` indices for one batch (rq.py:39-56, vq.py:63-99):
    levels with ε>0 take the Sinkhorn argmax, the rest the plain argmin.  group_order: the batch is one small
    collision group — ``matmul(latent, E.t())`` (vq.py:73) then runs in the reference's small-batch order."""
    r = np.ascontiguousarray(z, dtype=np.float32).copy()
    out = []
    for cb, eps in zip(codebooks, sk_epsilons):
        cb = np.ascontiguousarray(cb, dtype=np.float32)
        dk = small_batch_plan(r.shape[0], r.shape[1], cb.shape[0]) if group_order else 0
        idx, _, _, dist = quantize(r, [cb], want_xq=False, dist_level=0, threads=1, dot_kind=dk)
        ind = idx[:, 0]
        if eps > 0:
            Q = sinkhorn(center_distance(dist).astype(np.float64), eps, sk_iters)
            ind = np.argmax(Q, axis=-1).astype(np.int64)
            if gaps is not None:
                gaps.append(top2_gap(Q))
        q = cb[ind]
        xres = r + (q - r)
        r = r - xres
        out.append(ind)
    return np.stack(out, axis=-1).astype(np.int64)


# --------------------------------------------------------------------------- collisions / suffix

def suffix_dedup(codes: np.ndarray) -> np.ndarray:
    """Suffix column of infer.py:152-163: ``suffix[i] = #{j < i : codes[j] == codes[i]}``; returns
    ``[N, L+1]`` int64 like the array the reference saves (infer.py:177)."""
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    n, L = codes.shape
    out = np.empty((n, L + 1), dtype=np.int64)
    rc = lib().rq_oracle_suffix(codes.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), n, L,
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc != 0:
        raise MemoryError("rq_oracle_suffix")
    return out


def collision_groups(codes: np.ndarray) -> List[np.ndarray]:
    """``get_collision_item`` (infer.py:29-42): item-index groups sharing a full code, each in
    ascending item order, groups in order of first occurrence."""
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    _, inv, counts = np.unique(codes, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    order = np.argsort(inv, kind="stable")
    bounds = np.concatenate([[0], np.cumsum(counts)])
    groups = [order[bounds[g]:bounds[g + 1]] for g in range(len(counts)) if counts[g] > 1]
    groups.sort(key=lambda g: int(g[0]))
    return groups


def reencode_group(xg: np.ndarray, enc_w, enc_b, codebooks, eps, sk_iters: int, gaps: Optional[list] = None) -> np.ndarray:
    """``model.get_indices(data[collision_items], use_sk=True)`` (infer.py:120-122) on ONE collision group, exactly as
    the reference computes it: the encoder and every level run again on the group's rows alone, in the summation
    order its CPU GEMM uses for that batch size (small_batch_plan)."""
    zg = mlp_group(np.ascontiguousarray(xg, dtype=np.float32), enc_w, enc_b)
    return quantize_sk(zg, codebooks, eps, sk_iters, group_order=True, gaps=gaps)


def generate_codes(x: np.ndarray, enc_w, enc_b, codebooks, sk_epsilons, sk_iters: int,
                   max_rounds: int = 30, threads: Optional[int] = None, group_order: bool = False,
                   trace: Optional[list] = None, batch_size: Optional[int] = None) -> Tuple[np.ndarray, dict]:
    """The encode driver of infer.py:88-177 / generate_code.py:82-178: pass 1 (argmin codes),
    ≤30 rounds of per-group re-encoding with Sinkhorn on the last level only (infer.py:109-130),
    then the suffix column.  Returns ([N, L+1] int64, stats).

    group_order=False: every group is re-quantized from the catalogue-pass latents (a pure function of the item);
    group_order=True: the reference's literal computation — each group goes through the encoder again as its own
    small batch (reencode_group), which can differ from the former in the last bit of z."""
    L = len(codebooks)
    if L > 5:
        raise IndexError("list index out of range")      # prefix list has 5 entries (infer.py:90)
    z = mlp(x, enc_w, enc_b, threads=threads)
    codes = quantize(z, codebooks, want_xq=False, threads=threads)[0]
    tail = len(x) % int(batch_size) if batch_size else 0
    if 2 <= tail <= 15 and len(x) > tail:        # the reference's last DataLoader batch (infer.py:85,93-96) is a small batch
        zt = mlp_group(np.ascontiguousarray(x[len(x) - tail:], dtype=np.float32), enc_w, enc_b)
        r = zt.copy()
        for l, cb in enumerate(codebooks):
            cb = np.ascontiguousarray(cb, dtype=np.float32)
            idx = quantize(r, [cb], want_xq=False, threads=1, dot_kind=small_batch_plan(tail, r.shape[1], cb.shape[0]))[0][:, 0]
            codes[len(x) - tail:, l] = idx
            q = cb[idx]
            r = r - (r + (q - r))
    eps = [0.0] * (L - 1) + [float(sk_epsilons[-1])]
    rounds = 0
    if trace is not None:
        trace.append(codes.copy())
    while rounds < max_rounds:
        groups = collision_groups(codes)
        if not groups:
            break
        new = codes.copy()
        for g in groups:
            new[g] = (reencode_group(x[g], enc_w, enc_b, codebooks, eps, sk_iters) if group_order
                      else quantize_sk(z[g], codebooks, eps, sk_iters))
        codes = new
        rounds += 1
        if trace is not None:
            trace.append(codes.copy())
    uniq, counts = np.unique(codes, axis=0, return_counts=True)
    stats = {"rounds": rounds, "max_conflicts": int(counts.max()),
             "collision_rate": (len(codes) - len(uniq)) / len(codes)}
    return suffix_dedup(codes), stats


# --------------------------------------------------------------------------- k-means (Lloyd)

def kmeans_lloyd(x: np.ndarray, init: np.ndarray, iters: int, tol: float = 1e-4) -> np.ndarray:
    """Lloyd iterations from given initial centres — the algorithm scikit-learn's ``KMeans``
    (third-party, pinned scikit-learn==1.7.1 in the reference's requirements.txt:7; call site
    layers.py:77) runs after seeding, including its relocation of empty clusters (sklearn
    ``_relocate_empty_clusters_dense``: the n_empty samples farthest from their centre become the new
    centres and leave their old cluster).  Pinned against scikit-learn itself on given ``init`` arrays
    (oracle/make_golden_kmeans.py → tests/golden/kmeans_sklearn.npz).  The SEEDING stays unpinned: sklearn's
    k-means++ consumes numpy's global RNG.  fp64 accumulation (scikit-learn: fp32 on centred data)."""
    x = np.asarray(x, dtype=np.float32)
    c = np.asarray(init, dtype=np.float32).copy()
    xv = float(np.mean(np.var(x.astype(np.float64), axis=0)))
    x64 = x.astype(np.float64)
    xx = (x64 ** 2).sum(1)
    for _ in range(iters):
        c64 = c.astype(np.float64)
        d = xx[:, None] + (c64 ** 2).sum(1)[None] - 2.0 * x64 @ c64.T
        a = d.argmin(1)
        K = c.shape[0]
        sums = np.zeros((K, x.shape[1]), dtype=np.float64)
        np.add.at(sums, a, x64)
        counts = np.bincount(a, minlength=K).astype(np.float64)
        empty = np.nonzero(counts == 0)[0]
        if len(empty):
            dist = ((x - c[a]) ** 2).sum(axis=1)
            if dist.max() > 0:
                far = np.argpartition(dist, -len(empty))[:-len(empty) - 1:-1]
                for new_id, far_idx in zip(empty, far):
                    old_id = a[far_idx]
                    sums[old_id] -= x64[far_idx]
                    sums[new_id] = x64[far_idx]
                    counts[new_id] = 1
                    counts[old_id] -= 1
        new = c.copy()
        nz = counts > 0
        new[nz] = (sums[nz] / counts[nz, None]).astype(np.float32)
        shift = float(((new.astype(np.float64) - c64) ** 2).sum())
        c = new
        if tol > 0 and shift <= tol * xv:
            break
    return c


# --------------------------------------------------------------------------- consumers (tokens, TIGER splits)

def item_to_offset_code(data: np.ndarray, item_id: int, codebook_size: int) -> List[int]:
    """RQVAE-T5/data_read.ipynb cell 2, verbatim semantics: 1-indexed item id, token = c + i * K + 1."""
    raw_code = data[item_id - 1]
    return [int(c + i * codebook_size + 1) for i, c in enumerate(raw_code)]


def tiger_splits(user_ids, item_lists, data: np.ndarray, codebook_size: int):
    """The leave-one-out / teacher-forcing split of the same cell → (train_list, test_list) of dicts."""
    f = lambda iid: item_to_offset_code(data, int(iid), codebook_size)
    train_list, test_list = [], []
    for uid, user_seq in zip(user_ids, item_lists):
        seq_len = len(user_seq)
        if seq_len < 2:
            continue
        if seq_len == 2:
            train_list.append({"user_id": int(uid), "history": [f(i) for i in user_seq[:1]],
                               "target": [f(i) for i in user_seq[1:]]})
        else:
            test_list.append({"user_id": int(uid), "history": [f(i) for i in user_seq[:-1]],
                              "target": [f(i) for i in user_seq[-1:]]})
            train_list.append({"user_id": int(uid), "history": [f(i) for i in user_seq[:-2]],
                               "target": [f(i) for i in user_seq[1:-1]]})
    return train_list, test_list


# --------------------------------------------------------------------------- training step (tolerance parity)

def train_step_grads(x, sd, cfg):
    """One forward + backward of the reference training step (train.py:112-115) in numpy, dropout off:
    returns (losses dict, indices [B, L], grads dict keyed like the state_dict).  Restates rqvae.py:60-84, rq.py:39-56,
    vq.py:63-99 and their autograd: straight-through estimator (identity to the latent through level 0 only — deeper
    levels cancel because r_{l+1} = r_l − (r_l + (q_l − r_l).detach())), commitment gradient β·2(z − q_0)/(B·e·L),
    codebook gradient 2(q − r)/(B·e·L) scattered by code.  TEST INFRASTRUCTURE; pinned against
    tests/golden/train_steps.npz (the unmodified reference loop) in tests/test_train_host.py."""
    f32 = np.float32
    nl = len(cfg["layers"]) + 1
    Ks, e, beta, w = cfg["num_emb_list"], cfg["e_dim"], f32(cfg["beta"]), f32(cfg["quant_loss_weight"])
    L = len(Ks)
    x = np.asarray(x, dtype=f32)
    B = x.shape[0]

    def mlp_fwd(h, name):
        acts = [h]
        for i in range(nl):
            W, b = sd[f"{name}.mlp_layers.{1 + 3 * i}.weight"], sd[f"{name}.mlp_layers.{1 + 3 * i}.bias"]
            h = (h @ W.T + b).astype(f32)
            if i < nl - 1:
                h = np.maximum(h, f32(0))
            acts.append(h)
        return acts

    def mlp_bwd(acts, name, dy, grads):
        for i in range(nl - 1, -1, -1):
            W = sd[f"{name}.mlp_layers.{1 + 3 * i}.weight"]
            if i < nl - 1:
                dy = dy * (acts[i + 1] > 0)
            grads[f"{name}.mlp_layers.{1 + 3 * i}.weight"] = (dy.T @ acts[i]).astype(f32)
            grads[f"{name}.mlp_layers.{1 + 3 * i}.bias"] = dy.sum(0).astype(f32)
            dy = (dy @ W).astype(f32)
        return dy

    enc = mlp_fwd(x, "encoder")
    z = enc[-1]
    r = z.copy()
    xq = np.zeros_like(z)
    idxs, resid, level_loss = [], [], []
    for l in range(L):
        cb = np.ascontiguousarray(sd[f"rq.vq_layers.{l}.embedding.weight"], dtype=f32)
        idx, _, _, d = quantize(r, [cb], want_xq=False, dist_level=0, threads=1)
        ind = idx[:, 0]
        if cfg["sk_epsilons"][l] > 0:
            ind = sinkhorn_assign(d, cfg["sk_epsilons"][l], cfg["sk_iters"])
        q = cb[ind]
        mse = np.mean((q - r).astype(np.float64) ** 2)
        level_loss.append(mse + float(beta) * mse)
        resid.append(r)
        idxs.append(ind)
        xres = r + (q - r)
        r = r - xres
        xq = xq + xres
    rq_loss = float(np.mean(level_loss))
    dec = mlp_fwd(xq.astype(f32), "decoder")
    out = dec[-1]
    diff = (out - x).astype(np.float64)
    recon = float(np.mean(diff ** 2)) if cfg["loss_type"] == "mse" else float(np.mean(np.abs(diff)))
    total = recon + float(w) * rq_loss
    grads = {}
    d_out = (2.0 * diff / diff.size if cfg["loss_type"] == "mse" else np.sign(diff) / diff.size).astype(f32)
    g_xq = mlp_bwd(dec, "decoder", d_out, grads)
    base = 2.0 / (B * e) / L * float(w)
    for l in range(L):
        cb = sd[f"rq.vq_layers.{l}.embedding.weight"]
        dE = np.zeros_like(cb, dtype=np.float64)
        np.add.at(dE, idxs[l], (cb[idxs[l]] - resid[l]).astype(np.float64))
        grads[f"rq.vq_layers.{l}.embedding.weight"] = (base * dE).astype(f32)
    cb0 = sd["rq.vq_layers.0.embedding.weight"]
    dz = (g_xq + base * float(beta) * (z - cb0[idxs[0]])).astype(f32)
    mlp_bwd(enc, "encoder", dz, grads)
    return {"loss": total, "recon": recon, "rq": rq_loss}, np.stack(idxs, -1).astype(np.int64), grads


def adamw_clip_update(sd, grads, state, lr, weight_decay, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8):
    """clip_grad_norm_(…, max_norm) + torch.optim.AdamW.step() (train.py:116-117) in numpy, in place on `sd`;
    `state` = {"step": int, "m": {}, "v": {}}.  Returns the gradient norm before clipping."""
    gn = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    coef = min(1.0, max_norm / (gn + 1e-6)) if max_norm > 0 else 1.0
    state["step"] += 1
    t = state["step"]
    b1, b2 = betas
    for k, g in grads.items():
        g = g.astype(np.float64) * coef
        m = state["m"].get(k, np.zeros_like(g))
        v = state["v"].get(k, np.zeros_like(g))
        m = m + (g - m) * (1 - b1)
        v = v * b2 + g * g * (1 - b2)
        state["m"][k], state["v"][k] = m, v
        p = sd[k].astype(np.float64) * (1 - lr * weight_decay)
        p -= (lr / (1 - b1 ** t)) * (m / (np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps))
        sd[k] = p.astype(np.float32)
    return gn
