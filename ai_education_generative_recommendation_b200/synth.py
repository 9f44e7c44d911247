"""Synthetic catalogue + weights, numpy twin of csrc/synth.cu (SURVEY.md §8d "Synthetic inputs").

Everything is a pure function of (seed, row, column) built from integer hashing and single IEEE
roundings, so `synth_items` here and `rqb200_synth_items` on the GPU give identical bytes.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u64(x):
    return np.asarray(x, dtype=np.uint64)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = _u64(x) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def hash3(seed, a, b):
    with np.errstate(over="ignore"):
        return splitmix64(splitmix64(np.uint64(seed) ^ (_u64(a) * np.uint64(0xD6E8FEB86659FD93)))
                          ^ (_u64(b) * np.uint64(0xA0761D6478BD642F)))


def approx_normal(bits):
    bits = _u64(bits)
    s = ((bits & np.uint64(0xFFFF)) + ((bits >> np.uint64(16)) & np.uint64(0xFFFF))
         + ((bits >> np.uint64(32)) & np.uint64(0xFFFF)) + (bits >> np.uint64(48))).astype(np.float32)
    return (s - np.float32(131070.0)) * np.float32(2.6429153e-05)


def synth_items(seed: int, first_row: int, n: int, dim: int, n_total: int) -> np.ndarray:
    """Rows [first_row, first_row+n) of the synthetic catalogue of n_total items, float32 [n, dim]."""
    rows = np.arange(first_row, first_row + n, dtype=np.uint64)
    n_centres = np.uint64(max(n_total // 1000, 40))
    safe = np.where(rows == 0, np.uint64(1), rows)
    dup = (hash3(seed, 1, rows) % np.uint64(1000)) == 0
    src = np.where(dup, np.uint64(1) + hash3(seed, 2, rows) % safe, rows)
    c = hash3(seed, 3, src) % n_centres
    k = np.arange(dim, dtype=np.uint64)[None, :]
    mu = approx_normal(hash3(seed, (np.uint64(0x100000000) + c)[:, None], k))
    ep = approx_normal(hash3(seed, (np.uint64(0x200000000) + src)[:, None], k))
    x = np.float32(0.5) * mu + np.float32(0.1) * ep
    x[rows == 0] = 0.0
    return np.ascontiguousarray(x, dtype=np.float32)


def synth_matrix(seed: int, tag: int, rows: int, cols: int, scale: float) -> np.ndarray:
    """Deterministic approx-normal matrix (weights / biases / codebook seeds) of std `scale`."""
    r = np.arange(rows, dtype=np.uint64)[:, None]
    c = np.arange(cols, dtype=np.uint64)[None, :]
    return (approx_normal(hash3(seed, np.uint64(tag) * np.uint64(0x10000000) + r, c)) * np.float32(scale)).astype(np.float32)


def synth_state_dict(seed: int, in_dim: int, layers, e_dim: int, num_emb_list, bias_scale: float = 0.02):
    """A reference-format state_dict (numpy arrays) with xavier-scaled weights; codebooks are left to
    the caller (k-means init or `codebooks_from_latents`)."""
    dims = [in_dim] + list(layers) + [e_dim]
    sd = {}
    for name, dd in (("encoder", dims), ("decoder", dims[::-1])):
        for i, (a, b) in enumerate(zip(dd[:-1], dd[1:])):
            pos = 1 + 3 * i            # [Dropout, Linear, ReLU] per layer (bn=False), layers.py:18-32
            std = (2.0 / (a + b)) ** 0.5
            tag = (1 if name == "encoder" else 2) * 16 + i
            sd[f"{name}.mlp_layers.{pos}.weight"] = synth_matrix(seed, tag, b, a, std)
            sd[f"{name}.mlp_layers.{pos}.bias"] = synth_matrix(seed, tag + 8, 1, b, bias_scale)[0]
    for lvl, K in enumerate(num_emb_list):
        sd[f"rq.vq_layers.{lvl}.embedding.weight"] = np.zeros((K, e_dim), dtype=np.float32)
    return sd
