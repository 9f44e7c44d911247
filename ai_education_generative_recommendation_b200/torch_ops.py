"""torch.ops.rqvae_b200.* — the thin PyTorch C++ extension of SURVEY.md §8b (csrc/torch_ops.cpp → librqvae_b200_torch.so,
built in-tree by csrc/build.py).  The ops are a second binding of the SAME C ABI (include/rqvae_b200.h) the ctypes layer
uses: borrowed CUDA tensors in, fresh tensors from the caller's caching allocator out, kernels enqueued on the current
stream, TORCH_CHECK errors.  `handle(model)` turns a Python RQVAE into the integer the ops take."""
from __future__ import annotations

import os

import torch

from . import _cabi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librqvae_b200_torch.so")
_loaded = False


def load() -> None:
    """Registers the ops (idempotent).  Raises when the extension has not been built: there is no fallback op set."""
    global _loaded
    if _loaded:
        return
    if not os.path.exists(LIB_PATH):
        raise _cabi.RQB200Error(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first")
    _cabi.lib()                                   # librqvae_b200.so first: the op library links against it
    torch.ops.load_library(LIB_PATH)
    _loaded = True


_unavailable = False


def available() -> bool:
    """True once the op library is registered; a failed load is remembered (the ctypes binding of the same C ABI is then
    what the module mirror calls)."""
    global _unavailable
    if _loaded:
        return True
    if _unavailable:
        return False
    if os.environ.get("RQB200_LIB"):              # A/B build of the C-ABI library (tools/build_variant.sh): the op library is linked
        _unavailable = True                       # against the regular one, so such runs go through the ctypes binding
        return False
    try:
        load()
    except Exception:
        _unavailable = True
        return False
    return True


def handle(model) -> int:
    """The rqb200_model* of a Python RQVAE as the `int handle` argument of the ops (weights synchronised first)."""
    model._sync()
    h = model._handle
    return int(h.value if hasattr(h, "value") else h)


OPS = ("encode_indices", "encode_latents", "quantize", "sinkhorn_assign", "resolve_collisions", "collision_rate")
