"""ctypes binding of include/rqvae_b200.h — the only way the Python host reaches the CUDA kernels.

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
compute entry point is called, this module raises.  (Loading the library and listing its symbols
works without a GPU; that is what the CPU-side tests check.)
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RQB200_LIB") or os.path.join(_PKG, "librqvae_b200.so")   # RQB200_LIB: A/B builds (tools/)
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "rqvae_b200.h")

ENCODE_EXACT = 0
ENCODE_FAST = 1


class RQB200Error(RuntimeError):
    pass


_lib = None

_P = c_void_p          # device / host pointers travel as integers
_PROTOS = {
    "rqb200_abi_version": (c_int, []),
    "rqb200_last_error": (c_char_p, []),
    "rqb200_device_count": (c_int, []),
    "rqb200_launch_count": (ctypes.c_longlong, []),
    "rqb200_profile_enable": (c_int, [c_int]),
    "rqb200_debug_tc_trace": (c_int, [_P]),
    "rqb200_debug_tc_flags": (c_int, [c_int]),
    "rqb200_profile_read": (c_int, [POINTER(c_double), POINTER(ctypes.c_longlong), c_int]),
    "rqb200_model_create": (c_int, [POINTER(c_void_p), c_int, c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    "rqb200_model_destroy": (None, [c_void_p]),
    "rqb200_model_set_linear": (c_int, [c_void_p, c_int, c_int, _P, _P, POINTER(c_int), c_int]),
    "rqb200_model_set_codebook": (c_int, [c_void_p, c_int, _P]),
    "rqb200_model_get_codebook": (c_int, [c_void_p, c_int, _P]),
    "rqb200_model_set_gate": (c_int, [c_void_p, c_float, c_float]),
    "rqb200_model_set_screen": (c_int, [c_void_p, c_int, c_float]),
    "rqb200_model_last_tier_rows": (c_int, [c_void_p, POINTER(c_int64)]),
    "rqb200_debug_linear_tc": (c_int, [c_void_p, c_int, c_int, _P, c_int64, _P, c_int, c_int, _P]),
    "rqb200_debug_check_division": (c_int, [ctypes.c_uint64, c_int, ctypes.c_uint64, POINTER(ctypes.c_uint64), _P]),
    "rqb200_debug_mlp_tc": (c_int, [c_void_p, c_int, _P, c_int64, _P, c_int, _P]),
    "rqb200_mlp_tc": (c_int, [c_void_p, c_int, _P, c_int64, _P, _P]),
    "rqb200_mlp_exact": (c_int, [c_void_p, c_int, _P, _P, c_int64, _P, _P]),
    "rqb200_quantize": (c_int, [c_void_p, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "rqb200_get_indices": (c_int, [c_void_p, c_int, _P, c_int64, _P, _P, POINTER(c_int64), _P]),
    "rqb200_forward": (c_int, [c_void_p, _P, c_int64, _P, _P, _P, _P, _P]),
    "rqb200_sinkhorn_regroup": (c_int, [c_void_p, _P, _P, _P, c_int64, c_int, c_double, c_int, _P, _P]),
    "rqb200_reencode_groups": (c_int, [c_void_p, _P, c_int, _P, _P, c_int64, c_int64, _P, _P, _P]),
    "rqb200_debug_sinkhorn_variant": (c_int, [c_int]),
    "rqb200_model_levels": (c_int, [c_void_p]),
    "rqb200_model_e_dim": (c_int, [c_void_p]),
    "rqb200_reencode_classes": (c_int, [c_void_p]),
    "rqb200_reencode_groups_memo": (c_int, [c_void_p, _P, _P, _P, c_int64, c_int64, _P, _P, c_int64, _P, _P, _P, _P]),
    "rqb200_reencode_rows": (c_int, [c_void_p, _P, c_int, _P, _P, c_int64, _P, _P, _P]),
    "rqb200_sinkhorn_group_cap": (c_int, [c_void_p]),
    "rqb200_sinkhorn_regroup_large": (c_int, [c_void_p, _P, _P, _P, _P, _P, c_int64, _P, c_double, c_int, _P, _P]),
    "rqb200_sinkhorn": (c_int, [_P, c_int64, c_int, c_double, c_int, _P]),
    "rqb200_sinkhorn_assign": (c_int, [_P, c_int64, c_int, c_double, c_int, _P, _P, _P]),
    "rqb200_distances": (c_int, [c_void_p, c_int, _P, c_int64, _P, _P]),
    "rqb200_suffix_dedup": (c_int, [c_void_p, _P, c_int64, c_int, POINTER(c_int), _P, POINTER(c_int64),
                                    POINTER(c_int64), _P]),
    "rqb200_collision_groups": (c_int, [c_void_p, _P, c_int64, c_int, POINTER(c_int), _P, _P, POINTER(c_int64),
                                        POINTER(c_int64), POINTER(c_int64), _P]),
    "rqb200_collision_groups_changed": (c_int, [c_void_p, _P, c_int64, c_int, POINTER(c_int), _P, _P, c_int, _P, _P,
                                                POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), _P]),
    "rqb200_pack_keys": (c_int, [_P, c_int64, c_int, POINTER(c_int), _P, _P]),
    "rqb200_sort_pairs": (c_int, [c_void_p, _P, _P, c_int64, c_int, _P]),
    "rqb200_segment_rank": (c_int, [c_void_p, _P, c_int64, _P, _P]),
    "rqb200_shard_create": (c_int, [POINTER(c_void_p), c_void_p, c_int, c_int, c_int64, c_int64]),
    "rqb200_shard_handle_bytes": (c_int, []),
    "rqb200_shard_get_handle": (c_int, [c_void_p, _P]),
    "rqb200_shard_connect": (c_int, [c_void_p, _P]),
    "rqb200_shard_suffix_dedup": (c_int, [c_void_p, _P, c_int64, c_int, POINTER(c_int), _P, _P]),
    "rqb200_shard_destroy": (None, [c_void_p]),
    "rqb200_kmeans_assign": (c_int, [_P, c_int64, c_int, _P, c_int, _P, _P, _P]),
    "rqb200_kmeans_distances": (c_int, [_P, c_int64, c_int, _P, c_int, _P, _P, _P]),
    "rqb200_kmeans_accumulate": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P, _P, _P]),
    "rqb200_kmeans_update": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "rqb200_dropout": (c_int, [_P, c_int64, c_float, c_uint64, _P, _P]),
    "rqb200_dropout_dev": (c_int, [_P, c_int64, c_float, c_uint64, _P, _P, _P]),
    "rqb200_set_floats": (c_int, [_P, c_int, POINTER(c_float), _P]),
    "rqb200_adamw_clip_step_dev": (c_int, [_P, c_int, _P, _P, c_float, c_float, _P, _P]),
    "rqb200_linear_forward": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "rqb200_linear_backward_scratch_floats": (c_int64, [c_int64, c_int, c_int]),
    "rqb200_linear_backward": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, c_int64, _P]),
    "rqb200_rq_level_apply": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, _P]),
    "rqb200_vq_codebook_grad": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P]),
    "rqb200_rq_latent_grad": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, _P, _P, _P]),
    "rqb200_recon_loss": (c_int, [_P, _P, c_int64, _P, _P]),
    "rqb200_recon_grad": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P]),
    "rqb200_adamw_clip_step": (c_int, [_P, c_int, _P, _P, c_float, c_float, c_float, c_float, c_float, c_float,
                                       c_float, c_int64, _P]),
    "rqb200_offset_tokens": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "rqb200_gather_item_tokens": (c_int, [_P, c_int64, c_int, _P, c_int64, _P, _P, _P]),
    "rqb200_synth_items": (c_int, [c_uint64, c_int64, c_int64, c_int, c_int64, _P, _P]),
    "rqb200_generate_codes_host": (c_int, [c_void_p, c_int, _P, c_int64, c_int64, _P, POINTER(c_int64), _P]),
}


def declared_symbols(header_path: str = HEADER_PATH):
    """Names of every function include/rqvae_b200.h declares."""
    with open(header_path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rqb200_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    """Load librqvae_b200.so (built in-tree by csrc/build.py); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RQB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for this path)")
        import torch  # noqa: F401  (brings libcudart.so.12 into the process before our library needs it)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.rqb200_abi_version() != 1:
            raise RQB200Error("ABI version mismatch between _cabi.py and librqvae_b200.so")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().rqb200_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        if rc == -3:
            raise MemoryError(f"rqvae_b200: {msg}")
        raise RQB200Error(f"rqvae_b200 error {rc}: {msg}")


def require_cuda() -> None:
    if lib().rqb200_device_count() <= 0:
        raise RQB200Error("no CUDA device visible: the RQ-VAE encode path is CUDA-only (sm_100a), "
                          "there is no CPU fallback")


def ptr(t) -> int:
    """Device (or host) address of a contiguous torch tensor, or 0 for None."""
    if t is None:
        return 0
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def int_array(values):
    arr = (c_int * len(values))(*[int(v) for v in values])
    return arr
