"""B200-native RQ-VAE semantic-ID encode path (drop-in for the reference's RQ-VAE/models + infer.py).

    from ai_education_generative_recommendation_b200 import RQVAE, infer, generate_codes

Every compute call goes through `librqvae_b200.so` (hand-written sm_100a CUDA behind a C ABI);
there is no CPU fallback.
"""
from . import _cabi, torch_ops
from ._cabi import ENCODE_EXACT, ENCODE_FAST, RQB200Error
from .rqvae import (MLPLayers, RQVAE, ResidualVectorQuantizer, VectorQuantizer, activation_layer, kmeans,
                    sinkhorn_algorithm)
from .generate_code import collision_groups, encode_latents, generate_codes, infer, suffix_dedup
from .dataset import EmbDataset
from .consumers import build_tiger_splits, gather_item_tokens, offset_codes
from .trainer import DeviceBatches, Trainer, train, warmup_lambda
from .train_ops import FusedAdamW

__all__ = ["RQVAE", "MLPLayers", "ResidualVectorQuantizer", "VectorQuantizer", "activation_layer", "kmeans",
           "sinkhorn_algorithm", "generate_codes", "infer", "suffix_dedup", "collision_groups", "encode_latents",
           "EmbDataset", "offset_codes", "gather_item_tokens", "build_tiger_splits", "Trainer", "train", "DeviceBatches", "FusedAdamW", "warmup_lambda", "ENCODE_EXACT", "ENCODE_FAST", "RQB200Error"]
