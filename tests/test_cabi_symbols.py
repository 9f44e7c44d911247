"""CPU: the C-ABI library loads, exports every symbol the header declares, and the product never routes
through the oracle or a CPU path."""
import ctypes
import os
import re

import pytest
import torch

import ai_education_generative_recommendation_b200 as rq
from ai_education_generative_recommendation_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _cabi.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rqvae_b200.h but not exported"
    assert sorted(_cabi._PROTOS) == declared            # the ctypes table covers the whole header
    assert _cabi.lib().rqb200_abi_version() == 1


def test_torch_library_ops_are_registered():
    """csrc/torch_ops.cpp: the TORCH_LIBRARY layer over the C ABI loads without a GPU, declares the ops of SURVEY.md §8b with
    CUDA kernels only (a CPU tensor has no implementation to fall back on)."""
    from ai_education_generative_recommendation_b200 import torch_ops
    torch_ops.load()
    for name in torch_ops.OPS:
        assert hasattr(torch.ops.rqvae_b200, name), name
        dump = torch._C._dispatch_dump(f"rqvae_b200::{name}")
        assert "CUDA: registered" in dump and "CPU: registered" not in dump, dump
    with pytest.raises(NotImplementedError):
        torch.ops.rqvae_b200.sinkhorn_assign(torch.zeros(2, 8), 0.003, 50)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "ai_education_generative_recommendation_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "librqvae_oracle" not in text, f


def test_cpu_tensors_are_refused():
    m = rq.RQVAE(in_dim=64, num_emb_list=[8, 8], e_dim=16, layers=[32], sk_epsilons=[0.0, 0.0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.get_indices(torch.zeros(4, 64))
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 64), use_sk=False)
    with pytest.raises(RuntimeError):
        rq.sinkhorn_algorithm(torch.zeros(4, 8, dtype=torch.float64), 0.01, 5)
    with pytest.raises(RuntimeError):
        rq.suffix_dedup(None, torch.zeros(4, 3, dtype=torch.int64), [8, 8, 8])


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_model_create_fails_loudly_without_gpu():
    h = ctypes.c_void_p(None)
    rc = _cabi.lib().rqb200_model_create(ctypes.byref(h), 0, 1, _cabi.int_array([8, 8]), 1, _cabi.int_array([2]))
    assert rc != 0 and b"no CUDA device" in _cabi.lib().rqb200_last_error()
    m = rq.RQVAE(in_dim=64, num_emb_list=[8], e_dim=16, layers=[32], sk_epsilons=[0.0])
    with pytest.raises(RuntimeError):
        m._ensure_handle()


def test_state_dict_layout_matches_reference_checkpoints():
    m = rq.RQVAE(in_dim=768, num_emb_list=[8, 8, 8], e_dim=32, layers=[256, 128], sk_epsilons=[0.01] * 3)
    keys = list(m.state_dict().keys())
    expect = [f"encoder.mlp_layers.{i}.{p}" for i in (1, 4, 7) for p in ("weight", "bias")]
    expect += [f"rq.vq_layers.{l}.embedding.weight" for l in range(3)]
    expect += [f"decoder.mlp_layers.{i}.{p}" for i in (1, 4, 7) for p in ("weight", "bias")]
    assert keys == expect
    assert m.encoder.mlp_layers[1].weight.shape == (256, 768)
    assert m.decoder.mlp_layers[7].weight.shape == (768, 256)
    m.rq.vq_layers[0].sk_epsilon = 0.0                    # the attribute poke of infer.py:109-110
    mb = rq.RQVAE(in_dim=768, num_emb_list=[8], e_dim=32, layers=[256, 128], bn=True, sk_epsilons=[0.0])
    assert "encoder.mlp_layers.2.running_mean" in mb.state_dict() and "encoder.mlp_layers.5.weight" in mb.state_dict()
    with pytest.raises(ValueError, match="incompatible loss type"):
        rq.RQVAE(in_dim=8, num_emb_list=[2], e_dim=8, layers=[8], loss_type="huber", sk_epsilons=[0.0]).compute_loss(
            torch.zeros(2, 8), torch.zeros(()), xs=torch.zeros(2, 8))
