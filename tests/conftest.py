import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    cfg = json.loads(str(g["cfg"]))
    cbs = [g[f"codebook{l}"] for l in range(len(cfg["num_emb_list"]))]
    return g, cfg, cbs


def synth_weights(cfg, seed=2024):
    from ai_education_generative_recommendation_b200 import synth
    sd = synth.synth_state_dict(seed, cfg["in_dim"], cfg["layers"], cfg["e_dim"], cfg["num_emb_list"])
    n = len(cfg["layers"]) + 1
    enc = ([sd[f"encoder.mlp_layers.{1 + 3 * i}.weight"] for i in range(n)],
           [sd[f"encoder.mlp_layers.{1 + 3 * i}.bias"] for i in range(n)])
    dec = ([sd[f"decoder.mlp_layers.{1 + 3 * i}.weight"] for i in range(n)],
           [sd[f"decoder.mlp_layers.{1 + 3 * i}.bias"] for i in range(n)])
    return sd, enc, dec


def build_model(cfg, cbs, device="cuda:0", seed=2024):
    """Product RQVAE loaded exactly like a reference checkpoint would be (load_state_dict)."""
    import torch
    from ai_education_generative_recommendation_b200 import RQVAE
    sd, _, _ = synth_weights(cfg, seed)
    for l, c in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = c
    m = RQVAE(in_dim=cfg["in_dim"], num_emb_list=cfg["num_emb_list"], e_dim=cfg["e_dim"], layers=cfg["layers"],
              dropout_prob=0.0, bn=False, loss_type="mse", quant_loss_weight=cfg.get("quant_loss_weight", 1.0),
              kmeans_init=False, kmeans_iters=cfg.get("kmeans_iters", 10), sk_epsilons=list(cfg["sk_epsilons"]),
              sk_iters=cfg["sk_iters"])
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
    return m.to(device).eval()


def train_case_state(cfg, seed=2024):
    """(x[batch, in], reference-format state_dict as numpy arrays) of a tests/golden/train_steps.npz case — the same
    function oracle/make_golden_train.py fed to the reference."""
    from ai_education_generative_recommendation_b200 import synth
    sd = synth.synth_state_dict(seed, cfg["in_dim"], cfg["layers"], cfg["e_dim"], cfg["num_emb_list"])
    for l, K in enumerate(cfg["num_emb_list"]):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = synth.synth_matrix(seed, 200 + l, K, cfg["e_dim"], cfg["cb_scale"])
    x = synth.synth_items(seed, 1, cfg["batch"], cfg["in_dim"], 1_000_000)
    return x, sd


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
