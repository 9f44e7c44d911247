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


from ai_education_generative_recommendation_b200.fixtures import build_model, infer_fixture, load_golden, synth_weights  # noqa: E402,F401


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
