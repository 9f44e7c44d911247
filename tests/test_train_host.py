"""Host-side pieces of the training driver that need no GPU: the schedule restated from transformers
(train.py:80-91), the golden fixture's integrity, and the no-CPU-fallback rule of the training path."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, train_case_state


def test_warmup_lambda_equals_transformers_schedules():
    transformers = pytest.importorskip("transformers")
    from ai_education_generative_recommendation_b200.trainer import warmup_lambda
    for kind, make in (("linear", lambda o, w, t: transformers.get_linear_schedule_with_warmup(o, w, t)),
                       ("constant", lambda o, w, t: transformers.get_constant_schedule_with_warmup(o, w))):
        for warm, total in ((0, 10), (5, 40), (7, 7), (3, 100)):
            p = torch.nn.Parameter(torch.zeros(1))
            opt = torch.optim.SGD([p], lr=0.25)
            sched = make(opt, warm, total)
            lam = warmup_lambda(kind, warm, total)
            for step in range(total + 3):
                assert abs(sched.get_last_lr()[0] - 0.25 * lam(step)) < 1e-15, (kind, warm, total, step)
                opt.step()
                sched.step()


def test_training_fixture_is_consistent_with_its_generator_inputs():
    g = np.load(os.path.join(GOLD, "train_steps.npz"))
    cases = json.loads(str(g["cases"]))
    assert set(cases) == {"sk_mse", "argmin_l1", "c1_shape"}
    for name, cfg in cases.items():
        x, sd = train_case_state(cfg)
        assert x.shape == (cfg["batch"], cfg["in_dim"]) and x.dtype == np.float32
        names = [str(s) for s in g[f"{name}/names"]]
        assert names == list(sd.keys()) or set(names) == set(sd.keys())
        assert g[f"{name}/loss"].shape == (cfg["steps"],)
        assert g[f"{name}/lr"][0] == 0.0                       # LambdaLR starts the warmup at factor 0 (train.py:84-86)
        assert abs(g[f"{name}/loss"][0] - g[f"{name}/loss"][1]) < 1e-12    # … so the first step changes nothing
        assert g[f"{name}/codes"].shape == (cfg["steps"], cfg["batch"], len(cfg["num_emb_list"]))
        tot = g[f"{name}/recon"] + cfg["quant_loss_weight"] * g[f"{name}/rq"]        # rqvae.py:82
        assert np.allclose(tot, g[f"{name}/loss"], rtol=1e-6)


def test_training_forward_refuses_cpu_tensors():
    import ai_education_generative_recommendation_b200 as rq
    m = rq.RQVAE(in_dim=16, num_emb_list=[4, 4], e_dim=4, layers=[8], sk_epsilons=[0.0, 0.0]).train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(3, 16))
    with pytest.raises(ValueError, match="incompatible loss type"):
        m.loss_type = "huber"
        m.compute_loss(torch.zeros(1), torch.zeros(()), xs=torch.zeros(1))


def test_emb_dataset_npy_and_shards(tmp_path):
    """vision_data.py:9-30 contract on the .npy stand-in: float32 rows, dim, meta side file, contiguous shards."""
    from ai_education_generative_recommendation_b200 import EmbDataset
    x = np.arange(35 * 6, dtype=np.float64).reshape(35, 6)
    np.save(tmp_path / "embs.npy", x)
    (tmp_path / "embs_meta.json").write_text('{"num_items": 35}')
    ds = EmbDataset(str(tmp_path / "embs.npy"))
    assert len(ds) == 35 and ds.dim == 6 and ds.meta == {"num_items": 35}
    assert ds[3].dtype == torch.float32 and torch.equal(ds[3], torch.from_numpy(x[3].astype(np.float32)))
    parts = [ds.shard(r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), x.astype(np.float32)) and parts[0].dtype == np.float32


@pytest.mark.parametrize("name", ["sk_mse", "argmin_l1", "c1_shape"])
def test_oracle_training_step_matches_reference_golden(oracle, name):
    """The numpy restatement of the training step (forward, autograd algebra of the straight-through quantizer, clip +
    AdamW, warmup schedule) reproduces the unmodified reference loop recorded in train_steps.npz."""
    from ai_education_generative_recommendation_b200.trainer import warmup_lambda
    g = np.load(os.path.join(GOLD, "train_steps.npz"))
    cfg = json.loads(str(g["cases"]))[name]
    x, sd = train_case_state(cfg)
    init = {k: v.copy() for k, v in sd.items()}
    lam = warmup_lambda("linear", cfg["warmup_steps"], cfg["max_steps"])
    state = {"step": 0, "m": {}, "v": {}}
    names = [str(s) for s in g[f"{name}/names"]]
    for step in range(cfg["steps"]):
        losses, idx, grads = oracle.train_step_grads(x, sd, cfg)
        assert np.array_equal(idx, g[f"{name}/codes"][step].astype(np.int64)), step
        for key in ("loss", "recon", "rq"):
            assert abs(losses[key] - g[f"{name}/{key}"][step]) <= 1e-4 * abs(g[f"{name}/{key}"][step]), (key, step)
        if step == 0:
            for i, n in enumerate(names):
                if f"{name}/grad0/{n}" in g:
                    ref = g[f"{name}/grad0/{n}"]
                    assert np.abs(grads[n] - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-30), n
                else:
                    ref = g[f"{name}/grad0_norms"][i]
                    assert abs(np.linalg.norm(grads[n].astype(np.float64)) - ref) <= 1e-4 * max(ref, 1e-12), n
        gn = oracle.adamw_clip_update(sd, grads, state, cfg["lr"] * lam(step), cfg["weight_decay"])
        assert abs(gn - g[f"{name}/gnorm"][step]) <= (1e-4 if step <= 1 else 2e-3) * g[f"{name}/gnorm"][step]
    for i, n in enumerate(names):
        if f"{name}/final/{n}" in g:
            ref = g[f"{name}/final/{n}"].astype(np.float64)
            upd = np.linalg.norm(ref - init[n])
            assert np.linalg.norm(sd[n] - ref) <= 2e-3 * max(upd, 1e-12), n
        else:
            ref = g[f"{name}/delta_norms"][i]
            assert abs(np.linalg.norm(sd[n].astype(np.float64) - init[n]) - ref) <= 2e-3 * max(ref, 1e-12), n
