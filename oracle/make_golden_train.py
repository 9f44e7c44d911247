"""Generates tests/golden/train_steps.npz by running the UNMODIFIED reference training step
(/root/reference/RQ-VAE: models.rqvae.RQVAE + the loop body of train.py:108-121 with its optimizer and schedule)
on CPU in the build container.

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists):  python oracle/make_golden_train.py
The GPU box never runs this; it only reads the committed fixture.

Pinned per case (tolerance parity, north_star: 1e-4 relative): per-step total / reconstruction / quantizer loss, the
codes chosen (Sinkhorn arg-max or arg-min), every gradient of step 0 before clipping, the gradient norm per step and
every parameter after the last step (AdamW, clip 1.0, linear warmup schedule).  Inputs, encoder / decoder weights and
codebooks are regenerated from integer hashing (package synth.py), so the fixture only stores results.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/RQ-VAE"
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from ai_education_generative_recommendation_b200 import synth              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: model config, batch rows, steps, optimizer / schedule
    "sk_mse": dict(in_dim=96, layers=[64, 32], e_dim=16, num_emb_list=[16, 16, 16], sk_epsilons=[0.01, 0.01, 0.01],
                   sk_iters=50, loss_type="mse", quant_loss_weight=0.1, beta=0.25, batch=160, steps=4, lr=1e-3,
                   weight_decay=1e-4, warmup_steps=2, max_steps=8, cb_scale=0.15),
    "argmin_l1": dict(in_dim=128, layers=[32], e_dim=8, num_emb_list=[8, 8], sk_epsilons=[0.0, 0.0], sk_iters=50,
                      loss_type="l1", quant_loss_weight=1.0, beta=0.5, batch=96, steps=3, lr=2e-3, weight_decay=1e-2,
                      warmup_steps=1, max_steps=5, cb_scale=0.2),
    "c1_shape": dict(in_dim=768, layers=[256, 128], e_dim=32, num_emb_list=[8, 8, 8], sk_epsilons=[0.01, 0.01, 0.01],
                     sk_iters=50, loss_type="mse", quant_loss_weight=0.1, beta=0.25, batch=64, steps=3, lr=1e-3,
                     weight_decay=1e-4, warmup_steps=1, max_steps=6, cb_scale=0.1, summary_only=True),
}


sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import train_case_state as case_state                         # noqa: E402  (shared with the GPU tests)


def run_case(cfg):
    from models.rqvae import RQVAE
    from transformers import get_linear_schedule_with_warmup
    x_np, sd = case_state(cfg)
    torch.manual_seed(0)
    m = RQVAE(in_dim=cfg["in_dim"], num_emb_list=cfg["num_emb_list"], e_dim=cfg["e_dim"], layers=cfg["layers"],
              dropout_prob=0.0, bn=False, loss_type=cfg["loss_type"], quant_loss_weight=cfg["quant_loss_weight"],
              beta=cfg["beta"], kmeans_init=False, kmeans_iters=10, sk_epsilons=cfg["sk_epsilons"],
              sk_iters=cfg["sk_iters"])
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])        # train.py:75-78
    sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=cfg["warmup_steps"],
                                            num_training_steps=cfg["max_steps"])                    # train.py:84-86
    x = torch.from_numpy(x_np)
    names = [n for n, _ in m.named_parameters()]
    rec = {"loss": [], "recon": [], "rq": [], "gnorm": [], "lr": [], "codes": []}
    grads0 = None
    for step in range(cfg["steps"]):                       # loop body of train.py:108-121
        opt.zero_grad()
        out, rq_loss, indices = m(x)
        loss, loss_recon = m.compute_loss(out, rq_loss, xs=x)
        loss.backward()
        if step == 0:
            grads0 = {n: p.grad.detach().clone().numpy() for n, p in m.named_parameters()}
        gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        rec["lr"].append(sched.get_last_lr()[0])
        opt.step()
        sched.step()
        rec["loss"].append(loss.item()); rec["recon"].append(loss_recon.item()); rec["rq"].append(rq_loss.item())
        rec["gnorm"].append(float(gn)); rec["codes"].append(indices.numpy().copy())
    out = {"names": np.array(names), "loss": np.array(rec["loss"], np.float64), "recon": np.array(rec["recon"], np.float64),
           "rq": np.array(rec["rq"], np.float64), "gnorm": np.array(rec["gnorm"], np.float64),
           "lr": np.array(rec["lr"], np.float64), "codes": np.stack(rec["codes"]).astype(np.int16)}
    final = {n: p.detach().numpy() for n, p in m.named_parameters()}
    if cfg.get("summary_only"):                            # large shapes: norms instead of full tensors
        out["grad0_norms"] = np.array([np.linalg.norm(grads0[n].astype(np.float64)) for n in names])
        out["final_norms"] = np.array([np.linalg.norm(final[n].astype(np.float64)) for n in names])
        out["delta_norms"] = np.array([np.linalg.norm(final[n].astype(np.float64) - sd[n].astype(np.float64)) for n in names])
    else:
        for n in names:
            out["grad0/" + n] = grads0[n]
            out["final/" + n] = final[n]
    return out


def main():
    blob = {"cases": np.array(json.dumps(CASES))}
    for name, cfg in CASES.items():
        res = run_case(cfg)
        for k, v in res.items():
            blob[f"{name}/{k}"] = v
        print(name, "loss", res["loss"], "gnorm", res["gnorm"], "lr", res["lr"])
    path = os.path.join(GOLD, "train_steps.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
