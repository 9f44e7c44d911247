"""Seeded model / fixture helpers shared by bench.py, __graft_entry__.smoke(), tools/ and the tests.

Nothing here computes on the product path: these build a product RQVAE from the integer-hash weights of synth.py
(the same bytes the golden fixtures were generated from) and read the committed fixtures under tests/golden/.
"""
from __future__ import annotations

import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    """(npz, cfg dict, codebooks list) of tests/golden/<name>.npz."""
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    cfg = json.loads(str(g["cfg"]))
    cbs = [g[f"codebook{l}"] for l in range(len(cfg["num_emb_list"]))]
    return g, cfg, cbs


def synth_weights(cfg, seed=2024):
    """(state_dict, (encoder weights, biases), (decoder weights, biases)) as numpy arrays."""
    from . import synth
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
    from . import RQVAE
    sd, _, _ = synth_weights(cfg, seed)
    for l, c in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = c
    m = RQVAE(in_dim=cfg["in_dim"], num_emb_list=cfg["num_emb_list"], e_dim=cfg["e_dim"], layers=cfg["layers"],
              dropout_prob=0.0, bn=False, loss_type="mse", quant_loss_weight=cfg.get("quant_loss_weight", 1.0),
              kmeans_init=False, kmeans_iters=cfg.get("kmeans_iters", 10), sk_epsilons=list(cfg["sk_epsilons"]),
              sk_iters=cfg["sk_iters"])
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
    return m.to(device).eval()


def infer_fixture(name):
    """A reference `infer()` fixture (oracle/make_golden.py: make_infer): returns a dict with
    cfg, codebooks, x (the catalogue rows incl. the planted duplicates), semantic_ids (the reference's saved array),
    trace (list of [n, L] int64 codes: after pass 1 and after every re-encode round of the reference) and
    ties (per round: items whose Sinkhorn arg-max is an fp64 tie in the reference's own matrix)."""
    from . import synth
    g, cfg, cbs = load_golden(name)
    n, n_total = int(g["n"]) if "n" in g else int(g["n_total"]), int(g["n_total"])
    x = synth.synth_items(int(g["seed"]), 0, n, cfg["in_dim"], n_total)
    for dst, src, cnt in g["dup"].astype(np.int64).reshape(-1, 3):
        x[dst:dst + cnt] = x[src:src + cnt]
    cur = g["trace0"].astype(np.int64)
    trace = [cur]
    off, items, codes = g["chg_offsets"], g["chg_items"], g["chg_codes"]
    for t in range(int(g["rounds"])):
        cur = cur.copy()
        cur[items[off[t]:off[t + 1]]] = codes[off[t]:off[t + 1]]
        trace.append(cur)
    toff, titems = g["tie_offsets"], g["tie_items"]
    ties = [titems[toff[t]:toff[t + 1]].astype(np.int64) for t in range(len(toff) - 1)]
    return {"cfg": cfg, "codebooks": cbs, "x": x, "semantic_ids": g["semantic_ids"].astype(np.int64), "trace": trace,
            "ties": ties, "rounds": int(g["rounds"]), "n": n}
